"""Throughput of the BASELINE configs other than the one bench.py times (configs[2] Lyon-8 on
10 M candidates, configs[3] PFD shape, configs[4] survey 1024-bin shape) on one B200.  Seeded
synthetic data of the stated shapes; a few thousand generated candidates are tiled to the batch
size (host generation of millions of candidates would dominate the run).  Prints one JSON line
per config.  usage: python scripts/gpu_configs.py [cli] [lyon] [pfd] [survey]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from pulsarfeatureextractor_b200 import _lib, synth
from pulsarfeatureextractor_b200.engine import Engine

dev = torch.device("cuda", 0)
eng = Engine(0)
which = sys.argv[1:] or ["cli", "lyon", "pfd", "survey"]


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def tile(a, n):
    reps = (n + a.shape[0] - 1) // a.shape[0]
    return np.ascontiguousarray(np.concatenate([a] * reps, axis=0)[:n])


if "cli" in which:
    # configs[0]: Lyon-8 (and Bates-22) through the CLI classes on 1 000 synthetic HTRU .phcx.gz files:
    # directory walk, gunzip + scan, scoring, CSV rows; reading reported separately
    import shutil
    import tempfile
    from pulsarfeatureextractor_b200 import phcx_io
    from pulsarfeatureextractor_b200.DataProcessor import DataProcessor
    nfiles = 1000
    b = synth.generate(1, 0, nfiles)
    d = tempfile.mkdtemp(prefix="pfx_cli_")
    for i in range(nfiles):
        phcx_io.write_phcx(os.path.join(d, "cand_%05d.phcx.gz" % i), b.profile[i], b.subbands[i], b.dmcurve[i],
                           b.dmstat[i], b.params[i])
    paths = sorted(os.path.join(d, f) for f in os.listdir(d))
    t0 = time.perf_counter()
    sc = phcx_io.Scan(paths)
    t_native = time.perf_counter() - t0
    sc.close()
    t0 = time.perf_counter()
    for p_ in paths[:200]:
        phcx_io.read_phcx(p_)
    t_python = (time.perf_counter() - t0) * nfiles / 200
    res = {"config": "configs[0] CLI on %d synthetic .phcx.gz files" % nfiles, "files": nfiles,
           "read_native_scanner_s": t_native, "read_python_reader_s": t_python, "host_threads": os.cpu_count()}
    import contextlib
    import io
    for feats, method in (("lyon", "dmprofPHCX"), ("bates", "processPHCXCollectively")):
        out = os.path.join(d, "scores_%s.csv" % feats)
        dp = DataProcessor(engine=eng, features=feats, error_log=os.path.join(d, "CandidateErrorLog.txt"))
        with contextlib.redirect_stdout(io.StringIO()):
            for rep in range(2):          # first pass: kernels load lazily, scratch is allocated; second: steady state
                if os.path.exists(out):
                    os.remove(out)
                t0 = time.perf_counter()
                if feats == "lyon":
                    dp.dmprofPHCX(d, False, out, False, False)
                else:
                    dp.processPHCXCollectively(d, False, out, False, False, False)
                dt = time.perf_counter() - t0
                if rep == 0:
                    res["cli_%s_first_call_s" % feats] = dt
        rows = sum(1 for _ in open(out))
        res["cli_%s_s" % feats] = dt
        res["cli_%s_files_per_s" % feats] = nfiles / dt
        res["cli_%s_rows" % feats] = rows
    print(json.dumps(res))
    shutil.rmtree(d, ignore_errors=True)

if "lyon" in which:
    # configs[2]: Lyon-8, 10 M candidates, profile u8[128] + DM curve u8[256]
    n = 10_000_000
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    prof = torch.randint(0, 256, (n, 128), dtype=torch.uint8, device=dev, generator=g)
    ds = torch.randint(0, 256, (n, 256), dtype=torch.uint8, device=dev, generator=g)
    ms = timed(lambda: eng.lyon8(prof, ds), 5)
    # streamed: pinned host arrays in, rows out, chunks double-buffered inside pfx_score_host
    ns = 4_000_000
    hp = torch.empty((ns, 128), dtype=torch.uint8).pin_memory()
    hd = torch.empty((ns, 256), dtype=torch.uint8).pin_memory()
    hp.copy_(prof[:ns].cpu())
    hd.copy_(ds[:ns].cpu())
    from pulsarfeatureextractor_b200.engine import Scores
    out = Scores(torch.empty((ns, 8), dtype=torch.float64).pin_memory(), torch.zeros((ns,), dtype=torch.int32).pin_memory(), None)
    eng.score_host(hp, None, None, hd, None, features="lyon", out=out)
    t0 = time.perf_counter()
    eng.score_host(hp, None, None, hd, None, features="lyon", out=out)
    dt = time.perf_counter() - t0
    print(json.dumps({"config": "configs[2] Lyon-8 HTRU u8", "candidates": n, "kernel_ms": ms,
                      "device_resident_cand_per_s": n / ms * 1e3, "hbm_GBps": n * 448 / ms / 1e6,
                      "streamed_candidates": ns, "streamed_cand_per_s": ns / dt,
                      "streamed_h2d_GBps": ns * 384 / dt / 1e9}))
    del prof, ds, hp, hd

if "pfd" in which:
    # configs[3]: Bates-22 + Lyon-8, PFD shape (f64 profile[64], f64 sub-bands[64][64], f32 curve[256])
    n = 100_000
    b = synth.generate_pfd_shaped(4, 0, 2048)
    t = lambda a: torch.from_numpy(tile(a, n)).to(dev)
    prof, sub, dm, ds, par = t(b.profile), t(b.subbands), t(b.dmcurve), t(b.dmstat), t(b.params)

    def run():
        eng.pfd_params(prof, par)
        return eng.score_device(prof, sub, dm, ds, par, features="all", family=_lib.PFX_FAMILY_PFD)

    ms = timed(run, 2)
    eng.set_stage_timing(1)
    run()
    st = eng.last_timing()
    eng.set_stage_timing(0)
    print(json.dumps({"config": "configs[3] PFD shape f64 64/64x64/256, S/N and width on device", "candidates": n,
                      "ms": ms, "cand_per_s": n / ms * 1e3, "stage_ms": {k: round(v, 2) for k, v in st.items()}}))
    del prof, sub, dm, ds, par

if "survey" in which:
    # configs[4]: all 30 features, 1024-bin uint8 profiles, 16 sub-bands
    n = 50_000
    b = synth.generate(5, 0, 1024, nbin=1024)
    t = lambda a: torch.from_numpy(tile(a, n)).to(dev)
    prof, sub, dm, ds, par = t(b.profile), t(b.subbands), t(b.dmcurve), t(b.dmstat), t(b.params)
    run = lambda: eng.score_device(prof, sub, dm, ds, par, features="all")
    ms = timed(run, 2)
    eng.set_stage_timing(1)
    run()
    st = eng.last_timing()
    eng.set_stage_timing(0)
    print(json.dumps({"config": "configs[4] survey shape u8 1024/16x1024/256", "candidates": n, "ms": ms,
                      "cand_per_s": n / ms * 1e3, "stage_ms": {k: round(v, 2) for k, v in st.items()}}))
