#!/usr/bin/env python
"""
bench.py -- throughput of the per-candidate scoring path (22 Bates + 8 Lyon) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config NAME] [--n CANDS] [--impl reference]

Configurations (BASELINE.json `configs`; data are seeded, synthetic and UNTILED: every candidate of a
batch is a different draw of pulsarfeatureextractor_b200/synth.py, generated in worker processes):

  htru100k   (default; configs[1], the configuration the metric is quoted on) HTRU-shaped PHCX
             candidates -- profile u8[128], sub-bands u8[16][128], DM curve u8[256], DataBlock[0]
             u8[256], 8 float64 parameters -- 100 000 per GPU, all 30 features.  Weak scaling.
  target10m  BASELINE north_star: the same shape, 10 000 000 candidates IN TOTAL (split over the
             ranks: strong scaling), all 30 features.
  lyon10m    configs[2]: the 8 Lyon statistics over 10 000 000 candidates in total, streamed from
             pinned host memory (e2e) and device resident (value).
  pfd        configs[3]: Bates 22 on PFD-shaped candidates (float64 64-bin profile, 64 x 64 float64
             sub-bands, float32 256-trial curve, PFD constants), 100 000 per GPU.
  survey     configs[4] shape: all 30 features on 1024-bin uint8 profiles (16 x 1024 sub-bands),
             50 000 per GPU (the 100 M of the config do not fit one run of a few minutes).
  --total T  with htru100k: T candidates in total split over the ranks (strong scaling).

A "step" is one pass of the scoring path over the batch.  Under torchrun every rank scores its own
shard of the seeded stream (no collective: candidates are independent) and rank 0 prints one JSON line.

value   candidates/s with the batch resident in HBM (device pointers through pfx_score_dev)
e2e     candidates/s through pfx_score_host: pinned host buffers in, feature rows out, H2D and D2H
        copies inside the timed region
roofline       the dominant kernel pair (k_gauss_*, scores 8-11) against measured HBM bandwidth, as the
               metric demands, plus its share of the step; the pair is FP64/latency bound, see DESIGN.md
roofline_fp64  the same kernels against the FP64 pipe: flops per candidate counted by ncu
               (profiles/r02_fp64_flops.json) over the live kernel time, against the DFMA peak measured
               on this pool (profiles/r01_fp64_microbench.txt)
roofline_lyon  the HBM-bound kernel of the path (Lyon statistics), measured live
cpu_baseline   the oracle (reference-faithful Python/SciPy port) on this box's host cores, a bounded
               SAMPLE of the same stream
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: stream id, nbin, nsub, n_dm, n_dmstat, per-GPU default, total (strong) default, features, family
    "htru100k": dict(stream=2, nbin=128, nsub=16, n_dm=256, n_dmstat=256, per_gpu=100000, total=None,
                     features="all", family="phcx", es=1, es_dm=1,
                     workload="BASELINE configs[1]: HTRU PHCX shape, profile u8[128], sub-bands u8[16x128], "
                              "DM curve u8[256], DataBlock0 u8[256]; all 30 features (22 Bates + 8 Lyon)"),
    "target10m": dict(stream=2, nbin=128, nsub=16, n_dm=256, n_dmstat=256, per_gpu=None, total=10_000_000,
                      features="all", family="phcx", es=1, es_dm=1,
                      workload="BASELINE north_star target: 10M HTRU-shaped candidates in total, all 30 features"),
    "lyon10m": dict(stream=3, nbin=128, nsub=16, n_dm=256, n_dmstat=256, per_gpu=None, total=10_000_000,
                    features="lyon", family="phcx", es=1, es_dm=1,
                    workload="BASELINE configs[2]: Lyon 8 statistics, 10M candidates in total, profile u8[128] + "
                             "DataBlock0 u8[256], streamed from pinned host memory"),
    "pfd": dict(stream=4, nbin=64, nsub=64, n_dm=256, n_dmstat=256, per_gpu=100000, total=None,
                features="bates", family="pfd", es=8, es_dm=4,
                workload="BASELINE configs[3]: Bates 22 on PFD-shaped candidates, profile f64[64], sub-bands "
                         "f64[64x64], chi-square curve f32[256], 4-parameter DM model"),
    "survey": dict(stream=5, nbin=1024, nsub=16, n_dm=256, n_dmstat=256, per_gpu=50000, total=None,
                   features="all", family="phcx", es=1, es_dm=1,
                   workload="BASELINE configs[4] shape: all 30 features, profile u8[1024], sub-bands u8[16x1024]"),
}
NCOL = {"all": 30, "bates": 22, "lyon": 8}
LYON_N = 4_000_000            # candidates of the secondary, HBM-bound roofline entry
FP64_PEAK_TFLOPS = 34.0       # DFMA peak measured on this pool (profiles/r01_fp64_microbench.txt)


def bytes_per_cand(c):
    """Algorithmic bytes per candidate: every input array once in its storage dtype, the
    parameters, and the outputs (float64 features, 18 int8 ier, int32 status)."""
    f = c["features"]
    b_in = c["es"] * c["nbin"]
    if f != "lyon":
        b_in += c["es"] * c["nsub"] * c["nbin"] + c["es_dm"] * c["n_dm"] + 8 * 8
    if f != "bates":
        b_in += (c["es_dm"] if c["family"] == "pfd" else 1) * c["n_dmstat"]
    b_out = NCOL[f] * 8 + (18 + 4 if f != "lyon" else 0)
    return b_in, b_out


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as fh:
            return json.load(fh).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


def read_json(rel):
    p = os.path.join(ROOT, rel)
    if os.path.isfile(p):
        with open(p) as fh:
            return json.load(fh)
    return None


class ClockSampler(threading.Thread):
    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip().split(",")
                self.samples.append((float(out[0]), float(out[1])))
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(s[1] for s in self.samples),
                "reasons": sorted(self.reasons)}


# ---- synthetic data (worker processes) ---------------------------------------------------------
def _gen_part(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    from pulsarfeatureextractor_b200 import synth
    name, lo, hi = args
    c = CONFIGS[name]
    if c["family"] == "pfd":
        b = synth.generate_pfd_shaped(c["stream"], lo, hi - lo, nbin=c["nbin"], nsub=c["nsub"], n_dm=c["n_dm"])
    else:
        lyon = c["features"] == "lyon"
        b = synth.generate(c["stream"], lo, hi - lo, c["nbin"], c["nsub"], c["n_dm"], c["n_dmstat"],
                           want_subbands=not lyon, want_dm=True)
    return b.profile, b.subbands, b.dmcurve, b.dmstat, b.params


def make_batch(name, start, n, workers):
    """Candidates [start, start+n) of the config's seeded stream in PINNED host tensors, generated in
    `workers` processes (4 096-candidate chunks of the stream: identical bytes however it is split)."""
    import numpy as np
    import torch
    import multiprocessing as mp
    c = CONFIGS[name]
    step = 4096 * 4
    first = (start // 4096) * 4096
    parts = [(name, max(lo, start), min(lo + step, start + n)) for lo in range(first, start + n, step)]
    parts = [p for p in parts if p[2] > p[1]]
    pdt = torch.float64 if c["family"] == "pfd" else torch.uint8
    ddt = torch.float32 if c["family"] == "pfd" else torch.uint8
    lyon = c["features"] == "lyon"
    pin = torch.cuda.is_available()
    prof = torch.empty((n, c["nbin"]), dtype=pdt, pin_memory=pin)
    sub = None if lyon else torch.empty((n, c["nsub"], c["nbin"]), dtype=pdt, pin_memory=pin)
    dmc = None if lyon else torch.empty((n, c["n_dm"]), dtype=ddt, pin_memory=pin)
    dms = torch.empty((n, c["n_dmstat"]), dtype=torch.uint8, pin_memory=pin)
    par = torch.empty((n, 8), dtype=torch.float64, pin_memory=pin)

    def place(p, res):
        lo, hi = p[1] - start, p[2] - start
        prof[lo:hi] = torch.from_numpy(res[0])
        if not lyon:
            sub[lo:hi] = torch.from_numpy(res[1])
            dmc[lo:hi] = torch.from_numpy(res[2])
        dms[lo:hi] = torch.from_numpy(res[3])
        par[lo:hi] = torch.from_numpy(res[4])

    if workers > 1 and len(parts) > 1:
        with mp.get_context("fork").Pool(min(workers, len(parts))) as pool:
            for p, res in zip(parts, pool.imap(_gen_part, parts)):
                place(p, res)
    else:
        for p in parts:
            place(p, _gen_part(p))
    if c["family"] == "pfd":
        dms = dmc          # PFD family: the chi-square curve is also the array of the Lyon DM statistics
    return prof, sub, dmc, dms, par


# ---- CPU arm -----------------------------------------------------------------------------------
_CPU_CACHE = {}


def _cpu_worker_warm(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pfx_oracle  # noqa: F401  (import cost outside the timed region)
    _CPU_CACHE[args] = _gen_part(args)
    return 0


def _cpu_worker(args):
    import warnings
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pfx_oracle as O
    warnings.simplefilter("ignore")
    name, lo, hi = args
    c = CONFIGS[name]
    prof, sub, dmc, dms, par = _CPU_CACHE.get(args) or _gen_part(args)
    O.FAST_CORR = False
    for i in range(hi - lo):
        if c["features"] != "lyon":
            if c["family"] == "pfd":
                O.bates22_pfd_arrays(prof[i], sub[i], dmc[i], par[i])
            else:
                O.bates22_phcx(prof[i].astype(np.int64), sub[i].astype(np.int64), dmc[i].astype(np.int64), par[i])
        if c["features"] != "bates":
            O.lyon8(prof[i], dms[i])
    return hi - lo


def cpu_baseline(name, n_sample, workers):
    """Oracle (Python/SciPy port of the reference) on host cores over the first n_sample candidates of
    the same seeded stream; data generation and imports outside the timed region."""
    import multiprocessing as mp
    with mp.Pool(workers) as pool:
        parts = [(name, i * n_sample // workers, (i + 1) * n_sample // workers) for i in range(workers)]
        parts = [p for p in parts if p[2] > p[1]]
        pool.map(_cpu_worker_warm, parts)
        t0 = time.time()
        done = sum(pool.map(_cpu_worker, parts))
        dt = time.time() - t0
    return done / dt, dt


def cpu_sample_size(name, workers, asked):
    per = {"htru100k": 512, "target10m": 512, "lyon10m": 40000, "pfd": 128, "survey": 128}[name]
    return max(workers * 8, asked or per)


def port_note():
    d = read_json("profiles/r02_ref_vs_port_speed.json")
    if not d:
        return None
    return {"measured_in": "build container (the reference cannot travel to the GPU box)", **d}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is
    Python 2 source that cannot run on this image's Python 3, so the oracle -- its
    line-faithful Python 3 / SciPy restatement, pinned against the reference through
    oracle/refshim.py -- is timed, on all host cores, on a bounded SAMPLE of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.config
    workers = os.cpu_count() or 1
    n_sample = cpu_sample_size(name, workers, args.cpu_sample)
    vals = []
    for _ in range(max(1, args.steps)):
        v, dt = cpu_baseline(name, n_sample, workers)
        vals.append(v)
    value = sum(vals) / len(vals)
    cfg = workload_config(name, n_sample, 1)
    cfg["sample_note"] = ("the CPU arm times a %d-candidate SAMPLE of the stream per step (the GPU arm scores the whole "
                          "batch); candidates/s is size independent" % n_sample)
    line = {
        "impl": "reference", "metric": "candidates/sec (22 Bates + 8 Lyon)", "value": value,
        "unit": "candidates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * n_sample / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "candidates/s", "cores": workers, "kind": "port",
                         "sample": "first %d candidates of the bench stream per step, features: %s"
                                   % (n_sample, CONFIGS[name]["features"]),
                         "port_vs_shimmed_reference": port_note()},
        "e2e": {"value": value, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(name, n, world):
    c = CONFIGS[name]
    b_in, b_out = bytes_per_cand(c)
    return {"workload": c["workload"], "name": name, "candidates_per_gpu": n, "candidates_total": n * world,
            "features": c["features"], "data_generation": "seeded, untiled (every candidate a separate draw)",
            "l2": "inputs+outputs %.0f MB per step per GPU > 126 MB L2" % (n * (b_in + b_out) / 1e6)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="htru100k", choices=sorted(CONFIGS))
    ap.add_argument("--n", type=int, default=0, help="candidates per GPU (overrides the config)")
    ap.add_argument("--total", type=int, default=0, help="candidates in total, split over the ranks (strong scaling)")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--subbatches", type=int, default=-1, help="pfx_set_subbatches (development)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from pulsarfeatureextractor_b200 import _lib
    from pulsarfeatureextractor_b200.engine import Engine, Scores

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created: keep stdout for the
        # one JSON line (file descriptor 1 points at stderr until the first collective is through)
        sys.stdout.flush()
        keep = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            os.dup2(keep, 1)
            os.close(keep)
    name = args.config
    c = CONFIGS[name]
    total = args.total or (c["total"] if not args.n else 0)
    if total:
        lo, hi = total * rank // world, total * (rank + 1) // world
        n, start, scaling = hi - lo, lo, "strong"
    else:
        n = args.n or c["per_gpu"]
        start, scaling = rank * n, "weak"
    features, ncol = c["features"], NCOL[c["features"]]
    family = _lib.PFX_FAMILY_PFD if c["family"] == "pfd" else _lib.PFX_FAMILY_PHCX
    b_in, b_out = bytes_per_cand(c)

    # this rank's shard of the seeded stream, in pinned host memory
    workers = max(1, (os.cpu_count() or 1) // world)
    t_gen = time.time()
    h_prof, h_sub, h_dm, h_ds, h_par = make_batch(name, start, n, workers)
    t_gen = time.time() - t_gen
    to_dev = lambda x: None if x is None else x.to(dev)
    d_prof, d_sub, d_dm, d_par = to_dev(h_prof), to_dev(h_sub), to_dev(h_dm), to_dev(h_par)
    d_ds = d_dm if c["family"] == "pfd" else to_dev(h_ds)
    eng = Engine(local)
    if args.subbatches >= 0:
        eng.set_subbatches(args.subbatches)
    bates = features != "lyon"
    if bates:
        eng.reserve(eng.make_shape(d_prof, d_sub, d_dm, d_ds, family), n)
        if c["family"] == "pfd":
            eng.pfd_params(d_prof, d_par)       # S/N and width from the profile (PFDOperations.py:130-231)
            torch.cuda.synchronize()
            h_par.copy_(d_par)
    out_dev = Scores(torch.empty((n, ncol), dtype=torch.float64, device=dev),
                     torch.zeros((n,), dtype=torch.int32, device=dev) if bates else None,
                     torch.zeros((n, 18), dtype=torch.int8, device=dev) if bates else None)
    out_host = Scores(torch.empty((n, ncol), dtype=torch.float64).pin_memory(),
                      torch.zeros((n,), dtype=torch.int32).pin_memory() if bates else None,
                      torch.zeros((n, 18), dtype=torch.int8).pin_memory() if bates else None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev():
        eng.score_device(d_prof, d_sub, d_dm, d_ds, d_par, features=features, family=family, out=out_dev)

    def step_host():
        eng.score_host(h_prof, h_sub, h_dm, h_ds, h_par, features=features, family=family, out=out_host)

    # warm-up: at least three passes; for the multi-second configurations on a 100 000-candidate slice
    long_step = n > 400000 and bates
    if long_step:
        wn = 100000
        w_out = Scores(out_dev.features[:wn], out_dev.status[:wn], out_dev.ier[:wn])
        for _ in range(max(3, args.warmup)):
            eng.score_device(d_prof[:wn], d_sub[:wn], d_dm[:wn], d_ds[:wn], d_par[:wn], features=features,
                             family=family, out=w_out)
    else:
        for _ in range(max(3, args.warmup)):
            step_dev()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()

    # ---- device-resident timing (CUDA events on the launching stream) ----------------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_dev()
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    stage, trace, launches = {}, {}, 0
    if bates:
        launches = int(eng.last_timing().get("launches", 0)) * args.steps
        if not long_step:
            # one more step with the score groups serialised and timed individually (CUDA events inside
            # libpfx): kernel durations for the roofline entries and the per-stage table
            eng.set_stage_timing(True)
            step_dev()
            stage = eng.last_timing()
            # and one with the production stream layout, recording when each group completed
            eng.set_stage_timing(2)
            step_dev()
            trace = eng.last_timing()
            eng.set_stage_timing(False)
    else:
        launches = args.steps

    # ---- end to end through the host-buffer call ---------------------------------------------
    if not long_step:
        for _ in range(min(2, args.warmup)):
            step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1000.0
    barrier()
    # ---- secondary roofline entry: the HBM-bound kernel of the path (Lyon statistics) ---------
    lyon_roof = None
    if rank == 0:
        gen = torch.Generator(device=dev)
        gen.manual_seed(7)
        lp = torch.randint(0, 256, (LYON_N, 128), dtype=torch.uint8, device=dev, generator=gen)
        ld_ = torch.randint(0, 256, (LYON_N, 256), dtype=torch.uint8, device=dev, generator=gen)
        for _ in range(3):
            eng.lyon8(lp, ld_)
        torch.cuda.synchronize()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for _ in range(5):
            eng.lyon8(lp, ld_)          # 1.8 GB per launch: larger than L2
        l1.record()
        torch.cuda.synchronize()
        lyon_ms = l0.elapsed_time(l1) / 5
        peak_l, kind_l = read_peaks()
        lyon_bytes = 128 + 256 + 64
        ach = LYON_N * lyon_bytes / (lyon_ms / 1000.0) / 1e9
        prof_l = read_json("profiles/r02_lyon_summary.json") or {}
        lyon_roof = {"bound": "hbm", "kernel": "k_lyon8_u8<1,2> (8 Lyon statistics, uint8 rows)",
                     "achieved": ach, "peak": peak_l, "peak_kind": kind_l, "unit": "GB/s", "frac": ach / peak_l,
                     "traffic": LYON_N * prof_l.get("dram_bytes_per_candidate", 446.0), "candidates": LYON_N,
                     "kernel_ms": lyon_ms, "candidates_per_s": LYON_N / (lyon_ms / 1000.0),
                     "note": "traffic: ncu dram read+write per candidate (%s) against %d B algorithmic"
                             % (prof_l.get("source", "profiles/r01_lyon_r1_summary.csv"), lyon_bytes)}
        del lp, ld_
    sampler.stop_flag = True
    sampler.join(timeout=2)

    times = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(times[0]), float(times[1])
    ok = int(((out_host.status.numpy() & 0x3df) == 0).sum()) if bates else n
    fd, fh = out_dev.features.cpu(), out_host.features
    diff = ~((fd == fh) | (torch.isnan(fd) & torch.isnan(fh)))
    same = not bool(diff.any())
    if not same and rank == 0:
        sys.stderr.write("host/device rows differ: %d, columns %s\n" % (int(diff.any(1).sum()),
                         (torch.nonzero(diff.any(0)).flatten() + 1).tolist()))

    if rank == 0:
        peak, peak_kind = read_peaks()
        total_n = n * world if not total else total
        value = total_n * args.steps / (ms_dev / 1000.0)
        e2e = total_n * args.steps / (ms_e2e / 1000.0)
        line = {
            "metric": "candidates/sec (22 Bates + 8 Lyon)", "value": value, "unit": "candidates/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(name, n, world),
            "e2e": {"value": e2e, "unit": "candidates/s", "h2d_bytes_per_step": n * b_in,
                    "d2h_bytes_per_step": n * b_out, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline_lyon": lyon_roof,
            "clocks": sampler.summary(),
            "rows_valid": ok, "host_equals_device": same, "data_generation_s": t_gen,
        }
        step_gbs = n * (b_in + b_out) / (ms_dev / args.steps / 1000.0) / 1e9
        if bates:
            gauss_ms = stage.get("gauss", 0.0)
            prof_g = read_json("profiles/r02_gauss_summary.json") or {}
            traffic = prof_g.get("dram_bytes_per_candidate", 337.7)
            # algorithmic bytes per candidate of the path x candidates per launch, over the
            # duration of the dominant kernel pair
            achieved = n * (b_in + b_out) / (gauss_ms / 1000.0) / 1e9 if gauss_ms > 0 else step_gbs
            line["roofline"] = {
                "bound": "hbm", "kernel": "k_gauss_rounds_s + k_gauss_final_s (scores 8-11, 10 LM solves per candidate)",
                "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": achieved / peak, "traffic": n * traffic,
                "kernel_ms": gauss_ms, "kernel_share_of_step": gauss_ms / (ms_dev / args.steps) if gauss_ms else None,
                "whole_step_gbs": step_gbs, "whole_step_frac": step_gbs / peak,
                "note": "the Bates path is FP64/latency bound (~1e3 flop per input byte, 10 Levenberg-Marquardt solves "
                        "per candidate in this kernel pair), so its HBM fraction is tiny by construction: see "
                        "roofline_fp64 for the pipe that bounds it and roofline_lyon for the HBM-bound kernel of the "
                        "path.  traffic: ncu dram read+write per candidate (%s)"
                        % prof_g.get("source", "profiles/r02_fp64_ops.csv")}
            flops = read_json("profiles/r02_fp64_flops.json")
            if flops and gauss_ms > 0 and name in ("htru100k", "target10m"):
                per_cand = flops["gauss_pair_fp64_flop_per_candidate"]
                ach = n * per_cand / (gauss_ms / 1000.0) / 1e12
                line["roofline_fp64"] = {
                    "bound": "fp64", "kernel": "k_gauss_rounds_s + k_gauss_final_s", "achieved": ach,
                    "peak": FP64_PEAK_TFLOPS, "peak_kind": "measured DFMA micro-benchmark (profiles/r01_fp64_microbench.txt)",
                    "unit": "TFLOP/s", "frac": ach / FP64_PEAK_TFLOPS, "flop_per_candidate": per_cand,
                    "kernel_ms": gauss_ms,
                    "note": "flops = 2 x DFMA + DADD + DMUL thread instructions counted by ncu (%s), "
                            "x candidates of this step / live kernel time" % flops.get("source", "profiles/")}
            line["stage_ms_serialised"] = {k: v for k, v in stage.items() if k != "launches"}
            line["stage_done_ms_overlapped"] = {k: v for k, v in trace.items() if k not in ("launches", "subband")}
            line["stage_note"] = ("stage_ms_serialised: groups one after the other on one stream; stage_done_ms_overlapped: "
                                  "side streams as in production but the final-fit kernel after the rounds kernel; the timed "
                                  "steps (value, ms_per_step) also overlap the tails of the two Gaussian kernels through the "
                                  "hand-over list (DESIGN 4.2)")
        else:
            line["roofline"] = {"bound": "hbm", "kernel": "k_lyon8_u8 (the whole step is this kernel)",
                                "achieved": step_gbs, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                                "frac": step_gbs / peak, "traffic": n * 446.0, "kernel_ms": ms_dev / args.steps,
                                "note": "e2e is bound by the host link (PCIe), value by HBM"}
        if not args.no_cpu and world == 1:
            workers = os.cpu_count() or 1
            ns = cpu_sample_size(name, workers, args.cpu_sample)
            v, dt = cpu_baseline(name, ns, workers)
            line["cpu_baseline"] = {"value": v, "unit": "candidates/s", "cores": workers, "kind": "port",
                                    "sample": "first %d candidates of the same stream, features %s, %.1f s"
                                              % (ns, features, dt),
                                    "port_vs_shimmed_reference": port_note()}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
