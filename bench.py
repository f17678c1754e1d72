#!/usr/bin/env python
"""
bench.py -- throughput of the per-candidate scoring path (22 Bates + 8 Lyon) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n CANDS] [--impl reference]

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): HTRU-shaped PHCX
candidates -- profile u8[128], sub-bands u8[16][128], DM curve u8[256], DataBlock[0] u8[256],
8 float64 parameters -- 100 000 per GPU, all 30 features.  A "step" is one pass of the whole
scoring path over that batch.  Under torchrun every rank scores its own shard of the seeded
stream (weak scaling, no collective: candidates are independent) and rank 0 prints one JSON line.

value   candidates/s with the batch resident in HBM (device pointers through pfx_score_dev)
e2e     candidates/s through pfx_score_host: pinned host buffers in, feature rows out, H2D and D2H
        copies inside the timed region
roofline  the dominant kernel (k_gauss, scores 8-11) against measured HBM bandwidth, as the metric
        demands, plus its share of the step; the kernel is FP64/latency bound, see DESIGN.md
cpu_baseline  the oracle (reference-faithful Python/SciPy port) on this box's host cores
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

NBIN, NSUB, NDM, NDMSTAT = 128, 16, 256, 256
CONFIG_ID = 2
# algorithmic bytes per candidate: every input array once in its storage dtype, the
# parameters, and the outputs (30 float64 features, 18 int8 ier, int32 status)
BYTES_IN = NBIN + NSUB * NBIN + NDM + NDMSTAT + 8 * 8
BYTES_OUT = 30 * 8 + 18 + 4
BYTES_PER_CAND = BYTES_IN + BYTES_OUT
LYON_N = 4_000_000            # candidates of the secondary, HBM-bound roofline entry
LYON_BYTES_PER_CAND = NBIN + NDMSTAT + 64


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as fh:
            return json.load(fh).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


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


def cpu_baseline(n_sample, workers):
    """Oracle (Python/SciPy port of the reference) on host cores over the first n_sample
    candidates of the same seeded stream."""
    import multiprocessing as mp
    t0 = time.time()
    with mp.Pool(workers) as pool:
        parts = [(CONFIG_ID, i * n_sample // workers, (i + 1) * n_sample // workers) for i in range(workers)]
        pool.map(_cpu_worker_warm, parts)          # import + generate outside the timed region
        t0 = time.time()
        done = sum(pool.map(_cpu_worker, parts))
        dt = time.time() - t0
    return done / dt, dt


_CPU_CACHE = {}


def _cpu_worker_warm(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from pulsarfeatureextractor_b200 import synth
    cfg, lo, hi = args
    _CPU_CACHE[(cfg, lo, hi)] = synth.generate(cfg, lo, hi - lo, NBIN, NSUB, NDM, NDMSTAT)
    return 0


def _cpu_worker(args):
    import warnings
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pfx_oracle as O
    from pulsarfeatureextractor_b200 import synth
    warnings.simplefilter("ignore")
    cfg, lo, hi = args
    b = _CPU_CACHE.get((cfg, lo, hi))
    if b is None:
        b = synth.generate(cfg, lo, hi - lo, NBIN, NSUB, NDM, NDMSTAT)
    for i in range(hi - lo):
        O.bates22_phcx(b.profile[i].astype(np.int64), b.subbands[i].astype(np.int64),
                       b.dmcurve[i].astype(np.int64), b.params[i])
        O.lyon8(b.profile[i], b.dmstat[i])
    return hi - lo


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is
    Python 2 source that cannot run on this image's Python 3, so the oracle -- its
    line-faithful Python 3 / SciPy restatement, pinned against the reference through
    oracle/refshim.py -- is timed, on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = os.cpu_count() or 1
    n_sample = max(workers * 8, args.cpu_sample)
    vals = []
    for _ in range(max(1, args.steps)):
        v, dt = cpu_baseline(n_sample, workers)
        vals.append(v)
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "candidates/sec (22 Bates + 8 Lyon)", "value": value,
        "unit": "candidates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * n_sample / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n_sample),
        "cpu_baseline": {"value": value, "unit": "candidates/s", "cores": workers, "kind": "port",
                         "sample": "%d candidates of the bench stream per step, all 30 features" % n_sample},
        "e2e": {"value": value, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n):
    return {"workload": "BASELINE configs[1]: HTRU PHCX shape, profile u8[128], sub-bands u8[16x128], "
                        "DM curve u8[256], DataBlock0 u8[256]; all 30 features (22 Bates + 8 Lyon)",
            "candidates_per_gpu": n, "l2": "inputs+outputs %.0f MB per step > 126 MB L2" % (n * BYTES_PER_CAND / 1e6)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=100000, help="candidates per GPU")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-sample", type=int, default=512)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from pulsarfeatureextractor_b200 import synth
    from pulsarfeatureextractor_b200.engine import Engine, Scores

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n = args.n

    # this rank's shard of the seeded stream, in pinned host memory
    b = synth.generate(CONFIG_ID, rank * n, n, NBIN, NSUB, NDM, NDMSTAT)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_prof, h_sub, h_dm, h_ds, h_par = pin(b.profile), pin(b.subbands), pin(b.dmcurve), pin(b.dmstat), pin(b.params)
    d_prof, d_sub, d_dm, d_ds, d_par = (x.to(dev) for x in (h_prof, h_sub, h_dm, h_ds, h_par))
    eng = Engine(local)
    out_dev = Scores(torch.empty((n, 30), dtype=torch.float64, device=dev),
                     torch.zeros((n,), dtype=torch.int32, device=dev),
                     torch.zeros((n, 18), dtype=torch.int8, device=dev))
    out_host = Scores(torch.empty((n, 30), dtype=torch.float64).pin_memory(),
                      torch.zeros((n,), dtype=torch.int32).pin_memory(),
                      torch.zeros((n, 18), dtype=torch.int8).pin_memory())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev():
        eng.score_device(d_prof, d_sub, d_dm, d_ds, d_par, features="all", out=out_dev)

    def step_host():
        eng.score_host(h_prof, h_sub, h_dm, h_ds, h_par, features="all", out=out_host)

    for _ in range(args.warmup):
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
    # one more step with the score groups serialised and timed individually (CUDA events inside
    # libpfx): kernel durations for the roofline entry and the per-stage table
    eng.set_stage_timing(True)
    step_dev()
    stage = eng.last_timing()
    # and one with the production stream layout, recording when each group completed
    eng.set_stage_timing(2)
    step_dev()
    trace = eng.last_timing()
    eng.set_stage_timing(False)
    launches = int(stage.get("launches", 0)) * args.steps

    # ---- end to end through the host-buffer call ---------------------------------------------
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
        lp = torch.randint(0, 256, (LYON_N, NBIN), dtype=torch.uint8, device=dev, generator=gen)
        ld_ = torch.randint(0, 256, (LYON_N, NDMSTAT), dtype=torch.uint8, device=dev, generator=gen)
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
        ach = LYON_N * LYON_BYTES_PER_CAND / (lyon_ms / 1000.0) / 1e9
        lyon_roof = {"bound": "hbm", "kernel": "k_lyon8_u8<1,2> (8 Lyon statistics, uint8 rows)",
                     "achieved": ach, "peak": peak_l, "peak_kind": kind_l, "unit": "GB/s", "frac": ach / peak_l,
                     "traffic": LYON_N * 446.0, "candidates": LYON_N, "kernel_ms": lyon_ms,
                     "candidates_per_s": LYON_N / (lyon_ms / 1000.0),
                     "note": "traffic: ncu dram read+write of the 10M-candidate capture, 446 B/candidate "
                             "(profiles/r01_lyon_r1_summary.csv) against 448 B algorithmic"}
        del lp, ld_
    sampler.stop_flag = True
    sampler.join(timeout=2)

    times = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(times[0]), float(times[1])
    ok = int(((out_host.status.numpy() & 0x3df) == 0).sum())
    fd, fh = out_dev.features.cpu(), out_host.features
    diff = ~((fd == fh) | (torch.isnan(fd) & torch.isnan(fh)))
    same = not bool(diff.any())
    if not same and rank == 0:
        sys.stderr.write("host/device rows differ: %d, columns %s\n" % (int(diff.any(1).sum()),
                         (torch.nonzero(diff.any(0)).flatten() + 1).tolist()))

    if rank == 0:
        peak, peak_kind = read_peaks()
        total = n * world
        value = total * args.steps / (ms_dev / 1000.0)
        e2e = total * args.steps / (ms_e2e / 1000.0)
        gauss_ms = stage.get("gauss", 0.0)
        # algorithmic bytes per candidate of the path x candidates per launch, over the
        # duration of the dominant kernel pair
        achieved = n * BYTES_PER_CAND / (gauss_ms / 1000.0) / 1e9 if gauss_ms > 0 else 0.0
        step_gbs = n * BYTES_PER_CAND / (ms_dev / args.steps / 1000.0) / 1e9
        line = {
            "metric": "candidates/sec (22 Bates + 8 Lyon)", "value": value, "unit": "candidates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n),
            "e2e": {"value": e2e, "unit": "candidates/s", "h2d_bytes_per_step": n * BYTES_IN,
                    "d2h_bytes_per_step": n * BYTES_OUT, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "k_gauss_rounds + k_gauss_final (scores 8-11, 10 LM solves per candidate)",
                         "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": n * 337.7,
                         "kernel_ms": gauss_ms, "kernel_share_of_step": gauss_ms / (ms_dev / args.steps),
                         "whole_step_gbs": step_gbs, "whole_step_frac": step_gbs / peak,
                         "note": "the Bates path is FP64/latency bound (~1e3 flop per input byte, 10 Levenberg-"
                                 "Marquardt solves per candidate in this kernel pair), so its HBM fraction is "
                                 "tiny by construction; the HBM-bound kernel of the path is reported in roofline_lyon. "
                                 "traffic: ncu dram read+write of k_gauss_rounds at 100k candidates, 337.7 B/candidate "
                                 "(profiles/r01_gauss_final_build_summary.csv: profile row in, carry record and scores "
                                 "out; no re-reads); same capture: issue slots 25.7% busy, FP64 pipe 19.8%, 14.9 of 32 "
                                 "lanes active per instruction, top stall = fixed-latency dependency (wait)"},
            "roofline_lyon": lyon_roof,
            "stage_ms_serialised": {k: v for k, v in stage.items() if k != "launches"},
            "stage_done_ms_overlapped": {k: v for k, v in trace.items() if k not in ("launches", "subband")},
            "clocks": sampler.summary(),
            "rows_valid": ok, "host_equals_device": same,
        }
        if not args.no_cpu and world == 1:
            workers = os.cpu_count() or 1
            ns = max(workers * 8, args.cpu_sample)
            v, dt = cpu_baseline(ns, workers)
            line["cpu_baseline"] = {"value": v, "unit": "candidates/s", "cores": workers, "kind": "port",
                                    "sample": "first %d candidates of the same stream, all 30 features, %.1f s" % (ns, dt)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
