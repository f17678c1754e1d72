"""GPU: per-score parity table of the least-squares scores (SURVEY 8c tiers) against the oracle and its
two re-associated twins, for the HTRU shape (bench stream), the PFD family shape and 1024-bin
profiles.  Writes profiles/r02_parity.json (copy of gpurun_out/r02_parity.json).
usage: python scripts/gpu_parity_table.py [n_htru] [n_pfd] [n_1024]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle_pool as OP
from pulsarfeatureextractor_b200 import synth, _lib
from pulsarfeatureextractor_b200.engine import Engine

n_htru = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
n_pfd = int(sys.argv[2]) if len(sys.argv) > 2 else 512
n_1k = int(sys.argv[3]) if len(sys.argv) > 3 else 256
eng = Engine(0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")


def gpu(b, family=_lib.PFX_FAMILY_PHCX):
    out = eng.score_device(t(b.profile), t(b.subbands), t(b.dmcurve), t(b.dmstat), t(b.params), features="all", family=family)
    torch.cuda.synchronize()
    return out.features.cpu().numpy(), out.status.cpu().numpy(), out.ier.cpu().numpy()


def section(name, b, family, fast_corr):
    t0 = time.time()
    f, s, i = gpu(b, _lib.PFX_FAMILY_PFD if family == "pfd" else _lib.PFX_FAMILY_PHCX)
    rows = OP.cached_run("parity_%s_%d" % (name, len(b)), b, family=family, twins=True, fast_corr=fast_corr)
    tab = OP.tier_table(rows, f[:, :22], i)
    fail_o, fail_g = (rows.status & 0x3df) != 0, (s & 0x3df) != 0
    ok = ~fail_o
    det = OP.rel(rows.feat[ok][:, OP.DET], f[ok][:, OP.DET])
    tab.update({"candidates": len(b), "failure_flags_identical": bool(np.array_equal(fail_o, fail_g)),
                "status_mismatches": int((fail_o != fail_g).sum()),
                "score3_exact": bool(np.array_equal(rows.feat[ok, 2], f[ok, 2])),
                "deterministic_max_rel": float(det.max()) if det.size else 0.0,
                "lyon_max_rel": float(OP.rel(rows.lyon, f[:, 22:30]).max()),
                "seconds": time.time() - t0, "host_cores": os.cpu_count()})
    print(name, "done in %.0f s" % (time.time() - t0), flush=True)
    return tab


res = {"what": "GPU (pfx_score_dev) against oracle/pfx_oracle.py and its twins; Tier W = twin 1 reproduces the oracle to 1e-8"}
full = synth.generate(2, 0, 100000)
idx = np.arange(0, 100000, 100000 // n_htru)[:n_htru]          # stratified sample of the bench stream
ff, fs, fi = gpu(full)
sub = synth.Batch(full.profile[idx], full.subbands[idx], full.dmcurve[idx], full.dmstat[idx], full.params[idx],
                  full.klass[idx], full.edge[idx])
t0 = time.time()
rows = OP.cached_run("parity_htru_%d" % len(sub), sub, family="phcx", twins=True)
tab = OP.tier_table(rows, ff[idx][:, :22], fi[idx])
fail_o, fail_g = (rows.status & 0x3df) != 0, (fs[idx] & 0x3df) != 0
ok = ~fail_o
det = OP.rel(rows.feat[ok][:, OP.DET], ff[idx][ok][:, OP.DET])
tab.update({"candidates": int(len(idx)), "of_batch": 100000, "failure_flags_identical": bool(np.array_equal(fail_o, fail_g)),
            "status_mismatches": int((fail_o != fail_g).sum()), "score3_exact": bool(np.array_equal(rows.feat[ok, 2], ff[idx][ok, 2])),
            "deterministic_max_rel": float(det.max()), "lyon_max_rel": float(OP.rel(rows.lyon, ff[idx][:, 22:30]).max()),
            "seconds": time.time() - t0, "host_cores": os.cpu_count()})
res["htru_bench_stream_stratified"] = tab
print("htru done in %.0f s" % (time.time() - t0), flush=True)
res["pfd_family_shape"] = section("pfd", synth.generate_pfd_shaped(4, 0, n_pfd), "pfd", True)
res["survey_1024_bins"] = section("1024", synth.generate(5, 0, n_1k, nbin=1024), "phcx", True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "r02_parity.json"), "w") as fh:
    json.dump(res, fh, indent=1)
for sec in ("htru_bench_stream_stratified", "pfd_family_shape", "survey_1024_bins"):
    print("==", sec, res[sec]["candidates"], "status mismatches", res[sec]["status_mismatches"], "det max rel", res[sec]["deterministic_max_rel"])
    for k, v in res[sec]["scores"].items():
        print("  %-8s tierW %5d  gpu %.4f twin2 %.4f  gpu p50/90/99 %s  twin1 %s" % (
            k, v["tier_w"], v["gpu_within_1e-6_in_tier_w"] or -1, v["twin2_within_1e-6_in_tier_w"] or -1,
            ["%.1e" % x for x in v["gpu_p50_p90_p99"]], ["%.1e" % x for x in v["twin1_p50_p90_p99"]]))
    print("  ier all slots:", res[sec]["ier_agreement_all_slots"])
