"""Development helper (GPU): score the same batch under two environments (e.g. PFX_HIST_U8=0 / 1) in
separate processes, compare the rows bit for bit and print the stage times.
usage: gpu_ab.py N CONFIG VAR=a[,VAR2=c] VAR=b ...     (CONFIG: htru | superb64); rows are compared with the first set"""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    from pulsarfeatureextractor_b200 import synth
    from pulsarfeatureextractor_b200.engine import Engine
    n, cfg, out = int(sys.argv[2]), sys.argv[3], sys.argv[4]
    b = synth.generate(2, 0, n) if cfg == "htru" else synth.generate(3, 0, n, nbin=64)
    eng = Engine(0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")
    args = (t(b.profile), t(b.subbands), t(b.dmcurve), t(b.dmstat), t(b.params))
    for _ in range(2):
        r = eng.score_device(*args, features="all")
    torch.cuda.synchronize()
    eng.set_stage_timing(1)
    r = eng.score_device(*args, features="all")
    torch.cuda.synchronize()
    tm = eng.last_timing()
    np.savez(out, feat=r.features.cpu().numpy(), status=r.status.cpu().numpy(), ier=r.ier.cpu().numpy())
    print(json.dumps({k: round(v, 2) if isinstance(v, float) else v for k, v in tm.items() if k != "stages_done_ms"}))
else:
    import numpy as np
    n, cfg = sys.argv[1], sys.argv[2]
    outs = []
    for k, kv in enumerate(sys.argv[3:]):
        env = dict(x.split("=") for x in kv.split(","))
        out = "/tmp/ab_%d.npz" % k
        print(kv, end=" ", flush=True)
        rc = subprocess.run([sys.executable, __file__, "--one", n, cfg, out], env=dict(os.environ, **env)).returncode
        if rc != 0:
            print("FAILED rc=%d" % rc)
            continue
        outs.append((kv, np.load(out)))
    base = outs[0][1]
    for kv, b in outs[1:]:
        print("--", kv, "against", outs[0][0])
        for f in ("feat", "status", "ier"):
            same = np.array_equal(base[f], b[f], equal_nan=True) if f == "feat" else np.array_equal(base[f], b[f])
            print(f, "identical" if same else "DIFFERENT: %d rows" % int((base[f] != b[f]).reshape(len(base[f]), -1).any(axis=1).sum()))
            if not same and f == "feat":
                x, y = base[f], b[f]
                for c in range(x.shape[1]):
                    d = ~((x[:, c] == y[:, c]) | (np.isnan(x[:, c]) & np.isnan(y[:, c])))
                    if d.any():
                        rel = np.abs(x[d, c] - y[d, c]) / np.maximum(np.abs(x[d, c]), 1e-300)
                        print("   col %d: %d rows differ, rel diff median %.1e p90 %.1e max %.1e, > 1e-6: %d" % (
                            c, int(d.sum()), np.nanmedian(rel), np.nanpercentile(rel, 90), np.nanmax(rel), int((rel > 1e-6).sum())))
