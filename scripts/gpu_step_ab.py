"""Development helper (GPU): step time of pfx_score_dev in its production stream layout under several
environments (separate processes), rows compared with the first.
usage: gpu_step_ab.py N STEPS VAR=a[,VAR2=b] VAR=c ..."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    from pulsarfeatureextractor_b200 import synth
    from pulsarfeatureextractor_b200.engine import Engine
    n, steps, out = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    b = synth.generate(2, 0, n)
    eng = Engine(0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")
    args = [t(x) for x in (b.profile, b.subbands, b.dmcurve, b.dmstat, b.params)]
    for _ in range(3):
        r = eng.score_device(*args, features="all")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r = eng.score_device(*args, features="all")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    np.savez(out, feat=r.features.cpu().numpy(), status=r.status.cpu().numpy(), ier=r.ier.cpu().numpy())
    print("%.1f ms/step  %.0f cand/s" % (ms, n / ms * 1e3))
else:
    import numpy as np
    n, steps = sys.argv[1], sys.argv[2]
    outs = []
    for k, kv in enumerate(sys.argv[3:]):
        env = dict(x.split("=") for x in kv.split(",")) if kv != "-" else {}
        out = "/tmp/sab_%d.npz" % k
        print(kv, end="  ", flush=True)
        rc = subprocess.run([sys.executable, __file__, "--one", n, steps, out], env=dict(os.environ, **env)).returncode
        if rc != 0:
            print("FAILED rc=%d" % rc)
            continue
        outs.append((kv, np.load(out)))
    base = outs[0][1]
    for kv, b in outs[1:]:
        same = all(np.array_equal(base[f], b[f], equal_nan=(f == "feat")) for f in ("feat", "status", "ier"))
        print("--", kv, "rows identical to", outs[0][0], ":", same)
