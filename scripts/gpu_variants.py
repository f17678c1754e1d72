"""Development helper: time one score group (gaussian | sinusoid | dmfit | all) with each libpfx
variant given on the command line (built by scripts/build_variants.sh); one subprocess per library.
usage: gpu_variants.py N OP lib1.so lib2.so ..."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    from pulsarfeatureextractor_b200 import synth
    from pulsarfeatureextractor_b200.engine import Engine
    n = int(sys.argv[2]); op = sys.argv[3]
    b = synth.generate(2, 0, n)
    eng = Engine(0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")
    prof, sub, dm, ds, par = t(b.profile), t(b.subbands), t(b.dmcurve), t(b.dmstat), t(b.params)
    run = {"gaussian": lambda: eng.gaussian(prof), "sinusoid": lambda: eng.sinusoid(prof),
           "dmfit": lambda: eng.dmfit(prof, dm, par),
           "all": lambda: eng.score_device(prof, sub, dm, ds, par, features="all")}[op]
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = run()
    e1.record()
    torch.cuda.synchronize()
    f = r.features.cpu().numpy()
    f = f[np.isfinite(f).all(axis=1)]
    print("%s n=%d %s %.1f ms  checksum %.6e" % (os.path.basename(os.environ.get("PFX_LIB", "libpfx.so")), n, op, e0.elapsed_time(e1), float(np.abs(f).clip(0, 1e6).sum())))
else:
    n, op = sys.argv[1], sys.argv[2]
    for lib in sys.argv[3:]:
        env = dict(os.environ, PFX_LIB=os.path.join(ROOT, "pulsarfeatureextractor_b200", lib))
        subprocess.run([sys.executable, __file__, "--one", n, op], env=env)
