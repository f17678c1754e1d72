"""Development helper: time the histogram + Gaussian group with each libpfx variant given on the
command line (built by scripts/build_variants.sh); one subprocess per library."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    from pulsarfeatureextractor_b200 import synth
    from pulsarfeatureextractor_b200.engine import Engine
    n = int(sys.argv[2])
    b = synth.generate(2, 0, n)
    eng = Engine(0)
    prof = torch.from_numpy(np.ascontiguousarray(b.profile)).to("cuda:0")
    for _ in range(2):
        eng.gaussian(prof)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = eng.gaussian(prof)
    e1.record()
    torch.cuda.synchronize()
    f = r.features.cpu().numpy()
    print("%s n=%d hist+gauss %.1f ms  checksum %.6e" % (os.path.basename(os.environ.get("PFX_LIB", "libpfx.so")), n, e0.elapsed_time(e1), float(np.nansum(np.abs(f[:, 3:7])))))
else:
    n = sys.argv[1]
    for lib in sys.argv[2:]:
        env = dict(os.environ, PFX_LIB=os.path.join(ROOT, "pulsarfeatureextractor_b200", lib))
        subprocess.run([sys.executable, __file__, "--one", n], env=env)
