"""Development helper: LM phase breakdown (needs libpfx_prof.so, PFX_LIB env)."""
import os, sys, ctypes, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PFX_LIB"] = os.path.join(ROOT, "pulsarfeatureextractor_b200", "libpfx_prof.so")
import numpy as np, torch
from pulsarfeatureextractor_b200 import synth, _lib
from pulsarfeatureextractor_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
b = synth.generate(2, 0, n)
eng = Engine(0); dev = torch.device("cuda", 0)
lib = eng._lib
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
prof, sub, dm, ds, par = t(b.profile), t(b.subbands), t(b.dmcurve), t(b.dmstat), t(b.params)
buf = (ctypes.c_ulonglong * 16)()
names = ["f0", "fdjac", "qrfac", "qtf", "gnorm+publish", "lmpar", "feval", "update", "fits", "outer", "inner"]
def run(name, fn):
    lib.pfx_prof_dump(buf, 1)
    torch.cuda.synchronize(); t0 = time.time(); fn(); torch.cuda.synchronize(); dt = time.time() - t0
    lib.pfx_prof_dump(buf, 1)
    v = np.array(list(buf), dtype=np.float64)
    tot = v[:8].sum()
    print("%s: %.3f s; fits %d outer %d inner %d; cycles/fit %.0f" % (name, dt, v[8], v[9], v[10], tot / max(v[8], 1)))
    print("   " + "  ".join("%s %.1f%%" % (names[i], 100 * v[i] / tot) for i in range(8)))
    print("   cycles per call: fdjac %.0f qrfac %.0f qtf %.0f lmpar %.0f feval %.0f update %.0f" % (v[1] / v[9], v[2] / v[9], v[3] / v[9], v[5] / v[10], v[6] / v[10], v[7] / v[10]))
run("sinusoid", lambda: eng.sinusoid(prof))
run("sinusoid", lambda: eng.sinusoid(prof))
run("hist+gauss", lambda: eng.gaussian(prof))
run("dmfit", lambda: eng.dmfit(prof, dm, par))
