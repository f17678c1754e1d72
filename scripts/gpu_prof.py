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

def run(name, fn):
    lib.pfx_prof_dump(buf, 1)
    torch.cuda.synchronize(); t0 = time.time(); fn(); torch.cuda.synchronize(); dt = time.time() - t0
    lib.pfx_prof_dump(buf, 1)
    v = np.array(list(buf), dtype=np.float64)
    print("%s: %.3f s" % (name, dt))
    for k, nm in enumerate(["setup", "jac", "trial", "fin"]):
        print("   %-6s lane-executions %10d  avg cycles %10.0f  share of lane-cycles %5.1f%%" % (nm, v[8 + k], v[k] / max(v[8 + k], 1), 100 * v[k] / max(v[:4].sum(), 1)))
    print("   warp trips", v[12])
eng.gaussian(prof)
run("hist+gauss", lambda: eng.gaussian(prof))
