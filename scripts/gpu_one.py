"""Development helper: run one scoring call (for ncu).  usage: gpu_one.py <n> <all|gauss|lyon>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pulsarfeatureextractor_b200 import synth
from pulsarfeatureextractor_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
what = sys.argv[2] if len(sys.argv) > 2 else "all"
eng = Engine(0); dev = torch.device("cuda", 0)
if what == "lyon":
    g = torch.Generator(device=dev); g.manual_seed(1)
    prof = torch.randint(0, 256, (n, 128), dtype=torch.uint8, device=dev, generator=g)
    ds = torch.randint(0, 256, (n, 256), dtype=torch.uint8, device=dev, generator=g)
    eng.lyon8(prof, ds); torch.cuda.synchronize()
    eng.lyon8(prof, ds); torch.cuda.synchronize()
    print("done")
    sys.exit(0)
b = synth.generate(2, 0, n)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
prof, sub, dm, ds, par = t(b.profile), t(b.subbands), t(b.dmcurve), t(b.dmstat), t(b.params)
if what == "gauss":
    eng.gaussian(prof)
elif what == "lyon":
    eng.lyon8(prof, ds)
else:
    eng.score_device(prof, sub, dm, ds, par, features="all")
torch.cuda.synchronize()
print("done", eng.last_timing())
