"""Development helper: one serialised pass of every Bates kernel over n candidates (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pulsarfeatureextractor_b200 import synth
from pulsarfeatureextractor_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
b = synth.generate(2, 0, n)
eng = Engine(0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")
prof, sub, dm, ds, par = t(b.profile), t(b.subbands), t(b.dmcurve), t(b.dmstat), t(b.params)
eng.set_stage_timing(1)
for _ in range(2):
    eng.score_device(prof, sub, dm, ds, par, features="all")
torch.cuda.synchronize()
print(eng.last_timing())
