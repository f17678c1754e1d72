"""Development helper: run the histogram + Gaussian group alone (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pulsarfeatureextractor_b200 import synth
from pulsarfeatureextractor_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
b = synth.generate(2, 0, n)
eng = Engine(0)
prof = torch.from_numpy(np.ascontiguousarray(b.profile)).to("cuda:0")
for _ in range(reps):
    eng.gaussian(prof)
torch.cuda.synchronize()
print("done")
