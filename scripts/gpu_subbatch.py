"""Development helper (GPU): step time of pfx_score_dev against the number of sub-batches of the
histogram/Gaussian chain.  usage: python scripts/gpu_subbatch.py [N] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pulsarfeatureextractor_b200 import synth
from pulsarfeatureextractor_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
b = synth.generate(2, 0, n)
eng = Engine(0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")
args = [t(x) for x in (b.profile, b.subbands, b.dmcurve, b.dmstat, b.params)]
ref = None
for s in (1, 2, 3, 4):
    eng.set_subbatches(s)
    for _ in range(2):
        out = eng.score_device(*args, features="all")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = eng.score_device(*args, features="all")
    e1.record()
    torch.cuda.synchronize()
    f = out.features.cpu().numpy()
    if ref is None:
        ref = f
    same = np.array_equal(ref, f, equal_nan=True)
    ms = e0.elapsed_time(e1) / steps
    print("n=%d subbatches=%d  %.1f ms/step  %.0f cand/s  rows identical to S=1: %s" % (n, s, ms, n / ms * 1e3, same))
