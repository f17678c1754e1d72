"""Development helper (GPU): time the DM-fit group on slices of the bench stream to find slow candidates.
usage: gpu_dm_outliers.py N CHUNK"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pulsarfeatureextractor_b200 import synth
from pulsarfeatureextractor_b200.engine import Engine
n = int(sys.argv[1]); chunk = int(sys.argv[2])
b = synth.generate(2, 0, n)
eng = Engine(0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")
prof, dm, par = t(b.profile), t(b.dmcurve), t(b.params)
def run(lo, hi):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    r = eng.dmfit(prof[lo:hi], dm[lo:hi], par[lo:hi])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), r
run(0, chunk)
res = []
for lo in range(0, n, chunk):
    ms, r = run(lo, min(n, lo + chunk))
    res.append((ms, lo))
res.sort(reverse=True)
print("slowest chunks:", [(round(a, 2), b_) for a, b_ in res[:6]], "median", round(float(np.median([a for a, _ in res])), 2))
ms, lo = res[0]
# bisect the slowest chunk
hi = min(n, lo + chunk)
while hi - lo > 1:
    mid = (lo + hi) // 2
    a, _ = run(lo, mid); c, _ = run(mid, hi)
    if a >= c: hi = mid
    else: lo = mid
ms, r = run(lo, lo + 1)
print("slowest candidate", lo, "ms", round(ms, 2), "ier", r.ier[0].cpu().numpy().tolist(), "scores", r.features[0].cpu().numpy().tolist())
print("params", b.params[lo].tolist()); print("dmcurve", b.dmcurve[lo].tolist())
