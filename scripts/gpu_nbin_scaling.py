"""Development helper: Gaussian-stage time against the number of profile bins (same generator,
same candidates per batch): separates the per-row cost from the per-iteration cost."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pulsarfeatureextractor_b200 import synth
from pulsarfeatureextractor_b200.engine import Engine
eng = Engine(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
for nbin in (32, 64, 128, 256):
    b = synth.generate(2, 0, 4096, nbin=nbin)
    reps = (n + 4095) // 4096
    t = lambda a: torch.from_numpy(np.ascontiguousarray(np.concatenate([a] * reps)[:n])).to("cuda:0")
    prof, sub, dm, ds, par = t(b.profile), t(b.subbands), t(b.dmcurve), t(b.dmstat), t(b.params)
    eng.set_stage_timing(1)
    for _ in range(2):
        eng.score_device(prof, sub, dm, ds, par, features="all")
    torch.cuda.synchronize()
    st = eng.last_timing()
    print("nbin %4d  sinusoid %7.1f  hist %7.1f  gauss %7.1f  dmfit %6.1f  subband %6.1f ms" % (
        nbin, st["sinusoid"], st["histgauss"], st["gauss"], st["dmfit"], st["subband"]))
