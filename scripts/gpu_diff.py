"""Development helper: is the device path deterministic / equal to the chunked host path?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pulsarfeatureextractor_b200 import synth
from pulsarfeatureextractor_b200.engine import Engine
n = 2000
b = synth.generate(1, 0, n)
eng = Engine(0); dev = torch.device("cuda", 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
args = (t(b.profile), t(b.subbands), t(b.dmcurve), t(b.dmstat), t(b.params))
a = eng.score_device(*args, features="all"); torch.cuda.synchronize(); fa = a.features.cpu().numpy()
c = eng.score_device(*args, features="all"); torch.cuda.synchronize(); fc = c.features.cpu().numpy()
h = eng.score_host(b.profile, b.subbands, b.dmcurve, b.dmstat, b.params, features="all", chunk=300).features
def diff(x, y, name):
    d = ~((x == y) | (np.isnan(x) & np.isnan(y)))
    print(name, "rows differing", d.any(1).sum(), "cols", np.nonzero(d.any(0))[0] + 1)
    r = np.nonzero(d.any(1))[0][:5]
    for i in r: print("  row", i, "class", b.klass[i], b.edge[i], "cols", np.nonzero(d[i])[0] + 1, x[i][d[i]][:4], y[i][d[i]][:4])
diff(fa, fc, "dev vs dev")
diff(fa, h, "dev vs host")
