import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import pfx_oracle as O
from pulsarfeatureextractor_b200 import synth
from pulsarfeatureextractor_b200.engine import Engine
warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 100
b = synth.generate(1, 0, n)
eng = Engine(0); dev = torch.device("cuda", 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
args = (t(b.profile), t(b.subbands), t(b.dmcurve), t(b.dmstat), t(b.params))
out = eng.score_device(*args, features="all"); torch.cuda.synchronize()
feat = out.features.cpu().numpy(); ier = out.ier.cpu().numpy()
for i in range(min(n, 0)):
    prm = b.params[i]; snr, dm, bp, w, ds, de = prm[:6]
    log = O.FitLog()
    o = np.array(O.dm_fittings_phcx(b.dmcurve[i].astype(np.int64), 128 * np.arange(256) - 1, 128 * 256, snr, dm, bp * 1000, w, ds, de, log))
    c = feat[i, 15:19]
    if (np.isnan(o) != np.isnan(c)).any() or (np.isinf(o) != np.isinf(c)).any():
        print(i, "edge", b.edge[i], "oracle", o, "gpu", c, "ier", log.ier[15], ier[i, 15], "params", prm[:6])
h = eng.score_host(b.profile, b.subbands, b.dmcurve, b.dmstat, b.params, features="all", chunk=chunk)
d = ~((h.features == feat) | (np.isnan(h.features) & np.isnan(feat)))
print("host vs dev: rows differing", d.any(1).sum(), "cols", np.nonzero(d.any(0))[0] + 1)
for i in np.nonzero(d.any(1))[0][:6]:
    print("  row", i, "cols", np.nonzero(d[i])[0] + 1, h.features[i][d[i]][:4], feat[i][d[i]][:4], "ier", h.ier[i], ier[i])
out2 = eng.score_device(*args, features="all"); torch.cuda.synchronize()
f2 = out2.features.cpu().numpy()
d2 = ~((f2 == feat) | (np.isnan(f2) & np.isnan(feat)))
print("dev vs dev: rows differing", d2.any(1).sum(), "cols", np.nonzero(d2.any(0))[0] + 1)
