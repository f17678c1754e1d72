"""Development helper: GPU vs oracle error table on a synthetic batch (run on the GPU box)."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import pfx_oracle as O
from pulsarfeatureextractor_b200 import synth
from pulsarfeatureextractor_b200.engine import Engine
warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
b = synth.generate(1, 0, n)
eng = Engine(0)
dev = torch.device("cuda", 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
tp, ts, td, tds, tpar = t(b.profile), t(b.subbands), t(b.dmcurve), t(b.dmstat), t(b.params)
out = eng.score_device(tp, ts, td, tds, tpar, features="all"); torch.cuda.synchronize()
t0 = time.time(); out = eng.score_device(tp, ts, td, tds, tpar, features="all"); torch.cuda.synchronize(); dt = time.time() - t0
print("gpu time %.4f s for %d cands; timing %s" % (dt, n, eng.last_timing()))
feat = out.features.cpu().numpy(); status = out.status.cpu().numpy(); ier = out.ier.cpu().numpy()
host = eng.score_host(b.profile, b.subbands, b.dmcurve, b.dmstat, b.params, features="all", chunk=100)
print("host-path == device-path:", np.array_equal(host.features, feat, equal_nan=True), np.array_equal(host.status, status), np.array_equal(host.ier, ier))
rel = np.zeros((n, 30)); nbad = 0; iersame = np.zeros(18)
t0 = time.time()
for i in range(n):
    r = O.bates22_phcx(b.profile[i].astype(np.int64), b.subbands[i].astype(np.int64), b.dmcurve[i].astype(np.int64), b.params[i])
    ly = O.lyon8(b.profile[i], b.dmstat[i])
    a = np.concatenate([r.features, ly]); c = feat[i]
    if ((r.status & O.ST_FAIL_MASK) != 0) != ((status[i] & O.ST_FAIL_MASK) != 0) or (r.status & O.ST_DG_SENTINEL) != (status[i] & O.ST_DG_SENTINEL):
        nbad += 1; print(i, "STATUS oracle", hex(r.status), "gpu", hex(status[i]), "class", b.klass[i], b.edge[i])
    with np.errstate(all="ignore"):
        e = np.abs(a - c) / np.maximum(np.abs(a), 1e-12)
    e[(a == c) | (np.isnan(a) & np.isnan(c))] = 0
    e[np.isnan(e)] = np.inf
    rel[i] = e; iersame += (r.ier == ier[i])
print("oracle time %.1f s" % (time.time() - t0))
for k in range(30):
    e = rel[:, k]
    print("col %2d: exact %4d <1e-9 %4d <1e-6 %4d >1e-6 %4d  p50 %.1e p90 %.1e p99 %.1e max %.1e" % (k + 1, (e == 0).sum(), (e < 1e-9).sum(), (e < 1e-6).sum(), (e >= 1e-6).sum(), np.quantile(e, .5), np.quantile(e, .9), np.quantile(e, .99), e.max()))
print("status mismatches", nbad, "ier same frac", np.round(iersame / n, 3))
