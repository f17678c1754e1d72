"""Lyon-8 kernel at scale: device-resident GB/s and streamed (pinned host) candidates/s."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
from pulsarfeatureextractor_b200.engine import Engine, Scores
import pfx_oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nbin, nds = 128, 256
eng = Engine(0); dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
prof = torch.randint(0, 256, (n, nbin), dtype=torch.uint8, device=dev, generator=g)
dms = torch.randint(0, 256, (n, nds), dtype=torch.uint8, device=dev, generator=g)
out = eng.lyon8(prof, dms); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(5):
    e0.record(); out = eng.lyon8(prof, dms); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = min(ts)
byts = n * (nbin + nds + 64)
print("device-resident: %.3f ms, %.1f M cand/s, %.1f GB/s (algorithmic %d B/cand) = %.1f%% of 6543 GB/s" % (ms, n / ms / 1e3, byts / ms / 1e6, nbin + nds + 64, 100 * byts / ms / 1e6 / 6543.1))
f = out.features[:64].cpu().numpy(); p = prof[:64].cpu().numpy(); d = dms[:64].cpu().numpy()
ref = np.array([O.lyon8(p[i], d[i]) for i in range(64)])
print("max rel err vs oracle on 64 rows: %.2e" % np.max(np.abs(f - ref) / np.maximum(np.abs(ref), 1e-12)))
# streamed from pinned host memory
hp = torch.empty((n, nbin), dtype=torch.uint8).pin_memory(); hp.copy_(prof)
hd = torch.empty((n, nds), dtype=torch.uint8).pin_memory(); hd.copy_(dms)
res = Scores(torch.empty((n, 8), dtype=torch.float64).pin_memory(), None, None)
eng.score_host(hp, None, None, hd, None, features="lyon", out=res)
t0 = time.perf_counter(); eng.score_host(hp, None, None, hd, None, features="lyon", out=res); dt = time.perf_counter() - t0
print("streamed from pinned host: %.3f s, %.2f M cand/s, H2D %.1f GB/s + D2H %.1f GB/s" % (dt, n / dt / 1e6, n * 384 / dt / 1e9, n * 64 / dt / 1e9))
print("streamed == resident:", bool(torch.equal(torch.nan_to_num(res.features, nan=-7.0), torch.nan_to_num(out.features.cpu(), nan=-7.0))))
