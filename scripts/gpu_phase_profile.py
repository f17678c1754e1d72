"""Development helper (GPU): per-phase cycle / active-lane accounting of k_gauss_rounds from a
-DPFX_PROFILE build (scripts/build_variants.sh "pfx_k_gauss prof -DPFX_PROFILE").
usage: PFX_LIB=.../libpfx_prof.so python scripts/gpu_phase_profile.py [N] > profiles/r02_phase_profile.txt"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pulsarfeatureextractor_b200 import synth, _lib
from pulsarfeatureextractor_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
b = synth.generate(2, 0, n)
eng = Engine(0)
lib = ctypes.CDLL(_lib.LIB_PATH)
prof = torch.from_numpy(b.profile).to("cuda:0")
eng.gaussian(prof)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 32)()
assert lib.pfx_debug_profile(buf, 1) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.gaussian(prof)
e1.record()
torch.cuda.synchronize()
assert lib.pfx_debug_profile(buf, 0) == 0
names = ["setup", "jac rows + QR", "jac_end (n x n)", "lmpar (n x n)", "trial rows", "trial_end", "finish",
         "queue / flush", "TOTAL (warp life)", "tail (queue empty .. exit)"]
names8 = ["final: start", "final: jac rows + QR", "final: jac_end (n x n)", "final: trial (lmpar + rows)",
          "final: finish", "final: TOTAL (warp life)"]
v = np.array(list(buf), dtype=np.float64).reshape(16, 2)
tot = v[8, 0]
print("k_gauss_rounds phase profile, %d candidates, hist+rounds+final %.1f ms (instrumented build)" % (n, e0.elapsed_time(e1)))
print("%-28s %10s %8s %12s" % ("phase", "warp Gcyc", "share", "lanes active"))
for k, nm in enumerate(names):
    c, lc = v[k]
    print("%-28s %10.3f %7.1f%% %12.2f" % (nm, c / 1e9, 100 * c / tot, lc / c if c else 0))
busy = v[:8, 1].sum()
print("lane-cycles in phases / (32 x warp life) = %.3f" % (busy / (32 * tot)))
tot8 = v[15, 0]
if tot8:
    print("k_gauss_final")
    for k, nm in enumerate(names8):
        c, lc = v[10 + k]
        print("%-28s %10.3f %7.1f%% %12.2f" % (nm, c / 1e9, 100 * c / tot8, lc / c if c else 0))
