"""Development helper: row-loop micro-benchmarks of the lane kernels (needs libpfx_prof.so)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PFX_LIB"] = os.path.join(ROOT, "pulsarfeatureextractor_b200", "libpfx_prof.so")
import numpy as np, torch
from pulsarfeatureextractor_b200 import synth
lib = ctypes.CDLL(os.environ["PFX_LIB"])
lib.pfx_micro_rows.restype = ctypes.c_float
lib.pfx_micro_rows.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
b = synth.generate(2, 0, 40000, want_subbands=False, want_dm=False)
prof = torch.from_numpy(b.profile).cuda()
reps = 50
for bps in (1, 2):
    for what, name in ((0, "jac pass (128 rows, 3 exp/row + QR + jac_end)"), (1, "lmpar + trial rows (1 exp/row)"), (2, "lmpar only")):
        ms = lib.pfx_micro_rows(ctypes.c_void_p(prof.data_ptr()), 0, 40000, 128, reps, what, bps)
        cyc = ms * 1e-3 * 1.965e9 / reps
        print("blocks/SM %d  %-48s %.3f ms total, %.0f cycles per pass per warp-slot (8 KB..)" % (bps, name, ms, cyc))
