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
what = int(sys.argv[1])
print(lib.pfx_micro_rows(ctypes.c_void_p(prof.data_ptr()), 0, 40000, 128, 50, what, 2))
