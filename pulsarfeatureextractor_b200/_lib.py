"""
ctypes binding of libpfx (include/pfx.h).  There is no CPU path: if the shared library was not
built (``python -c 'import __graft_entry__ as g; g.build()'`` or ``make -C
pulsarfeatureextractor_b200/csrc``) importing this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PFX_LIB") or os.path.join(_HERE, "libpfx.so")

PFX_U8, PFX_F32, PFX_F64 = 0, 1, 2
PFX_FAMILY_PHCX, PFX_FAMILY_PFD = 0, 1
PFX_FEAT_BATES, PFX_FEAT_LYON, PFX_FEAT_ALL = 1, 2, 3
PFX_NPARAM, PFX_NBATES, PFX_NLYON, PFX_NALL, PFX_NIER, PFX_NFIT = 8, 22, 8, 30, 18, 16

ST_HIST_ZERO_BINS = 1 << 0
ST_LSQ_IMPROPER = 1 << 1
ST_INDEX_OOB = 1 << 2
ST_SUBBAND_NO_PAIRS = 1 << 3
ST_DM_NDEG_ZERO = 1 << 4
ST_DG_SENTINEL = 1 << 5
ST_SUBBAND_BAD_WIDTH = 1 << 6
ST_HIST_OVERFLOW = 1 << 7
ST_OTHER = 1 << 8
ST_EMPTY_SEQUENCE = 1 << 9
ST_FAIL_MASK = (ST_HIST_ZERO_BINS | ST_LSQ_IMPROPER | ST_INDEX_OOB | ST_SUBBAND_NO_PAIRS |
                ST_DM_NDEG_ZERO | ST_SUBBAND_BAD_WIDTH | ST_HIST_OVERFLOW | ST_OTHER |
                ST_EMPTY_SEQUENCE)

# every symbol include/pfx.h declares
EXPORTS = [
    "pfx_version", "pfx_create", "pfx_destroy", "pfx_last_error", "pfx_device_info",
    "pfx_lyon8_dev", "pfx_pfd_params_dev", "pfx_sinusoid_dev", "pfx_gaussian_dev", "pfx_dmfit_dev", "pfx_subband_dev",
    "pfx_score_dev", "pfx_score_host", "pfx_last_timing", "pfx_set_stage_timing",
    "pfx_reserve", "pfx_set_subbatches", "pfx_set_dm_exact",
]


class Shape(ctypes.Structure):
    _fields_ = [("nbin", ctypes.c_int32), ("nsub", ctypes.c_int32), ("n_dm", ctypes.c_int32),
                ("n_dmstat", ctypes.c_int32), ("dtype", ctypes.c_int32), ("dm_dtype", ctypes.c_int32),
                ("dmstat_dtype", ctypes.c_int32), ("family", ctypes.c_int32)]


class Timing(ctypes.Structure):
    _fields_ = [("lyon", ctypes.c_float), ("sinusoid", ctypes.c_float), ("histgauss", ctypes.c_float),
                ("gauss", ctypes.c_float), ("dmfit", ctypes.c_float), ("subband", ctypes.c_float),
                ("total", ctypes.c_float), ("launches", ctypes.c_int32)]


class LibraryMissing(RuntimeError):
    pass


def load():
    if not os.path.isfile(LIB_PATH):
        raise LibraryMissing(
            "libpfx.so not found at %s -- build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()').  "
            "pulsarfeatureextractor_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i64, dp = ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p
    sp = ctypes.POINTER(Shape)
    lib.pfx_version.restype = ctypes.c_int
    lib.pfx_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    lib.pfx_destroy.argtypes = [vp]
    lib.pfx_destroy.restype = None
    lib.pfx_last_error.argtypes = [vp]
    lib.pfx_last_error.restype = ctypes.c_char_p
    lib.pfx_device_info.argtypes = [vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                    ctypes.POINTER(ctypes.c_size_t)]
    lib.pfx_lyon8_dev.argtypes = [vp, sp, i64, vp, vp, dp, i64, vp]
    lib.pfx_sinusoid_dev.argtypes = [vp, sp, i64, vp, dp, i64, vp, vp, vp]
    lib.pfx_gaussian_dev.argtypes = [vp, sp, i64, vp, dp, i64, vp, vp, vp]
    lib.pfx_dmfit_dev.argtypes = [vp, sp, i64, vp, vp, vp, dp, i64, vp, vp, vp]
    lib.pfx_pfd_params_dev.argtypes = [vp, sp, i64, vp, vp, vp]
    lib.pfx_subband_dev.argtypes = [vp, sp, i64, vp, vp, vp, dp, i64, vp, vp]
    lib.pfx_score_dev.argtypes = [vp, sp, i64, ctypes.c_uint32, vp, vp, vp, vp, vp, dp, i64, vp, vp, vp]
    lib.pfx_score_host.argtypes = [vp, sp, i64, ctypes.c_uint32, vp, vp, vp, vp, vp, dp, i64, vp, vp, i64]
    lib.pfx_last_timing.argtypes = [vp, ctypes.POINTER(Timing)]
    lib.pfx_set_stage_timing.argtypes = [vp, ctypes.c_int]
    lib.pfx_reserve.argtypes = [vp, sp, i64]
    lib.pfx_set_subbatches.argtypes = [vp, ctypes.c_int]
    lib.pfx_set_dm_exact.argtypes = [vp, ctypes.c_int]
    for name in EXPORTS:
        getattr(lib, name)
    return lib
