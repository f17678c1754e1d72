"""
Batch engine: Python host side of libpfx.  PyTorch is used only for device / pinned buffers
and streams; every score is computed by the CUDA kernels behind the C ABI (include/pfx.h).

    eng = Engine(device=0)
    out = eng.score_host(profile, subbands, dmcurve, dmstat, params, features="all")
    out.features   [N, 30] float64   Bates 1..22 then the 8 Lyon statistics
    out.status     [N] int32         which reference exception would have fired (0 = row is valid)
    out.ier        [N, 18] int8      MINPACK info per least-squares call site
"""
import ctypes

import numpy as np

from . import _lib

_NP_DTYPE = {np.dtype(np.uint8): _lib.PFX_U8, np.dtype(np.float32): _lib.PFX_F32,
             np.dtype(np.float64): _lib.PFX_F64}
_FEATURES = {"bates": _lib.PFX_FEAT_BATES, "lyon": _lib.PFX_FEAT_LYON, "all": _lib.PFX_FEAT_ALL}


class PfxError(RuntimeError):
    pass


class Scores(object):
    __slots__ = ("features", "status", "ier")

    def __init__(self, features, status, ier):
        self.features = features
        self.status = status
        self.ier = ier

    def ok(self):
        """Boolean mask of candidates whose row the reference would have written."""
        if self.status is None:
            return np.ones(len(self.features), dtype=bool)
        st = self.status.cpu().numpy() if hasattr(self.status, "cpu") else self.status
        return (st & _lib.ST_FAIL_MASK) == 0


def _code(dt):
    import torch
    if isinstance(dt, torch.dtype):
        return {torch.uint8: _lib.PFX_U8, torch.float32: _lib.PFX_F32, torch.float64: _lib.PFX_F64}[dt]
    return _NP_DTYPE[np.dtype(dt)]


def _ptr(a):
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return ctypes.c_void_p(a.data_ptr())
    return ctypes.c_void_p(a.ctypes.data)


class Engine(object):
    def __init__(self, device=0):
        import torch
        if not torch.cuda.is_available():
            raise PfxError("no CUDA device: pulsarfeatureextractor_b200 has no CPU path")
        self._lib = _lib.load()
        self.device = int(device)
        self._torch = torch
        h = ctypes.c_void_p()
        rc = self._lib.pfx_create(self.device, ctypes.byref(h))
        if rc != 0 or not h:
            raise PfxError("pfx_create(device=%d) failed with %d" % (self.device, rc))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pfx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise PfxError("libpfx error %d: %s" % (rc, self._lib.pfx_last_error(self._h).decode()))

    def device_info(self):
        sm, smem, scratch = ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
        self._check(self._lib.pfx_device_info(self._h, ctypes.byref(sm), ctypes.byref(smem), ctypes.byref(scratch)))
        return {"sm_count": sm.value, "smem_per_block": smem.value, "scratch_bytes": scratch.value}

    def reserve(self, shape, n_max):
        """Size the context's scratch for batches of up to n_max candidates of this shape
        (pfx_reserve): later score_device calls neither allocate nor synchronise."""
        self._check(self._lib.pfx_reserve(self._h, ctypes.byref(shape), int(n_max)))

    def set_subbatches(self, nsub):
        """Sub-batches of the histogram/Gaussian chain per call (0 = default, 1 = off)."""
        self._check(self._lib.pfx_set_subbatches(self._h, int(nsub)))

    def set_dm_exact(self, on):
        """DM-curve fit in MINPACK's own summation order, bit-exact against scipy (default on)."""
        self._check(self._lib.pfx_set_dm_exact(self._h, int(on)))

    def set_stage_timing(self, on):
        """Serialise and time the score groups of score_device (see pfx_set_stage_timing)."""
        self._check(self._lib.pfx_set_stage_timing(self._h, int(on)))

    def last_timing(self):
        t = _lib.Timing()
        self._check(self._lib.pfx_last_timing(self._h, ctypes.byref(t)))
        return {k: getattr(t, k) for k, _ in _lib.Timing._fields_}

    # -------------------------------------------------------------------------------------
    @staticmethod
    def make_shape(profile, subbands=None, dmcurve=None, dmstat=None, family=_lib.PFX_FAMILY_PHCX):
        s = _lib.Shape()
        s.nbin = profile.shape[1]
        s.dtype = _code(profile.dtype)
        s.nsub = subbands.shape[1] if subbands is not None else 0
        s.n_dm = dmcurve.shape[1] if dmcurve is not None else 0
        s.dm_dtype = _code(dmcurve.dtype) if dmcurve is not None else 0
        s.n_dmstat = dmstat.shape[1] if dmstat is not None else 0
        s.dmstat_dtype = _code(dmstat.dtype) if dmstat is not None else 0
        s.family = family
        return s

    def _stream(self):
        return ctypes.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def score_device(self, profile, subbands=None, dmcurve=None, dmstat=None, params=None,
                     features="all", family=_lib.PFX_FAMILY_PHCX, out=None, want_ier=True):
        """All inputs are CUDA tensors on ``self.device`` (contiguous).  Asynchronous on the
        current torch stream; returns device tensors."""
        torch = self._torch
        fbits = _FEATURES[features] if isinstance(features, str) else int(features)
        n = profile.shape[0]
        ncol = (22 if fbits & 1 else 0) + (8 if fbits & 2 else 0)
        dev = profile.device
        if out is None:
            feat = torch.empty((n, ncol), dtype=torch.float64, device=dev)
            status = torch.zeros((n,), dtype=torch.int32, device=dev)
            ier = torch.zeros((n, _lib.PFX_NIER), dtype=torch.int8, device=dev) if (want_ier and fbits & 1) else None
        else:
            feat, status, ier = out.features, out.status, out.ier
        for t in (profile, subbands, dmcurve, dmstat, params):
            if t is not None and not t.is_contiguous():
                raise PfxError("inputs must be contiguous")
        shape = self.make_shape(profile, subbands, dmcurve, dmstat, family)
        rc = self._lib.pfx_score_dev(self._h, ctypes.byref(shape), n, fbits, _ptr(profile), _ptr(subbands),
                                     _ptr(dmcurve), _ptr(dmstat), _ptr(params), _ptr(feat), feat.shape[1],
                                     _ptr(ier), _ptr(status), self._stream())
        self._check(rc)
        return Scores(feat, status, ier)

    def score_host(self, profile, subbands=None, dmcurve=None, dmstat=None, params=None,
                   features="all", family=_lib.PFX_FAMILY_PHCX, out=None, chunk=0, want_ier=True):
        """Inputs are host arrays (numpy, or pinned torch CPU tensors for asynchronous copies).
        Streams chunks H2D, scores them and copies the rows back; returns numpy arrays (or the
        tensors passed in ``out``)."""
        fbits = _FEATURES[features] if isinstance(features, str) else int(features)
        n = profile.shape[0]
        ncol = (22 if fbits & 1 else 0) + (8 if fbits & 2 else 0)
        for a in (profile, subbands, dmcurve, dmstat, params):
            if a is None:
                continue
            ok = a.is_contiguous() if hasattr(a, "is_contiguous") else a.flags.c_contiguous
            if not ok:
                raise PfxError("inputs must be C-contiguous")
        if out is None:
            # pinned result rows: a device-to-host copy into pageable memory is synchronous and
            # would serialise the double-buffered chunks (measured: 19 instead of 52 GB/s on the
            # Lyon-only stream); the numpy views keep the pinned tensors alive
            torch = self._torch
            feat = torch.empty((n, ncol), dtype=torch.float64, pin_memory=True).numpy()
            status = torch.zeros((n,), dtype=torch.int32, pin_memory=True).numpy()
            ier = (torch.zeros((n, _lib.PFX_NIER), dtype=torch.int8, pin_memory=True).numpy()
                   if (want_ier and fbits & 1) else None)
        else:
            feat, status, ier = out.features, out.status, out.ier
        shape = self.make_shape(profile, subbands, dmcurve, dmstat, family)
        rc = self._lib.pfx_score_host(self._h, ctypes.byref(shape), n, fbits, _ptr(profile), _ptr(subbands),
                                      _ptr(dmcurve), _ptr(dmstat), _ptr(params), _ptr(feat), feat.shape[1],
                                      _ptr(ier), _ptr(status), int(chunk))
        self._check(rc)
        return Scores(feat, status, ier)

    # ---- one entry per operator-API method (ProfileOperationsInterface.py:69-130) ---------
    def _group(self, n):
        torch = self._torch
        dev = torch.device("cuda", self.device)
        feat = torch.full((n, 22), float("nan"), dtype=torch.float64, device=dev)
        status = torch.zeros((n,), dtype=torch.int32, device=dev)
        ier = torch.zeros((n, _lib.PFX_NIER), dtype=torch.int8, device=dev)
        return feat, status, ier

    def sinusoid(self, profile):
        feat, status, ier = self._group(profile.shape[0])
        shape = self.make_shape(profile)
        self._check(self._lib.pfx_sinusoid_dev(self._h, ctypes.byref(shape), profile.shape[0], _ptr(profile),
                                               _ptr(feat), 22, _ptr(ier), _ptr(status), self._stream()))
        return Scores(feat[:, 0:4], status, ier)

    def gaussian(self, profile):
        feat, status, ier = self._group(profile.shape[0])
        shape = self.make_shape(profile)
        self._check(self._lib.pfx_gaussian_dev(self._h, ctypes.byref(shape), profile.shape[0], _ptr(profile),
                                               _ptr(feat), 22, _ptr(ier), _ptr(status), self._stream()))
        return Scores(feat[:, 4:11], status, ier)

    def dmfit(self, profile, dmcurve, params, family=_lib.PFX_FAMILY_PHCX):
        feat, status, ier = self._group(profile.shape[0])
        shape = self.make_shape(profile, None, dmcurve, None, family)
        self._check(self._lib.pfx_dmfit_dev(self._h, ctypes.byref(shape), profile.shape[0], _ptr(profile),
                                            _ptr(dmcurve), _ptr(params), _ptr(feat), 22, _ptr(ier), _ptr(status),
                                            self._stream()))
        return Scores(feat[:, 11:19], status, ier)

    def subband(self, profile, subbands, params):
        feat, status, ier = self._group(profile.shape[0])
        shape = self.make_shape(profile, subbands)
        self._check(self._lib.pfx_subband_dev(self._h, ctypes.byref(shape), profile.shape[0], _ptr(profile),
                                              _ptr(subbands), _ptr(params), _ptr(feat), 22, _ptr(status),
                                              self._stream()))
        return Scores(feat[:, 19:22], status, None)

    def pfd_params(self, profile, params):
        """PFD family: fill params[:, 0] (S/N) and params[:, 3] (width) from the profile on the device
        (PFDOperations.py:130-231).  In place; returns params."""
        shape = self.make_shape(profile)
        self._check(self._lib.pfx_pfd_params_dev(self._h, ctypes.byref(shape), profile.shape[0], _ptr(profile),
                                                 _ptr(params), self._stream()))
        return params

    def score_pfd_host(self, profile, subbands, dmcurve, params, features="all"):
        """PFD family from host arrays (what pfd_io gives): S/N and width are derived on the device
        (pfx_pfd_params_dev), then all scores; the chi-square curve is also the array of the Lyon
        DM statistics (PFDFile.py:553-583).  Returns numpy arrays."""
        torch = self._torch
        dev = torch.device("cuda", self.device)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        prof, sub, dmc, par = t(profile), t(subbands), t(dmcurve), t(params)
        if profile.shape[0] == 0:
            ncol = {"bates": 22, "lyon": 8, "all": 30}[features]
            return Scores(np.empty((0, ncol)), np.zeros((0,), np.int32), np.zeros((0, _lib.PFX_NIER), np.int8))
        self.pfd_params(prof, par)
        out = self.score_device(prof, sub, dmc, dmc, par, features=features, family=_lib.PFX_FAMILY_PFD)
        torch.cuda.synchronize(dev)
        return Scores(out.features.cpu().numpy(), out.status.cpu().numpy(),
                      out.ier.cpu().numpy() if out.ier is not None else None)

    def lyon8(self, profile, dmstat):
        torch = self._torch
        n = profile.shape[0]
        feat = torch.empty((n, 8), dtype=torch.float64, device=profile.device)
        shape = self.make_shape(profile, None, None, dmstat)
        self._check(self._lib.pfx_lyon8_dev(self._h, ctypes.byref(shape), n, _ptr(profile), _ptr(dmstat),
                                            _ptr(feat), 8, self._stream()))
        return Scores(feat, None, None)


def shard_ranges(n, ngpu, chunk=65536):
    """How a batch of n candidates is dealt to ngpu devices: candidates are independent, so
    any split gives the same rows.  Small batches: one contiguous range per device.  Streamed batches
    (at least two chunks per device): chunks of `chunk` candidates dealt round robin, which balances
    the skew in Levenberg-Marquardt iterations between regions of a survey.  Returns, per device, a
    list of (lo, hi) ranges in ascending order; together they tile [0, n) exactly once."""
    ngpu = max(1, int(ngpu))
    if n <= 0:
        return [[] for _ in range(ngpu)]
    if n < 2 * chunk * ngpu:
        return [[(n * r // ngpu, n * (r + 1) // ngpu)] if n * (r + 1) // ngpu > n * r // ngpu else []
                for r in range(ngpu)]
    out = [[] for _ in range(ngpu)]
    for k, lo in enumerate(range(0, n, chunk)):
        out[k % ngpu].append((lo, min(n, lo + chunk)))
    return out


class MultiEngine(object):
    """The same host-array interface as Engine over several GPUs of one box: one context, one host
    thread and one pair of staging buffers per device, no collective.  Each device scores its ranges
    of the batch and its device-to-host copies land at the rows' own offsets of ONE shared pinned
    result matrix, so the caller sees rows in input order whatever the split."""

    def __init__(self, devices):
        import concurrent.futures
        self.devices = [int(d) for d in devices]
        if not self.devices:
            raise PfxError("MultiEngine needs at least one device")
        self.engines = [Engine(d) for d in self.devices]
        self._torch = self.engines[0]._torch
        self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=len(self.engines))
        self.device = self.devices[0]

    def close(self):
        self._pool.shutdown(wait=True)
        for e in self.engines:
            e.close()

    def device_info(self):
        return self.engines[0].device_info()

    def _fan_out(self, n, work):
        """work(engine, lo, hi) for every range of every device, one thread per device."""
        plan = shard_ranges(n, len(self.engines))

        def run(r):
            for lo, hi in plan[r]:
                work(self.engines[r], lo, hi)
        futs = [self._pool.submit(run, r) for r in range(len(self.engines)) if plan[r]]
        for f in futs:
            f.result()          # re-raises the first PfxError of a device

    def _outputs(self, n, fbits, want_ier):
        torch = self._torch
        ncol = (22 if fbits & 1 else 0) + (8 if fbits & 2 else 0)
        feat = torch.empty((n, ncol), dtype=torch.float64, pin_memory=True).numpy()
        status = torch.zeros((n,), dtype=torch.int32, pin_memory=True).numpy()
        ier = (torch.zeros((n, _lib.PFX_NIER), dtype=torch.int8, pin_memory=True).numpy()
               if (want_ier and fbits & 1) else None)
        return Scores(feat, status, ier)

    def score_host(self, profile, subbands=None, dmcurve=None, dmstat=None, params=None,
                   features="all", family=_lib.PFX_FAMILY_PHCX, out=None, chunk=0, want_ier=True):
        fbits = _FEATURES[features] if isinstance(features, str) else int(features)
        n = profile.shape[0]
        if out is None:
            out = self._outputs(n, fbits, want_ier)
        cut = lambda a, lo, hi: None if a is None else a[lo:hi]

        def work(eng, lo, hi):
            part = Scores(out.features[lo:hi], out.status[lo:hi], cut(out.ier, lo, hi))
            eng.score_host(cut(profile, lo, hi), cut(subbands, lo, hi), cut(dmcurve, lo, hi), cut(dmstat, lo, hi),
                           cut(params, lo, hi), features=features, family=family, out=part, chunk=chunk)
        self._fan_out(n, work)
        return out

    def score_pfd_host(self, profile, subbands, dmcurve, params, features="all"):
        fbits = _FEATURES[features]
        n = profile.shape[0]
        out = self._outputs(n, fbits, True)
        if out.ier is None:
            out.ier = np.zeros((n, _lib.PFX_NIER), np.int8)

        def work(eng, lo, hi):
            r = eng.score_pfd_host(profile[lo:hi], subbands[lo:hi], dmcurve[lo:hi], params[lo:hi], features=features)
            out.features[lo:hi] = r.features
            out.status[lo:hi] = r.status
            if r.ier is not None:
                out.ier[lo:hi] = r.ier
        self._fan_out(n, work)
        return out


def make_engine(gpus=1, first=0):
    """Engine for one device, MultiEngine for several (ScoreGenerator --gpus N)."""
    gpus = int(gpus)
    return Engine(first) if gpus <= 1 else MultiEngine(range(first, first + gpus))
