"""
PRESTO ``.pfd`` candidate files: reader (packer) and a writer for synthetic candidates.

What the reference does with a ``.pfd`` before any score is computed (PFDFile.py:96-456,
PFDOperations.py:242-252,474-525), reproduced operation for operation because the scores are
judged on the arrays it produces:

* header and the fold cube ``profs[npart][nsub][proflen]`` (float64) with its ``foldstats``
  (PFDFile.py:107-201); ``avgprof = (profs / proflen).sum()`` (numpy pairwise summation over the
  flattened cube) and ``varprof = sum of prof_var`` are taken BEFORE any rotation;
* de-dispersion at the best DM by whole-bin left rotations of every sub-band
  (``floor(delaybins + 0.5) % proflen``, PFDFile.py:329-374), delays from
  ``DM / (0.000241 f f)`` relative to the highest sub-band, sub-band frequencies with the
  Python-2 integer division ``numchan / nsub`` (PFDFile.py:177);
* profile = ``sumprof - min``, divided by its mean, scaled to [0, 255] with the reference's own
  formula ``0 (1 - t) + 255 t`` (PFDFile.py:269-310);
* sub-bands = ``profs.sum(0)`` after de-dispersion (PFDFile.py:442-456);
* the reduced chi-square versus DM curve: 100 trial DMs spanning ``dms[0] .. dms[-1]``, rotations
  accumulated from trial to trial on the time-summed sub-bands, every value rounded to float32
  (PFDFile.py:378-423); its x axis enters the DM fit as ``dm_start = DMs[1]``, ``dm_end = DMs[-1]``,
  ``len = 100`` (PFDOperations.py:337-338).

``read_pfd`` is the single-file reader in numpy; the bulk path is the native scanner
(``pfx_pfd_scan`` in include/pfx_io.h, ``PfdScan`` below).  tests/test_pfd_io.py holds both against
the reference's own reader run under oracle/refshim.py.
"""
import ctypes
import os
import struct

import numpy as np

from . import phcx_io

N_TRIAL_DM = 100


class PFDCandidate(object):
    __slots__ = ("profile", "subbands", "dmcurve", "dmstat", "params", "dms")


def _pairwise_all(a):
    # numpy's own reduction: the reference calls ndarray.sum() / numpy.mean
    return np.add.reduce(np.ascontiguousarray(a).reshape(-1))


def read_pfd(path):
    with open(path, "rb") as fh:
        raw = fh.read()
    pos = 0

    def take(fmt):
        nonlocal pos
        size = struct.calcsize(fmt)
        if pos + size > len(raw):
            raise ValueError("truncated .pfd file")
        v = struct.unpack_from(fmt, raw, pos)
        pos += size
        return v

    sw = "<"
    test5 = struct.unpack_from("<5i", raw, 0)
    if np.abs(np.asarray(test5, dtype=np.float64)).max() > 100000:
        sw = ">"
    numdms, numperiods, numpdots, nsub, npart = take(sw + "5i")
    proflen, numchan, _pstep, _pdstep, _dmstep, _ndmfact, _npfact = take(sw + "7i")
    for _ in range(4):
        (ln,) = take(sw + "i")
        if ln < 0 or pos + ln > len(raw):
            raise ValueError("bad string length in .pfd header")
        pos += ln
    test = raw[pos:pos + 16]
    if len(test) < 16:
        raise ValueError("truncated .pfd file")
    if all(chr(c) in "0123456789:.-\0" for c in test):
        pos += 32
    _dt, _startT = take(sw + "2d")
    _endT, _tepoch, _bepoch, _avgvoverc, lofreq, chan_wid, bestdm = take(sw + "7d")
    take(sw + "2f")
    take(sw + "3d")
    take(sw + "2f")
    bary_p1, _b2, _b3 = take(sw + "3d")
    take(sw + "2f")
    fold_p1, _f2, _f3 = take(sw + "3d")
    take(sw + "7d")
    if min(numdms, numperiods, numpdots) < 0 or min(nsub, npart, proflen) <= 0:
        raise ValueError("bad dimensions in .pfd header")
    dms = np.asarray(take(sw + "%dd" % numdms), dtype=np.float64)
    take(sw + "%dd" % numperiods)
    take(sw + "%dd" % numpdots)
    ncube = npart * nsub * proflen
    if pos + 8 * (ncube + npart * nsub * 7) > len(raw):
        raise ValueError("truncated .pfd file")
    profs = np.frombuffer(raw, dtype=sw + "f8", count=ncube, offset=pos).astype(np.float64).reshape(npart, nsub, proflen)
    pos += 8 * ncube
    stats = np.frombuffer(raw, dtype=sw + "f8", count=npart * nsub * 7, offset=pos).astype(np.float64).reshape(npart, nsub, 7)
    if numdms < 2:
        raise ValueError("the DM curve needs dms[0] and dms[-1]")      # dms becomes a scalar, PFDFile.py:157

    binspersec = fold_p1 * proflen
    chanpersub = numchan // nsub                      # Python 2 int / int, PFDFile.py:177
    subdeltafreq = chan_wid * chanpersub
    losubfreq = lofreq + subdeltafreq - chan_wid
    subfreqs = np.arange(nsub, dtype="d") * subdeltafreq + losubfreq
    avgprof = (profs / proflen).sum()
    varprof = 0.0
    for part in range(npart):
        for sub in range(nsub):
            varprof += stats[part][sub][5]

    def delay_from_dm(dm):
        with np.errstate(all="ignore"):
            return np.where(subfreqs > 0.0, dm / (0.000241 * subfreqs * subfreqs), 0.0)

    # dedisperse(): whole-bin rotations at the best DM
    profs = profs.copy()
    subdelays = delay_from_dm(bestdm)
    subdelays = subdelays - subdelays[-1]
    delaybins = subdelays * binspersec - np.zeros(nsub)
    new_bins = np.floor(delaybins + 0.5)
    for ii in range(nsub):
        rot = int(new_bins[ii]) % proflen
        if rot:
            sub = profs[:, ii, :]
            profs[:, ii] = np.concatenate((sub[:, rot:], sub[:, :rot]), 1)
    subdelays_bins = np.zeros(nsub) + new_bins
    sumprof = profs.sum(0).sum(0)

    # getprofile() and scale()
    normprof = sumprof - min(sumprof)
    with np.errstate(all="ignore"):
        s = normprof / np.mean(normprof)
        mn, mx = min(s), max(s)
        profile = np.array([(0 * (1 - ((v - mn) / (mx - mn)))) + (255 * ((v - mn) / (mx - mn))) for v in s],
                           dtype=np.float64)

    # plot_subbands()
    subbands = profs.sum(0)

    # plot_chi2_vs_DM(dms[0], dms[-1], 100)
    work = profs.sum(0)
    lo, hi = dms[0], dms[-1]
    trial = lo + (hi - lo) * np.arange(N_TRIAL_DM) / (N_TRIAL_DM - 1)
    chis = np.zeros(N_TRIAL_DM, dtype="f")
    sdb = subdelays_bins.copy()
    for ii, dm in enumerate(trial):
        sd = delay_from_dm(dm)
        sd = sd - sd[-1]
        db = sd * binspersec - sdb
        nb = np.floor(db + 0.5)
        for jj in range(nsub):
            r = int(nb[jj]) % proflen
            if r:
                work[jj] = np.concatenate((work[jj][r:], work[jj][:r]))
        sdb += nb
        sp = work.sum(0)
        with np.errstate(all="ignore"):
            chis[ii] = ((sp - avgprof) ** 2.0 / varprof).sum() / (len(sp) - 1.0)

    c = PFDCandidate()
    c.profile = profile
    c.subbands = np.ascontiguousarray(subbands)
    c.dmcurve = chis
    c.dmstat = chis
    c.dms = trial
    c.params = np.array([0.0, bestdm, bary_p1, 0.0, float(trial[1]), float(trial[-1]), float(N_TRIAL_DM), 0.0],
                        dtype=np.float64)
    return c


def write_pfd(path, profs, stats, dms, lofreq, chan_wid, numchan, bestdm, fold_p1, bary_p1,
              dt=6.4e-5, with_position=True, big_endian=False):
    """A synthetic .pfd with the layout PFDFile.load() reads."""
    sw = ">" if big_endian else "<"
    npart, nsub, proflen = profs.shape
    periods = np.linspace(bary_p1 * 0.999, bary_p1 * 1.001, 5)
    pdots = np.zeros(3)

    def s(b):
        return struct.pack(sw + "i", len(b)) + b

    out = struct.pack(sw + "5i", len(dms), len(periods), len(pdots), nsub, npart)
    out += struct.pack(sw + "7i", proflen, numchan, 1, 2, 2, 1, 1)
    out += s(b"synthetic.fil") + s(b"PSR_cand") + s(b"Parkes") + s(b"/XWIN")
    if with_position:
        out += b"12:34:56.7890\0\0\0" + b"-12:34:56.789\0\0\0"
    out += struct.pack(sw + "2d", dt, 0.0)
    out += struct.pack(sw + "7d", 1.0, 55000.0, 55000.01, 1e-5, lofreq, chan_wid, bestdm)
    for p1 in (bary_p1 * 1.00001, bary_p1, fold_p1):
        out += struct.pack(sw + "2f", 100.0, 0.0) + struct.pack(sw + "3d", p1, 0.0, 0.0)
    out += struct.pack(sw + "7d", *([0.0] * 7))
    dt8 = sw + "f8"
    out += np.asarray(dms, dt8).tobytes() + np.asarray(periods, dt8).tobytes() + np.asarray(pdots, dt8).tobytes()
    out += np.ascontiguousarray(profs, dt8).tobytes() + np.ascontiguousarray(stats, dt8).tobytes()
    with open(path, "wb") as fh:
        fh.write(out)


def synth_pfd(seed, npart=8, nsub=16, proflen=64, kind=None):
    """Fold cube of a dispersed Gaussian pulse (or noise) with the header values to go with it."""
    rng = np.random.default_rng(seed)
    kind = rng.integers(0, 4) if kind is None else kind
    period = float(10 ** rng.uniform(-2.3, 0.3))
    bestdm = float(rng.uniform(20, 300))
    lofreq, chan_wid, numchan = 1200.0 + float(rng.uniform(0, 100)), 0.5, nsub * int(rng.integers(4, 33))
    fold_p1 = 1.0 / period
    x = np.arange(proflen)
    chanpersub = numchan // nsub
    subfreqs = np.arange(nsub) * chan_wid * chanpersub + lofreq + chan_wid * chanpersub - chan_wid
    delay = bestdm / (0.000241 * subfreqs * subfreqs)
    delay = (delay - delay[-1]) * fold_p1 * proflen
    width = float(rng.uniform(0.02, 0.15)) * proflen
    amp = 0.0 if kind == 0 else float(rng.uniform(0.5, 8.0))
    pos = float(rng.uniform(0, proflen))
    profs = np.empty((npart, nsub, proflen))
    for j in range(nsub):
        d = (x - pos - delay[j]) % proflen
        d = np.minimum(d, proflen - d)
        shape = np.exp(-0.5 * (d / width) ** 2)
        if kind == 3:       # double pulse
            d2 = (x - pos - delay[j] - 0.4 * proflen) % proflen
            d2 = np.minimum(d2, proflen - d2)
            shape = shape + 0.6 * np.exp(-0.5 * (d2 / width) ** 2)
        profs[:, j, :] = 100.0 + amp * shape[None, :] + rng.normal(0, 1.0, (npart, proflen))
    stats = np.zeros((npart, nsub, 7))
    stats[:, :, 0] = 1e5
    stats[:, :, 1] = 100.0
    stats[:, :, 2] = 1.0
    stats[:, :, 3] = proflen
    stats[:, :, 4] = 100.0
    stats[:, :, 5] = rng.uniform(0.8, 1.2, (npart, nsub))
    stats[:, :, 6] = 1.0
    dms = np.linspace(max(bestdm - 40.0, 0.0), bestdm + 40.0, 41)
    return dict(profs=profs, stats=stats, dms=dms, lofreq=lofreq, chan_wid=chan_wid, numchan=numchan,
                bestdm=bestdm, fold_p1=fold_p1, bary_p1=period)


# ---- native scanner --------------------------------------------------------------------------
class PfdInfo(ctypes.Structure):
    _fields_ = [("ok", ctypes.c_int32), ("nbin", ctypes.c_int32), ("nsub", ctypes.c_int32),
                ("n_dm", ctypes.c_int32), ("reserved0", ctypes.c_int32), ("reserved1", ctypes.c_int32),
                ("params", ctypes.c_double * 8)]


class PfdScan(object):
    """Native reader for a list of .pfd files (include/pfx_io.h): per-file ok / shape, and
    ``gather(indices)`` -> (profile f64, subbands f64, dmcurve f32, params f64) rows of files that
    share a shape, in pinned host arrays when CUDA is available."""

    def __init__(self, paths, threads=0):
        lib = phcx_io._load_io()
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        lib.pfx_pfd_scan.argtypes = [ctypes.POINTER(ctypes.c_char_p), i64, ctypes.c_int, ctypes.POINTER(PfdInfo),
                                     ctypes.POINTER(vp)]
        lib.pfx_pfd_gather.argtypes = [vp, ctypes.POINTER(i64), i64, vp, vp, vp, vp]
        lib.pfx_pfd_free.argtypes = [vp]
        self._lib = lib
        n = len(paths)
        arr = (ctypes.c_char_p * n)(*[os.fsencode(p) for p in paths])
        self.info = (PfdInfo * n)()
        self._h = vp()
        if lib.pfx_pfd_scan(arr, n, int(threads), self.info, ctypes.byref(self._h)) != 0:
            raise RuntimeError("pfx_pfd_scan failed")
        self.n = n

    def ok(self, i):
        return bool(self.info[i].ok)

    def shape_key(self, i):
        f = self.info[i]
        return (f.nbin, f.nsub, f.n_dm)

    def gather(self, indices):
        k = len(indices)
        nbin, nsub, n_dm = self.shape_key(indices[0])
        buf = phcx_io.Scan._buffer
        prof = buf((k, nbin), np.float64)
        sub = buf((k, nsub, nbin), np.float64)
        dmc = buf((k, n_dm), np.float32)
        par = buf((k, 8), np.float64)
        idx = (ctypes.c_int64 * k)(*indices)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        rc = self._lib.pfx_pfd_gather(self._h, idx, k, p(prof), p(sub), p(dmc), p(par))
        if rc != 0:
            raise RuntimeError("pfx_pfd_gather failed (%d)" % rc)
        return prof, sub, dmc, par

    def close(self):
        if self._h:
            self._lib.pfx_pfd_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:       # noqa: BLE001 - interpreter shutdown
            pass
