"""
pfx_oracle -- CPU restatement of PulsarFeatureExtractor's per-candidate scoring path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this file, and
only as the checker / CPU baseline.  The product (``pulsarfeatureextractor_b200``) never
imports it and has no CPU fallback.

What it restates (citations are file:line under
``/root/reference/PulsarFeatureExtractor/src/``; PO = ProfileOperations.py,
POI = ProfileOperationsInterface.py, PHXO = PHCXOperations.py, PFDO = PFDOperations.py,
PHX = PHCXFile.py, CFI = CandidateFileInterface.py, DP = DataProcessor.py):

  Lyon-8   PHX:320-379 (+ PHXO:528-539)             -> lyon4 / lyon8
  1-4      PO:190-587                               -> sinusoid_fittings
  5-11     PO:595-1483, POI:138-207                 -> gaussian_fittings
  12-15    PHXO:81-112, PFDO:93-233, CFI:84-112     -> cand_params_phcx / cand_params_pfd
  16-19    PHXO:121-259, PFDO:274-393               -> dm_fittings_phcx / dm_fittings_pfd
  20-22    PO:1585-1686, PHXO:387-415, PFDO:445-466 -> subband_scores / profile_corr
  writer   DP:305-325,405-447                       -> format_row

Third-party arithmetic: exactly the calls the reference makes -- ``scipy.optimize.leastsq``
(MINPACK lmdif, default tolerances), ``scipy.stats.skew/kurtosis/scoreatpercentile``,
``numpy.histogram/corrcoef/mean/std/argmax/delete`` -- from the numpy 2.3 / scipy 1.18 in
this image (the reference pins no versions; README.md:194-199).

Pinning: the reference ships no tests or golden vectors; its only
known answers are docstring examples (PO:213-221 score 4 = 32.5, PO:271-289 two peaks,
CFI:97-112 filterScore).  Those are checked in tests/test_oracle_kat.py.  Beyond that the
oracle is pinned against THE REFERENCE'S OWN CODE executed in this container through
``oracle/refshim.py`` (mechanical Python 2 -> 3 semantics shim, sources read in place):
tests/test_oracle_vs_reference.py compares every score group on seeded synthetic and
edge-case candidates, and tests/golden/*.npz holds reference outputs generated the same way
(tests/golden/make_golden.py) so the check travels to machines without /root/reference.

Python 2 semantics pinned here: int/int division sites
POI:206 (always-rotate), PO:1106/1166 (cut = n//2), PO:1041 (bins//2 - 1), PO:1628 (w//2),
PFDO:202 (n//2), POI:158 (range // 60 when the IQR is 0 and the data are integers).  ``PY2_DIV`` switches them (default True = reference-faithful).
"""
import math
import warnings

import numpy as np
from scipy import stats as _stats
from scipy.optimize import leastsq as _leastsq

PY2_DIV = True
# one corrcoef over the whole matrix instead of nsub(nsub-1)/2 pair calls (ulp-level
# differences only); the pair loop of PO:1666-1678 is the default.
FAST_CORR = False
# "Perturbed twin" (SURVEY.md section 8c tolerances): the same oracle with the residual expressions
# re-associated (algebraically identical, a few ulp different).  oracle-vs-twin scatter is the
# noise floor any independent implementation is judged against on ill-conditioned fits.
TWIN = False      # False, 1 (first twin) or 2 (second twin)

# --- per-candidate status bits (mirrored in include/pfx.h) --------------------------------
ST_HIST_ZERO_BINS = 1 << 0    # numpy.histogram(bins=0) -> ValueError        (PO:657,678)
ST_LSQ_IMPROPER = 1 << 1      # leastsq N > M -> TypeError                   (PO:1045)
ST_INDEX_OOB = 1 << 2         # IndexError outside the guarded block          (PO:972,974)
ST_SUBBAND_NO_PAIRS = 1 << 3  # ZeroDivisionError, m == 0 or width_bins == 0  (PO:1681,1684)
ST_DM_NDEG_ZERO = 1 << 4      # ZeroDivisionError in the PFD chi-square        (PFDO:385)
ST_DG_SENTINEL = 1 << 5       # soft: except IndexError -> 1e6 sentinels       (PO:762-764)
ST_SUBBAND_BAD_WIDTH = 1 << 6 # width_bins > nbin -> max_bin unbound           (PO:1630)
ST_HIST_OVERFLOW = 1 << 7     # implementation limit of the CUDA engine (never set here)
ST_OTHER = 1 << 8             # any other exception
ST_EMPTY_SEQUENCE = 1 << 9    # argmax/max of an empty sequence -> ValueError  (PO:1361)
ST_FAIL_MASK = (ST_HIST_ZERO_BINS | ST_LSQ_IMPROPER | ST_INDEX_OOB | ST_SUBBAND_NO_PAIRS |
                ST_DM_NDEG_ZERO | ST_SUBBAND_BAD_WIDTH | ST_HIST_OVERFLOW | ST_OTHER |
                ST_EMPTY_SEQUENCE)

# --- fit slots: one ``ier`` per leastsq call site, in call order ---------------------------
FIT_SINE = 0            # PO:449
FIT_SINESQ = 1          # PO:549
FIT_DYHIST = 2          # PO:949 on the derivative histogram (last retry)
FIT_PROFHIST = 3        # PO:949 on the profile histogram (last retry)
FIT_FIXED = 4           # PO:1045
FIT_T1 = 5              # PO:1246
FIT_DG_ROUND0 = 6       # PO:1379 rounds 1..8 -> slots 6..13
FIT_DG_FINAL = 14       # PO:1473
FIT_DM = 15             # PHXO:212 / PFDO:369
NFIT = 16
FIT_DYHIST_COUNT = 16   # number of fits done by the retry loops (1..6), extra bookkeeping
FIT_PROFHIST_COUNT = 17
NIER = 18

FWHM_K = 2.0 * math.sqrt(2.0 * math.log(2.0))


class CandidateFailure(Exception):
    def __init__(self, status):
        Exception.__init__(self, "candidate failed, status=0x%x" % status)
        self.status = status


class FitLog(object):
    """Collects ier / nfev of every least-squares solve of one candidate."""

    def __init__(self):
        self.ier = np.zeros(NIER, dtype=np.int8)
        self.nfev = np.zeros(NFIT, dtype=np.int32)
        self.cost = np.full(NFIT, np.nan)
        self.params = {}

    def record(self, slot, x, info, ier):
        self.ier[slot] = ier
        self.nfev[slot] = info["nfev"]
        self.cost[slot] = float(np.sum(np.asarray(info["fvec"], dtype=np.float64) ** 2))
        self.params[slot] = np.array(x, dtype=np.float64)


def _div(a, b):
    """int/int division at the sites SURVEY 8c D1 lists."""
    return a // b if PY2_DIV else a / b


def _is_intlike(v):
    return isinstance(v, (int, np.integer)) or (isinstance(v, np.ndarray) and v.dtype.kind in "iub")


def _classic_div(a, b):
    """Python 2 ``/``: floor division when both operands are integers, true division otherwise."""
    if PY2_DIV and _is_intlike(a) and _is_intlike(b):
        return a // b
    return a / b


def _seqsum(a):
    """Left-to-right floating point sum (what ``for ...: s += a[i]`` does)."""
    a = np.asarray(a, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.cumsum(a)[-1])


def _chisq(y, fit, n, div=None, mask=None):
    """The reference's scalar accumulation loops, e.g. PO:453-458: ``c += (y[i]-fit[i])**2``.
    Kept scalar because numpy's scalar ``**2`` (libm pow) and array ``**2`` (multiply) can
    differ by an ulp.  ``div`` divides every term (PO:1481), ``mask`` selects terms."""
    c = 0
    for i in range(n):
        if mask is not None and not mask[i]:
            continue
        t = (y[i] - fit[i]) ** 2
        if div is not None:
            t = t / div
        c += t
    return c


def _lsq(resid, p0, args, log, slot):
    """scipy.optimize.leastsq with the reference's defaults; full_output only adds reporting."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            x, _cov, info, _msg, ier = _leastsq(resid, p0, args=args, full_output=True)
    x = np.atleast_1d(x)
    if log is not None:
        log.record(slot, x, info, ier)
    return x


def _builtin_max(seq):
    """Python's ``max`` on a sequence: sequential ``>`` (NaN survives only when first)."""
    it = iter(seq)
    best = next(it)
    for v in it:
        if v > best:
            best = v
    return best


def _builtin_min(seq):
    it = iter(seq)
    best = next(it)
    for v in it:
        if v < best:
            best = v
    return best


# ==========================================================================================
# Lyon statistics                                                         PHX:320-379
# ==========================================================================================
def lyon4(values):
    """[mean, std(ddof 0), skew(bias), excess kurtosis(bias)] -- PHX:337-341 / PHX:367-371."""
    bins = [float(v) for v in values]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            return [float(np.mean(bins)), float(np.std(bins)),
                    float(_stats.skew(bins)), float(_stats.kurtosis(bins))]


def lyon4_asis(values):
    """The same four statistics on the array AS IT IS: for .pfd candidates the reference passes the
    float32 chi-square curve straight to numpy / scipy (PFDFile.py:567-574), so mean, std, skew and
    kurtosis are evaluated in float32 arithmetic there."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            return [float(np.mean(values)), float(np.std(values)),
                    float(_stats.skew(values)), float(_stats.kurtosis(values))]


def lyon8_pfd(profile, chis):
    """PFD family: profile statistics on Python floats (PFDFile.py:531-541), DM-curve statistics on
    the float32 array as is."""
    return np.array(lyon4(profile) + lyon4_asis(np.asarray(chis)), dtype=np.float64)


def lyon8(profile, dm_data):
    """Profile statistics followed by DM-curve statistics (column order DP:884-886,918-926)."""
    return np.array(lyon4(profile) + lyon4(dm_data), dtype=np.float64)


# ==========================================================================================
# Scores 1-4                                                               PO:190-587
# ==========================================================================================
def count_peaks(normalised):
    """Block scan of PO:294-326.  Returns the list of peak indexes."""
    peaks = []
    blk_idx, blk_val = [], []
    zeros = 0
    for i, v in enumerate(normalised):
        if v != 0:
            blk_val.append(v)
            blk_idx.append(i)
        elif zeros < 4:
            blk_val.append(v)
            blk_idx.append(i)
            zeros += 1
        else:
            # PO:314 -- max() of an empty block cannot happen: a flush is only reached
            # after at least four zeros were appended to the block.
            if _builtin_max(blk_val) != 0:
                peaks.append(blk_idx[int(np.argmax(blk_val))])
            blk_idx, blk_val = [], []
            zeros = 0
    if blk_val:
        if _builtin_max(blk_val) != 0:
            peaks.append(blk_idx[int(np.argmax(blk_val))])
    return peaks


def _phi0_sine(y0, background, frequency):
    if y0 == background:
        return 0
    if y0 < background:
        return -1 / (4 * frequency) if frequency != 0 else -1.0 / (4.0 * 0.00000000001)
    if y0 > background:
        return +1 / (4 * frequency) if frequency != 0 else +1.0 / (4.0 * 0.00000000001)
    # NaN in the data: the reference leaves phi0 unbound -> UnboundLocalError
    raise UnboundLocalError("phi0")


def fit_sine(y, maxima, log=None):
    """PO:380-491.  chi-square of |A| sin(2 pi f x + phi) + |bg| with A = bg = |max-min|/2."""
    n = len(y)
    x = np.array(range(n))
    amp = abs(_builtin_max(y) - _builtin_min(y)) / 2.
    freq = float(maxima / (n - 1.))
    bg = abs(_builtin_max(y) - _builtin_min(y)) / 2.

    def resid(p, x, y, amp, bg):
        f, phi = p
        if TWIN == 2:
            return y - (abs(amp) * np.sin(2 * np.pi * (f * x) + phi) + abs(bg))
        if TWIN:
            return (y - abs(bg)) - abs(amp) * np.sin(phi + (2 * np.pi * x) * f)
        return y - (abs(amp) * np.sin(2 * np.pi * f * x + phi) + abs(bg))

    phi0 = _phi0_sine(y[0], bg, freq)
    p = _lsq(resid, (freq, phi0), (x, y, amp, bg), log, FIT_SINE)
    f, phi = p
    with np.errstate(all="ignore"):
        fit = abs(amp) * np.sin(2 * np.pi * f * x + phi) + abs(bg)
        return _chisq(y, fit, n) / n


def fit_sine_sqr(y, maxima, log=None):
    """PO:495-587.  NOTE the residual adds |bg| *outside* the bracket (PO:522) while the
    evaluated curve has it inside (PO:529); both are reproduced."""
    n = len(y)
    x = np.array(range(n))
    amp = abs(_builtin_max(y) - _builtin_min(y)) / 2.
    freq = float(maxima / (n - 1.) / 2.)
    bg = abs(_builtin_max(y) - _builtin_min(y)) / 2.

    def resid(p, x, y, amp, bg):
        f, phi = p
        if TWIN == 2:
            sn = np.sin(2 * np.pi * (f * x) + phi)
            return y - abs(amp) * (sn * sn) + abs(bg)
        if TWIN:
            sn = np.sin(phi + (2 * np.pi * x) * f)
            return (y + abs(bg)) - abs(amp) * (sn * sn)
        return y - (abs(amp) * pow(np.sin(2 * np.pi * f * x + phi), 2)) + abs(bg)

    if y[0] == 0:
        phi0 = 0
    else:
        phi0 = -1 / (4 * freq) if freq != 0 else -1.0 / (4.0 * 0.00000000001)
    p = _lsq(resid, (freq, phi0), (x, y, amp, bg), log, FIT_SINESQ)
    f, phi = p
    with np.errstate(all="ignore"):
        fit = abs(amp) * pow(np.sin(2 * np.pi * f * x + phi), 2) + abs(bg)
        return _chisq(y, fit, n) / n


def sinusoid_fittings(profile, log=None):
    """Scores 1-4 (PO:190-376).  Returns [s1, s2, s3, s4] as floats."""
    profile = np.asarray(profile)
    p_mean = profile.mean()
    p_std = profile.std()
    p_max = profile.max()
    p_min = profile.min()
    half = abs(p_max - p_min) / 2.
    # PO:246-247: s += half - p[i], left to right
    s4 = 0
    for v in profile:
        s4 += half - v
    norm = profile - p_mean - p_std
    norm[norm < 0] = 0
    peaks = count_peaks(norm)
    maxima = len(peaks)
    n_diff = max(maxima - 1, 0)          # len(delete(peaks,0) - delete(peaks,maxima-1)), PO:348-351
    with np.errstate(all="ignore"):
        # numpy float64 / int 0 -> inf or nan (never ZeroDivisionError), PO:373-374
        s1 = np.float64(fit_sine(profile, maxima, log)) / maxima
        s2 = np.float64(fit_sine_sqr(profile, maxima, log)) / maxima
    return [float(s1), float(s2), float(n_diff), float(s4)]


# ==========================================================================================
# Scores 5-11                                                              PO:595-1483
# ==========================================================================================
def freedman_diaconis_bins(data):
    """POI:138-166.  The exponent literal is -0.3333333, not -1/3."""
    iqr = _stats.scoreatpercentile(data, 75) - _stats.scoreatpercentile(data, 25)
    binwidth = 2 * iqr * pow(len(data), -0.3333333)
    if binwidth <= 0:
        binwidth = 60
    rnge = _builtin_max(data) - _builtin_min(data)
    # POI:158: with the integer fallback width 60 and integer data this is int/int under
    # Python 2 (floor), e.g. range 252 -> 4 bins, not ceil(4.2) = 5.
    return int(np.ceil(_classic_div(rnge, binwidth)))


def derivative(y):
    """POI:170-186: dy[i] = y[i] - y[i+1]."""
    return [y[i] - y[i + 1] for i in range(len(y) - 1)]


def _histogram(data, nbins):
    try:
        return np.histogram(data, nbins)
    except ValueError:
        raise CandidateFailure(ST_HIST_ZERO_BINS)


def _gexp(x, mu, sigma):
    """exp(-((x-mu)/sigma)^2 / 2) as the reference writes it, or re-associated for the twin."""
    if TWIN == 2:
        return np.exp(-((x - mu) ** 2) / (2 * sigma * sigma))
    if TWIN:
        return np.exp(-0.5 * ((x - mu) * (1.0 / sigma)) ** 2)
    return np.exp((-((x - mu) / sigma) ** 2) / 2)


def _gauss3(x, sigma, mu, maximum):
    return abs(maximum) * _gexp(x, mu, sigma)


def fit_gaussian_hist(edges, counts, log=None, slot=FIT_PROFHIST):
    """PO:774-983: 3-parameter Gaussian on histogram (left edges, counts) with the
    retry loop that moves the start position (at most 6 fits).  Returns (params, fwhm, chisq)."""
    x = np.asarray(edges)
    y = np.asarray(counts)
    first_max = int(np.argmax(y))
    expect = x[first_max]
    sigma0 = np.std(y)
    max0 = _builtin_max(y)
    meansq = np.mean(y) ** 2
    temp = y
    if len(x) == len(y) + 1:
        x = x[0:-1]
    n0 = len(x)
    counter = 0
    nfits = 0
    while True:
        p0 = [sigma0, expect, max0]
        if len(p0) > len(x):                          # PO:943-947 zero padding
            for _ in range(len(p0) - len(x)):
                x = np.append(x, 0)
                y = np.append(y, 0)

        def resid(p, x, y):
            s, mu, mx = p
            return y - _gauss3(x, s, mu, mx)

        p = _lsq(resid, p0, (x, y), log, slot)
        nfits += 1
        with np.errstate(all="ignore"):
            fwhm = abs(FWHM_K_NP * p[0])
            fit = _gauss3(x, p[0], p[1], p[2])
            chisq = _chisq(y, fit, n0) / len(y)     # PO:962 vs PO:965 (D5)
        if (chisq > meansq * n0) & (p[0] < 0.2 * n0):
            counter += 1
            try:
                temp = np.delete(temp, first_max)
                pos = int(np.argmax(temp))
                expect = x[pos + counter]
            except IndexError:
                raise CandidateFailure(ST_INDEX_OOB)
            except ValueError:
                raise CandidateFailure(ST_EMPTY_SEQUENCE)
            if counter > 5:
                break
        else:
            break
    if log is not None:
        log.ier[FIT_DYHIST_COUNT if slot == FIT_DYHIST else FIT_PROFHIST_COUNT] = nfits
    return p, fwhm, chisq


FWHM_K_NP = 2 * np.sqrt(2 * np.log(2))


def fit_gaussian_fixed(edges, counts, bins, log=None):
    """PO:988-1057: Gaussian with the expectation value fixed to x[bins//2 - 1]."""
    x = np.asarray(edges)
    y = np.asarray(counts)
    if len(x) == len(y) + 1:
        x = x[0:-1]
    sigma0 = np.std(y)
    max0 = _builtin_max(y)
    xmax = x[int(_div(bins, 2)) - 1]

    def resid(p, x, y, xmax):
        s, mx = p
        return y - (abs(mx) * _gexp(x, xmax, s))

    if len(y) < 2:
        raise CandidateFailure(ST_LSQ_IMPROPER)          # leastsq: N=2 must not exceed M
    p = _lsq(resid, [sigma0, max0], (x, y, xmax), log, FIT_FIXED)
    with np.errstate(all="ignore"):
        fwhm = abs(FWHM_K_NP * p[0])
        fit = abs(p[1]) * _gexp(x, xmax, p[0])
        chisq = _chisq(y, fit, len(y)) / len(y)
    return p, fwhm, chisq, fit, xmax


def _rotate_half(y):
    """PO:1098-1110 / 1158-1170.  Under Python 2 ``scale()`` (POI:206) is int/int = 0 so the
    'near border' test is always true and the data are always rotated by cut = n//2."""
    n = len(y)
    xmax = int(np.argmax(y))
    if PY2_DIV:
        t = (0 * (1 - ((xmax - 0) // (n - 0)))) + (128 * ((xmax - 0) // (n - 0)))
    else:
        t = (0 * (1 - ((xmax - 0) / (n - 0)))) + (128 * ((xmax - 0) / (n - 0)))
    if (t < 15) or (t >= 112):
        cut = int(np.ceil(_div(n, 2)))
        return list(y[cut:]) + list(y[:cut]), cut
    return y, 0


def fit_gaussian_bg(y, log=None, slot=FIT_T1):
    """PO:1194-1264: Gaussian + signed background, 4 parameters, x = 0..n-1."""
    n = len(y)
    x = np.array(range(n))
    expect = int(np.argmax(y))
    max0 = y[expect]
    sigma0 = np.std(y)

    def model(x, p):
        s, mu, mx, bg = p
        return abs(mx) * _gexp(x, mu, abs(s)) + (bg)

    def resid(p, x, y):
        return y - model(x, p)

    p = _lsq(resid, [sigma0, expect, max0, 1.], (x, y), log, slot)
    with np.errstate(all="ignore"):
        fwhm = abs(FWHM_K_NP * p[0])
        fit = model(x, p)
        chisq = _chisq(y, fit, n) / n
    return p, fwhm, chisq, fit


def fit_gaussian_t1(y, log=None):
    """PO:1061-1132."""
    yrot, _cut = _rotate_half(y)
    return fit_gaussian_bg(yrot, log)


def _del(lst, idx):
    """numpy.delete(arr, idx) on a list: negative wraps, out of range -> IndexError."""
    n = len(lst)
    if idx < -n or idx >= n:
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (idx, n))
    del lst[idx]


def _dg_strip_peak(y):
    """PO:1305-1354: remove the main peak (and its flanks) from (x, y).  ``y`` is a Python
    list here (always rotated), so negative indexes wrap and y[n] raises IndexError."""
    n = len(y)
    pos = int(np.argmax(y))
    newx = list(range(n))
    newy = list(y)
    _del(newx, pos)
    _del(newy, pos)
    tol, limit = 0, 5
    y = list(y)
    i = 1
    while i < n:
        if ((pos - i) > 0) & ((pos + i) < n):
            if (y[pos - i] < y[pos - i + 1]) & (y[pos + i] < y[pos + i - 1]):
                for _ in range(2):
                    _del(newx, pos - i)
                    _del(newy, pos - i)
            elif (y[pos - i] >= y[pos - i + 1]) or ((y[pos + i] >= y[pos + i - 1]) & (tol < limit)):
                for _ in range(2):
                    _del(newx, pos - i)
                    _del(newy, pos - i)
                tol += 1
            else:
                break
        elif (pos - i) < 0:
            if y[pos + i] < y[pos + i - 1]:
                _del(newx, pos - i + 1)
                _del(newy, pos - i + 1)
            elif (y[pos + i] >= y[pos + i - 1]) & (tol < limit):
                _del(newx, pos - i + 1)
                _del(newy, pos - i + 1)
                tol += 1
            else:
                break
        elif (pos + i) > n:
            if y[pos - i] < y[pos - i + 1]:
                _del(newx, pos - i + 1)
                _del(newy, pos - i + 1)
            elif (y[pos - i] >= y[pos - i + 1]) & (tol < limit):
                _del(newx, pos - i)
                _del(newy, pos - i)
                tol += 1
            else:
                break
        i += 1
    return newx, newy


def _gauss4abs(x, p):
    s, mu, mx, bg = p
    return abs(mx) * _gexp(x, mu, s) + abs(bg)


def _gauss8(x, p):
    s1, m1, a1, b1, s2, m2, a2, b2 = p
    return ((abs(a1) * _gexp(x, m1, abs(s1))) +
            (abs(a2) * _gexp(x, m2, abs(s2))) + (abs(b1) + abs(b2)) / 2)


def fit_double_gaussian(y, log=None):
    """PO:1268-1483.  Returns (params8, fwhm_a, chisq, fit, fwhm_b); only min(fwhm_a, fwhm_b)
    is used downstream, so the differing tuple layouts of PO:1426/1428 collapse."""
    n = len(y)
    xfull = list(range(n))
    xarr = np.array(xfull)
    yarr = np.asarray(y, dtype=np.float64)
    newx, newy = _dg_strip_peak(y)
    newx = np.array(newx)
    newy = np.array(newy, dtype=np.float64) if len(newy) else np.array(newy)
    store = {}

    def resid(p, x, y):
        return y - _gauss4abs(x, p)

    for rnd in range(8):
        try:
            npos = int(np.argmax(newy))
        except ValueError:
            raise CandidateFailure(ST_EMPTY_SEQUENCE)
        np0 = [np.std(newy), newx[npos], _builtin_max(newy), np.mean(newy)]
        if len(np0) > len(newx):
            for _ in range(len(np0) - len(newx)):
                newx = np.append(newx, 0)
                newy = np.append(newy, 0)
        p = _lsq(resid, np0, (newx, newy), log, FIT_DG_ROUND0 + rnd)
        with np.errstate(all="ignore"):
            nfwhm = abs(FWHM_K_NP * p[0])
            # PO:1391 evaluates point by point; numpy's scalar and SIMD exp differ by an ulp
            # now and then, so the scalar form is kept.
            ev = np.array([_gauss4abs(xi, p) for xi in xfull])
            lo = p[2] - (1.5 * nfwhm) / 2
            hi = p[2] + (1.5 * nfwhm) / 2
            below = ev <= yarr
            inwin = (ev > yarr) & (xarr > lo) & (xarr < hi)
            newy = np.where(below, yarr - ev + p[3], np.where(inwin, p[3], yarr))
        newx = xarr
        if rnd == 6:
            store[2] = p
        elif rnd == 7:
            store[1] = p
    p1, p2 = store[1], store[2]
    p = list(p1) + list(p2)

    def resid8(q, x, y):
        return y - _gauss8(x, q)

    q = _lsq(resid8, np.array(p), (xfull, yarr), log, FIT_DG_FINAL)
    with np.errstate(all="ignore"):
        fwhm1 = abs(FWHM_K_NP * q[0])
        fwhm2 = abs(FWHM_K_NP * q[4])
        fit = _gauss8(xarr, q)
        mask = fit >= 1.
        chisq = _chisq(y, fit, n, div=n, mask=mask)
        fit1 = _gauss4abs(xarr, p1)
        fit2 = _gauss4abs(xarr, p2)
        combi = fit1 + fit2 - p1[3] - p2[3] + (p1[3] + p2[3]) / 2
        cmask = combi >= 1.
        combi_chisq = _chisq(y, combi, n, div=n, mask=cmask)
        combi_fwhm1 = abs(FWHM_K_NP * p[0])
        combi_fwhm2 = abs(FWHM_K_NP * p[4])
    if chisq <= combi_chisq:
        return q, fwhm1, chisq, fit, fwhm2
    return np.array(p), combi_fwhm2, combi_chisq, combi, combi_fwhm1


def fit_double_gaussian_t2(y, log=None):
    """PO:1136-1190."""
    yrot, _cut = _rotate_half(y)
    return fit_double_gaussian(yrot, log)


def gaussian_fittings(profile, log=None):
    """Scores 5-11 (PO:595-770).  Returns ([s5..s11], status_bits)."""
    profile = np.asarray(profile)
    status = 0
    bins = freedman_diaconis_bins(profile)
    dy = derivative(profile)
    dy_bins = freedman_diaconis_bins(dy)
    h_dy = _histogram(dy, dy_bins)
    (dy_sigma, dy_expect, dy_max), _, _ = fit_gaussian_hist(h_dy[1], h_dy[0], log, FIT_DYHIST)
    h_p = _histogram(profile, bins)
    (ph_sigma, ph_expect, ph_max), _, _ = fit_gaussian_hist(h_p[1], h_p[0], log, FIT_PROFHIST)
    (fx_sigma, fx_max), _fx_fwhm, _fx_chi, _fx_fit, fx_xmax = fit_gaussian_fixed(h_p[1], h_p[0], bins, log)
    with np.errstate(all="ignore"):
        s5 = abs(fx_xmax - ph_expect)
        s6 = abs(fx_max / ph_max)
        s7 = abs(dy_expect - ph_expect)
    p_mean = profile.mean()
    minbg = p_mean if p_mean < ph_expect else ph_expect      # builtin min(ph_expect, mean)
    p_std = profile.std()
    if minbg > 0.:
        temp = []
        for v in profile:
            newy = v - minbg + p_std
            if newy < 0.:
                newy = 0.
            temp.append(newy)
    else:
        temp = profile
    t1_p, t1_fwhm, t1_chi, t1_fit = fit_gaussian_t1(temp, log)
    s8, s9 = t1_fwhm, t1_chi
    try:
        _q, fw_a, dg_chi, dg_fit, fw_b = fit_double_gaussian_t2(profile, log)
        with np.errstate(all="ignore"):
            diff = dg_fit - (t1_fit + minbg - p_std)
            dstd = float(abs(diff.std()))
        if dstd < 3.:
            s10 = t1_fwhm
        else:
            s10 = float(fw_b if fw_b < fw_a else fw_a)        # builtin min(fw_a, fw_b)
        s11 = dg_chi
    except IndexError:
        s10 = 1000000
        s11 = 1000000
        status |= ST_DG_SENTINEL
    return [float(s5), float(s6), float(s7), float(s8), float(s9), float(s10), float(s11)], status


# ==========================================================================================
# Scores 12-19
# ==========================================================================================
def filter_score(s, value, epsilon=0.000005):
    """CFI:84-149."""
    if s in (13, 14):
        if abs(value - 0.0) > epsilon and value < 0.0:
            return 0.0
        return value
    if s == 18:
        return float(abs(value))
    return value


def cand_params_phcx(snr, dm, bary_period_s, width):
    """PHXO:81-112 + PHX:547-585.  Returns [s12..s15] and the hidden state (period_ms, ...)."""
    period = float(bary_period_s) * 1000
    return [float(period), filter_score(13, float(snr)), filter_score(14, float(dm)), float(width)]


def dm_curve_reduce(block):
    """PHXO:237-259: max of the first 127 values of every 128-block; x index = 128k - 1."""
    result, x, temp = [], [], []
    for i in range(len(block)):
        if (i + 1) % 128 == 0:
            result.append(_builtin_max(temp))
            x.append(i - 128)
            temp = []
        else:
            temp.append(block[i])
    return np.array(result), np.array(x)


def dm_fittings_phcx(curve_y, curve_x, length_all, snr, dm, period_ms, width, dm_start, dm_end,
                     log=None):
    """Scores 16-19 for PHCX (PHXO:121-233).  ``curve_y, curve_x`` = dm_curve_reduce() output."""
    y = np.asarray(curve_y)
    length = len(y)
    dm_step = abs(dm_start - dm_end) / length_all
    wint = (width * period_ms) ** 2
    kdm = 8.3 * 10 ** 6
    df = 400
    f = 1374
    with np.errstate(all="ignore"):
        peak = snr / np.sqrt((period_ms - np.sqrt(wint)) / np.sqrt(wint))
        x = np.array([dm_start + curve_x[i] * dm_step for i in range(length)])
        helper = []
        for i in range(length):
            weff = np.sqrt(wint + pow(kdm * abs(dm - x[i]) * df / pow(f, 3), 2))
            helper.append(float(np.sqrt((period_ms - weff) / weff)))
        hmax = _builtin_max(helper)
        theo = (255. / hmax) * np.array(helper)

        def model(x, p):
            amp, prop, shift = p
            if TWIN == 2:
                weff = np.sqrt(wint + pow(prop * (kdm * abs((dm + shift) - x) * df) / pow(f, 3), 2))
                return amp * (np.sqrt(period_ms - weff) / np.sqrt(weff))
            if TWIN:
                weff = np.sqrt(wint + pow((prop * (kdm * df / pow(f, 3))) * abs((dm + shift) - x), 2))
                return amp * np.sqrt(period_ms / weff - 1.0)
            weff = np.sqrt(wint + pow(prop * kdm * abs((dm + shift) - x) * df / pow(f, 3), 2))
            return amp * np.sqrt((period_ms - weff) / weff)

        def resid(p, x, y):
            return y - model(x, p)

        p = _lsq(resid, (255. / hmax, 1, 0), (x, y), log, FIT_DM)
        fit = model(x, p)
        mask = fit >= 1.
        chi_theo = _chisq(y, theo, length, mask=mask) / length
        s17 = abs(1 - p[1])
        s18 = filter_score(18, float(p[2]))
    return [float(peak), float(s17), float(s18), float(chi_theo)]


def cand_params_pfd(profile, bary_p1, bestdm):
    """PFDO:93-233: period, SNR (3-sigma clipped), DM, half-max width after FFT rotation.
    Returns ([s12..s15 after filterScore], (period_ms, snr, dm, width))."""
    profile = np.asarray(profile, dtype=np.float64)
    period = bary_p1 * 1000
    avg = profile.mean()
    var = profile.var()
    sigma = np.sqrt(var)
    keep = (profile > avg - 3 * sigma) & (profile < avg + 3 * sigma)
    clipped = profile[keep]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            avg = clipped.mean()
            var = clipped.var()
            snr = ((profile - avg) / np.sqrt(var)).sum()
    if snr < 0:
        snr = 0.1
    n = len(profile)
    peak = int(profile.argmax())
    shift = peak - _div(n, 2)
    freqs = np.arange(_div(n, 2) + 1, dtype=float)
    phasor = np.exp(complex(0.0, (2.0 * np.pi)) * freqs * shift / float(n))
    rot = np.fft.irfft(phasor * np.fft.rfft(profile)) - _builtin_min(profile)
    peak = int(rot.argmax())
    half = _builtin_max(rot) / 2
    left = peak
    while left > 0:
        if rot[left] < half:
            break
        left -= 1
    right = peak
    while right < len(rot):
        if rot[right] < half:
            break
        right += 1
    width = (1.0 * (right - left - 1.0)) / len(rot)
    scores = [float(period), filter_score(13, float(snr)), filter_score(14, float(bestdm)), float(width)]
    return scores, (float(period), float(snr), float(bestdm), float(width))


def dm_fittings_pfd(chi_curve, dms, snr, dm, period_ms, width, log=None):
    """Scores 16-19 for PFD (PFDO:274-393).  ``chi_curve`` = reduced chi-square per trial DM
    (float32, PFD:393), ``dms`` the trial DMs."""
    yv = np.asarray(chi_curve)
    with np.errstate(all="ignore"):
        y = 255. / _builtin_max(yv) * yv
    length = len(y)
    dm_start, dm_end = float(dms[1]), float(dms[len(dms) - 1])
    dm_step = abs(dm_start - dm_end) / length
    wint = (width * period_ms) ** 2
    kdm = 8.3 * 10 ** 6
    df = 32
    f = 135
    with np.errstate(all="ignore"):
        peak = snr / np.sqrt((period_ms - np.sqrt(wint)) / np.sqrt(wint))
        x = np.array([dm_start + i * dm_step for i in range(length)])
        helper = []
        for i in range(length):
            weff = np.sqrt(wint + pow(kdm * abs(dm - x[i]) * df / pow(f, 3), 2))
            if weff > period_ms:
                weff = period_ms
            helper.append(float(np.sqrt((period_ms - weff) / weff)))
        hmax = _builtin_max(helper)
        theo = (255. / hmax) * np.array(helper)

        def model(x, p):
            amp, prop, shift, up = p
            if TWIN == 2:
                weff = np.sqrt(wint + pow(prop * (kdm * abs((dm + shift) - x) * df) / pow(f, 3), 2))
                weff = np.where(weff > period_ms, period_ms, weff)
                return up + amp * (np.sqrt(period_ms - weff) / np.sqrt(weff))
            if TWIN:
                weff = np.sqrt(wint + pow((prop * (kdm * df / pow(f, 3))) * abs((dm + shift) - x), 2))
                weff = np.where(weff > period_ms, period_ms, weff)
                return up + amp * np.sqrt(period_ms / weff - 1.0)
            weff = np.sqrt(wint + pow(prop * kdm * abs((dm + shift) - x) * df / pow(f, 3), 2))
            weff = np.where(weff > period_ms, period_ms, weff)
            return up + amp * np.sqrt((period_ms - weff) / weff)

        def resid(p, x, y):
            return y - model(x, p)

        p = _lsq(resid, (255. / hmax, 1, 0, 0), (x, y), log, FIT_DM)
        mask = theo > 0
        ndeg = int(np.count_nonzero(mask))
        if ndeg == 0:
            raise CandidateFailure(ST_DM_NDEG_ZERO)
        chi_theo = _seqsum(((y - theo) ** 2 / theo)[mask]) / ndeg
        s17 = abs(1 - p[1])
        s18 = filter_score(18, float(p[2]))
    return [float(peak), float(s17), float(s18), float(chi_theo)]


# ==========================================================================================
# Scores 20-22                                                             PO:1585-1686
# ==========================================================================================
def boxcar_sums(subbands, nbin, nsub, width_bins):
    """PO:1608-1618: S[i][j] = sum_{b<w} sub[i][j+b] accumulated left to right."""
    out = []
    for i in range(nsub):
        row = subbands[i]
        vec = []
        for j in range(nbin - width_bins + 1):
            s = 0
            for b in range(width_bins):
                s += row[j + b]
            vec.append(s)
        out.append(vec)
    return out


def subband_scores(subbands, nbin, nsub, best_width):
    """Scores 20, 21 (PO:1585-1686)."""
    w = int(np.ceil(best_width * nbin))
    sums = boxcar_sums(subbands, nbin, nsub, w)
    max_bins = []
    for i in range(len(sums)):
        max_sum = -10000.0
        max_bin = None
        for j in range(len(sums[i])):
            if sums[i][j] > max_sum:
                max_sum = sums[i][j]
                max_bin = j + _div(w, 2)
        if max_bin is None:
            if not max_bins:
                raise CandidateFailure(ST_SUBBAND_BAD_WIDTH)      # UnboundLocalError, PO:1630
            max_bin = max_bins[-1]                                # stale value of previous band
        max_bins.append(float(max_bin))
    med = np.array(max_bins).mean()
    count = 0
    var_med = 0.0
    for mb in max_bins:
        if abs(mb - med) <= float(w):
            count += 1
            var_med += pow(mb - med, 2)
    if count > 1:
        var = var_med / float(count - 1)
    else:
        mean = 0
        var = 0
        for mb in max_bins:
            mean += mb
        mean /= float(len(max_bins))
        for mb in max_bins:
            var += pow(mb - mean, 2)
        if len(max_bins) - 1 == 0:
            raise CandidateFailure(ST_SUBBAND_NO_PAIRS)
        var /= float(len(max_bins) - 1)
    stdev = np.sqrt(var)
    m = 0
    total = 0.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            cc_all = None
            if FAST_CORR and len(sums) > 1 and len(sums[0]) > 0:
                cc_all = np.corrcoef(np.asarray(sums, dtype=np.float64))
            for i in range(len(sums)):
                for k in range(i + 1, len(sums)):
                    if cc_all is not None:
                        cc = cc_all[i][k]
                    else:
                        cc = np.corrcoef(sums[i], sums[k])[0][1]
                    if str(cc) == "nan":
                        continue
                    total += cc
                    m += 1
    if m == 0 or w == 0:
        raise CandidateFailure(ST_SUBBAND_NO_PAIRS)              # ZeroDivisionError PO:1681/1684
    return float(stdev / float(w)), float(total / float(m))


def profile_corr(subbands, nsub, profile, pairwise=True):
    """Score 22 (PHXO:387-415 / PFDO:445-466 + PHXO:341-345)."""
    total = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            for j in range(nsub):
                c = abs(np.corrcoef(subbands[j], profile))[0][1]
                if c > 0.0055:
                    total += c
    return float(total)


# ==========================================================================================
# Whole-candidate drivers                                                 PHX:383-409
# ==========================================================================================
class Result(object):
    __slots__ = ("features", "status", "ier", "nfev", "cost", "params")

    def __init__(self, nfeat):
        self.features = np.full(nfeat, np.nan)
        self.status = 0
        self.ier = np.zeros(NIER, dtype=np.int8)
        self.nfev = np.zeros(NFIT, dtype=np.int32)
        self.cost = np.full(NFIT, np.nan)
        self.params = {}


def _exc_status(e):
    if isinstance(e, CandidateFailure):
        return e.status
    if isinstance(e, IndexError):
        return ST_INDEX_OOB
    if isinstance(e, ZeroDivisionError):
        return ST_SUBBAND_NO_PAIRS
    if isinstance(e, TypeError):
        return ST_LSQ_IMPROPER
    return ST_OTHER


def bates22_phcx(profile, subbands, dm_curve_y, params, log=None):
    """All 22 scores for one PHCX-family candidate.

    profile      int array [nbin]
    subbands     int array [nsub][nbin]
    dm_curve_y   reduced DM curve [n_dm] (dm_curve_reduce output; x index = 128k-1,
                 length_all = 128*n_dm unless params[6] > 0 gives it explicitly)
    params       (snr, dm, bary_period_s, width, dm_start, dm_end[, length_all])
    """
    res = Result(22)
    log = log if log is not None else FitLog()
    profile = np.asarray(profile)
    snr, dm, bary_p, width, dm_start, dm_end = [float(v) for v in params[:6]]
    n_dm = len(dm_curve_y)
    length_all = int(params[6]) if len(params) > 6 and params[6] > 0 else 128 * n_dm
    try:
        res.features[0:4] = sinusoid_fittings(profile.copy(), log)
        g, st = gaussian_fittings(profile.copy(), log)
        res.features[4:11] = g
        res.status |= st
        cp = cand_params_phcx(snr, dm, bary_p, width)
        res.features[11:15] = cp
        curve_x = 128 * np.arange(n_dm) - 1
        res.features[15:19] = dm_fittings_phcx(np.asarray(dm_curve_y), curve_x, length_all, snr, dm,
                                               bary_p * 1000, width, dm_start, dm_end, log)
        nsub = len(subbands)
        nbin = len(subbands[0])
        s20, s21 = subband_scores(np.asarray(subbands), nbin, nsub, width)
        s22 = profile_corr(np.asarray(subbands), nsub, profile)
        res.features[19:22] = [s20, s21, s22]
    except Exception as e:       # DP:517-523: the row is omitted and the path logged
        res.status |= _exc_status(e)
    res.ier = log.ier
    res.nfev = log.nfev
    res.cost = log.cost
    res.params = log.params
    return res


def bates22_pfd(profile, subbands, chi_curve, dms, bary_p1, bestdm, log=None):
    """All 22 scores for one PFD-family candidate (arrays already extracted from the .pfd:
    scaled profile PFD:256-310, sub-bands PFD:442-456, chi-square curve PFD:378-423)."""
    res = Result(22)
    log = log if log is not None else FitLog()
    profile = np.asarray(profile, dtype=np.float64)
    try:
        res.features[0:4] = sinusoid_fittings(profile.copy(), log)
        g, st = gaussian_fittings(profile.copy(), log)
        res.features[4:11] = g
        res.status |= st
        cp, (period, snr, dm, width) = cand_params_pfd(profile, bary_p1, bestdm)
        res.features[11:15] = cp
        res.features[15:19] = dm_fittings_pfd(chi_curve, dms, snr, dm, period, width, log)
        nsub = len(subbands)
        nbin = len(subbands[0])
        s20, s21 = subband_scores(np.asarray(subbands), nbin, nsub, width)
        s22 = profile_corr(np.asarray(subbands), nsub, profile)
        res.features[19:22] = [s20, s21, s22]
    except Exception as e:
        res.status |= _exc_status(e)
    res.ier = log.ier
    res.nfev = log.nfev
    res.cost = log.cost
    res.params = log.params
    return res


def bates22_pfd_arrays(profile, subbands, chi_curve, params, log=None):
    """PFD-family scoring of already extracted arrays with the candidate scalars GIVEN in ``params``
    (snr, dm, bary_period_s, width, dm_start, dm_end, ...), i.e. everything of PFD.compute()
    (PFDFile.py:587-613) except the derivation of period/SNR/width from the profile
    (PFDOperations.py:93-233, see cand_params_pfd).  This is what the engine's PFX_FAMILY_PFD path
    computes: float64 profile and sub-bands, float32 chi-square curve, the 4-parameter DM model with
    df = 32, f = 135 and the theo > 0 chi-square (PFDOperations.py:274-393)."""
    res = Result(22)
    log = log if log is not None else FitLog()
    profile = np.asarray(profile, dtype=np.float64)
    snr, dm, bary_p, width, dm_start, dm_end = [float(v) for v in params[:6]]
    try:
        res.features[0:4] = sinusoid_fittings(profile.copy(), log)
        g, st = gaussian_fittings(profile.copy(), log)
        res.features[4:11] = g
        res.status |= st
        res.features[11:15] = [bary_p * 1000, filter_score(13, snr), filter_score(14, dm), width]
        n = len(chi_curve)
        # trial DMs whose element 1 and last are dm_start / dm_end (PFDO:337)
        dms = np.concatenate([[dm_start], np.linspace(dm_start, dm_end, n - 1)]) if n > 2 else np.array([dm_start, dm_end])
        res.features[15:19] = dm_fittings_pfd(np.asarray(chi_curve), dms, snr, dm, bary_p * 1000, width, log)
        nsub = len(subbands)
        nbin = len(subbands[0])
        s20, s21 = subband_scores(np.asarray(subbands), nbin, nsub, width)
        s22 = profile_corr(np.asarray(subbands), nsub, profile)
        res.features[19:22] = [s20, s21, s22]
    except Exception as e:
        res.status |= _exc_status(e)
    res.ier = log.ier
    res.nfev = log.nfev
    res.cost = log.cost
    res.params = log.params
    return res


# ==========================================================================================
# Output rows                                                             DP:305-325,405-447
# ==========================================================================================
def fmt_float(v, py2=True):
    """Python 2 ``str(float)`` = '%.12g' (SURVEY 8c D3); integers-valued floats keep '.0'."""
    v = float(v)
    if not py2:
        return repr(v)
    if v != v:
        return "nan"
    if v in (float("inf"), float("-inf")):
        return "inf" if v > 0 else "-inf"
    s = "%.12g" % v
    if "." not in s and "e" not in s and "n" not in s:
        s += ".0"
    return s


def format_row(candidate, scores, arff=False, py2=True):
    """DP:321-325 (CSV) / DP:421-425 (ARFF): join, then replace 'nan' and 'inf' by '0'
    over the whole line (this also mangles paths containing those substrings)."""
    joined = ",".join(fmt_float(s, py2) for s in scores)
    line = (joined + ",?%" + candidate) if arff else (candidate + "," + joined)
    return line.replace("nan", "0").replace("inf", "0")
