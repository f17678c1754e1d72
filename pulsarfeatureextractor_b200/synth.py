"""
Seeded synthetic candidates (SURVEY.md section 8d "Value distributions & seeds").

Candidates are produced in chunks of ``CHUNK`` = 4 096 from
``numpy.random.Generator(Philox(key=[0xB2000000 + config_id, chunk_index]))`` so the CPU
oracle and every GPU rank see identical bytes no matter how a batch is sharded.

Class mix per candidate: 45 % noise, 30 % single Gaussian pulse, 10 % double pulse,
10 % sinusoidal RFI, 5 % edge cases (the corpus of SURVEY.md section 8c).  PHCX-family arrays are
min-max scaled to 0..255 and rounded to ``uint8`` exactly as the file format stores them
(reference PHCXFile.py:169-176 reads two hex digits per sample).

Layout of one batch (structure of arrays, C order):
    profile   [N, nbin]         uint8 (PHCX family) or float64 (PFD family)
    subbands  [N, nsub, nbin]   same dtype as profile
    dmcurve   [N, n_dm]         uint8: DM-SNR curve after the max-of-127 reduction of
                                PHCXOperations.py:237-259 (PHCX) / float32 chi-square curve (PFD)
    dmstat    [N, n_dmstat]     uint8: bytes of DataBlock[0], input of the Lyon DM statistics
                                (PHCXOperations.py:528-539)
    params    [N, 8] float64    snr, dm, bary_period_s, width, dm_start, dm_end, dm_len_all, 0
"""
import numpy as np

CHUNK = 4096
KEY_BASE = 0xB2000000

CLASS_NOISE, CLASS_PULSE, CLASS_DOUBLE, CLASS_RFI, CLASS_EDGE = 0, 1, 2, 3, 4
N_EDGE_KINDS = 14

PARAM_SNR, PARAM_DM, PARAM_PERIOD, PARAM_WIDTH, PARAM_DM_START, PARAM_DM_END, PARAM_DM_LEN_ALL = range(7)
NPARAM = 8


class Batch(object):
    __slots__ = ("profile", "subbands", "dmcurve", "dmstat", "params", "klass", "edge")

    def __init__(self, profile, subbands, dmcurve, dmstat, params, klass, edge):
        self.profile = profile
        self.subbands = subbands
        self.dmcurve = dmcurve
        self.dmstat = dmstat
        self.params = params
        self.klass = klass
        self.edge = edge

    def __len__(self):
        return self.profile.shape[0]

    def slice(self, lo, hi):
        return Batch(self.profile[lo:hi],
                     None if self.subbands is None else self.subbands[lo:hi],
                     None if self.dmcurve is None else self.dmcurve[lo:hi],
                     None if self.dmstat is None else self.dmstat[lo:hi],
                     self.params[lo:hi], self.klass[lo:hi], self.edge[lo:hi])


def _scale_u8(a):
    """min-max scale the last axis to 0..255 and round (constant rows -> 0)."""
    lo = a.min(axis=-1, keepdims=True)
    hi = a.max(axis=-1, keepdims=True)
    rng = hi - lo
    rng = np.where(rng > 0, rng, 1)
    return np.rint((a - lo) / rng * 255.0).astype(np.uint8)


def _dm_shape(period_ms, width, dm, x):
    """Theoretical S/N-vs-DM shape of PHCXOperations.py:200-206 (PHCX constants)."""
    wint = (width * period_ms) ** 2
    weff = np.sqrt(wint + (8.3e6 * np.abs(dm - x) * 400.0 / 1374.0 ** 3) ** 2)
    with np.errstate(all="ignore"):
        h = np.sqrt(np.clip((period_ms - weff) / weff, 0, None))
    return h


def _chunk(config_id, chunk_index, count, nbin, nsub, n_dm, n_dmstat, want_subbands, want_dm):
    rng = np.random.Generator(np.random.Philox(key=[KEY_BASE + config_id, chunk_index]))
    n = count
    f32 = np.float32
    u = rng.random(n)
    klass = np.select([u < 0.45, u < 0.75, u < 0.85, u < 0.95], [0, 1, 2, 3], default=4).astype(np.int8)
    edge = np.where(klass == CLASS_EDGE, rng.integers(0, N_EDGE_KINDS, n), -1).astype(np.int8)

    phase = rng.random(n)
    duty = rng.uniform(0.01, 0.2, n)
    snr_amp = rng.uniform(5, 50, n)
    phase2 = rng.random(n)
    duty2 = rng.uniform(0.01, 0.2, n)
    amp2 = rng.uniform(0.3, 1.0, n)
    cycles = rng.integers(1, 5, n)
    rfi_amp = rng.uniform(2, 10, n)
    rfi_ph = rng.uniform(0, 2 * np.pi, n)

    x = (np.arange(nbin, dtype=f32) / nbin)[None, :]

    def pulse(ph, dt, amp):
        d = x - ph[:, None].astype(f32)
        d = d - np.rint(d)                      # wrap to [-0.5, 0.5)
        sig = (dt[:, None] / 2.355).astype(f32)
        return amp[:, None].astype(f32) * np.exp(-0.5 * (d / sig) ** 2)

    shape = np.zeros((n, nbin), dtype=f32)
    is_p = (klass == CLASS_PULSE) | (klass == CLASS_DOUBLE) | (klass == CLASS_EDGE)
    shape += np.where(is_p[:, None], pulse(phase, duty, snr_amp), 0)
    is_d = klass == CLASS_DOUBLE
    shape += np.where(is_d[:, None], pulse(phase2, duty2, snr_amp * amp2), 0)
    is_r = klass == CLASS_RFI
    shape += np.where(is_r[:, None],
                      rfi_amp[:, None].astype(f32) * np.sin(2 * np.pi * cycles[:, None] * x + rfi_ph[:, None]).astype(f32), 0)

    prof_f = shape + rng.standard_normal((n, nbin), dtype=f32)
    profile = _scale_u8(prof_f)

    # scalars
    period_s = np.exp(rng.uniform(np.log(1e-3), np.log(5.0), n))
    dm = rng.uniform(0, 1000, n)
    width = duty.copy()
    snr = rng.uniform(5, 100, n)
    dm_start = np.zeros(n)
    dm_end = rng.uniform(dm, 2 * dm + 50)

    # ---- edge cases (SURVEY 8c corpus) -------------------------------------------------
    e = edge
    idx = np.nonzero(e == 0)[0]                 # constant profile
    profile[idx] = 77
    idx = np.nonzero(e == 1)[0]                 # dip-only profile
    if idx.size:
        profile[idx] = 200
        profile[idx, nbin // 3: nbin // 3 + 3] = 10
    idx = np.nonzero(e == 2)[0]                 # peak at bin 0
    if idx.size:
        profile[idx] = _scale_u8(pulse(np.zeros(idx.size), duty[idx], snr_amp[idx]) +
                                 rng.standard_normal((idx.size, nbin), dtype=f32))
    idx = np.nonzero(e == 3)[0]                 # peak at bin n-1
    if idx.size:
        profile[idx] = _scale_u8(pulse(np.full(idx.size, (nbin - 1) / nbin), duty[idx], snr_amp[idx]) +
                                 rng.standard_normal((idx.size, nbin), dtype=f32))
    idx = np.nonzero(e == 4)[0]                 # two equal maxima
    if idx.size:
        profile[idx, nbin // 4] = 255
        profile[idx, (3 * nbin) // 4] = 255
    idx = np.nonzero(e == 5)[0]                 # single spike on a flat floor (IQR = 0)
    if idx.size:
        profile[idx] = 3
        profile[idx, nbin // 2] = 255
    idx = np.nonzero(e == 6)[0]                 # width = 0
    width[idx] = 0.0
    idx = np.nonzero(e == 7)[0]                 # width * nbin > nbin
    width[idx] = 1.5
    idx = np.nonzero(e == 10)[0]                # millisecond pulsar: weff > P at the far-DM edge
    period_s[idx] = 0.002
    width[idx] = 0.05
    dm[idx] = 300.0
    dm_end[idx] = 650.0
    idx = np.nonzero(e == 12)[0]                # slightly negative snr / dm
    snr[idx] = -1e-6
    dm[idx] = -2.5
    idx = np.nonzero(e == 13)[0]                # two-level profile: tiny IQR -> very many histogram bins
    if idx.size:
        profile[idx] = 1
        profile[idx, ::3] = 2
        profile[idx, nbin // 2] = 255

    subbands = None
    if want_subbands:
        sb = shape[:, None, :] / np.sqrt(f32(nsub)) + rng.standard_normal((n, nsub, nbin), dtype=f32)
        subbands = _scale_u8(sb)
        idx = np.nonzero(e == 8)[0]             # identical sub-bands
        if idx.size:
            subbands[idx] = subbands[idx, :1, :]
        idx = np.nonzero(e == 9)[0]             # constant sub-bands (all pair correlations NaN)
        subbands[idx] = 9

    dmcurve = None
    dmstat = None
    if want_dm:
        length_all = 128 * n_dm
        step = np.abs(dm_start - dm_end) / length_all
        k = np.arange(n_dm)[None, :]
        xdm = dm_start[:, None] + (128 * k - 1) * step[:, None]
        h = _dm_shape((period_s * 1000)[:, None], width[:, None], dm[:, None], xdm)
        hmax = np.nanmax(h, axis=1, keepdims=True)
        hmax = np.where(hmax > 0, hmax, 1)
        pulsed = ((klass == CLASS_PULSE) | (klass == CLASS_DOUBLE) | (klass == CLASS_EDGE))[:, None]
        noise = rng.standard_normal((n, n_dm)) * 8.0
        curve = np.where(pulsed, 235.0 * h / hmax + noise, 30.0 + noise)
        dmcurve = np.clip(np.rint(curve), 0, 255).astype(np.uint8)
        idx = np.nonzero(e == 11)[0]            # flat DM curve
        dmcurve[idx] = 0
        # section-0 block (FFT section): coarse version of the same curve + noise
        ks = (np.arange(n_dmstat) * n_dm // max(n_dmstat, 1))
        stat = curve[:, ks] * 0.8 + rng.standard_normal((n, n_dmstat)) * 6.0
        dmstat = np.clip(np.rint(stat), 0, 255).astype(np.uint8)

    params = np.zeros((n, NPARAM), dtype=np.float64)
    params[:, PARAM_SNR] = snr
    params[:, PARAM_DM] = dm
    params[:, PARAM_PERIOD] = period_s
    params[:, PARAM_WIDTH] = width
    params[:, PARAM_DM_START] = dm_start
    params[:, PARAM_DM_END] = dm_end
    params[:, PARAM_DM_LEN_ALL] = 128 * n_dm
    c = np.ascontiguousarray
    return Batch(c(profile), None if subbands is None else c(subbands), None if dmcurve is None else c(dmcurve),
                 None if dmstat is None else c(dmstat), c(params), klass, edge)


def generate(config_id, start, count, nbin=128, nsub=16, n_dm=256, n_dmstat=256,
             want_subbands=True, want_dm=True):
    """Candidates [start, start+count) of the stream ``config_id`` (PHCX family, uint8)."""
    parts = []
    pos = start
    end = start + count
    while pos < end:
        c = pos // CHUNK
        lo = pos - c * CHUNK
        hi = min(CHUNK, end - c * CHUNK)
        full = _chunk(config_id, c, CHUNK, nbin, nsub, n_dm, n_dmstat, want_subbands, want_dm)
        parts.append(full.slice(lo, hi) if (lo, hi) != (0, CHUNK) else full)
        pos = c * CHUNK + hi
    if len(parts) == 1:
        return parts[0]

    def cat(name):
        vals = [getattr(p, name) for p in parts]
        return None if vals[0] is None else np.concatenate(vals, axis=0)

    return Batch(cat("profile"), cat("subbands"), cat("dmcurve"), cat("dmstat"),
                 cat("params"), cat("klass"), cat("edge"))


def generate_pfd_shaped(config_id, start, count, nbin=64, nsub=64, n_dm=256):
    """PFD-shaped candidates (BASELINE configs[3]): float64 profile scaled to [0, 255]
    (reference PFDFile.py:282-310), float64 sub-bands (not rescaled: profs.sum(0), PFDFile.py:442-456),
    float32 reduced chi-square curve over n_dm trial DMs (PFDFile.py:393), params with
    dm_start = first trial DM, dm_end = last, dm_len_all = n_dm."""
    b = generate(config_id, start, count, nbin=nbin, nsub=nsub, n_dm=n_dm, n_dmstat=n_dm)
    rng = np.random.Generator(np.random.Philox(key=[KEY_BASE + 7000 + config_id, start]))
    n = len(b)
    prof = b.profile.astype(np.float64) + rng.uniform(-0.4, 0.4, (n, nbin))
    lo = prof.min(axis=1, keepdims=True)
    hi = prof.max(axis=1, keepdims=True)
    prof = np.where(hi > lo, (prof - lo) / np.where(hi > lo, hi - lo, 1) * 255.0, prof)
    sub = b.subbands.astype(np.float64) * 0.37 + rng.normal(0, 0.2, (n, nsub, nbin)) + 1.25
    const = b.edge == 9
    sub[const] = 3.5
    curve = (b.dmcurve.astype(np.float32) / np.float32(17.0) + np.float32(0.5) +
             rng.uniform(0, 0.05, (n, n_dm)).astype(np.float32))
    params = b.params.copy()
    params[:, PARAM_DM_LEN_ALL] = n_dm
    c = np.ascontiguousarray
    return Batch(c(prof), c(sub), c(curve), b.dmstat, c(params), b.klass, b.edge)


def to_pfd_family(batch, n_trials=None):
    """Re-express a uint8 batch as PFD-shaped arrays: float64 profile scaled to [0, 255]
    (reference PFDFile.py:282-310), float64 sub-bands, float32 chi-square curve
    (PFDFile.py:393) and linearly spaced trial DMs from dm_start to dm_end."""
    prof = batch.profile.astype(np.float64)
    sub = None if batch.subbands is None else batch.subbands.astype(np.float64) * 0.37 + 1.25
    curve = None
    if batch.dmcurve is not None:
        curve = (batch.dmcurve.astype(np.float32) / np.float32(17.0) + np.float32(0.5))
    return Batch(prof, sub, curve, batch.dmstat, batch.params.copy(), batch.klass, batch.edge)
