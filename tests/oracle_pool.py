"""
TEST INFRASTRUCTURE: run the oracle (oracle/pfx_oracle.py) and its two re-associated twins over a
batch in a process pool.  Used by the GPU parity tests, scripts/gpu_parity_table.py and nothing in
the product.

    rows = run(batch, family="phcx", twins=True, fast_corr=False, workers=None)
    rows.feat [n, 22], rows.status [n], rows.ier [n, 18], rows.lyon [n, 8]
    rows.t1 / rows.t2: the same for the twins (features and ier), NaN / 0 where not computed
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


class Rows(object):
    pass


def _one(args):
    import pfx_oracle as O
    warnings.simplefilter("ignore")
    family, twins, fast_corr, items = args
    O.FAST_CORR = bool(fast_corr)
    out = []
    for prof, sub, dmc, dms, par in items:
        if family == "phcx":
            call = lambda: O.bates22_phcx(prof.astype(np.int64), sub.astype(np.int64), dmc.astype(np.int64), par)
        else:
            call = lambda: O.bates22_pfd_arrays(prof, sub, dmc, par)
        O.TWIN = False
        r = call()
        rec = [r.features, r.status, np.asarray(r.ier, dtype=np.int8), np.asarray(O.lyon8(prof, dms))]
        for level in (1, 2):
            if twins and not (r.status & O.ST_FAIL_MASK):
                O.TWIN = level
                try:
                    t = call()
                finally:
                    O.TWIN = False
                rec += [t.features, np.asarray(t.ier, dtype=np.int8)]
            else:
                rec += [np.full(22, np.nan), np.zeros(18, np.int8)]
        out.append(rec)
    return out


def run(batch, family="phcx", twins=True, fast_corr=False, workers=None, index=None):
    import multiprocessing as mp
    idx = list(range(len(batch))) if index is None else [int(i) for i in index]
    workers = workers or min(len(idx), os.cpu_count() or 1)
    items = [(batch.profile[i], batch.subbands[i], batch.dmcurve[i], batch.dmstat[i], batch.params[i]) for i in idx]
    step = max(1, min(16, (len(items) + workers - 1) // workers))
    jobs = [(family, twins, fast_corr, items[a:a + step]) for a in range(0, len(items), step)]
    if workers > 1:
        with mp.get_context("fork").Pool(workers) as pool:
            parts = pool.map(_one, jobs)
    else:
        parts = [_one(j) for j in jobs]
    recs = [r for p in parts for r in p]
    rows = Rows()
    rows.index = np.array(idx)
    rows.feat = np.array([r[0] for r in recs])
    rows.status = np.array([r[1] for r in recs], dtype=np.int64)
    rows.ier = np.array([r[2] for r in recs])
    rows.lyon = np.array([r[3] for r in recs])
    rows.t1 = np.array([r[4] for r in recs])
    rows.t1_ier = np.array([r[5] for r in recs])
    rows.t2 = np.array([r[6] for r in recs])
    rows.t2_ier = np.array([r[7] for r in recs])
    return rows


CACHE_DIR = os.path.join(ROOT, "tests", "golden", "_cache")
_FIELDS = ("index", "feat", "status", "ier", "lyon", "t1", "t1_ier", "t2", "t2_ier")


def cached_run(key, batch, **kw):
    """run() with the result kept in tests/golden/_cache/<key>.npz (git-ignored; it travels to the GPU
    box with the snapshot, so the minutes of oracle time are spent once in the build container).
    The cache is keyed by name and checked against a digest of the inputs; absent or stale -> computed."""
    import hashlib
    idx = list(range(len(batch))) if kw.get("index") is None else [int(i) for i in kw["index"]]
    h = hashlib.sha1()
    for a in (batch.profile[idx], batch.dmcurve[idx], batch.params[idx]):
        h.update(np.ascontiguousarray(a).tobytes())
    h.update(repr(sorted((k, v) for k, v in kw.items() if k not in ("index", "workers"))).encode())
    digest = h.hexdigest()
    path = os.path.join(CACHE_DIR, key + ".npz")
    if os.path.isfile(path):
        z = np.load(path)
        if str(z["digest"]) == digest:
            rows = Rows()
            for f in _FIELDS:
                setattr(rows, f, z[f])
            return rows
    rows = run(batch, **kw)
    os.makedirs(CACHE_DIR, exist_ok=True)
    np.savez_compressed(path, digest=digest, **{f: getattr(rows, f) for f in _FIELDS})
    return rows


def rel(a, c):
    """element-wise relative error with an absolute floor of 1e-12; equal (or both NaN) -> 0"""
    a = np.asarray(a, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    with np.errstate(all="ignore"):
        e = np.abs(a - c) / np.maximum(np.abs(a), 1e-12)
    e[(a == c) | (np.isnan(a) & np.isnan(c))] = 0
    e[np.isnan(e)] = np.inf
    return e


LM = [0, 1, 4, 5, 6, 7, 8, 9, 10, 16, 17]         # least-squares derived columns (score number - 1)
DET = [2, 3, 11, 12, 13, 14, 15, 18, 19, 20, 21]  # deterministic columns


def tier_table(rows, gpu_feat, gpu_ier, cols=LM):
    """Per least-squares score: Tier-W size (twin 1 reproduces the oracle to 1e-8), the fraction of
    Tier W within 1e-6 for the GPU and for twin 2, and p50/p90/p99 of the relative error of the GPU
    and of both twins over all candidates both twins were computed for."""
    ok = (rows.status & 0x3df) == 0
    ok &= ~np.isnan(rows.t1).all(axis=1)
    e_g, e_1, e_2 = rel(rows.feat[ok], gpu_feat[ok]), rel(rows.feat[ok], rows.t1[ok]), rel(rows.feat[ok], rows.t2[ok])
    table = {}
    for k in cols:
        w = e_1[:, k] <= 1e-8
        q = lambda e: [float(np.quantile(e[:, k], p)) if len(e) else None for p in (0.5, 0.9, 0.99)]
        table["score%d" % (k + 1)] = {
            "n": int(ok.sum()), "tier_w": int(w.sum()),
            "gpu_within_1e-6_in_tier_w": float(np.mean(e_g[w, k] <= 1e-6)) if w.any() else None,
            "twin2_within_1e-6_in_tier_w": float(np.mean(e_2[w, k] <= 1e-6)) if w.any() else None,
            "gpu_p50_p90_p99": q(e_g), "twin1_p50_p90_p99": q(e_1), "twin2_p50_p90_p99": q(e_2),
            "gpu_within_1e-6_all": float(np.mean(e_g[:, k] <= 1e-6)), "twin1_within_1e-6_all": float(np.mean(e_1[:, k] <= 1e-6)),
            "gpu_bit_exact": float(np.mean(e_g[:, k] == 0)), "twin1_bit_exact": float(np.mean(e_1[:, k] == 0)),
        }
    slots = {}
    for s in range(16):
        slots["slot%d" % s] = {"gpu": float(np.mean(gpu_ier[ok, s] == rows.ier[ok, s])),
                               "twin1": float(np.mean(rows.t1_ier[ok, s] == rows.ier[ok, s])),
                               "twin2": float(np.mean(rows.t2_ier[ok, s] == rows.ier[ok, s]))}
    allier = {"gpu": float(np.mean((gpu_ier[ok, :16] == rows.ier[ok, :16]).all(axis=1))),
              "twin1": float(np.mean((rows.t1_ier[ok, :16] == rows.ier[ok, :16]).all(axis=1))),
              "twin2": float(np.mean((rows.t2_ier[ok, :16] == rows.ier[ok, :16]).all(axis=1)))}
    return {"scores": table, "ier_agreement_per_slot": slots, "ier_agreement_all_slots": allier}
