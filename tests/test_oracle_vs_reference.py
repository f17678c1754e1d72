"""
Pins the oracle against the reference.

1. test_golden_*: against tests/golden/ref_phcx_v1.npz, outputs of the reference itself
   (executed through oracle/refshim.py by tests/golden/make_golden.py).  Travels everywhere.
2. test_live_*: when /root/reference is present (the build container), run the reference's own
   code through the shim on fresh seeds and edge cases and compare.

The oracle calls the same numpy/scipy routines as the reference, in the same order, so the
comparison is exact (bit for bit), not toleranced.
"""
import os
import warnings

import numpy as np
import pytest

import pfx_oracle as O
import refshim

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_phcx_v1.npz")


def _same(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.array_equal(a, b, equal_nan=True)


def test_golden_bates_and_lyon():
    g = np.load(GOLD)
    n = len(g["index"])
    assert n >= 150 and g["failed"].sum() >= 5
    warnings.simplefilter("ignore")
    for k in range(0, n, 2):          # every other candidate keeps the CPU suite short
        r = O.bates22_phcx(g["profile"][k].astype(np.int64), g["subbands"][k].astype(np.int64),
                           g["dmcurve"][k].astype(np.int64), g["params"][k])
        assert bool(r.status & O.ST_FAIL_MASK) == bool(g["failed"][k]), (k, hex(r.status))
        if not g["failed"][k]:
            assert _same(r.features, g["bates"][k]), (k, r.features, g["bates"][k])
        assert _same(O.lyon8(g["profile"][k], g["dmstat"][k]), g["lyon"][k]), k


@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present")
def test_live_profile_scores_random_and_edges():
    from pulsarfeatureextractor_b200 import synth
    warnings.simplefilter("ignore")
    mods = refshim.load_reference()
    po = mods["PHCXOperations"].PHCXOperations(False)
    b = synth.generate(21, 0, 4096)
    pick = list(range(24)) + [int(np.nonzero(b.edge == k)[0][0]) for k in range(synth.N_EDGE_KINDS)]
    for i in pick:
        prof = b.profile[i].astype(np.int64)
        try:
            ref = [float(v) for v in po.getSinusoidFittings(prof.copy())] + \
                  [float(v) for v in po.getGaussianFittings(prof.copy())]
            err = None
        except Exception as e:      # noqa: BLE001 - the reference raises generic exceptions
            ref, err = None, e
        try:
            got = O.sinusoid_fittings(prof.copy()) + O.gaussian_fittings(prof.copy())[0]
            gerr = None
        except Exception as e:      # noqa: BLE001
            got, gerr = None, e
        assert (err is None) == (gerr is None), (i, repr(err), repr(gerr))
        if err is None:
            assert _same(ref, got), (i, ref, got)


@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present")
def test_live_subband_scores():
    warnings.simplefilter("ignore")
    mods = refshim.load_reference()
    po = mods["PHCXOperations"].PHCXOperations(False)
    rng = np.random.default_rng(3)
    for width in (0.01, 0.05, 0.2):
        sub = np.round(rng.uniform(0, 255, (16, 128))).astype(np.int64)
        ref = po.getSubband_scores(sub, 128, 16, width)
        got = O.subband_scores(sub, 128, 16, width)
        assert _same(ref, got)
