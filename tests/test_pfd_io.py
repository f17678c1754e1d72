"""CPU tests of the .pfd reader: the numpy reader and the native scanner reproduce the
arrays the REFERENCE's PFDFile code produces (golden vectors generated with the reference itself,
tests/golden/make_golden_pfd.py; live comparison when /root/reference is present), bit for bit; the
oracle scores those arrays like the reference's PFD.compute() does."""
import contextlib
import io
import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import pfx_oracle as O
import refshim
from pulsarfeatureextractor_b200 import pfd_io, phcx_io

GOLD = os.path.join(ROOT, "tests", "golden", "ref_pfd_v1.npz")


def _files(tmp_path):
    g = np.load(GOLD)
    paths = []
    for k, (seed, npart, nsub, proflen, posn, big) in enumerate(g["cases"]):
        p = str(tmp_path / ("cand_%02d.pfd" % k))
        pfd_io.write_pfd(p, with_position=bool(posn), big_endian=bool(big),
                         **pfd_io.synth_pfd(int(seed), npart=int(npart), nsub=int(nsub), proflen=int(proflen)))
        paths.append(p)
    return g, paths


def _ensure_io():
    if not phcx_io.native_available():
        import __graft_entry__ as ge
        ge.build()


def test_readers_reproduce_reference_arrays(tmp_path):
    _ensure_io()
    g, paths = _files(tmp_path)
    scan = pfd_io.PfdScan(paths, threads=3)
    for k, p in enumerate(paths):
        r = pfd_io.read_pfd(p)
        assert scan.ok(k)
        prof, sub, dmc, par = scan.gather([k])
        for name, a, b in (("profile", r.profile, prof[0]), ("subbands", r.subbands, sub[0]), ("chis", r.dmcurve, dmc[0])):
            ref = g["%s_%d" % (name, k)]
            assert a.dtype == ref.dtype and np.array_equal(a, ref), (k, name, "numpy reader")
            assert b.dtype == ref.dtype and np.array_equal(b, ref), (k, name, "native scanner")
        dms = g["dms_%d" % k]
        head = g["head_%d" % k]
        for q in (r.params, par[0]):
            assert q[1] == head[0] and q[2] == head[1]                 # best DM, barycentric period
            assert q[4] == dms[1] and q[5] == dms[-1] and q[6] == 100   # PFDOperations.py:337-338
    scan.close()


def test_malformed_files_fail_in_both_readers(tmp_path):
    _ensure_io()
    g, paths = _files(tmp_path)
    raw = open(paths[0], "rb").read()
    bad = []
    for name, data in (("short_header", raw[:60]), ("truncated_cube", raw[: len(raw) // 2]), ("empty", b""),
                       ("one_dm", None)):
        p = str(tmp_path / (name + ".pfd"))
        if data is None:
            kw = pfd_io.synth_pfd(3)
            kw["dms"] = kw["dms"][:1]
            pfd_io.write_pfd(p, **kw)
        else:
            with open(p, "wb") as fh:
                fh.write(data)
        bad.append(p)
    scan = pfd_io.PfdScan(bad + paths[:2])
    for i, p in enumerate(bad):
        with pytest.raises(Exception):
            pfd_io.read_pfd(p)
        assert not scan.ok(i), p
    assert scan.ok(len(bad)) and scan.ok(len(bad) + 1)
    scan.close()


def test_oracle_scores_the_arrays_like_the_reference(tmp_path):
    """PFD.compute() of the reference (golden) against the oracle fed with the reader's arrays:
    deterministic scores to 1e-9, the fits on the usual noise floor."""
    warnings.simplefilter("ignore")
    g, paths = _files(tmp_path)
    det = [2, 3, 11, 12, 13, 14, 15, 19, 20, 21]
    close = 0
    n = 0
    for k, p in enumerate(paths):
        r = pfd_io.read_pfd(p)
        ref = g["bates_%d" % k]
        res = O.bates22_pfd(r.profile, r.subbands, r.dmcurve, r.dms, r.params[2], r.params[1])
        failed = ("fail_%d" % k) in g.files
        assert failed == bool(res.status & O.ST_FAIL_MASK), k
        # the reference evaluates the DM-curve statistics of a .pfd in float32 (the chi-square curve goes
        # to numpy / scipy as it is, PFDFile.py:567-574): the oracle does the same and is exact; the
        # float64 evaluation the kernels make of the same float32 samples agrees to float32 rounding
        # (mean, std) and to its cancellation noise (skew, kurtosis)
        gl = g["lyon_%d" % k]
        assert np.array_equal(O.lyon8_pfd(r.profile, r.dmcurve), gl, equal_nan=True), k
        ly = O.lyon8(r.profile, r.dmcurve)
        assert np.allclose(ly[:6], gl[:6], rtol=2e-6, atol=0) and np.allclose(ly[6:], gl[6:], rtol=0, atol=5e-3), k
        if failed:
            continue
        n += 1
        for c in det:
            a, b = ref[c], res.features[c]
            assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-9 * abs(a) + 1e-12, (k, c + 1, a, b)
        e = np.abs(ref - res.features) / np.maximum(np.abs(ref), 1e-12)
        close += int(np.nanmax(e) <= 1e-6)
    assert n >= 6 and close >= n - 1, (n, close)


@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present")
def test_live_against_the_reference_reader(tmp_path):
    _ensure_io()
    warnings.simplefilter("ignore")
    PFD = refshim.load_reference()["PFDFile"].PFD
    paths = []
    shapes = [(4, 16, 64), (8, 32, 128), (8, 64, 50), (16, 64, 64), (3, 8, 200)]
    for seed in range(100, 115):
        npart, nsub, proflen = shapes[seed % len(shapes)]
        p = str(tmp_path / ("live_%d.pfd" % seed))
        pfd_io.write_pfd(p, with_position=(seed % 4 != 0), big_endian=(seed % 5 == 0),
                         **pfd_io.synth_pfd(seed, npart=npart, nsub=nsub, proflen=proflen))
        paths.append(p)
    scan = pfd_io.PfdScan(paths)
    for k, p in enumerate(paths):
        with contextlib.redirect_stdout(io.StringIO()):
            c = PFD(False, p)
            prof = np.array(c.profile)
            sub = np.array(c.plot_subbands())
            chis, dms = c.plot_chi2_vs_DM(c.dms[0], c.dms[-1])
        a, b, d, par = scan.gather([k])
        assert np.array_equal(a[0], prof) and np.array_equal(b[0], sub) and np.array_equal(d[0], chis), k
        assert par[0][4] == dms[1] and par[0][5] == dms[-1] and par[0][1] == c.bestdm and par[0][2] == c.bary_p1
    scan.close()
