"""
CPU tests against the second set of golden files made with the reference itself
(tests/golden/make_golden_v2.py, reference executed through oracle/refshim.py):

* ref_superb_v1.npz  -- SUPERB candidates (SUPERBPHCXFile.py:80,102-113,242: uncompressed, section 0,
  64-bin profiles): the oracle reproduces SUPERBPHCX.compute() and the Lyon statistics bit for bit;
  the Python reader and the native scanner give the stored arrays from the same files.
* ref_writer_v1.npz  -- the text DataProcessor.storeScore / storeScoreARFF / outputScores emit
  (DataProcessor.py:305-325,405-447): pfx_format_rows and the oracle's format_row reproduce it
  byte for byte.
"""
import os
import warnings

import numpy as np
import pytest

import pfx_oracle as O
import refshim
from pulsarfeatureextractor_b200 import phcx_io, synth

HERE = os.path.dirname(os.path.abspath(__file__))
SUPERB = os.path.join(HERE, "golden", "ref_superb_v1.npz")
WRITER = os.path.join(HERE, "golden", "ref_writer_v1.npz")


def _same(a, b):
    return np.array_equal(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), equal_nan=True)


def test_oracle_reproduces_superb_reference():
    g = np.load(SUPERB)
    n = len(g["index"])
    assert n >= 90 and g["failed"].sum() >= 4 and g["valid"].all()
    warnings.simplefilter("ignore")
    for k in range(0, n, 2):
        r = O.bates22_phcx(g["profile"][k].astype(np.int64), g["subbands"][k].astype(np.int64),
                           g["dmcurve"][k].astype(np.int64), g["params"][k])
        assert bool(r.status & O.ST_FAIL_MASK) == bool(g["failed"][k]), (k, hex(r.status))
        if not g["failed"][k]:
            assert _same(r.features, g["bates"][k]), (k, r.features, g["bates"][k])
        assert _same(O.lyon8(g["profile"][k], g["dmstat"][k]), g["lyon"][k]), k


def test_readers_on_superb_files(tmp_path):
    """Files written like the golden generator wrote them: read_phcx and the native scanner select
    section 0 for an uncompressed .phcx and return the golden arrays."""
    g = np.load(SUPERB)
    b = synth.generate(13, 0, 2048, nbin=64)
    paths = []
    pick = [0, 5, 17, 40]
    for k in pick:
        i = int(g["index"][k])
        p = str(tmp_path / ("cand_%05d.phcx" % i))
        phcx_io.write_phcx(p, b.profile[i], b.subbands[i], b.dmcurve[i], b.dmstat[i], b.params[i], superb=True)
        paths.append(p)
        c = phcx_io.read_phcx(p)
        assert np.array_equal(c.profile, g["profile"][k]) and np.array_equal(c.subbands, g["subbands"][k])
        assert np.array_equal(c.dmcurve, g["dmcurve"][k]) and np.array_equal(c.dmstat, g["dmstat"][k])
        assert np.array_equal(c.params, g["params"][k])
    if not phcx_io.native_available():
        pytest.skip("libpfxio.so not built")
    scan = phcx_io.Scan(paths)
    assert all(scan.ok(i) for i in range(len(paths)))
    prof, sub, dmc, dms, par = scan.gather(list(range(len(paths))))
    assert np.array_equal(prof, g["profile"][pick]) and np.array_equal(sub, g["subbands"][pick])
    assert np.array_equal(dmc, g["dmcurve"][pick]) and np.array_equal(dms, g["dmstat"][pick])
    assert np.array_equal(par, g["params"][pick])
    scan.close()


def test_row_text_equals_reference_writer():
    """The three writers of the reference, byte for byte: native formatter and oracle."""
    g = np.load(WRITER)
    names = [str(s) for s in g["names"]]
    n = len(names)
    assert n >= 380
    for i in range(n):
        row = g["scores"][i, :int(g["ncol"][i])]
        assert O.format_row(names[i], row) == str(g["csv"][i]), i
        assert O.format_row(names[i], row, arff=True) == str(g["arff"][i]), i
    if not phcx_io.native_available():
        pytest.skip("libpfxio.so not built")
    for ncol in sorted(set(int(c) for c in g["ncol"])):
        sel = np.nonzero(g["ncol"] == ncol)[0]
        feat = np.ascontiguousarray(g["scores"][sel, :ncol])
        nm = [names[i] for i in sel]
        assert phcx_io.format_rows(feat, nm, None, "csv", True, threads=2) == "".join(str(g["csv"][i]) + "\n" for i in sel)
        assert phcx_io.format_rows(feat, nm, None, "arff", True, threads=2) == "".join(str(g["arff"][i]) + "\n" for i in sel)
        assert phcx_io.format_rows(feat, nm, None, "bare", True, threads=2) == "".join(str(g["dat"][i]) + "\n" for i in sel)


@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present")
def test_live_reference_writer_random_rows(tmp_path, monkeypatch):
    """Fresh random rows through the reference's own storeScore under the shim."""
    monkeypatch.chdir(tmp_path)
    mods = refshim.load_reference()
    dp = mods["DataProcessor"].DataProcessor(False)
    rng = np.random.default_rng(5)
    vals = rng.normal(size=(60, 22)) * 10.0 ** rng.integers(-20, 20, (60, 22))
    ints = rng.random((60, 22)) < 0.1
    vals[ints] = rng.integers(-50, 50, int(ints.sum())).astype(np.float64)
    vals[3, 4] = np.nan
    vals[7, 0] = np.inf
    vals[9, 21] = -np.inf
    names = ["/x/inf/nan_%d.phcx.gz" % i for i in range(60)]
    for i in range(60):
        dp.storeScore(names[i], [float(v) for v in vals[i]], "unused")
    assert [O.format_row(names[i], vals[i]) for i in range(60)] == dp.scoreStore
    if phcx_io.native_available():
        assert phcx_io.format_rows(vals, names, None, "csv", True) == "\n".join(dp.scoreStore) + "\n"


def test_profile_mode_rows_equal_reference(tmp_path):
    """--profile (raw profile bins instead of scores, PHCXFile.py:291-304): the CLI classes write the rows
    the reference's computeProfileScores + storeScore / storeScoreARFF produce (tests/golden/
    ref_profile_mode_v1.npz), for HTRU .phcx.gz (section 1, 128 bins) and SUPERB .phcx (section 0, 64 bins)."""
    from pulsarfeatureextractor_b200.DataProcessor import DataProcessor
    g = np.load(os.path.join(HERE, "golden", "ref_profile_mode_v1.npz"))
    for kind, nbin, method in (("phcx", 128, "processPHCXCollectively"), ("superb", 64, "processSUPERBCollectively")):
        d = tmp_path / kind
        d.mkdir()
        b = synth.generate(17, 0, 8, nbin=nbin)
        sel = [i for i in range(len(g["names"])) if str(g["kinds"][i]) == kind]
        for i in sel:
            k = int(str(g["names"][i]).split("_")[2].split(".")[0])
            phcx_io.write_phcx(str(d / str(g["names"][i])), b.profile[k], b.subbands[k], b.dmcurve[k], b.dmstat[k],
                               b.params[k], superb=(kind == "superb"))
        for arff in (False, True):
            out = tmp_path / ("%s_%d.txt" % (kind, arff))
            dp = DataProcessor(False, engine=None, error_log=str(tmp_path / "err.txt"))
            ok, bad = getattr(dp, method)(str(d) + "/", False, str(out), arff, True, False)
            assert (ok, bad) == (len(sel), 0)
            lines = [l for l in open(out).read().split("\n") if l and not l.startswith("@")]
            want = {}
            for i in sel:
                want[str(g["names"][i])] = str(g["arff" if arff else "csv"][i])
            got = {}
            for l in lines:
                name = os.path.basename(l.split(",")[0] if not arff else l.split("?%")[1])
                got[name] = l.replace(str(d) + "/", "")
            assert got == want, (kind, arff)
            if arff:
                head = open(out).read()
                assert head.count("@attribute Score") == nbin and "@attribute class {0,1}" in head
