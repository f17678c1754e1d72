"""CPU tests of the host side of the CLI (DataProcessor / Candidate / ScoreGenerator mirrors): file
discovery, the native readers, batching by shape, failure logging and the three writers, with the
GPU engine replaced by a stand-in that answers with the oracle's scores.  The stand-in lives in this
test only; the package itself has no CPU path."""
import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import pfx_oracle as O
from pulsarfeatureextractor_b200 import pfd_io, phcx_io, synth
from pulsarfeatureextractor_b200.DataProcessor import DataProcessor
from pulsarfeatureextractor_b200.engine import Scores


class OracleEngine(object):
    """Same methods as Engine.score_host / score_pfd_host, scores from oracle/pfx_oracle.py."""

    def __init__(self):
        self.calls = []

    def score_host(self, profile, subbands=None, dmcurve=None, dmstat=None, params=None, features="all", **_kw):
        warnings.simplefilter("ignore")
        n = profile.shape[0]
        self.calls.append(("phcx", n, features))
        ncol = {"bates": 22, "lyon": 8, "all": 30}[features]
        feat = np.full((n, ncol), np.nan)
        status = np.zeros(n, np.int32)
        for i in range(n):
            col = 0
            if features in ("bates", "all"):
                r = O.bates22_phcx(profile[i].astype(np.int64), subbands[i].astype(np.int64),
                                   dmcurve[i].astype(np.int64), params[i])
                feat[i, :22] = r.features
                status[i] = r.status
                col = 22
            if features in ("lyon", "all"):
                feat[i, col:col + 8] = O.lyon8(profile[i], dmstat[i])
        return Scores(feat, status, None)

    def score_pfd_host(self, profile, subbands, dmcurve, params, features="all"):
        warnings.simplefilter("ignore")
        n = profile.shape[0]
        self.calls.append(("pfd", n, features))
        ncol = {"bates": 22, "lyon": 8, "all": 30}[features]
        feat = np.full((n, ncol), np.nan)
        status = np.zeros(n, np.int32)
        for i in range(n):
            col = 0
            if features in ("bates", "all"):
                r = O.bates22_pfd_arrays(profile[i], subbands[i], dmcurve[i], params[i])
                feat[i, :22] = r.features
                status[i] = r.status
                col = 22
            if features in ("lyon", "all"):
                feat[i, col:col + 8] = O.lyon8(profile[i], dmcurve[i])
        return Scores(feat, status, None)


@pytest.fixture(scope="module")
def cand_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("cands")
    b = synth.generate(51, 0, 4096)
    const = int(np.nonzero(b.edge == 0)[0][0])           # constant profile: the reference raises (zero bins)
    pick = list(range(10)) + [const]
    for i in pick:
        phcx_io.write_phcx(str(d / ("cand_%04d.phcx.gz" % i)), b.profile[i], b.subbands[i], b.dmcurve[i],
                           b.dmstat[i], b.params[i])
    with open(str(d / "broken_0001.phcx.gz"), "wb") as fh:
        fh.write(b"not a gzip file")
    (d / "sub").mkdir()
    phcx_io.write_phcx(str(d / "sub" / "cand_9000.phcx"), b.profile[3], b.subbands[3], b.dmcurve[3], b.dmstat[3],
                       b.params[3], gz=False)               # SUPERB-style, found by the *.phcx pattern only
    for k in range(3):
        pfd_io.write_pfd(str(d / ("pfd_%02d.pfd" % k)), **pfd_io.synth_pfd(k + 1))
    return d, b, pick


def test_collective_csv_and_error_log(cand_dir, tmp_path):
    d, b, pick = cand_dir
    eng = OracleEngine()
    out, log = tmp_path / "scores.csv", tmp_path / "err.txt"
    dp = DataProcessor(False, engine=eng, error_log=str(log), batch=4)
    ok, bad = dp.processPHCXCollectively(str(d) + "/", False, str(out), False, False, False)
    assert (ok, bad) == (10, 2)                       # the constant profile and the broken file fail
    lines = open(out).read().strip().split("\n")
    assert len(lines) == 10
    got = {ln.split(",")[0]: ln for ln in lines}
    for i in pick[:10]:
        path = os.path.join(str(d) + "/", "cand_%04d.phcx.gz" % i)
        r = O.bates22_phcx(b.profile[i].astype(np.int64), b.subbands[i].astype(np.int64),
                           b.dmcurve[i].astype(np.int64), phcx_io.read_phcx(path).params)
        assert got[path] == O.format_row(path, r.features), path
    logged = open(log).read().split()
    assert len(logged) == 2 and any("broken_0001" in p for p in logged)
    assert all(n <= 4 for _f, n, _k in eng.calls)     # batches of at most `batch` candidates


def test_arff_dmprof_and_separate_modes(cand_dir, tmp_path):
    d, b, pick = cand_dir
    eng = OracleEngine()
    out = tmp_path / "lyon.arff"
    dp = DataProcessor(False, engine=eng, error_log=str(tmp_path / "e1.txt"))
    ok, bad = dp.dmprofPHCX(str(d) + "/", False, str(out), True, False)
    assert ok == 11 and bad == 1                      # Lyon statistics never raise for a readable file
    text = open(out).read()
    head, rows = text.split("@data\n")
    assert "@attribute Profile_mean numeric" in head and "@attribute DM_kurtosis numeric" in head
    rows = rows.strip().split("\n")
    assert len(rows) == 11 and all(r.count(",") == 8 and ",?%" in r for r in rows)
    i = pick[0]
    path = os.path.join(str(d) + "/", "cand_%04d.phcx.gz" % i)
    assert O.format_row(path, O.lyon8(b.profile[i], b.dmstat[i]), arff=True) in rows
    # SUPERB pattern finds the un-gzipped file in the sub-directory (section 0 holds everything)
    dp2 = DataProcessor(False, engine=eng, error_log=str(tmp_path / "e2.txt"))
    ok, bad = dp2.processSUPERBCollectively(str(d) + "/", False, str(tmp_path / "superb.csv"), False, False, False)
    assert (ok, bad) == (1, 0)
    # separate mode: one <candidate>.dat per candidate, no newline, no candidate name
    dp3 = DataProcessor(False, engine=eng, error_log=str(tmp_path / "e3.txt"))
    dp3.superb = True
    ok, bad = dp3.processSeparately(str(d / "sub") + "/", False, [dp3.superbRegex], False)
    dat = open(str(d / "sub" / "cand_9000.phcx.dat")).read()
    assert ok == 1 and "\n" not in dat and dat.count(",") == 21


def test_pfd_files_through_the_cli_classes(cand_dir, tmp_path):
    d, b, pick = cand_dir
    eng = OracleEngine()
    out = tmp_path / "pfd.csv"
    dp = DataProcessor(False, engine=eng, error_log=str(tmp_path / "e.txt"))
    ok, bad = dp.processPFDCollectively(str(d) + "/", False, str(out), False, False, False)
    assert ok + bad == 3 and [k for k, _n, _f in eng.calls] == ["pfd"] * len(eng.calls)
    rows = open(out).read().strip().split("\n") if ok else []
    assert len(rows) == ok
    for r in rows:
        path = r.split(",")[0]
        c = pfd_io.read_pfd(path)
        ref = O.bates22_pfd_arrays(c.profile, c.subbands, c.dmcurve, c.params)
        assert r == O.format_row(path, ref.features)


def test_dmprof_accepts_files_whose_bates_tags_are_broken(tmp_path):
    """--dmprof touches only the profile of the section and DataBlock[0] (PHCXFile.py:320-379,
    PHCXOperations.py:528-539): a file whose <Width> is not a number or whose <SubBands> is missing gets its
    eight Lyon statistics, while the Bates modes log it as a failure -- in the native scanner, in the Python
    reader and through DataProcessor."""
    import gzip
    d = tmp_path / "cands"
    d.mkdir()
    b = synth.generate(52, 0, 8)
    good = str(d / "cand_good.phcx.gz")
    phcx_io.write_phcx(good, b.profile[0], b.subbands[0], b.dmcurve[0], b.dmstat[0], b.params[0])
    text = gzip.open(good, "rb").read()
    assert text.count(b"<Width") == 2 and text.count(b"<SubBands") == 2
    bad_width = str(d / "cand_badwidth.phcx.gz")
    with gzip.open(bad_width, "wb") as fh:
        fh.write(text.replace(b"<Width units='periods'>", b"<Width units='periods'>x"))
    no_sub = str(d / "cand_nosub.phcx.gz")
    with gzip.open(no_sub, "wb") as fh:
        fh.write(text.replace(b"<SubBands", b"<SubBandz").replace(b"</SubBands>", b"</SubBandz>"))
    no_prof = str(d / "cand_noprofile.phcx.gz")
    with gzip.open(no_prof, "wb") as fh:
        fh.write(text.replace(b"<Profile", b"<Profilx").replace(b"</Profile>", b"</Profilx>"))
    paths = [good, bad_width, no_sub, no_prof]
    # Python reader
    for p in (bad_width, no_sub):
        with pytest.raises(Exception):
            phcx_io.read_phcx(p)
        c = phcx_io.read_phcx(p, lyon_only=True)
        assert np.array_equal(c.profile, b.profile[0]) and np.array_equal(c.dmstat, b.dmstat[0])
    with pytest.raises(Exception):
        phcx_io.read_phcx(no_prof, lyon_only=True)
    # native scanner
    if phcx_io.native_available():
        scan = phcx_io.Scan(paths)
        assert [scan.ok(k) for k in range(4)] == [True, False, False, False]
        assert [scan.ok(k, True) for k in range(4)] == [True, True, True, False]
        prof, dms = scan.gather_lyon([0, 1, 2])
        assert all(np.array_equal(prof[k], b.profile[0]) and np.array_equal(dms[k], b.dmstat[0]) for k in range(3))
        with pytest.raises(RuntimeError):
            scan.gather([0, 1])            # a full gather refuses the file that is only lyon_ok
        scan.close()
    # through the CLI classes
    eng = OracleEngine()
    dp = DataProcessor(False, engine=eng, error_log=str(tmp_path / "e_lyon.txt"))
    ok, bad = dp.dmprofPHCX(str(d) + "/", False, str(tmp_path / "lyon.csv"), False, False)
    assert (ok, bad) == (3, 1)
    rows = open(tmp_path / "lyon.csv").read().strip().split("\n")
    want = O.format_row("x", O.lyon8(b.profile[0], b.dmstat[0])).split(",", 1)[1]
    assert len(rows) == 3 and all(r.split(",", 1)[1] == want for r in rows)
    dp2 = DataProcessor(False, engine=eng, error_log=str(tmp_path / "e_bates.txt"))
    ok, bad = dp2.processPHCXCollectively(str(d) + "/", False, str(tmp_path / "bates.csv"), False, False, False)
    assert (ok, bad) == (1, 3)
