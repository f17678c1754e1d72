"""
Generates two more golden files from the REFERENCE ITSELF (executed through oracle/refshim.py,
sources read in place under /root/reference).  Run in the build container only:

    python tests/golden/make_golden_v2.py

ref_superb_v1.npz   SUPERB candidates (uncompressed two-section .phcx, 64-bin profile, 16 x 64
                    sub-bands, section 0 for everything: SUPERBPHCXFile.py:80,102-113,242): input
                    arrays as the packer reads them, the 22 scores of SUPERBPHCX.compute(), the 8
                    Lyon statistics, the failure flag.
ref_writer_v1.npz   text the reference's writer emits: DataProcessor.storeScore (DP:305-325),
                    storeScoreARFF (DP:405-425) and outputScores (DP:427-447) fed with the score
                    rows of ref_phcx_v1.npz (Python floats, as PHCX.compute() returns them), with
                    its Lyon rows (numpy.float64, as DP:886 concatenates them) and with crafted
                    rows and candidate names that exercise the formatting corners ('nan'/'inf'
                    substring replacement over the whole line, -0.0, exponents, integers).
                    Number formatting is Python 2's str(float) (see refshim docstring).
"""
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import refshim  # noqa: E402
from pulsarfeatureextractor_b200 import synth, phcx_io  # noqa: E402


def superb(mods):
    SUPERB = mods["SUPERBPHCXFile"].SUPERBPHCX
    b = synth.generate(13, 0, 2048, nbin=64)
    pick = list(range(72))
    for kind in range(synth.N_EDGE_KINDS):
        pick += list(np.nonzero(b.edge == kind)[0][:2])
    pick = sorted(set(int(i) for i in pick))
    n = len(pick)
    bates = np.full((n, 22), np.nan)
    lyon = np.full((n, 8), np.nan)
    failed = np.zeros(n, dtype=bool)
    valid = np.zeros(n, dtype=bool)
    tmp = tempfile.mkdtemp()
    prof = np.zeros((n, 64), np.uint8)
    sub = np.zeros((n, 16, 64), np.uint8)
    dmc = np.zeros((n, 256), np.uint8)
    dms = np.zeros((n, 256 * 128), np.uint8)
    par = np.zeros((n, 8))
    for k, i in enumerate(pick):
        path = os.path.join(tmp, "cand_%05d.phcx" % i)
        phcx_io.write_phcx(path, b.profile[i], b.subbands[i], b.dmcurve[i], b.dmstat[i], b.params[i], superb=True)
        c = phcx_io.read_phcx(path)
        prof[k], sub[k], dmc[k], dms[k], par[k] = c.profile, c.subbands, c.dmcurve, c.dmstat, c.params
        cand = SUPERB(False, path)
        try:
            bates[k] = np.array(cand.compute(), dtype=np.float64)
        except Exception:
            failed[k] = True
        lyon[k] = np.array(SUPERB(False, path).computeProfileStatScores() +
                           SUPERB(False, path).computeDMCurveStatScores(), dtype=np.float64)
        c2 = SUPERB(False, path)
        c2.load()
        valid[k] = bool(c2.isValid())
    np.savez_compressed(os.path.join(HERE, "ref_superb_v1.npz"), index=np.array(pick), profile=prof, subbands=sub,
                        dmcurve=dmc, dmstat=dms, params=par, bates=bates, lyon=lyon, failed=failed, valid=valid,
                        klass=b.klass[pick], edge=b.edge[pick])
    print("superb: wrote", n, "candidates;", int(failed.sum()), "reference failures;", int(valid.sum()), "valid")


def writer(mods):
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)                     # the constructor creates CandidateErrorLog.txt in the cwd
    try:
        DP = mods["DataProcessor"].DataProcessor
        g = np.load(os.path.join(HERE, "ref_phcx_v1.npz"))
        ok = ~g["failed"]
        names, rows, csv, arff, dat = [], [], [], [], []

        def emit(name, scores_for_reference, row):
            dp = DP(False)
            dp.storeScore(name, scores_for_reference, "unused")
            dp.storeScoreARFF(name, scores_for_reference, "unused")
            dp.outputScores(scores_for_reference, os.path.join(tmp, "one"))
            names.append(name)
            rows.append(np.asarray(row, dtype=np.float64))
            csv.append(dp.scoreStore[0])
            arff.append(dp.scoreStore[1])
            dat.append(open(os.path.join(tmp, "one.dat")).read())

        for k in np.nonzero(ok)[0]:
            emit("/data/cands/cand_%05d.phcx.gz" % int(g["index"][k]), [float(v) for v in g["bates"][k]], g["bates"][k])
        n_bates = len(names)
        for k in range(len(g["lyon"])):
            row = np.concatenate((np.array(list(g["lyon"][k, :4])), np.array(list(g["lyon"][k, 4:]))))   # DP:886
            emit("/data/cands/cand_%05d.phcx.gz" % int(g["index"][k]), row, row)
        n_lyon = len(names) - n_bates
        crafted = [1.0, 1e16, 1e-5, 123456789012345.0, float("nan"), float("inf"), -float("inf"), 0.1 + 0.2,
                   1.0 / 3.0, -0.0, 5e-324, 1e22, 100000000000.0, 1e11 + 0.5, 2.5, 999999999999.5, 0.0001,
                   0.00001234, -1e-300, 1.7976931348623157e308, 12345678901.25, 1e15, 123456789012.0]
        for nm in ("/data/inferno/nancy_001.phcx.gz", "plain.phcx", "/tmp/x/inf.pfd", "nan"):
            emit(nm, list(crafted), crafted)
        width = max(len(r) for r in rows)
        mat = np.full((len(rows), width), np.nan)
        ncol = np.array([len(r) for r in rows])
        for i, r in enumerate(rows):
            mat[i, :len(r)] = r
        np.savez_compressed(os.path.join(HERE, "ref_writer_v1.npz"), names=np.array(names), scores=mat, ncol=ncol,
                            csv=np.array(csv), arff=np.array(arff), dat=np.array(dat),
                            n_bates=n_bates, n_lyon=n_lyon)
        print("writer: wrote", len(names), "rows (%d Bates, %d Lyon, %d crafted)" % (n_bates, n_lyon, len(names) - n_bates - n_lyon))
    finally:
        os.chdir(cwd)


def profile_mode(mods):
    """--profile mode (raw profile bins as the "scores", PHCXFile.py:291-304, SUPERBPHCXFile.py:291-304):
    the reference's computeProfileScores and the CSV / ARFF text its writer makes of them."""
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)
    try:
        DP = mods["DataProcessor"].DataProcessor
        rows, names, csv, arff, kinds = [], [], [], [], []
        for kind, cls, nbin, ext in (("phcx", mods["PHCXFile"].PHCX, 128, ".phcx.gz"),
                                     ("superb", mods["SUPERBPHCXFile"].SUPERBPHCX, 64, ".phcx")):
            b = synth.generate(17, 0, 8, nbin=nbin)
            for i in range(4):
                path = os.path.join(tmp, "cand_%s_%d%s" % (kind, i, ext))
                phcx_io.write_phcx(path, b.profile[i], b.subbands[i], b.dmcurve[i], b.dmstat[i], b.params[i],
                                   superb=(kind == "superb"))
                scores = cls(False, path).computeProfileScores()
                dp = DP(False)
                dp.storeScore(os.path.basename(path), scores, "unused")
                dp.storeScoreARFF(os.path.basename(path), scores, "unused")
                rows.append(np.array(scores, dtype=np.float64))
                names.append(os.path.basename(path))
                csv.append(dp.scoreStore[0])
                arff.append(dp.scoreStore[1])
                kinds.append(kind)
        width = max(len(r) for r in rows)
        mat = np.full((len(rows), width), np.nan)
        for i, r in enumerate(rows):
            mat[i, :len(r)] = r
        np.savez_compressed(os.path.join(HERE, "ref_profile_mode_v1.npz"), names=np.array(names), kinds=np.array(kinds),
                            scores=mat, ncol=np.array([len(r) for r in rows]), csv=np.array(csv), arff=np.array(arff))
        print("profile mode: wrote", len(rows), "rows")
    finally:
        os.chdir(cwd)


def main():
    warnings.simplefilter("ignore")
    mods = refshim.load_reference()
    superb(mods)
    writer(mods)
    profile_mode(mods)


if __name__ == "__main__":
    main()
