"""Golden vectors for the .pfd reader: outputs of the REFERENCE's own PFDFile / PFDOperations code
(run under oracle/refshim.py, needs /root/reference) on synthetic .pfd files that
pulsarfeatureextractor_b200.pfd_io.synth_pfd / write_pfd regenerate from the seeds stored here.

Per file: the scaled profile (PFD.getprofile), the sub-bands (PFD.plot_subbands), the reduced
chi-square curve and its trial DMs (PFD.plot_chi2_vs_DM), the 22 scores of PFD.compute() and the 8
Lyon statistics (computeProfileStatScores + computeDMCurveStatScores).

usage: python tests/golden/make_golden_pfd.py   -> tests/golden/ref_pfd_v1.npz
"""
import contextlib
import io
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

CASES = [  # seed, npart, nsub, proflen, with_position, big_endian
    (1, 8, 16, 64, True, False), (2, 8, 32, 64, False, False), (3, 4, 64, 64, True, False),
    (4, 8, 16, 128, True, True), (5, 6, 16, 50, True, False), (6, 8, 32, 64, True, False),
    (7, 8, 16, 64, True, False), (8, 16, 16, 64, False, True), (9, 8, 64, 64, True, False),
    (10, 8, 16, 100, True, False), (11, 8, 16, 64, True, False), (12, 8, 32, 64, True, False),
]


def main():
    import refshim
    from pulsarfeatureextractor_b200 import pfd_io
    warnings.simplefilter("ignore")
    PFD = refshim.load_reference()["PFDFile"].PFD
    d = tempfile.mkdtemp()
    out = {"cases": np.array(CASES, dtype=np.int64)}
    for k, (seed, npart, nsub, proflen, posn, big) in enumerate(CASES):
        p = os.path.join(d, "cand_%02d.pfd" % k)
        pfd_io.write_pfd(p, with_position=bool(posn), big_endian=bool(big),
                         **pfd_io.synth_pfd(seed, npart=npart, nsub=nsub, proflen=proflen))
        with contextlib.redirect_stdout(io.StringIO()):
            c = PFD(False, p)
            out["profile_%d" % k] = np.array(c.profile, dtype=np.float64)
            out["subbands_%d" % k] = np.array(c.plot_subbands(), dtype=np.float64)
            chis, dms = c.plot_chi2_vs_DM(c.dms[0], c.dms[-1])
            out["chis_%d" % k] = np.array(chis)
            out["dms_%d" % k] = np.array(dms, dtype=np.float64)
            out["head_%d" % k] = np.array([c.bestdm, c.bary_p1], dtype=np.float64)
            try:
                out["bates_%d" % k] = np.array(PFD(False, p).compute(), dtype=np.float64)
            except Exception as e:          # noqa: BLE001 - the candidate fails in the reference
                out["bates_%d" % k] = np.full(22, np.nan)
                out["fail_%d" % k] = np.array([1])
            out["lyon_%d" % k] = np.array(PFD(False, p).computeProfileStatScores() +
                                          PFD(False, p).computeDMCurveStatScores(), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "ref_pfd_v1.npz"), **out)
    print("wrote ref_pfd_v1.npz:", len(CASES), "files")


if __name__ == "__main__":
    main()
