"""
Generates tests/golden/ref_phcx_v1.npz: outputs of the REFERENCE ITSELF (executed through
oracle/refshim.py, sources read in place under /root/reference) on seeded synthetic candidates
plus the edge-case corpus.  Run in the build container only:

    python tests/golden/make_golden.py

Stored per candidate: the input arrays (so the file is self-contained), the 22 Bates scores of
PHCX.compute() (PHCXFile.py:383-409), the 8 Lyon statistics of computeProfileStatScores /
computeDMCurveStatScores (PHCXFile.py:320-379), a failure flag (the reference raised) and the
text row DataProcessor.storeScore would emit (DataProcessor.py:305-325) under '%.12g' formatting.
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


def main():
    warnings.simplefilter("ignore")
    mods = refshim.load_reference()
    PHCX = mods["PHCXFile"].PHCX
    # 160 candidates of stream 11 + every edge kind at least twice
    b = synth.generate(11, 0, 4096)
    pick = list(range(160))
    for kind in range(synth.N_EDGE_KINDS):
        pick += list(np.nonzero(b.edge == kind)[0][:3])
    pick = sorted(set(int(i) for i in pick))
    n = len(pick)
    bates = np.full((n, 22), np.nan)
    lyon = np.full((n, 8), np.nan)
    failed = np.zeros(n, dtype=bool)
    tmp = tempfile.mkdtemp()
    prof = np.zeros((n, 128), np.uint8)
    sub = np.zeros((n, 16, 128), np.uint8)
    dmc = np.zeros((n, 256), np.uint8)
    dms = np.zeros((n, 256), np.uint8)
    par = np.zeros((n, 8))
    for k, i in enumerate(pick):
        path = os.path.join(tmp, "cand_%05d.phcx.gz" % i)
        phcx_io.write_phcx(path, b.profile[i], b.subbands[i], b.dmcurve[i], b.dmstat[i], b.params[i])
        c = phcx_io.read_phcx(path)      # what the packer hands to the engine (values as stored in the file)
        prof[k], sub[k], dmc[k], dms[k], par[k] = c.profile, c.subbands, c.dmcurve, c.dmstat, c.params
        try:
            bates[k] = np.array(PHCX(False, path).compute(), dtype=np.float64)
        except Exception:
            failed[k] = True
        lyon[k] = np.array(PHCX(False, path).computeProfileStatScores() +
                           PHCX(False, path).computeDMCurveStatScores(), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "ref_phcx_v1.npz"), index=np.array(pick), profile=prof,
                        subbands=sub, dmcurve=dmc, dmstat=dms, params=par, bates=bates, lyon=lyon,
                        failed=failed, klass=b.klass[pick], edge=b.edge[pick])
    print("wrote", n, "candidates;", int(failed.sum()), "reference failures")


if __name__ == "__main__":
    main()
