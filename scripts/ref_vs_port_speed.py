"""Build container only (needs /root/reference): is the oracle port a fair stand-in for the reference on
the CPU arm?  Times, on ONE core and on the same 64 synthetic .phcx.gz candidates,
  * the reference's own classes under oracle/refshim.py: PHCX(path).compute() + the two Lyon calls
    (gunzip + DOM parse + scoring: what ScoreGenerator.py does per file), and its scoring alone
    (PHCXOperations methods on arrays already in memory);
  * the oracle port (oracle/pfx_oracle.py) on the same arrays.
Writes profiles/r02_ref_vs_port_speed.json; bench.py quotes it beside cpu_baseline (kind "port")."""
import json, os, sys, tempfile, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
os.environ["OMP_NUM_THREADS"] = "1"
import numpy as np
import pfx_oracle as O
import refshim
from pulsarfeatureextractor_b200 import synth, phcx_io

warnings.simplefilter("ignore")
n = 64
b = synth.generate(2, 0, n)
tmp = tempfile.mkdtemp()
paths = []
for i in range(n):
    p = os.path.join(tmp, "cand_%04d.phcx.gz" % i)
    phcx_io.write_phcx(p, b.profile[i], b.subbands[i], b.dmcurve[i], b.dmstat[i], b.params[i])
    paths.append(p)
mods = refshim.load_reference()
PHCX = mods["PHCXFile"].PHCX
os.chdir(tmp)
t0 = time.time()
for p in paths:
    try:
        PHCX(False, p).compute()
    except Exception:
        pass
    PHCX(False, p).computeProfileStatScores()
    PHCX(False, p).computeDMCurveStatScores()
t_ref_files = time.time() - t0
cands = [phcx_io.read_phcx(p) for p in paths]
t0 = time.time()
for c in cands:
    O.bates22_phcx(c.profile.astype(np.int64), c.subbands.astype(np.int64), c.dmcurve.astype(np.int64), c.params)
    O.lyon8(c.profile, c.dmstat)
t_port = time.time() - t0
# the reference's scoring alone: parse once, then time compute() on the loaded object again
objs = []
for p in paths:
    objs.append(PHCX(False, p))
t0 = time.time()
for o in objs:
    try:
        o.scores = []
        o.compute()
    except Exception:
        pass
t_ref_score = time.time() - t0
res = {"candidates": n, "cores": 1,
       "reference_under_shim_files_per_s": n / t_ref_files,
       "reference_under_shim_bates_only_cand_per_s": n / t_ref_score,
       "oracle_port_all30_cand_per_s": n / t_port,
       "note": "same 64 HTRU-shaped candidates of the bench stream; the reference times include gunzip + minidom parsing "
               "(three parses per file for Bates + Lyon, Candidate.py:141-286); the port scores arrays in memory"}
with open(os.path.join(ROOT, "profiles", "r02_ref_vs_port_speed.json"), "w") as fh:
    json.dump(res, fh, indent=1)
print(json.dumps(res, indent=1))
