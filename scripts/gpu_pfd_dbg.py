"""Development helper: per-column relative error of the .pfd end-to-end path against the golden reference scores."""
import os, sys, tempfile, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from pulsarfeatureextractor_b200 import pfd_io
from pulsarfeatureextractor_b200.engine import Engine
warnings.simplefilter("ignore")
np.set_printoptions(linewidth=250, precision=2)
g = np.load(os.path.join(ROOT, "tests", "golden", "ref_pfd_v1.npz"))
eng = Engine(0)
d = tempfile.mkdtemp()
for k, (seed, npart, nsub, proflen, posn, big) in enumerate(g["cases"]):
    if ("fail_%d" % k) in g.files:
        continue
    p = os.path.join(d, "c%d.pfd" % k)
    pfd_io.write_pfd(p, with_position=bool(posn), big_endian=bool(big), **pfd_io.synth_pfd(int(seed), npart=int(npart), nsub=int(nsub), proflen=int(proflen)))
    r = pfd_io.read_pfd(p)
    out = eng.score_pfd_host(r.profile[None], r.subbands[None], r.dmcurve[None], r.params[None], features="bates")
    ref = g["bates_%d" % k]
    f = out.features[0]
    e = np.abs(ref - f) / np.maximum(np.abs(ref), 1e-12)
    print(k, "ier", out.ier[0][:16], "\n   relerr", e)
