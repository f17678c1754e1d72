"""Fills tests/golden/_cache/ (git-ignored, travels with the gpurun snapshot) with the oracle + twin
rows scripts/gpu_parity_table.py needs, and writes the committed fixture
tests/golden/oracle_bench5k_v1.npz: oracle features / status / ier for a stratified sample of 5 000
candidates of the bench stream (synth.generate(2, 0, 100000)[::20]), used by
tests/test_gpu_parity.py::test_bench_stream_sample_against_oracle.
    python tests/golden/make_oracle_cache.py [n_htru] [n_pfd] [n_1024]"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_pool as OP
from pulsarfeatureextractor_b200 import synth

n_htru = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
n_pfd = int(sys.argv[2]) if len(sys.argv) > 2 else 512
n_1k = int(sys.argv[3]) if len(sys.argv) > 3 else 256
full = synth.generate(2, 0, 100000)
idx = np.arange(0, 100000, 100000 // n_htru)[:n_htru]
sub = synth.Batch(full.profile[idx], full.subbands[idx], full.dmcurve[idx], full.dmstat[idx], full.params[idx],
                  full.klass[idx], full.edge[idx])
rows = OP.cached_run("parity_htru_%d" % len(sub), sub, family="phcx", twins=True)
if n_htru == 5000:
    np.savez_compressed(os.path.join(HERE, "oracle_bench5k_v1.npz"), index=idx, feat=rows.feat, status=rows.status,
                        ier=rows.ier, lyon=rows.lyon)
print("htru", len(sub), flush=True)
OP.cached_run("parity_pfd_%d" % n_pfd, synth.generate_pfd_shaped(4, 0, n_pfd), family="pfd", twins=True, fast_corr=True)
print("pfd", n_pfd, flush=True)
OP.cached_run("parity_1024_%d" % n_1k, synth.generate(5, 0, n_1k, nbin=1024), family="phcx", twins=True, fast_corr=True)
print("1024", n_1k, flush=True)
