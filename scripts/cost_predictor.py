"""Which cheap statistics predict the Levenberg-Marquardt work of a candidate?  (development aid:
oracle nfev per fit slot against profile statistics, for ordering the lane queues longest first)"""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))

def work(args):
    lo, hi = args
    import pfx_oracle as O
    from pulsarfeatureextractor_b200 import synth
    warnings.simplefilter("ignore")
    b = synth.generate(2, lo, hi - lo, 128, 16, 256, 256)
    out = []
    for i in range(hi - lo):
        log = O.FitLog()
        p = b.profile[i].astype(np.int64)
        try:
            g, st = O.gaussian_fittings(p.copy(), log)
        except Exception:
            continue
        x = p.astype(float)
        sd = x.std()
        z = (x - x.mean()) / sd if sd > 0 else x * 0
        out.append([log.nfev[5], log.nfev[6:14].sum(), log.nfev[14], log.nfev[2] + log.nfev[3] + log.nfev[4],
                    x.mean(), sd, (z ** 3).mean(), (z ** 4).mean(), x.max() - np.median(x), g[2], g[3], g[4], st])
    return out

if __name__ == "__main__":
    import multiprocessing as mp
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3200
    parts = [(i * n // 32, (i + 1) * n // 32) for i in range(32)]
    with mp.Pool(16) as pool:
        rows = sum(pool.map(work, parts), [])
    a = np.array(rows, dtype=float)
    np.save("/tmp/cost.npy", a)
    names = ["t1", "rounds", "final", "hist", "mean", "sd", "skew", "kurt", "peak-med", "s7", "s8", "s9", "st"]
    tot = a[:, 0] + a[:, 1]
    print("n", len(a), "rounds nfev mean", tot.mean(), "p99", np.percentile(tot, 99), "max", tot.max())
    print("final nfev mean", a[:, 2].mean(), "p99", np.percentile(a[:, 2], 99), "max", a[:, 2].max())
    from scipy.stats import spearmanr
    for j in range(4, 12):
        print(names[j], "spearman vs rounds %.3f  vs final %.3f" % (spearmanr(a[:, j], tot)[0], spearmanr(a[:, j], a[:, 2])[0]))
    print("rounds vs final %.3f" % spearmanr(tot, a[:, 2])[0])
