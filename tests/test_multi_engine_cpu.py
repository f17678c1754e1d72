"""CPU tests of the multi-GPU host logic (no CUDA): how a batch is dealt to the devices and that the
per-device results land at their own rows of one result matrix, in input order."""
import threading

import numpy as np

from pulsarfeatureextractor_b200 import engine as E


def test_shard_ranges_tile_the_batch_exactly_once():
    for n in (0, 1, 5, 7, 8, 1000, 65536, 200000, 65536 * 16 + 123, 10_000_000):
        for g in (1, 2, 3, 4, 8):
            plan = E.shard_ranges(n, g)
            assert len(plan) == g
            seen = np.zeros(n, dtype=np.int8) if n <= 2_000_000 else None
            total = 0
            for ranges in plan:
                assert ranges == sorted(ranges)
                for lo, hi in ranges:
                    assert 0 <= lo < hi <= n
                    total += hi - lo
                    if seen is not None:
                        seen[lo:hi] += 1
            assert total == n
            if seen is not None:
                assert np.all(seen == 1)
            if n >= g:
                sizes = [sum(hi - lo for lo, hi in r) for r in plan]
                assert max(sizes) - min(sizes) <= 65536        # balanced to one chunk


def test_streamed_batches_are_dealt_round_robin():
    plan = E.shard_ranges(65536 * 9, 2)
    assert plan[0][0] == (0, 65536) and plan[1][0] == (65536, 131072) and plan[0][1] == (131072, 196608)
    assert len(plan[0]) == 5 and len(plan[1]) == 4


class _FakeEngine(object):
    """Stands in for Engine: 'scores' are a function of the input rows, and the device id is recorded
    so the test can see every device worked."""

    def __init__(self, dev):
        self.device = dev
        self.calls = []

    def score_host(self, profile, subbands=None, dmcurve=None, dmstat=None, params=None, features="all",
                   family=0, out=None, chunk=0, want_ier=True):
        self.calls.append((threading.get_ident(), profile.shape[0]))
        out.features[:, 0] = profile.sum(axis=1)
        out.features[:, 1] = params[:, 0] * 2
        out.status[:] = self.device + 1
        return out


def _multi(n_dev):
    m = E.MultiEngine.__new__(E.MultiEngine)
    import concurrent.futures
    m.devices = list(range(n_dev))
    m.engines = [_FakeEngine(d) for d in range(n_dev)]
    m._pool = concurrent.futures.ThreadPoolExecutor(max_workers=n_dev)
    m.device = 0
    return m


def test_results_land_at_their_rows_in_input_order(monkeypatch):
    rng = np.random.default_rng(0)
    for n in (10, 1001):
        prof = rng.integers(0, 255, (n, 16)).astype(np.uint8)
        par = rng.random((n, 8))
        for g in (1, 2, 4):
            m = _multi(g)
            out = E.Scores(np.zeros((n, 30)), np.zeros(n, np.int32), np.zeros((n, 18), np.int8))
            m.score_host(prof, prof[:, None, :], prof, prof, par, features="all", out=out)
            assert np.array_equal(out.features[:, 0], prof.sum(axis=1))
            assert np.array_equal(out.features[:, 1], par[:, 0] * 2)
            assert set(out.status.tolist()) == set(range(1, g + 1))         # every device did its share
            assert np.all(np.diff(out.status) >= 0)                          # contiguous ranges, ascending devices
            assert all(len(e.calls) == 1 for e in m.engines)                   # one call per device here
            m._pool.shutdown()
