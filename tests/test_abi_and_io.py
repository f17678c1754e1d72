"""CPU tests: the C-ABI library loads and exports what include/pfx.h declares (no compute calls),
the PHCX packer reproduces the reader quirks, the synthetic stream is shard-invariant."""
import ctypes
import os
import re
import tempfile

import numpy as np
import pytest

from pulsarfeatureextractor_b200 import _lib, phcx_io, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_lib():
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()


def test_header_symbols_exported():
    _ensure_lib()
    hdr = open(os.path.join(ROOT, "include", "pfx.h")).read()
    declared = set(re.findall(r"\b(pfx_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"pfx_ctx"}
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.EXPORTS)
    assert lib.pfx_version() == 100


def test_status_and_slot_constants_match_header_and_oracle():
    import pfx_oracle as O
    hdr = open(os.path.join(ROOT, "include", "pfx.h")).read()
    for name, val in re.findall(r"#define PFX_ST_([A-Z_]+)\s+\(1u << (\d+)\)", hdr):
        assert getattr(O, "ST_" + name) == 1 << int(val)
        assert getattr(_lib, "ST_" + name) == 1 << int(val)
    for name, val in re.findall(r"#define PFX_FIT_([A-Z0-9_]+)\s+(\d+)", hdr):
        assert getattr(O, "FIT_" + name) == int(val)
    assert O.ST_FAIL_MASK == _lib.ST_FAIL_MASK


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libpfx.so")
    with pytest.raises(_lib.LibraryMissing):
        _lib.load()


def test_engine_refuses_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pulsarfeatureextractor_b200.engine import Engine, PfxError
    with pytest.raises(PfxError):
        Engine(0)


def test_phcx_roundtrip_and_quirks():
    b = synth.generate(5, 0, 8)
    d = tempfile.mkdtemp()
    for i in range(4):
        p = os.path.join(d, "c%d.phcx.gz" % i)
        phcx_io.write_phcx(p, b.profile[i], b.subbands[i], b.dmcurve[i], b.dmstat[i], b.params[i])
        c = phcx_io.read_phcx(p)
        assert np.array_equal(c.profile, b.profile[i])          # section 1, not the decoy of section 0
        assert np.array_equal(c.subbands, b.subbands[i])
        assert np.array_equal(c.dmcurve, b.dmcurve[i])          # max of the first 127 of each 128-block
        assert np.array_equal(c.dmstat, b.dmstat[i])            # DataBlock[0]
        assert c.params[6] == 128 * 256
        assert c.params[0] == b.params[i][0] and c.params[2] == b.params[i][2]
    # SUPERB-style single section, un-gzipped: section index 0
    p = os.path.join(d, "s.phcx")
    phcx_io.write_phcx(p, b.profile[0], b.subbands[0], b.dmcurve[0], b.dmstat[0], b.params[0], gz=False)
    c = phcx_io.read_phcx(p)
    assert np.array_equal(c.profile, b.profile[0]) and c.dmstat.size == 128 * 256


def test_dm_curve_reduce_ignores_128th_value():
    blk = np.zeros(256, dtype=np.uint8)
    blk[127] = 200      # ignored (PHCXOperations.py:251-257)
    blk[5] = 7
    blk[128 + 126] = 9
    assert list(phcx_io.dm_curve_reduce(blk)) == [7, 9]


def test_synthetic_stream_is_shard_invariant():
    a = synth.generate(9, 0, 6000)
    lo = synth.generate(9, 0, 2500)
    hi = synth.generate(9, 2500, 3500)
    for name in ("profile", "subbands", "dmcurve", "dmstat", "params"):
        assert np.array_equal(getattr(a, name), np.concatenate([getattr(lo, name), getattr(hi, name)]))
    assert set(np.unique(a.edge)) >= set(range(synth.N_EDGE_KINDS))
