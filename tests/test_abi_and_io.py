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


# ---- native PHCX scanner (include/pfx_io.h) against the Python reader ---------------------------
def _ensure_io():
    if not phcx_io.native_available():
        import __graft_entry__ as ge
        ge.build()


def test_io_header_symbols_exported():
    _ensure_io()
    hdr = open(os.path.join(ROOT, "include", "pfx_io.h")).read()
    declared = set(re.findall(r"\b(pfx_[a-z0-9_]+)\s*\(", hdr))
    lib = ctypes.CDLL(phcx_io._IO_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(phcx_io.IO_EXPORTS)


def test_native_scanner_equals_python_reader(tmp_path):
    _ensure_io()
    b = synth.generate(1, 0, 48)
    paths = []
    for i in range(48):
        gz = i % 4 != 0
        p = str(tmp_path / ("c%03d.phcx%s" % (i, ".gz" if gz else "")))
        phcx_io.write_phcx(p, b.profile[i], b.subbands[i], b.dmcurve[i], b.dmstat[i], b.params[i], gz=gz)
        paths.append(p)
    # malformed files: what makes the Python reader (and the reference's) raise must give ok = 0
    import gzip
    with gzip.open(paths[1], "rb") as fh:
        good = fh.read()

    def variant(name, data, gz=True):
        p = str(tmp_path / name)
        with (gzip.open(p, "wb") if gz else open(p, "wb")) as fh:
            fh.write(data)
        paths.append(p)

    variant("not_gzip.phcx.gz", b"plain text, no gzip header", gz=False)
    variant("truncated.phcx.gz", good[: len(good) // 2])
    variant("no_profile.phcx.gz", good.replace(b"<Profile", b"<Profilx").replace(b"</Profile>", b"</Profilx>"))
    variant("one_section.phcx.gz", good[: good.index(b"</Section>") + 10] + b"</phcf>")
    variant("bad_number.phcx.gz", good.replace(b"<Snr>", b"<Snr>abc"))
    variant("subband_size.phcx.gz", good.replace(b"nSub='16'", b"nSub='15'"))
    variant("short_dmindex.phcx.gz", re.sub(rb"(<DmIndex[^>]*>)\n.*?(</DmIndex>)", rb"\g<1>1.0\n\g<2>", good, flags=re.S))
    variant("lower_hex_and_junk.phcx.gz", good.replace(b"<Profile nBins='128' format='02X' min='0' max='1'>\n",
                                                      b"<Profile nBins='128' format='02X' min='0' max='1'>\n").lower()
            .replace(b"<profile", b"<Profile").replace(b"</profile>", b"</Profile>").replace(b"<subbands", b"<SubBands")
            .replace(b"</subbands>", b"</SubBands>").replace(b"<datablock", b"<DataBlock").replace(b"</datablock>", b"</DataBlock>")
            .replace(b"<dmindex", b"<DmIndex").replace(b"</dmindex>", b"</DmIndex>").replace(b"<snr>", b"<Snr>")
            .replace(b"</snr>", b"</Snr>").replace(b"<dm ", b"<Dm ").replace(b"</dm>", b"</Dm>")
            .replace(b"<baryperiod", b"<BaryPeriod").replace(b"</baryperiod>", b"</BaryPeriod>")
            .replace(b"<width", b"<Width").replace(b"</width>", b"</Width>").replace(b"nbins=", b"nBins=").replace(b"nsub=", b"nSub="))
    scan = phcx_io.Scan(paths, threads=3)
    nbad = 0
    for i, p in enumerate(paths):
        try:
            r = phcx_io.read_phcx(p)
            if r.subbands.shape[1] != r.profile.size:
                r = None
        except Exception:       # noqa: BLE001 - any exception = the candidate is skipped
            r = None
        assert scan.ok(i) == (r is not None), p
        if r is None:
            nbad += 1
            continue
        prof, sub, dmc, dms, par = scan.gather([i])
        assert np.array_equal(prof[0], r.profile) and np.array_equal(sub[0], r.subbands), p
        assert np.array_equal(dmc[0], r.dmcurve) and np.array_equal(dms[0], r.dmstat), p
        assert np.array_equal(par[0], r.params), p
    assert nbad == 7
    # one gather for all files of the most common shape, rows in the order asked for
    idx = [i for i in range(48) if scan.ok(i) and i % 4 != 0][::-1]       # the two-section .gz files share a shape
    prof, sub, dmc, dms, par = scan.gather(idx)
    assert np.array_equal(prof, b.profile[idx]) and np.array_equal(sub, b.subbands[idx])
    assert np.array_equal(dmc, b.dmcurve[idx]) and np.array_equal(dms, b.dmstat[idx])
    assert np.array_equal(par[:, :4], b.params[idx, :4])     # dm_start / dm_end pass through '%.4f' in the file
    scan.close()


def test_native_row_formatter_equals_reference_format(tmp_path):
    """pfx_format_rows against the oracle's row writer (DataProcessor.py:321-325,421-425): Python-2
    '%.12g' numbers and Python-3 repr, CSV / ARFF / bare, 'nan' and 'inf' -> '0' over the whole line
    (the candidate name included), skipped rows."""
    _ensure_io()
    import pfx_oracle as O
    rng = np.random.default_rng(0)
    special = [0.0, -0.0, 1.0, -1.0, 5.0, 1e15, 1e16, 1e17, 123456789012.0, 1234567890123.0, 0.1, 1 / 3, 1e-4, 1e-5,
               5e-324, 1.7976931348623157e308, float("nan"), float("inf"), float("-inf"), 2.5e-7, 100.0,
               99999999999.5, 999999999999.5, 0.000123456789012345, 1e22, 1e21, 123456.789e3, 1000000.0]
    vals = np.concatenate([special, rng.normal(size=3000) * 10.0 ** rng.integers(-30, 30, 3000),
                           rng.integers(-1000, 1000, 500).astype(float), rng.random(472)])
    feat = vals.reshape(-1, 10)
    n = feat.shape[0]
    names = ["/data/nan_dir/cand_%d_inf.phcx.gz" % i for i in range(n)]
    skip = np.zeros(n, np.int8)
    skip[[3, n - 1]] = 1
    for py2 in (True, False):
        for mode in ("csv", "arff", "bare"):
            got = phcx_io.format_rows(feat, names, skip, mode, py2, threads=3)
            exp = []
            for i in range(n):
                if skip[i]:
                    continue
                if mode == "bare":
                    exp.append(",".join(O.fmt_float(v, py2) for v in feat[i]).replace("nan", "0").replace("inf", "0"))
                else:
                    exp.append(O.format_row(names[i], feat[i], arff=(mode == "arff"), py2=py2))
            assert got == "\n".join(exp) + "\n", (py2, mode)
