"""
PHCX candidate files: writer for synthetic candidates and the host-side packer (reader).

The reference reads a candidate with ``gzip`` + ``xml.dom.minidom`` and per-character Python
hex loops (PHCXFile.py:90-186, PHCXOperations.py:172-182,237-297,353-383) and, for the Lyon
statistics, parses every file twice (Candidate.py:218,277).  Here a file is scanned once.  Two
readers with the same behaviour: ``Scan`` drives the native multi-threaded scanner (libpfxio.so,
include/pfx_io.h: zlib inflate -> tag scan -> hex decode, rows gathered into pinned
structure-of-arrays buffers) and is what DataProcessor uses for bulk work; ``read_phcx`` is the
single-file reader in Python (compiled regular expressions and ``bytes.fromhex``) behind the
Candidate facade.  tests/test_abi_and_io.py holds the two against each other.  Quirks kept:

* section index 1 for HTRU ``.phcx.gz`` and 0 for SUPERB ``.phcx``
  (PHCXFile.py:80, SUPERBPHCXFile.py:80); the Lyon DM statistics always use DataBlock[0]
  (PHCXOperations.py:538);
* ``DmIndex``: tokens are the text between new-lines; dm_start is token 1, dm_end the last
  *complete* token (PHCXOperations.py:172-182);
* the DM curve is the max of the first 127 values of every 128-block of DataBlock[section]
  and the x index of block k is 128k-1 (PHCXOperations.py:237-259);
* hex decoding skips new-lines and stops at the first pair that is not hexadecimal
  (PHCXFile.py:166-184).
"""
import ctypes
import gzip
import os
import re

import numpy as np

# ---- native scanner (include/pfx_io.h, csrc/pfx_io.cpp) ------------------------------------
_IO_PATH = os.environ.get("PFX_IO_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpfxio.so")
IO_EXPORTS = ["pfx_phcx_scan", "pfx_phcx_gather", "pfx_phcx_free", "pfx_pfd_scan", "pfx_pfd_gather", "pfx_pfd_free",
              "pfx_format_rows", "pfx_io_free", "pfx_io_version"]
_io = None


class PhcxInfo(ctypes.Structure):
    _fields_ = [("ok", ctypes.c_int32), ("nbin", ctypes.c_int32), ("nsub", ctypes.c_int32),
                ("n_dm", ctypes.c_int32), ("n_dmstat", ctypes.c_int32), ("lyon_ok", ctypes.c_int32),
                ("params", ctypes.c_double * 8)]


def native_available():
    return os.path.isfile(_IO_PATH)


def _load_io():
    global _io
    if _io is None:
        lib = ctypes.CDLL(_IO_PATH)
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        lib.pfx_phcx_scan.argtypes = [ctypes.POINTER(ctypes.c_char_p), i64, ctypes.c_int, ctypes.c_int,
                                      ctypes.POINTER(PhcxInfo), ctypes.POINTER(vp)]
        lib.pfx_phcx_gather.argtypes = [vp, ctypes.POINTER(i64), i64, vp, vp, vp, vp, vp]
        lib.pfx_phcx_free.argtypes = [vp]
        lib.pfx_io_version.restype = ctypes.c_char_p
        lib.pfx_format_rows.argtypes = [vp, i64, i64, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), vp,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp),
                                        ctypes.POINTER(i64)]
        lib.pfx_io_free.argtypes = [vp]
        _io = lib
    return _io


def format_rows(features, names=None, skip=None, mode="csv", py2=True, threads=0):
    """Output text of a block of feature rows (native formatter, include/pfx_io.h): one line per
    row with skip[i] == 0, '<name>,s1,...' (csv), 's1,...,?%<name>' (arff) or 's1,...' (bare),
    Python-2 number formatting by default, 'nan' / 'inf' -> '0' over the whole line as the
    reference does (DataProcessor.py:321-325,421-425).  Returns ``str``."""
    lib = _load_io()
    feat = np.ascontiguousarray(features, dtype=np.float64)
    n, ncol = feat.shape
    m = {"csv": 0, "arff": 1, "bare": 2}[mode]
    arr = None
    if names is not None:
        arr = (ctypes.c_char_p * n)(*[os.fsencode(p) for p in names])
    sk = None
    if skip is not None:
        sk = np.ascontiguousarray(skip, dtype=np.int8)
    out, out_len = ctypes.c_void_p(), ctypes.c_int64()
    rc = lib.pfx_format_rows(feat.ctypes.data_as(ctypes.c_void_p), n, ncol, ncol, arr,
                             sk.ctypes.data_as(ctypes.c_void_p) if sk is not None else None, m,
                             1 if py2 else 0, int(threads), ctypes.byref(out), ctypes.byref(out_len))
    if rc != 0:
        raise RuntimeError("pfx_format_rows failed")
    try:
        return ctypes.string_at(out, out_len.value).decode("utf-8", "surrogateescape")
    finally:
        lib.pfx_io_free(out)


class Scan(object):
    """Result of scanning a list of PHCX files with the native reader: per-file ``ok`` and shape,
    and ``gather(indices)`` -> (profile, subbands, dmcurve, dmstat, params) rows of files that
    share a shape, written straight into pinned host arrays when torch is importable."""

    def __init__(self, paths, section=-1, threads=0):
        lib = _load_io()
        n = len(paths)
        arr = (ctypes.c_char_p * n)(*[os.fsencode(p) for p in paths])
        self.info = (PhcxInfo * n)()
        self._h = ctypes.c_void_p()
        if lib.pfx_phcx_scan(arr, n, int(section), int(threads), self.info, ctypes.byref(self._h)) != 0:
            raise RuntimeError("pfx_phcx_scan failed")
        self.n = n

    def ok(self, i, lyon_only=False):
        """Whole file readable; with lyon_only, what the Lyon statistics need is readable (profile of the
        section and DataBlock[0]: the reference touches nothing else in --dmprof mode, PHCXFile.py:320-379)."""
        return bool(self.info[i].ok) or (lyon_only and bool(self.info[i].lyon_ok))

    def shape_key(self, i, lyon_only=False):
        f = self.info[i]
        return (f.nbin, f.n_dmstat) if lyon_only else (f.nbin, f.nsub, f.n_dm, f.n_dmstat)

    def gather_lyon(self, indices):
        """(profile, dmstat) rows of files that share (nbin, n_dmstat); accepts files that are only lyon_ok."""
        lib = _load_io()
        k = len(indices)
        nbin, n_dms = self.shape_key(indices[0], True)
        prof = self._buffer((k, nbin), np.uint8)
        dms = self._buffer((k, n_dms), np.uint8)
        idx = (ctypes.c_int64 * k)(*indices)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        rc = lib.pfx_phcx_gather(self._h, idx, k, p(prof), None, None, p(dms), None)
        if rc != 0:
            raise RuntimeError("pfx_phcx_gather failed (%d)" % rc)
        return prof, dms

    @staticmethod
    def _buffer(shape, dtype):
        try:
            import torch
            t = torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name),
                            pin_memory=torch.cuda.is_available())
            return t.numpy()
        except ImportError:
            return np.empty(shape, dtype=dtype)

    def gather(self, indices):
        lib = _load_io()
        k = len(indices)
        nbin, nsub, n_dm, n_dms = self.shape_key(indices[0])
        prof = self._buffer((k, nbin), np.uint8)
        sub = self._buffer((k, nsub, nbin), np.uint8)
        dmc = self._buffer((k, n_dm), np.uint8)
        dms = self._buffer((k, n_dms), np.uint8)
        par = self._buffer((k, 8), np.float64)
        idx = (ctypes.c_int64 * k)(*indices)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        rc = lib.pfx_phcx_gather(self._h, idx, k, p(prof), p(sub), p(dmc), p(dms), p(par))
        if rc != 0:
            raise RuntimeError("pfx_phcx_gather failed (%d)" % rc)
        return prof, sub, dmc, dms, par

    def close(self):
        if self._h:
            _load_io().pfx_phcx_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:       # noqa: BLE001 - interpreter shutdown
            pass

_HEX_RE = re.compile(rb"^(?:[0-9A-Fa-f]{2})*")


def _tag_blocks(text, tag):
    """All (attribute string, body) of <tag ...>body</tag>, in document order."""
    pat = re.compile(rb"<" + tag + rb"(\s[^>]*)?>(.*?)</" + tag + rb">", re.S)
    return [(m.group(1) or b"", m.group(2)) for m in pat.finditer(text)]


def _attr(attrs, name):
    m = re.search(name + rb"\s*=\s*['\"]([^'\"]*)['\"]", attrs)
    if m is None:
        raise KeyError(name.decode())
    return m.group(1)


def _hex_to_u8(body):
    body = body.replace(b"\n", b"")
    m = _HEX_RE.match(body)
    return np.frombuffer(bytes.fromhex(m.group(0).decode("ascii")), dtype=np.uint8)


def dm_curve_reduce(block):
    """max of the first 127 values of each complete 128-block (PHCXOperations.py:237-259)."""
    nblk = block.size // 128
    if nblk == 0:
        return np.zeros(0, dtype=np.uint8)
    return block[: nblk * 128].reshape(nblk, 128)[:, :127].max(axis=1)


class PHCXCandidate(object):
    __slots__ = ("profile", "subbands", "dmcurve", "dmstat", "params")


def read_phcx(path, section=None, lyon_only=False):
    """Parse one candidate.  ``section`` defaults to 1 for ``.gz`` names, else 0
    (Candidate.py:141-150 dispatch).  With lyon_only only the profile and DataBlock[0] are read (what
    the reference touches in --dmprof mode, PHCXFile.py:320-379): the other arrays stay None."""
    if section is None:
        section = 1 if ".gz" in path else 0
    if ".gz" in path:
        with gzip.open(path, "rb") as fh:
            text = fh.read()
    else:
        with open(path, "rb") as fh:
            text = fh.read()
    c = PHCXCandidate()
    c.profile = _hex_to_u8(_tag_blocks(text, b"Profile")[section][1])
    if lyon_only:
        c.dmstat = _hex_to_u8(_tag_blocks(text, b"DataBlock")[0][1])
        c.subbands = c.dmcurve = c.params = None
        return c
    attrs, body = _tag_blocks(text, b"SubBands")[section]
    nbins = int(_attr(attrs, b"nBins"))
    nsub = int(_attr(attrs, b"nSub"))
    c.subbands = _hex_to_u8(body).reshape(nsub, nbins)
    blocks = _tag_blocks(text, b"DataBlock")
    block = _hex_to_u8(blocks[section][1])
    c.dmcurve = dm_curve_reduce(block)
    c.dmstat = _hex_to_u8(blocks[0][1])
    tokens = _tag_blocks(text, b"DmIndex")[section][1].split(b"\n")[:-1]
    dm_start = float(tokens[1])
    dm_end = float(tokens[len(tokens) - 1])

    def scalar(tag):
        return float(_tag_blocks(text, tag)[section][1])

    c.params = np.array([scalar(b"Snr"), scalar(b"Dm"), scalar(b"BaryPeriod"), scalar(b"Width"),
                         dm_start, dm_end, float(block.size), 0.0], dtype=np.float64)
    return c


def _hex_lines(a, per_line=40):
    h = bytes(np.ascontiguousarray(a, dtype=np.uint8)).hex().upper()
    step = per_line * 2
    return "\n".join(h[i:i + step] for i in range(0, len(h), step))


def _section(name, snr, dm, period_s, width, profile, subbands, dm_index, datablock):
    nsub, nbin = subbands.shape
    idx = "\n".join("%.4f" % v for v in dm_index)
    return (
        " <Section name='%s'>\n"
        "  <BestValues>\n"
        "   <TopoPeriod units='seconds'>%.12f</TopoPeriod>\n"
        "   <BaryPeriod units='seconds'>%r</BaryPeriod>\n"
        "   <Dm units='cm^-3pc'>%r</Dm>\n"
        "   <Accn units='m/s/s'>0.0</Accn>\n"
        "   <Jerk units='m/s/s/s'>0.0</Jerk>\n"
        "   <Snr>%r</Snr>\n"
        "   <Width units='periods'>%r</Width>\n"
        "  </BestValues>\n"
        "  <SubIntegrations nBins='%d' nSub='1' format='02X' min='0' max='1'>\n%s\n\t</SubIntegrations>\n"
        "  <SubBands nBins='%d' nSub='%d' format='02X' min='0' max='1'>\n%s\n\t</SubBands>\n"
        "  <Profile nBins='%d' format='02X' min='0' max='1'>\n%s\n\t</Profile>\n"
        "  <SnrBlock>\n"
        "   <PeriodIndex nVals='1' format='16.12f'>\n%.12f\n</PeriodIndex>\n"
        "   <DmIndex nVals='%d' format='6.4f'>\n%s\n</DmIndex>\n"
        "   <AccnIndex nVals='1' format='6.4f'>\n0.0000\n</AccnIndex>\n"
        "   <JerkIndex nVals='1' format='6.4f'>\n0.0000\n</JerkIndex>\n"
        "   <DataBlock format='02X' min='0' max='1'>\n%s\n\t</DataBlock>\n"
        "  </SnrBlock>\n"
        " </Section>\n"
    ) % (name, period_s, period_s, float(dm), float(snr), float(width),
         nbin, _hex_lines(profile),
         nbin, nsub, _hex_lines(subbands.reshape(-1)),
         nbin, _hex_lines(profile),
         period_s, len(dm_index), idx, _hex_lines(datablock))


def expand_dm_block(curve):
    """A DataBlock whose max-of-127 reduction gives ``curve`` back: every 128-block holds the
    curve value once and smaller values elsewhere; the 128th value (ignored by the reference,
    PHCXOperations.py:251-257) is set to 255 to exercise that quirk."""
    curve = np.asarray(curve, dtype=np.uint8)
    blk = (curve[:, None].astype(np.int32) * (np.arange(128)[None, :] % 7) // 7).astype(np.uint8)
    blk[np.arange(curve.size), (np.arange(curve.size) * 5) % 127] = curve
    blk[:, 127] = 255
    return blk.reshape(-1)


def write_phcx(path, profile, subbands, dmcurve, dmstat, params, gz=True, superb=False):
    """Write a two-section PHCX file.  Section 0 ('FFT') carries ``dmstat`` as its DataBlock
    (Lyon DM statistics read it, PHCXOperations.py:538) and a decoy profile; section 1
    ('FFT-Pdmpd') carries the candidate the Bates scores read.  With ``gz=False`` a
    single-section file is written (section index 0 holds everything).  With ``superb=True`` an
    uncompressed two-section SUPERB file is written (SUPERBPHCXFile.py:80,102-103): section 0 holds
    the candidate every score reads -- its DataBlock is both the source of the DM curve (max-of-127
    reduction) and the array of the Lyon DM statistics, so ``dmstat`` is ignored -- and section 1 is
    a decoy that only has to pass the validity test (SUPERBPHCXFile.py:212-252)."""
    snr, dm, period_s, width, dm_start, dm_end = [float(v) for v in params[:6]]
    block1 = expand_dm_block(dmcurve)
    n_idx = 64
    dm_index = np.linspace(dm_start, dm_end, n_idx)
    parts = ["<?xml version='1.0'?>\n<phcf>\n <head>\n  <SourceID>SYNTHETIC</SourceID>\n </head>\n"]
    if superb:
        gz = False
        decoy = (255 - np.asarray(profile, dtype=np.uint8))
        parts.append(_section("FFT", snr, dm, period_s, width, profile, np.asarray(subbands), dm_index, block1))
        parts.append(_section("FFT-Pdmpd", snr * 0.5, dm + 1.0, period_s * 1.0001, width * 2, decoy,
                              np.asarray(subbands)[::-1], dm_index * 0.5, block1[::-1]))
    elif gz:
        decoy = (255 - np.asarray(profile, dtype=np.uint8))
        parts.append(_section("FFT", snr * 0.5, dm + 1.0, period_s * 1.0001, width * 2, decoy,
                              np.asarray(subbands)[::-1], dm_index * 0.5, dmstat))
        parts.append(_section("FFT-Pdmpd", snr, dm, period_s, width, profile, np.asarray(subbands),
                              dm_index, block1))
    else:
        parts.append(_section("FFT-Pdmpd", snr, dm, period_s, width, profile, np.asarray(subbands),
                              dm_index, block1))
    parts.append("</phcf>\n")
    data = "".join(parts).encode("ascii")
    if gz:
        with gzip.open(path, "wb", compresslevel=1) as fh:
            fh.write(data)
    else:
        with open(path, "wb") as fh:
            fh.write(data)
