// pfx_io.cpp -- libpfxio.so: PHCX scanner, PRESTO .pfd reader and output-row formatter (include/pfx_io.h).
//
// One pass per file: inflate (zlib) -> tag scan -> hex decode -> arrays.  The scan follows the
// document-order, n-th-occurrence rule the reference gets from minidom's getElementsByTagName
// (PHCXFile.py:160-163, PHCXOperations.py:107-110,172,280-281,334-339), i.e. "section k" of a tag
// is simply its k-th occurrence in the file; pulsarfeatureextractor_b200/phcx_io.py is the same
// reader in Python and tests/test_abi_and_io.py holds the two against each other.
#include <zlib.h>

#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pfx_io.h"

namespace {

struct Cand {
    std::vector<uint8_t> profile, subbands, dmcurve, dmstat;
};

struct Span {
    const char* p;
    size_t n;
};

struct Block {
    Span attrs, body;
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

// k-th match of  <tag(\s[^>]*)?>(.*?)</tag>  in document order (non-overlapping, non-greedy)
bool tag_block(const std::string& text, const char* tag, int k, Block* out) {
    const size_t tl = strlen(tag);
    std::string open = "<";
    open += tag;
    std::string close = "</";
    close += tag;
    close += ">";
    size_t pos = 0;
    int seen = 0;
    for (;;) {
        const size_t a = text.find(open, pos);
        if (a == std::string::npos) return false;
        size_t q = a + 1 + tl;
        if (q >= text.size()) return false;
        size_t body_at;
        Span attrs = {text.data() + q, 0};
        if (text[q] == '>') {
            body_at = q + 1;
        } else if (is_space(text[q])) {
            const size_t gt = text.find('>', q);
            if (gt == std::string::npos) return false;
            attrs.n = gt - q;
            body_at = gt + 1;
        } else {            // a longer tag name (<DmIndex> while looking for <Dm>)
            pos = a + 1;
            continue;
        }
        const size_t e = text.find(close, body_at);
        if (e == std::string::npos) {      // no closing tag after this opening: the regex would retry later openings
            pos = a + 1;
            continue;
        }
        if (seen == k) {
            out->attrs = attrs;
            out->body = {text.data() + body_at, e - body_at};
            return true;
        }
        ++seen;
        pos = e + close.size();
    }
}

// name\s*=\s*['"]([^'"]*)['"]  -> integer (Python int(): optional sign, digits, surrounding blanks)
bool attr_int(Span attrs, const char* name, long* value) {
    const std::string s(attrs.p, attrs.n);
    const size_t nl = strlen(name);
    size_t pos = 0;
    for (;;) {
        const size_t a = s.find(name, pos);
        if (a == std::string::npos) return false;
        size_t q = a + nl;
        while (q < s.size() && is_space(s[q])) ++q;
        if (q < s.size() && s[q] == '=') {
            ++q;
            while (q < s.size() && is_space(s[q])) ++q;
            if (q < s.size() && (s[q] == '\'' || s[q] == '"')) {
                const size_t b = q + 1;
                size_t e = b;
                while (e < s.size() && s[e] != '\'' && s[e] != '"') ++e;
                if (e < s.size()) {
                    std::string v = s.substr(b, e - b);
                    size_t i = 0, j = v.size();
                    while (i < j && is_space(v[i])) ++i;
                    while (j > i && is_space(v[j - 1])) --j;
                    if (i == j) return false;
                    char* endp = nullptr;
                    errno = 0;
                    const std::string t = v.substr(i, j - i);
                    const long r = strtol(t.c_str(), &endp, 10);
                    if (errno != 0 || endp != t.c_str() + t.size()) return false;
                    *value = r;
                    return true;
                }
            }
        }
        pos = a + 1;
    }
}

inline int hexval(char c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}

// new-lines are dropped, then the longest prefix of complete hexadecimal pairs is decoded
// (PHCXFile.py:166-184: the reference's loop stops at the first pair that does not parse)
void hex_decode(Span body, std::vector<uint8_t>* out) {
    out->clear();
    out->reserve(body.n / 2);
    int hi = -1;
    for (size_t i = 0; i < body.n; ++i) {
        const char c = body.p[i];
        if (c == '\n') continue;
        const int v = hexval(c);
        if (v < 0) break;
        if (hi < 0) {
            hi = v;
        } else {
            out->push_back((uint8_t)(hi * 16 + v));
            hi = -1;
        }
    }
}

// Python float() on the text: blanks around one number, nothing else
bool parse_float(const char* p, size_t n, double* value) {
    size_t i = 0, j = n;
    while (i < j && is_space(p[i])) ++i;
    while (j > i && is_space(p[j - 1])) --j;
    if (i == j) return false;
    const std::string t(p + i, j - i);
    if (t.find_first_of("xXpP") != std::string::npos) return false;    // strtod's hex floats are not Python floats
    char* endp = nullptr;
    const double r = strtod(t.c_str(), &endp);
    if (endp != t.c_str() + t.size()) return false;
    *value = r;
    return true;
}

bool read_file(const char* path, bool gz, std::string* text) {
    text->clear();
    if (gz) {
        gzFile f = gzopen(path, "rb");
        if (!f) return false;
        gzbuffer(f, 1 << 18);
        char buf[1 << 16];
        for (;;) {
            const int got = gzread(f, buf, sizeof(buf));
            if (got < 0) {
                gzclose(f);
                return false;
            }
            if (got == 0) break;
            text->append(buf, (size_t)got);
        }
        return gzclose(f) == Z_OK;
    }
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) text->append(buf, got);
    const bool ok = !ferror(f);
    fclose(f);
    return ok;
}

bool parse_one(const char* path, int section, Cand* c, pfx_phcx_info* info, std::string* text) {
    const bool gz = strstr(path, ".gz") != nullptr;
    const int sec = section >= 0 ? section : (gz ? 1 : 0);
    if (!read_file(path, gz, text)) return false;
    Block b;
    if (!tag_block(*text, "Profile", sec, &b)) return false;
    hex_decode(b.body, &c->profile);
    // What the Lyon statistics need (PHCXFile.py:320-379, PHCXOperations.py:528-539): this profile and
    // DataBlock[0].  A file that has them is usable in --dmprof mode even when a tag of the Bates scores is
    // missing or malformed below (the reference never touches those tags in that mode).
    if (tag_block(*text, "DataBlock", 0, &b)) {
        hex_decode(b.body, &c->dmstat);
        info->lyon_ok = 1;
        info->nbin = (int32_t)c->profile.size();
        info->n_dmstat = (int32_t)c->dmstat.size();
    }
    if (!tag_block(*text, "SubBands", sec, &b)) return false;
    long nbins = 0, nsub = 0;
    if (!attr_int(b.attrs, "nBins", &nbins) || !attr_int(b.attrs, "nSub", &nsub)) return false;
    hex_decode(b.body, &c->subbands);
    if (nbins <= 0 || nsub <= 0 || (size_t)nbins * (size_t)nsub != c->subbands.size()) return false;
    if ((size_t)nbins != c->profile.size()) return false;      // one nbin per candidate in the batch layout
    std::vector<uint8_t> block;
    if (!tag_block(*text, "DataBlock", sec, &b)) return false;
    hex_decode(b.body, &block);
    if (!info->lyon_ok) return false;       // (no DataBlock[0])
    // max of the first 127 values of every complete 128-block (PHCXOperations.py:237-259)
    const size_t nblk = block.size() / 128;
    c->dmcurve.resize(nblk);
    for (size_t k = 0; k < nblk; ++k) {
        uint8_t m = 0;
        const uint8_t* row = block.data() + k * 128;
        for (int i = 0; i < 127; ++i) m = row[i] > m ? row[i] : m;
        c->dmcurve[k] = m;
    }
    // DmIndex: tokens between new-lines, the text after the last new-line is dropped; dm_start is
    // token 1, dm_end the last token (PHCXOperations.py:172-182)
    if (!tag_block(*text, "DmIndex", sec, &b)) return false;
    std::vector<Span> tok;
    {
        size_t s0 = 0;
        for (size_t i = 0; i < b.body.n; ++i)
            if (b.body.p[i] == '\n') {
                tok.push_back({b.body.p + s0, i - s0});
                s0 = i + 1;
            }
    }
    if (tok.size() < 2) return false;
    double dm_start, dm_end, snr, dm, period, width;
    if (!parse_float(tok[1].p, tok[1].n, &dm_start)) return false;
    if (!parse_float(tok.back().p, tok.back().n, &dm_end)) return false;
    if (!tag_block(*text, "Snr", sec, &b) || !parse_float(b.body.p, b.body.n, &snr)) return false;
    if (!tag_block(*text, "Dm", sec, &b) || !parse_float(b.body.p, b.body.n, &dm)) return false;
    if (!tag_block(*text, "BaryPeriod", sec, &b) || !parse_float(b.body.p, b.body.n, &period)) return false;
    if (!tag_block(*text, "Width", sec, &b) || !parse_float(b.body.p, b.body.n, &width)) return false;
    info->nbin = (int32_t)c->profile.size();
    info->nsub = (int32_t)nsub;
    info->n_dm = (int32_t)nblk;
    info->n_dmstat = (int32_t)c->dmstat.size();
    info->params[0] = snr;
    info->params[1] = dm;
    info->params[2] = period;
    info->params[3] = width;
    info->params[4] = dm_start;
    info->params[5] = dm_end;
    info->params[6] = (double)block.size();
    info->params[7] = 0.0;
    return true;
}

}  // namespace

struct pfx_scan {
    std::vector<Cand> cands;
    std::vector<pfx_phcx_info> info;
};

extern "C" int pfx_phcx_scan(const char* const* paths, int64_t n, int section, int threads,
                             pfx_phcx_info* info, pfx_scan** out) {
    if (!paths || n < 0 || !info || !out || section > 1) return -1;
    pfx_scan* h = new (std::nothrow) pfx_scan;
    if (!h) return -1;
    try {
        h->cands.resize((size_t)n);
        h->info.resize((size_t)n);
    } catch (...) {
        delete h;
        return -1;
    }
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if ((int64_t)nt > n) nt = (int)(n > 0 ? n : 1);
    std::atomic<int64_t> next(0);
    auto work = [&]() {
        std::string text;
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n) break;
            pfx_phcx_info fi;
            memset(&fi, 0, sizeof(fi));
            bool ok = false;
            try {
                ok = paths[i] != nullptr && parse_one(paths[i], section, &h->cands[(size_t)i], &fi, &text);
            } catch (...) {
                ok = false;
            }
            fi.ok = ok ? 1 : 0;
            if (!ok) {
                Cand& cd = h->cands[(size_t)i];
                if (fi.lyon_ok) {        // keep what the Lyon statistics need, drop the rest
                    cd.subbands.clear();
                    cd.dmcurve.clear();
                    const int32_t nb = fi.nbin, nd = fi.n_dmstat;
                    memset(&fi, 0, sizeof(fi));
                    fi.lyon_ok = 1;
                    fi.nbin = nb;
                    fi.n_dmstat = nd;
                } else {
                    cd = Cand();
                    memset(&fi, 0, sizeof(fi));
                }
            }
            h->info[(size_t)i] = fi;
            info[i] = fi;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    *out = h;
    return 0;
}

extern "C" int pfx_phcx_gather(const pfx_scan* scan, const int64_t* idx, int64_t count, uint8_t* profile,
                               uint8_t* subbands, uint8_t* dmcurve, uint8_t* dmstat, double* params) {
    if (!scan || !idx || count < 0) return -1;
    if (count == 0) return 0;
    const int64_t n = (int64_t)scan->cands.size();
    if (idx[0] < 0 || idx[0] >= n) return -1;
    const pfx_phcx_info& f0 = scan->info[(size_t)idx[0]];
    // a gather of profile and dmstat alone (Lyon statistics) also accepts files that are only lyon_ok
    const bool lyon_only = !subbands && !dmcurve && !params;
    for (int64_t r = 0; r < count; ++r) {
        if (idx[r] < 0 || idx[r] >= n) return -1;
        const pfx_phcx_info& f = scan->info[(size_t)idx[r]];
        if (lyon_only) {
            if (!(f.ok || f.lyon_ok) || f.nbin != f0.nbin || f.n_dmstat != f0.n_dmstat) return -2;
        } else if (!f.ok || f.nbin != f0.nbin || f.nsub != f0.nsub || f.n_dm != f0.n_dm || f.n_dmstat != f0.n_dmstat) {
            return -2;
        }
    }
    for (int64_t r = 0; r < count; ++r) {
        const Cand& c = scan->cands[(size_t)idx[r]];
        const pfx_phcx_info& f = scan->info[(size_t)idx[r]];
        if (profile) memcpy(profile + (size_t)r * f.nbin, c.profile.data(), (size_t)f.nbin);
        if (subbands) memcpy(subbands + (size_t)r * f.nsub * f.nbin, c.subbands.data(), (size_t)f.nsub * f.nbin);
        if (dmcurve && f.n_dm) memcpy(dmcurve + (size_t)r * f.n_dm, c.dmcurve.data(), (size_t)f.n_dm);
        if (dmstat && f.n_dmstat) memcpy(dmstat + (size_t)r * f.n_dmstat, c.dmstat.data(), (size_t)f.n_dmstat);
        if (params) memcpy(params + (size_t)r * 8, f.params, sizeof(double) * 8);
    }
    return 0;
}

extern "C" void pfx_phcx_free(pfx_scan* scan) { delete scan; }

// ---- PRESTO .pfd files ---------------------------------------------------------------------------
namespace {

struct PfdCand {
    std::vector<double> profile, subbands;
    std::vector<float> chis;
};

// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src), which ndarray.sum() and
// numpy.mean use on contiguous data: the reference's avgprof, mean(normprof) and chi-square sums
double np_pairwise(const double* a, size_t n) {
    if (n < 8) {
        double r = 0.0;
        for (size_t i = 0; i < n; ++i) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        size_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    size_t n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise(a, n2) + np_pairwise(a + n2, n - n2);
}

struct Reader {
    const unsigned char* p;
    size_t n, pos;
    bool swap, bad;
    void get(void* dst, size_t size, size_t count) {
        if (bad || pos + size * count > n) {
            bad = true;
            memset(dst, 0, size * count);
            return;
        }
        memcpy(dst, p + pos, size * count);
        pos += size * count;
        if (swap && size > 1) {
            unsigned char* d = (unsigned char*)dst;
            for (size_t i = 0; i < count; ++i)
                for (size_t b = 0; b < size / 2; ++b) {
                    const unsigned char t = d[i * size + b];
                    d[i * size + b] = d[i * size + size - 1 - b];
                    d[i * size + size - 1 - b] = t;
                }
        }
    }
    int32_t i32() {
        int32_t v;
        get(&v, 4, 1);
        return v;
    }
    double f64() {
        double v;
        get(&v, 8, 1);
        return v;
    }
    void skip(size_t k) {
        if (bad || pos + k > n) bad = true;
        else pos += k;
    }
};

inline long pymod(long a, long m) {
    const long r = a % m;
    return r < 0 ? r + m : r;
}

void rotate_left(double* row, int len, long bins, std::vector<double>* tmp) {
    const long r = pymod(bins, len);
    if (!r) return;
    tmp->assign(row, row + len);
    for (int k = 0; k < len; ++k) row[k] = (*tmp)[(size_t)((k + r) % len)];
}

bool parse_pfd(const char* path, PfdCand* c, pfx_pfd_info* info) {
    std::string raw;
    if (!read_file(path, false, &raw)) return false;
    Reader rd = {(const unsigned char*)raw.data(), raw.size(), 0, false, false};
    if (raw.size() < 20) return false;
    {   // endianness test of PFDFile.py:118-124: none of the first five ints should be large
        int32_t t[5];
        memcpy(t, raw.data(), 20);
        double mx = 0.0;
        for (int k = 0; k < 5; ++k) mx = std::fmax(mx, std::fabs((double)t[k]));
        rd.swap = mx > 100000.0;
    }
    const int32_t numdms = rd.i32(), numperiods = rd.i32(), numpdots = rd.i32(), nsub = rd.i32(), npart = rd.i32();
    const int32_t proflen = rd.i32(), numchan = rd.i32();
    for (int k = 0; k < 5; ++k) rd.i32();
    for (int k = 0; k < 4; ++k) {
        const int32_t ln = rd.i32();
        if (ln < 0) return false;
        rd.skip((size_t)ln);
    }
    if (rd.bad || rd.pos + 16 > rd.n) return false;
    bool has_posn = true;
    for (int k = 0; k < 16; ++k) {
        const char ch = (char)rd.p[rd.pos + k];
        if (!((ch >= '0' && ch <= '9') || ch == ':' || ch == '.' || ch == '-' || ch == '\0')) {
            has_posn = false;
            break;
        }
    }
    if (has_posn) rd.skip(32);
    rd.f64();                      // dt
    rd.f64();                      // startT
    rd.f64();                      // endT
    rd.f64();                      // tepoch
    rd.f64();                      // bepoch
    rd.f64();                      // avgvoverc
    const double lofreq = rd.f64(), chan_wid = rd.f64(), bestdm = rd.f64();
    rd.skip(8);
    for (int k = 0; k < 3; ++k) rd.f64();              // topo
    rd.skip(8);
    const double bary_p1 = rd.f64();
    rd.f64();
    rd.f64();
    rd.skip(8);
    const double fold_p1 = rd.f64();
    rd.f64();
    rd.f64();
    for (int k = 0; k < 7; ++k) rd.f64();              // orbit
    if (rd.bad || numdms < 2 || numperiods < 0 || numpdots < 0 || nsub <= 0 || npart <= 0 || proflen <= 0 ||
        nsub > 65536 || npart > 65536 || proflen > 65535)
        return false;
    std::vector<double> dms((size_t)numdms);
    rd.get(dms.data(), 8, (size_t)numdms);
    rd.skip(8 * (size_t)numperiods);
    rd.skip(8 * (size_t)numpdots);
    const size_t ncube = (size_t)npart * nsub * proflen;
    std::vector<double> profs(ncube), stats((size_t)npart * nsub * 7);
    rd.get(profs.data(), 8, ncube);
    rd.get(stats.data(), 8, stats.size());
    if (rd.bad) return false;

    const double binspersec = fold_p1 * proflen;
    const long chanpersub = numchan / nsub;            // Python 2 int / int (both positive here or the file is refused below)
    if (numchan < 0) return false;
    const double subdeltafreq = chan_wid * (double)chanpersub;
    const double losubfreq = lofreq + subdeltafreq - chan_wid;
    std::vector<double> subfreqs((size_t)nsub);
    for (int j = 0; j < nsub; ++j) subfreqs[(size_t)j] = (double)j * subdeltafreq + losubfreq;
    // avgprof = (profs / proflen).sum(), varprof = sum of prof_var -- before any rotation
    double avgprof;
    {
        std::vector<double> q(ncube);
        for (size_t k = 0; k < ncube; ++k) q[k] = profs[k] / (double)proflen;
        avgprof = np_pairwise(q.data(), ncube);
    }
    double varprof = 0.0;
    for (size_t k = 0; k < (size_t)npart * nsub; ++k) varprof += stats[k * 7 + 5];

    auto delays = [&](double dm, std::vector<double>* out) {
        out->resize((size_t)nsub);
        for (int j = 0; j < nsub; ++j) {
            const double f = subfreqs[(size_t)j];
            (*out)[(size_t)j] = (f > 0.0) ? dm / ((0.000241 * f) * f) : 0.0;
        }
        const double hi = (*out)[(size_t)nsub - 1];
        for (int j = 0; j < nsub; ++j) (*out)[(size_t)j] -= hi;
    };
    std::vector<double> sd, sdb((size_t)nsub, 0.0), tmp;
    // dedisperse() at the best DM
    delays(bestdm, &sd);
    for (int j = 0; j < nsub; ++j) {
        const double nb = std::floor((sd[(size_t)j] * binspersec - 0.0) + 0.5);
        if (!(std::fabs(nb) < 9e15)) return false;      // NaN / inf delay: int() raises in the reference
        for (int i = 0; i < npart; ++i) rotate_left(&profs[((size_t)i * nsub + j) * proflen], proflen, (long)nb, &tmp);
        sdb[(size_t)j] = 0.0 + nb;
    }
    // time-summed sub-bands (profs.sum(0)) and the summed profile (.sum(0) again), both sequential
    std::vector<double> sub((size_t)nsub * proflen, 0.0);
    for (int i = 0; i < npart; ++i)
        for (size_t k = 0; k < (size_t)nsub * proflen; ++k) {
            const double v = profs[(size_t)i * nsub * proflen + k];
            sub[k] = (i == 0) ? v : sub[k] + v;
        }
    std::vector<double> sumprof((size_t)proflen);
    for (int j = 0; j < nsub; ++j)
        for (int k = 0; k < proflen; ++k) {
            const double v = sub[(size_t)j * proflen + k];
            sumprof[(size_t)k] = (j == 0) ? v : sumprof[(size_t)k] + v;
        }
    // getprofile() + scale(): builtin min / max keep the first of equal values and let a NaN win only
    // when it comes first
    double mn = sumprof[0];
    for (int k = 1; k < proflen; ++k)
        if (sumprof[(size_t)k] < mn) mn = sumprof[(size_t)k];
    std::vector<double> s((size_t)proflen);
    for (int k = 0; k < proflen; ++k) s[(size_t)k] = sumprof[(size_t)k] - mn;
    const double mean = np_pairwise(s.data(), (size_t)proflen) / (double)proflen;
    for (int k = 0; k < proflen; ++k) s[(size_t)k] = s[(size_t)k] / mean;
    double smin = s[0], smax = s[0];
    for (int k = 1; k < proflen; ++k) {
        if (s[(size_t)k] < smin) smin = s[(size_t)k];
        if (s[(size_t)k] > smax) smax = s[(size_t)k];
    }
    c->profile.resize((size_t)proflen);
    for (int k = 0; k < proflen; ++k) {
        const double t = (s[(size_t)k] - smin) / (smax - smin);
        c->profile[(size_t)k] = (0.0 * (1.0 - t)) + (255.0 * t);
    }
    c->subbands = sub;
    // plot_chi2_vs_DM(dms[0], dms[-1], 100)
    const int N = 100;
    c->chis.assign((size_t)N, 0.0f);
    const double lo = dms[0], hi = dms[(size_t)numdms - 1];
    std::vector<double> work = sub, term((size_t)proflen), sp((size_t)proflen);
    double dm1 = 0.0, dmlast = 0.0;
    for (int ii = 0; ii < N; ++ii) {
        const double dm = lo + ((hi - lo) * (double)ii) / (double)(N - 1);
        if (ii == 1) dm1 = dm;
        if (ii == N - 1) dmlast = dm;
        delays(dm, &sd);
        for (int j = 0; j < nsub; ++j) {
            const double nb = std::floor((sd[(size_t)j] * binspersec - sdb[(size_t)j]) + 0.5);
            if (!(std::fabs(nb) < 9e15)) return false;
            rotate_left(&work[(size_t)j * proflen], proflen, (long)nb, &tmp);
            sdb[(size_t)j] += nb;
        }
        for (int j = 0; j < nsub; ++j)
            for (int k = 0; k < proflen; ++k) {
                const double v = work[(size_t)j * proflen + k];
                sp[(size_t)k] = (j == 0) ? v : sp[(size_t)k] + v;
            }
        for (int k = 0; k < proflen; ++k) {
            const double d = sp[(size_t)k] - avgprof;
            term[(size_t)k] = (d * d) / varprof;
        }
        c->chis[(size_t)ii] = (float)(np_pairwise(term.data(), (size_t)proflen) / ((double)proflen - 1.0));
    }
    info->nbin = proflen;
    info->nsub = nsub;
    info->n_dm = N;
    info->params[0] = 0.0;
    info->params[1] = bestdm;
    info->params[2] = bary_p1;
    info->params[3] = 0.0;
    info->params[4] = dm1;
    info->params[5] = dmlast;
    info->params[6] = (double)N;
    info->params[7] = 0.0;
    return true;
}

}  // namespace

struct pfx_pfd_scan_impl {
    std::vector<PfdCand> cands;
    std::vector<pfx_pfd_info> info;
};

extern "C" int pfx_pfd_scan(const char* const* paths, int64_t n, int threads, pfx_pfd_info* info, pfx_scan** out) {
    if (!paths || n < 0 || !info || !out) return -1;
    pfx_pfd_scan_impl* h = new (std::nothrow) pfx_pfd_scan_impl;
    if (!h) return -1;
    try {
        h->cands.resize((size_t)n);
        h->info.resize((size_t)n);
    } catch (...) {
        delete h;
        return -1;
    }
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if ((int64_t)nt > n) nt = (int)(n > 0 ? n : 1);
    std::atomic<int64_t> next(0);
    auto work = [&]() {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n) break;
            pfx_pfd_info fi;
            memset(&fi, 0, sizeof(fi));
            bool ok = false;
            try {
                ok = paths[i] != nullptr && parse_pfd(paths[i], &h->cands[(size_t)i], &fi);
            } catch (...) {
                ok = false;
            }
            fi.ok = ok ? 1 : 0;
            if (!ok) h->cands[(size_t)i] = PfdCand();
            h->info[(size_t)i] = fi;
            info[i] = fi;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    *out = (pfx_scan*)h;
    return 0;
}

extern "C" int pfx_pfd_gather(const pfx_scan* scan, const int64_t* idx, int64_t count, double* profile,
                              double* subbands, float* dmcurve, double* params) {
    const pfx_pfd_scan_impl* h = (const pfx_pfd_scan_impl*)scan;
    if (!h || !idx || count < 0) return -1;
    if (count == 0) return 0;
    const int64_t n = (int64_t)h->cands.size();
    if (idx[0] < 0 || idx[0] >= n) return -1;
    const pfx_pfd_info& f0 = h->info[(size_t)idx[0]];
    for (int64_t r = 0; r < count; ++r) {
        if (idx[r] < 0 || idx[r] >= n) return -1;
        const pfx_pfd_info& f = h->info[(size_t)idx[r]];
        if (!f.ok || f.nbin != f0.nbin || f.nsub != f0.nsub || f.n_dm != f0.n_dm) return -2;
    }
    for (int64_t r = 0; r < count; ++r) {
        const PfdCand& c = h->cands[(size_t)idx[r]];
        const pfx_pfd_info& f = h->info[(size_t)idx[r]];
        if (profile) memcpy(profile + (size_t)r * f.nbin, c.profile.data(), sizeof(double) * (size_t)f.nbin);
        if (subbands) memcpy(subbands + (size_t)r * f.nsub * f.nbin, c.subbands.data(), sizeof(double) * (size_t)f.nsub * f.nbin);
        if (dmcurve) memcpy(dmcurve + (size_t)r * f.n_dm, c.chis.data(), sizeof(float) * (size_t)f.n_dm);
        if (params) memcpy(params + (size_t)r * 8, f.params, sizeof(double) * 8);
    }
    return 0;
}

extern "C" void pfx_pfd_free(pfx_scan* scan) { delete (pfx_pfd_scan_impl*)scan; }

// ---- output rows ---------------------------------------------------------------------------------
namespace {

// Python 2 str(float): '%.12g'; a result without '.', 'e' (and not nan/inf) gets '.0'
void fmt_py2(double v, std::string* out) {
    if (v != v) {
        out->append("nan");
        return;
    }
    if (std::isinf(v)) {
        out->append(v > 0 ? "inf" : "-inf");
        return;
    }
    char buf[40];
    const int n = snprintf(buf, sizeof(buf), "%.12g", v);
    out->append(buf, (size_t)n);
    if (!memchr(buf, '.', (size_t)n) && !memchr(buf, 'e', (size_t)n)) out->append(".0");
}

// Python 3 repr(float): the shortest digit string that round-trips, fixed notation for decimal
// exponents in [-4, 16), else d.ddde+XX (float_repr_style 'short')
void fmt_repr(double v, std::string* out) {
    if (v != v) {
        out->append("nan");
        return;
    }
    if (std::isinf(v)) {
        out->append(v > 0 ? "inf" : "-inf");
        return;
    }
    if (v == 0.0) {
        out->append(std::signbit(v) ? "-0.0" : "0.0");
        return;
    }
    char buf[48];
    int prec = 0;
    for (prec = 0; prec < 17; ++prec) {       // prec = digits after the point in scientific notation
        snprintf(buf, sizeof(buf), "%.*e", prec, v);
        if (strtod(buf, nullptr) == v) break;
    }
    // buf = [-]d[.ddd]e[+-]XX
    const char* p = buf;
    if (*p == '-') {
        out->push_back('-');
        ++p;
    }
    std::string digits;
    digits.push_back(*p++);
    if (*p == '.') {
        ++p;
        while (*p && *p != 'e') digits.push_back(*p++);
    }
    const int exp10 = atoi(p + 1);           // after 'e'
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const int nd = (int)digits.size();
    const int decpt = exp10 + 1;             // position of the decimal point relative to the digits
    if (decpt > -4 && decpt <= 16) {
        if (decpt <= 0) {
            out->append("0.");
            out->append((size_t)(-decpt), '0');
            out->append(digits);
        } else if (decpt >= nd) {
            out->append(digits);
            out->append((size_t)(decpt - nd), '0');
            out->append(".0");
        } else {
            out->append(digits, 0, (size_t)decpt);
            out->push_back('.');
            out->append(digits, (size_t)decpt, std::string::npos);
        }
    } else {
        out->push_back(digits[0]);
        if (nd > 1) {
            out->push_back('.');
            out->append(digits, 1, std::string::npos);
        }
        char e[16];
        snprintf(e, sizeof(e), "e%c%02d", exp10 < 0 ? '-' : '+', exp10 < 0 ? -exp10 : exp10);
        out->append(e);
    }
}

// line.replace("nan", "0").replace("inf", "0")
void replace_all(std::string* s, const char* what) {
    size_t pos = 0;
    while ((pos = s->find(what, pos)) != std::string::npos) {
        s->replace(pos, 3, "0");
        pos += 1;
    }
}

}  // namespace

extern "C" int pfx_format_rows(const double* feat, int64_t count, int64_t ld, int ncol, const char* const* names,
                               const int8_t* skip, int mode, int py2, int threads, char** out, int64_t* out_len) {
    if (!feat || count < 0 || ncol < 1 || ld < ncol || mode < 0 || mode > 2 || !out || !out_len) return -1;
    if (mode != 2 && !names) return -1;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if ((int64_t)nt * 256 > count) nt = (int)(count / 256 > 0 ? count / 256 : 1);
    std::vector<std::string> parts((size_t)nt);
    std::atomic<bool> failed(false);
    auto work = [&](int t) {
        try {
            const int64_t lo = count * t / nt, hi = count * (t + 1) / nt;
            std::string& acc = parts[(size_t)t];
            acc.reserve((size_t)(hi - lo) * (size_t)(ncol * 14 + 48));
            std::string line;
            for (int64_t i = lo; i < hi; ++i) {
                if (skip && skip[i]) continue;
                line.clear();
                if (mode == 0) {
                    line.append(names[i] ? names[i] : "");
                    line.push_back(',');
                }
                const double* row = feat + i * ld;
                for (int k = 0; k < ncol; ++k) {
                    if (k) line.push_back(',');
                    if (py2) fmt_py2(row[k], &line);
                    else fmt_repr(row[k], &line);
                }
                if (mode == 1) {
                    line.append(",?%");
                    line.append(names[i] ? names[i] : "");
                }
                replace_all(&line, "nan");
                replace_all(&line, "inf");
                line.push_back('\n');
                acc.append(line);
            }
        } catch (...) {
            failed = true;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    if (failed) return -1;
    size_t total = 0;
    for (auto& p : parts) total += p.size();
    char* buf = (char*)malloc(total + 1);
    if (!buf) return -1;
    size_t off = 0;
    for (auto& p : parts) {
        memcpy(buf + off, p.data(), p.size());
        off += p.size();
    }
    buf[total] = 0;
    *out = buf;
    *out_len = (int64_t)total;
    return 0;
}

extern "C" void pfx_io_free(void* p) { free(p); }

extern "C" const char* pfx_io_version(void) { return "pfxio 0.2 (PHCX scanner, row formatter)"; }
