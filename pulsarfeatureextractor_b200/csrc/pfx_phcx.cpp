// pfx_phcx.cpp -- libpfxio.so: streaming PHCX scanner (include/pfx_io.h).
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
    if (!tag_block(*text, "SubBands", sec, &b)) return false;
    long nbins = 0, nsub = 0;
    if (!attr_int(b.attrs, "nBins", &nbins) || !attr_int(b.attrs, "nSub", &nsub)) return false;
    hex_decode(b.body, &c->subbands);
    if (nbins <= 0 || nsub <= 0 || (size_t)nbins * (size_t)nsub != c->subbands.size()) return false;
    if ((size_t)nbins != c->profile.size()) return false;      // one nbin per candidate in the batch layout
    std::vector<uint8_t> block;
    if (!tag_block(*text, "DataBlock", sec, &b)) return false;
    hex_decode(b.body, &block);
    if (!tag_block(*text, "DataBlock", 0, &b)) return false;
    hex_decode(b.body, &c->dmstat);
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
            if (!ok) h->cands[(size_t)i] = Cand();
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
    for (int64_t r = 0; r < count; ++r) {
        if (idx[r] < 0 || idx[r] >= n) return -1;
        const pfx_phcx_info& f = scan->info[(size_t)idx[r]];
        if (!f.ok || f.nbin != f0.nbin || f.nsub != f0.nsub || f.n_dm != f0.n_dm || f.n_dmstat != f0.n_dmstat)
            return -2;
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
