// Host-side text <-> batch conversion around the hot path (plain C++17, multi-threaded; no
// device code -- it lives in this library so that the feature stream can be parsed straight
// into the page-locked buffers dsp_forward_host_submit copies from).
//
//   dsp_parse_features  feature file of `deepsignal_plant extract` (12 tab-separated columns,
//                       extract_features.py:381-395) -> the five float32 arrays ModelBiLSTM.forward
//                       takes, as _read_features_file builds them line by line
//                       (call_modifications.py:55-127) and FloatTensor converts them
//                       (utils/constants_torch.py:10-13: Python float = double, then float32).
//   dsp_format_calls    probabilities -> call_mods text lines (call_modifications.py:175-188):
//                       float32 renormalise / round(6), str() of a numpy float32 (shortest
//                       round-trip digits, scientific below 1e-4), argmax label, centre 5-mer.
#include "common.cuh"
#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstring>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

namespace dsp {
namespace {

// utils/process_utils.py:22-29 (base2code_dna); anything else is a KeyError in the reference
int base_code(char ch) {
    switch (ch) {
        case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
        case 'N': return 4; case 'W': return 5; case 'S': return 6; case 'M': return 7;
        case 'K': return 8; case 'R': return 9; case 'Y': return 10; case 'B': return 11;
        case 'V': return 12; case 'D': return 13; case 'H': return 14; case 'Z': return 15;
        default: return -1;
    }
}

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

// Python float(str): decimal -> nearest double; the float32 tensor then rounds the double.
// Fast path (Clinger): up to 15 significant digits and no exponent -- what `extract` writes
// (np.around(x, 6)) -- is an exact integer divided by an exact power of ten, one correctly rounded
// operation; everything else goes through std::from_chars (also correctly rounded).
bool parse_double(const char* b, const char* e, double* out) {
    static const double P10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                   1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    while (b < e && is_space(*b)) ++b;
    while (e > b && is_space(e[-1])) --e;
    if (b >= e) return false;
    {
        const char* p = b;
        bool neg = false;
        if (*p == '-') { neg = true; ++p; } else if (*p == '+') ++p;
        unsigned long long mant = 0;
        int ndig = 0, frac = -1;
        for (; p < e; ++p) {
            const unsigned c = (unsigned)(*p - '0');
            if (c <= 9) { mant = mant * 10 + c; ++ndig; if (frac >= 0) ++frac; }
            else if (*p == '.' && frac < 0) frac = 0;
            else break;
        }
        if (p == e && ndig >= 1 && ndig <= 15) {
            const double v = (double)mant / P10[frac < 0 ? 0 : frac];
            *out = neg ? -v : v;
            return true;
        }
    }
    if (*b == '+') ++b;
    if (b >= e) return false;
    auto r = std::from_chars(b, e, *out);
    if (r.ec == std::errc::result_out_of_range) { *out = (*b == '-') ? -HUGE_VAL : HUGE_VAL; return r.ptr == e; }   // float('1e999') == inf
    return r.ec == std::errc() && r.ptr == e;
}

bool parse_int(const char* b, const char* e, long long* out) {
    while (b < e && is_space(*b)) ++b;
    while (e > b && is_space(e[-1])) --e;
    if (b < e && *b == '+') ++b;
    if (b >= e) return false;
    auto r = std::from_chars(b, e, *out);
    return r.ec == std::errc() && r.ptr == e;
}

// `count` values separated by `sep` in [b, e)
template <typename F>
bool split_exact(const char* b, const char* e, char sep, int count, F&& each) {
    for (int i = 0; i < count; ++i) {
        const char* s = (i + 1 < count) ? (const char*)memchr(b, sep, (size_t)(e - b)) : e;
        if (s == nullptr) return false;
        if (i + 1 == count && memchr(b, sep, (size_t)(e - b)) != nullptr) return false;   // too many fields
        if (!each(i, b, s)) return false;
        b = s + 1;
    }
    return true;
}

// `count` floating-point values separated by `sep` filling [b, e) exactly, tokenised and converted in one
// forward scan: the common spelling -- [-]digits[.digits], at most 15 significant digits, which is what
// `extract` writes -- is an exact integer divided by an exact power of ten (one correctly rounded operation,
// like parse_double's fast path); anything else in a field goes through parse_double on that field.
// value of up to eight ASCII digits held in the TOP bytes of y (first digit in the lowest of them), '0' already
// subtracted, lower bytes zero: the usual three multiply-and-shift steps (pairs, quads, all eight)
inline uint32_t eight_digits(uint64_t y) {
    y = (y * 10) + (y >> 8);
    return (uint32_t)((((y & 0x000000FF000000FFull) * 0x000F424000000064ull) +
                       (((y >> 16) & 0x000000FF000000FFull) * 0x0000271000000001ull)) >> 32);
}

template <typename Store>
bool parse_float_list(const char* b, const char* e, char sep, int count, Store&& store) {
    static const double P10[16] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
    static const unsigned long long P10I[8] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull};
    const char* p = b;
    for (int i = 0; i < count; ++i) {
        const char* q = p;
        bool neg = false;
        if (q < e && *q == '-') { neg = true; ++q; }
        // the spelling `extract` writes, [-]d+.d{0,7} followed by the separator, with the fraction taken eight
        // bytes at a time (needs eight readable bytes after the point, i.e. not the last values of the list)
        {
            const char* r = q;
            unsigned long long ip = 0;
            int ni = 0;
            for (; r < e && (unsigned)(*r - '0') <= 9 && ni < 8; ++r, ++ni) ip = ip * 10 + (unsigned)(*r - '0');
            if (ni >= 1 && ni <= 7 && r + 9 <= e && *r == '.') {
                uint64_t x;
                memcpy(&x, r + 1, 8);
                const uint64_t ge0 = (x | 0x8080808080808080ull) - 0x3030303030303030ull;      // high bit: byte >= '0'
                const uint64_t nondigit = (~ge0 | (x + 0x4646464646464646ull) | x) & 0x8080808080808080ull;
                if (nondigit) {
                    const int n = __builtin_ctzll(nondigit) >> 3;                             // fraction digits, 0..7
                    if (r[1 + n] == sep) {
                        const int sh = 8 * (8 - n);
                        const unsigned long long frac = n ? eight_digits((x << sh) - (0x3030303030303030ull << sh)) : 0;
                        double v = (double)(ip * P10I[n] + frac) / P10[n];
                        if (neg) v = -v;
                        if (i + 1 == count) return false;                                     // a separator follows: too many fields
                        store(i, v);
                        p = r + 2 + n;
                        continue;
                    }
                }
            }
        }
        unsigned long long mant = 0;
        int nd = 0, frac = -1;
        for (; q < e; ++q) {
            const unsigned c = (unsigned)(*q - '0');
            if (c <= 9) { mant = mant * 10 + c; ++nd; if (frac >= 0) ++frac; }
            else if (*q == '.' && frac < 0) frac = 0;
            else break;
        }
        double v;
        if ((q == e || *q == sep) && nd >= 1 && nd <= 15) {
            v = (double)mant / P10[frac < 0 ? 0 : frac];
            if (neg) v = -v;
        } else {
            const char* s = (const char*)memchr(p, sep, (size_t)(e - p));
            if (s == nullptr) s = e;
            if (!parse_double(p, s, &v)) return false;
            q = s;
        }
        if (i + 1 < count) { if (q == e) return false; p = q + 1; }      // too few fields
        else if (q != e) return false;                                   // too many fields
        store(i, v);
    }
    return true;
}

struct ParseJob {
    const char* text; const int64_t* begin; const int64_t* end; int T, S;
    float *kmer, *means, *stds, *lens, *signals; int32_t* labels; int32_t* info_len;
};

// returns 0 or the 1-based index of the offending column
int parse_line(const ParseJob& j, int64_t i) {
    const char* b = j.text + j.begin[i];
    const char* e = j.text + j.end[i];
    const char* f[13];
    f[0] = b;
    int nf = 1;
    for (const char* p = b; nf <= 12;) {                       // 12 columns = 11 tabs; a 12th tab is an error
        const char* t = (const char*)memchr(p, '\t', (size_t)(e - p));
        if (t == nullptr) break;
        f[nf++] = t + 1;
        p = t + 1;
    }
    if (nf != 12) return 13;
    f[12] = e + 1;
    const int T = j.T, S = j.S;
    j.info_len[i] = (int32_t)(f[6] - 1 - b);
    if (f[7] - 1 - f[6] != T) return 7;
    for (int t = 0; t < T; ++t) {
        const int code = base_code(f[6][t]);
        if (code < 0) return 7;
        j.kmer[i * T + t] = (float)code;
    }
    if (!parse_float_list(f[7], f[8] - 1, ',', T, [&](int t, double d) { j.means[i * T + t] = (float)d; })) return 8;
    if (!parse_float_list(f[8], f[9] - 1, ',', T, [&](int t, double d) { j.stds[i * T + t] = (float)d; })) return 9;
    long long v;
    if (!split_exact(f[9], f[10] - 1, ',', T, [&](int t, const char* x, const char* y) {
            if (!parse_int(x, y, &v)) return false; j.lens[i * T + t] = (float)v; return true; })) return 10;
    if (!split_exact(f[10], f[11] - 1, ';', T, [&](int t, const char* x, const char* y) {
            return parse_float_list(x, y, ',', S, [&](int s, double d) { j.signals[(i * T + t) * S + s] = (float)d; }); })) return 11;
    if (!parse_int(f[11], e, &v)) return 12;
    j.labels[i] = (int32_t)v;
    return 0;
}

template <typename F>
void parallel_for(int64_t n, int nthreads, F&& body, int64_t serial_below = 256) {
    if (nthreads <= 1 || n < serial_below) { body(0, n); return; }
    std::vector<std::thread> th;
    const int64_t per = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        const int64_t lo = t * per, hi = lo + per < n ? lo + per : n;
        if (lo >= hi) break;
        th.emplace_back([=, &body] { body(lo, hi); });
    }
    for (auto& x : th) x.join();
}

// str(numpy.float32(x)): shortest digits that round-trip in float32 (Dragon4, unique), positional
// for 1e-4 <= |x| < 1e16 (compared as a value), scientific otherwise, at least one digit after a positional point,
// exponent with at least two digits
int format_f32(float x, char* out) {
    if (std::isnan(x)) { memcpy(out, "nan", 3); return 3; }
    if (std::isinf(x)) { if (x < 0) { memcpy(out, "-inf", 4); return 4; } memcpy(out, "inf", 3); return 3; }
    char* o = out;
    if (std::signbit(x)) { *o++ = '-'; x = -x; }
    if (x == 0.f) { memcpy(o, "0.0", 3); return (int)(o + 3 - out); }
    char buf[48];
    auto r = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);   // d[.ddd]e[+-]XX, shortest
    char digits[16]; int nd = 0; const char* p = buf;
    for (; p < r.ptr && *p != 'e'; ++p) if (*p != '.') digits[nd++] = *p;
    int ex = 0; { ++p; bool neg = (*p == '-'); ++p; for (; p < r.ptr; ++p) ex = ex * 10 + (*p - '0'); if (neg) ex = -ex; }
    // numpy decides on the VALUE (float32(1e-4) = 9.9999997e-05 prints as '1e-04'), not on the digits
    if ((double)x >= 1e-4 && (double)x < 1e16) {
        if (ex < 0) {
            *o++ = '0'; *o++ = '.';
            for (int k = 0; k < -ex - 1; ++k) *o++ = '0';
            memcpy(o, digits, nd); o += nd;
        } else {
            for (int k = 0; k <= ex; ++k) *o++ = k < nd ? digits[k] : '0';
            *o++ = '.';
            if (nd > ex + 1) { memcpy(o, digits + ex + 1, nd - ex - 1); o += nd - ex - 1; } else *o++ = '0';
        }
    } else {
        *o++ = digits[0];
        if (nd > 1) { *o++ = '.'; memcpy(o, digits + 1, nd - 1); o += nd - 1; }
        *o++ = 'e'; *o++ = ex < 0 ? '-' : '+';
        int a = ex < 0 ? -ex : ex;
        if (a >= 100) { *o++ = (char)('0' + a / 100); a %= 100; }
        *o++ = (char)('0' + a / 10); *o++ = (char)('0' + a % 10);
    }
    return (int)(o - out);
}

// str(numpy.float64(x)) == repr(float(x)): shortest digits that round-trip in float64, positional for
// 1e-4 <= |x| < 1e16, scientific otherwise, at least one digit after a positional point, exponent with at
// least two digits.  (What _features_to_str prints for means, stds and signals, extract_features.py:388-392.)
int format_f64(double x, char* out) {
    // fast path for what the feature file is made of: a value that IS k / 1e6 (np.around(., 6)) with
    // 1e-4 <= |x| and at most 15 significant digits prints as that decimal, trailing zeros dropped
    {
        const double ax = std::fabs(x);
        if (ax >= 1e-4 && ax < 1e9) {
            const double kd = std::nearbyint(ax * 1e6);
            if (kd / 1e6 == ax) {
                char* o = out;
                if (std::signbit(x)) *o++ = '-';
                unsigned long long k = (unsigned long long)kd;
                const unsigned long long ip = k / 1000000ull;
                unsigned fp = (unsigned)(k % 1000000ull);
                o = std::to_chars(o, o + 24, ip).ptr;
                *o++ = '.';
                if (fp == 0) { *o++ = '0'; return (int)(o - out); }
                char d[6];
                for (int i = 5; i >= 0; --i) { d[i] = (char)('0' + fp % 10); fp /= 10; }
                int nd = 6;
                while (d[nd - 1] == '0') --nd;
                memcpy(o, d, (size_t)nd);
                return (int)(o + nd - out);
            }
        }
    }
    if (std::isnan(x)) { memcpy(out, "nan", 3); return 3; }
    if (std::isinf(x)) { if (x < 0) { memcpy(out, "-inf", 4); return 4; } memcpy(out, "inf", 3); return 3; }
    char* o = out;
    if (std::signbit(x)) { *o++ = '-'; x = -x; }
    if (x == 0.0) { memcpy(o, "0.0", 3); return (int)(o + 3 - out); }
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);   // d[.ddd]e[+-]XX, shortest
    char digits[24]; int nd = 0; const char* p = buf;
    for (; p < r.ptr && *p != 'e'; ++p) if (*p != '.') digits[nd++] = *p;
    int ex = 0; { ++p; bool neg = (*p == '-'); ++p; for (; p < r.ptr; ++p) ex = ex * 10 + (*p - '0'); if (neg) ex = -ex; }
    if (x >= 1e-4 && x < 1e16) {
        if (ex < 0) {
            *o++ = '0'; *o++ = '.';
            for (int k = 0; k < -ex - 1; ++k) *o++ = '0';
            memcpy(o, digits, nd); o += nd;
        } else {
            for (int k = 0; k <= ex; ++k) *o++ = k < nd ? digits[k] : '0';
            *o++ = '.';
            if (nd > ex + 1) { memcpy(o, digits + ex + 1, nd - ex - 1); o += nd - ex - 1; } else *o++ = '0';
        }
    } else {
        *o++ = digits[0];
        if (nd > 1) { *o++ = '.'; memcpy(o, digits + 1, nd - 1); o += nd - 1; }
        *o++ = 'e'; *o++ = ex < 0 ? '-' : '+';
        int a = ex < 0 ? -ex : ex;
        if (a >= 100) { *o++ = (char)('0' + a / 100); a %= 100; }
        *o++ = (char)('0' + a / 10); *o++ = (char)('0' + a % 10);
    }
    return (int)(o - out);
}

// numpy.round(x, 6) on float32: rint(x * 1e6f) / 1e6f in float32
inline float round6(float x) { return rintf(x * 1e6f) / 1e6f; }

}  // namespace
}  // namespace dsp

using namespace dsp;

extern "C" {

int dsp_parse_features(const char* text, int64_t nbytes, int32_t is_final, int32_t seq_len, int32_t signal_len,
                       int64_t max_sites, float* kmer, float* means, float* stds, float* lens, float* signals,
                       int32_t* labels, char* info_text, int64_t info_cap, int64_t* info_off,
                       int64_t* n_sites, int64_t* consumed, int32_t nthreads) {
    DSP_REQUIRE(text && kmer && means && stds && lens && signals && labels && info_text && info_off &&
                n_sites && consumed, DSP_ERR_INVALID, "dsp_parse_features: null argument");
    DSP_REQUIRE(seq_len >= 1 && signal_len >= 1 && max_sites >= 0 && nbytes >= 0, DSP_ERR_INVALID,
                "dsp_parse_features: bad dimensions");
    // pass 1: complete lines (line.strip(): surrounding white space does not belong to a field)
    std::vector<int64_t> begins, ends;
    begins.reserve((size_t)(max_sites < (1 << 20) ? max_sites : (1 << 20)));
    ends.reserve(begins.capacity());
    // the newline positions of the block, found by all workers (this is also where a mapped file's pages are
    // first touched, so the page-cache faults are spread over the threads)
    std::vector<int64_t> newlines;
    {
        const int P = nthreads > 1 && nbytes >= (1 << 20) ? nthreads : 1;
        std::vector<std::vector<int64_t>> part((size_t)P);
        parallel_for(P, P, [&](int64_t a, int64_t b) {
            for (int64_t r = a; r < b; ++r) {
                const int64_t lo = nbytes * r / P, hi = nbytes * (r + 1) / P;
                std::vector<int64_t>& v = part[(size_t)r];
                v.reserve((size_t)((hi - lo) / 1024 + 16));
                for (int64_t q = lo; q < hi;) {
                    const char* nl = (const char*)memchr(text + q, '\n', (size_t)(hi - q));
                    if (!nl) break;
                    v.push_back(nl - text);
                    q = (nl - text) + 1;
                }
            }
        }, 2);
        size_t total = 0;
        for (auto& v : part) total += v.size();
        newlines.reserve(total);
        for (auto& v : part) newlines.insert(newlines.end(), v.begin(), v.end());
    }
    int64_t pos = 0, n = 0;
    size_t next_nl = 0;
    while (pos < nbytes && n < max_sites) {
        const bool has_nl = next_nl < newlines.size();
        const char* nl = has_nl ? text + newlines[next_nl++] : nullptr;
        int64_t stop;
        if (nl) stop = nl - text; else if (is_final) stop = nbytes; else break;
        int64_t b = pos, e = stop;
        while (b < e && is_space(text[b])) ++b;
        while (e > b && is_space(text[e - 1])) --e;
        pos = nl ? stop + 1 : nbytes;
        if (b == e) {
            // the reference dies on a blank line inside the file (IndexError); a trailing one is never read
            bool only_space_left = true;
            for (int64_t k = pos; k < nbytes; ++k) if (!is_space(text[k])) { only_space_left = false; break; }
            DSP_REQUIRE(only_space_left && is_final, DSP_ERR_INVALID, "feature file: empty line at byte %lld", (long long)b);
            pos = nbytes;
            break;
        }
        begins.push_back(b);
        ends.push_back(e);
        ++n;
    }
    std::vector<int32_t> info_len((size_t)n);
    ParseJob job{text, begins.data(), ends.data(), seq_len, signal_len, kmer, means, stds, lens, signals, labels, info_len.data()};
    std::atomic<int64_t> bad_line{-1};
    std::atomic<int> bad_col{0};
    parallel_for(n, nthreads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            const int col = parse_line(job, i);
            if (col) {
                int64_t expect = -1;
                if (bad_line.compare_exchange_strong(expect, i)) bad_col = col;
                return;
            }
        }
    });
    if (bad_line >= 0) {
        if (bad_col == 13) set_error("feature file: line %lld (of this block) does not have 12 tab-separated columns", (long long)bad_line.load() + 1);
        else set_error("feature file: line %lld (of this block), column %d is malformed for seq_len %d / signal_len %d",
                       (long long)bad_line.load() + 1, bad_col.load(), seq_len, signal_len);
        return DSP_ERR_INVALID;
    }
    // the first six columns of every line ("sampleinfo", :89), packed back to back
    info_off[0] = 0;
    for (int64_t i = 0; i < n; ++i) info_off[i + 1] = info_off[i] + info_len[i];
    DSP_REQUIRE(info_off[n] <= info_cap, DSP_ERR_NOMEM, "dsp_parse_features: sample info needs %lld bytes, buffer has %lld",
                (long long)info_off[n], (long long)info_cap);
    parallel_for(n, nthreads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) memcpy(info_text + info_off[i], text + begins[i], (size_t)info_len[i]);
    });
    *n_sites = n;
    *consumed = pos;
    return DSP_OK;
}

int dsp_format_calls(const char* info_text, const int64_t* info_off, const float* kmer, int32_t seq_len,
                     const float* probs, const int32_t* labels, int64_t n,
                     char* out, int64_t out_cap, int64_t* out_bytes, int32_t nthreads) {
    DSP_REQUIRE(info_text && info_off && kmer && probs && labels && out && out_bytes, DSP_ERR_INVALID,
                "dsp_format_calls: null argument");
    static const char BASES[] = "ACGTNWSMKRYBVDHZ";              // code2base_dna (utils/process_utils.py:29)
    const int c = seq_len / 2;                                    // call_modifications.py:181-184
    const int lo5 = c - 2 > 0 ? c - 2 : 0, hi5 = c + 3 < seq_len ? c + 3 : seq_len;
    struct Num { char s[2][20]; uint8_t len[2]; };
    std::vector<Num> nums((size_t)n);
    std::vector<int64_t> off((size_t)n + 1);
    parallel_for(n, nthreads, [&](int64_t a, int64_t b) {
        for (int64_t i = a; i < b; ++i) {
            const float p0 = probs[2 * i], p1 = probs[2 * i + 1];
            const float p0n = round6(p0 / (p0 + p1));             // :177-179, float32 arithmetic
            const float p1n = round6(1.0f - p0n);
            nums[i].len[0] = (uint8_t)format_f32(p0n, nums[i].s[0]);
            nums[i].len[1] = (uint8_t)format_f32(p1n, nums[i].s[1]);
            char lab[16];
            auto r = std::to_chars(lab, lab + 16, labels[i]);
            off[i + 1] = (info_off[i + 1] - info_off[i]) + 1 + nums[i].len[0] + 1 + nums[i].len[1] + 1 + (r.ptr - lab) + 1 + (hi5 - lo5) + 1;
        }
    });
    off[0] = 0;
    for (int64_t i = 0; i < n; ++i) off[i + 1] += off[i];
    *out_bytes = off[n];
    DSP_REQUIRE(off[n] <= out_cap, DSP_ERR_NOMEM, "dsp_format_calls: output needs %lld bytes, buffer has %lld",
                (long long)off[n], (long long)out_cap);
    parallel_for(n, nthreads, [&](int64_t a, int64_t b) {
        for (int64_t i = a; i < b; ++i) {
            char* o = out + off[i];
            const int64_t il = info_off[i + 1] - info_off[i];
            memcpy(o, info_text + info_off[i], (size_t)il); o += il; *o++ = '\t';
            memcpy(o, nums[i].s[0], nums[i].len[0]); o += nums[i].len[0]; *o++ = '\t';
            memcpy(o, nums[i].s[1], nums[i].len[1]); o += nums[i].len[1]; *o++ = '\t';
            o = std::to_chars(o, o + 16, labels[i]).ptr; *o++ = '\t';
            for (int t = lo5; t < hi5; ++t) { const int code = (int)kmer[i * seq_len + t]; *o++ = BASES[code & 15]; }
            *o++ = '\n';
        }
    });
    return DSP_OK;
}

// The six sample-info columns of sites extracted on the device (call_modifications.py:312:
// chrom \t pos \t alignstrand \t pos_in_strand \t readname \t strand), packed back to back in the layout
// dsp_parse_features produces and dsp_format_calls consumes.  Per READ: chrom and readname as packed text +
// offsets, alignstrand and strand as one character each; per SITE: read index, pos, pos_in_strand.
int dsp_format_sampleinfo(const char* chrom_text, const int64_t* chrom_off, const char* name_text, const int64_t* name_off,
                          const char* alignstrand, const char* strand,
                          const int32_t* site_read, const int64_t* pos, const int64_t* pos_in_strand, int64_t n,
                          char* info_text, int64_t info_cap, int64_t* info_off, int32_t nthreads) {
    DSP_REQUIRE(n >= 0 && info_off, DSP_ERR_INVALID, "dsp_format_sampleinfo: bad argument");
    info_off[0] = 0;
    if (n == 0) return DSP_OK;
    DSP_REQUIRE(chrom_text && chrom_off && name_text && name_off && alignstrand && strand && site_read && pos &&
                pos_in_strand && info_text, DSP_ERR_INVALID, "dsp_format_sampleinfo: null argument");
    auto digits = [](int64_t v) { int d = v < 0 ? 2 : 1; uint64_t u = v < 0 ? (uint64_t)(-(v + 1)) + 1 : (uint64_t)v; while (u >= 10) { u /= 10; ++d; } return d; };
    parallel_for(n, nthreads, [&](int64_t a, int64_t b) {
        for (int64_t i = a; i < b; ++i) {
            const int32_t r = site_read[i];
            info_off[i + 1] = (chrom_off[r + 1] - chrom_off[r]) + 1 + digits(pos[i]) + 1 + 1 + 1 + digits(pos_in_strand[i]) + 1 +
                              (name_off[r + 1] - name_off[r]) + 1 + 1;
        }
    });
    for (int64_t i = 0; i < n; ++i) info_off[i + 1] += info_off[i];
    DSP_REQUIRE(info_off[n] <= info_cap, DSP_ERR_NOMEM, "dsp_format_sampleinfo: output needs %lld bytes, buffer has %lld",
                (long long)info_off[n], (long long)info_cap);
    parallel_for(n, nthreads, [&](int64_t a, int64_t b) {
        for (int64_t i = a; i < b; ++i) {
            const int32_t r = site_read[i];
            char* o = info_text + info_off[i];
            const int64_t cl = chrom_off[r + 1] - chrom_off[r], nl = name_off[r + 1] - name_off[r];
            memcpy(o, chrom_text + chrom_off[r], (size_t)cl); o += cl; *o++ = '\t';
            o = std::to_chars(o, o + 24, (long long)pos[i]).ptr; *o++ = '\t';
            *o++ = alignstrand[r]; *o++ = '\t';
            o = std::to_chars(o, o + 24, (long long)pos_in_strand[i]).ptr; *o++ = '\t';
            memcpy(o, name_text + name_off[r], (size_t)nl); o += nl; *o++ = '\t';
            *o++ = strand[r];
        }
    });
    return DSP_OK;
}

// One feature-file line per site, as _features_to_str writes it (extract_features.py:381-395):
//   sampleinfo \t k_mer \t means \t stds \t lens \t signals \t label \n
// means / stds / signals are the float64 values dsp_extract_features_f64 produced with round_stats = 1,
// printed like str(numpy.float64); lens as integers; values within a column joined by ',', the rows of the
// signal rectangle by ';'.  kmer_letters is (n, seq_len) ASCII.  Two passes: format every site into a
// thread-local arena to learn its length, then copy into place.
int dsp_format_features(const char* info_text, const int64_t* info_off, const uint8_t* kmer_letters,
                        const double* means, const double* stds, const double* lens, const double* signals,
                        int32_t methy_label, int64_t n, int32_t seq_len, int32_t signal_len,
                        char* out, int64_t out_cap, int64_t* out_bytes, int32_t nthreads) {
    DSP_REQUIRE(n >= 0 && out_bytes, DSP_ERR_INVALID, "dsp_format_features: bad argument");
    *out_bytes = 0;
    if (n == 0) return DSP_OK;
    DSP_REQUIRE(info_text && info_off && kmer_letters && means && stds && lens && signals && out && seq_len > 0 && signal_len > 0,
                DSP_ERR_INVALID, "dsp_format_features: null argument");
    const int T = seq_len, S = signal_len;
    const size_t worst = (size_t)T * (size_t)(2 * 26 + 22 + S * 26) + 64;   // per site, without the sample info
    char lab[16];
    const int lab_len = (int)(std::to_chars(lab, lab + 16, methy_label).ptr - lab);
    auto one = [&](int64_t i, char* o) -> char* {
        const int64_t il = info_off[i + 1] - info_off[i];
        memcpy(o, info_text + info_off[i], (size_t)il); o += il; *o++ = '\t';
        memcpy(o, kmer_letters + i * T, (size_t)T); o += T; *o++ = '\t';
        for (int t = 0; t < T; ++t) { o += format_f64(means[i * T + t], o); *o++ = t + 1 < T ? ',' : '\t'; }
        for (int t = 0; t < T; ++t) { o += format_f64(stds[i * T + t], o); *o++ = t + 1 < T ? ',' : '\t'; }
        for (int t = 0; t < T; ++t) { o = std::to_chars(o, o + 24, (long long)lens[i * T + t]).ptr; *o++ = t + 1 < T ? ',' : '\t'; }
        for (int t = 0; t < T; ++t) {
            const double* row = signals + ((size_t)i * T + t) * S;
            for (int k = 0; k < S; ++k) { o += format_f64(row[k], o); *o++ = k + 1 < S ? ',' : (t + 1 < T ? ';' : '\t'); }
        }
        memcpy(o, lab, (size_t)lab_len); o += lab_len;
        *o++ = '\n';
        return o;
    };
    // one formatting pass: every worker appends the lines of its contiguous range to its own buffer, the
    // buffers are then copied into place in range order
    struct Part { int64_t a = 0, b = 0; std::vector<char> text; };
    const int64_t nparts = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)std::max(1, nthreads) * 4));
    std::vector<Part> parts((size_t)nparts);
    for (int64_t q = 0; q < nparts; ++q) { parts[q].a = n * q / nparts; parts[q].b = n * (q + 1) / nparts; }
    parallel_for(nparts, nthreads, [&](int64_t qa, int64_t qb) {
        for (int64_t q = qa; q < qb; ++q) {
            Part& pt = parts[q];
            size_t used = 0;
            pt.text.resize((size_t)(pt.b - pt.a) * (worst / 3) + worst + 4096);
            for (int64_t i = pt.a; i < pt.b; ++i) {
                const size_t need = worst + (size_t)(info_off[i + 1] - info_off[i]);
                if (pt.text.size() - used < need) pt.text.resize(pt.text.size() * 2 + need);
                used = (size_t)(one(i, pt.text.data() + used) - pt.text.data());
            }
            pt.text.resize(used);
        }
    }, 2);
    std::vector<int64_t> off((size_t)nparts + 1);
    off[0] = 0;
    for (int64_t q = 0; q < nparts; ++q) off[q + 1] = off[q] + (int64_t)parts[q].text.size();
    *out_bytes = off[nparts];
    DSP_REQUIRE(off[nparts] <= out_cap, DSP_ERR_NOMEM, "dsp_format_features: output needs %lld bytes, buffer has %lld",
                (long long)off[nparts], (long long)out_cap);
    parallel_for(nparts, nthreads, [&](int64_t qa, int64_t qb) {
        for (int64_t q = qa; q < qb; ++q) memcpy(out + off[q], parts[q].text.data(), parts[q].text.size());
    }, 2);
    return DSP_OK;
}

// positions of the first (up to) ten tabs of [b, e), eight bytes at a time: f[k] = start of field k; returns the
// number of fields seen (capped at 11)
inline int split_tabs(const char* b, const char* e, const char** f) {
    int nf = 1;
    f[0] = b;
    const char* p = b;
    const uint64_t TAB = 0x0909090909090909ull, LO = 0x0101010101010101ull, HI = 0x8080808080808080ull;
    while (e - p >= 8 && nf <= 10) {
        uint64_t w;
        memcpy(&w, p, 8);
        const uint64_t x = w ^ TAB;
        uint64_t hit = (x - LO) & ~x & HI;                       // 0x80 in every byte that was a tab
        while (hit && nf <= 10) {
            const int byte = __builtin_ctzll(hit) >> 3;
            f[nf++] = p + byte + 1;
            hit &= hit - 1;
        }
        p += 8;
    }
    for (; p < e && nf <= 10; ++p)
        if (*p == '\t') f[nf++] = p + 1;
    return nf;
}

// float() of the spelling call_mods prints for its probabilities -- one integer digit, a point, one to six
// fraction digits ("0.533558", "1.0") -- as exact integer / exact power of ten (one correctly rounded operation,
// Clinger's fast path); anything else goes through parse_double
inline bool parse_prob_fast(const char* b, const char* e, double* out) {
    static const double P10[7] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6};
    const int len = (int)(e - b);
    if (len >= 3 && len <= 8 && b[1] == '.') {
        unsigned mant = (unsigned)(b[0] - '0');
        unsigned bad = mant > 9;
        for (int k = 2; k < len; ++k) {
            const unsigned c = (unsigned)(b[k] - '0');
            bad |= c > 9;
            mant = mant * 10 + c;
        }
        if (!bad) {
            *out = (double)mant / P10[len - 2];
            return true;
        }
    }
    return parse_double(b, e, out);
}

// int() of a plain decimal field ([+-]digits, at most 18 of them); anything else goes through parse_int
inline bool parse_int_fast(const char* b, const char* e, long long* out) {
    const char* p = b;
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
    const int nd = (int)(e - p);
    if (nd >= 1 && nd <= 18) {
        unsigned long long v = 0;
        for (; p < e; ++p) {
            const unsigned c = (unsigned)(*p - '0');
            if (c > 9) return parse_int(b, e, out);
            v = v * 10 + c;
        }
        *out = neg ? -(long long)v : (long long)v;
        return true;
    }
    return parse_int(b, e, out);
}

// ---- call_mods lines -> columns (the input of call_freq) ----------------------------------------------------
// ModRecord.__init__ (utils/txt_formater.py:8-21): words = line.strip().split("\t"); chromosome = words[0],
// pos = int(words[1]), strand = words[2], pos_in_strand = int(words[3]), prob_0 = float(words[6]),
// prob_1 = float(words[7]), called_label = int(words[8]), k_mer = words[9].  Columns beyond the tenth are ignored,
// fewer than ten are an error (IndexError in the reference).  Chromosome names are interned: chrom_code[i]
// indexes the '\n'-separated list written to `names` (order of first appearance per worker, then merged).
// strand and k_mer are copied into fixed-width, zero-padded cells (STRAND_W / KMER_W bytes); a longer value
// returns DSP_ERR_UNSUPPORTED so that the caller can take its general path.  With max_records == 0 the call only
// counts the lines of the block (*n_records).
int dsp_parse_calls(const char* text, int64_t nbytes, int64_t max_records,
                    int32_t* chrom_code, int64_t* pos, char* strand, int64_t* pos_in_strand,
                    double* p0, double* p1, int32_t* label, char* kmer,
                    char* names, int64_t names_cap, int64_t* names_bytes, int32_t* n_names,
                    int64_t* n_records, int32_t nthreads) {
    constexpr int STRAND_W = 4, KMER_W = 24;
    DSP_REQUIRE(text && n_records && nbytes >= 0 && max_records >= 0, DSP_ERR_INVALID, "dsp_parse_calls: bad argument");
    *n_records = 0;
    // lines = text.split('\n') with each line strip()-ed; blank lines die in the reference (IndexError) except a trailing one.
    // trailing white space is not a line; everything before it is split at the newlines
    int64_t eff_end = nbytes;
    while (eff_end > 0 && is_space(text[eff_end - 1])) --eff_end;
    // newline positions below eff_end, found by all threads over byte ranges and kept per range (no merged copy: line i's
    // newline is part[r][i - prefix[r]]); the counting call only counts
    const int PN = nthreads > 1 && nbytes >= (1 << 20) ? nthreads : 1;
    std::vector<std::vector<int64_t>> part((size_t)PN);
    std::vector<int64_t> part_count((size_t)PN, 0);
    const bool count_only = max_records == 0;
    parallel_for(PN, PN, [&](int64_t a, int64_t b) {
        for (int64_t r = a; r < b; ++r) {
            const int64_t lo = eff_end * r / PN, hi = eff_end * (r + 1) / PN;
            int64_t cnt = 0;
            if (!count_only) part[(size_t)r].reserve((size_t)((hi - lo) / 48 + 16));
            for (int64_t q = lo; q < hi;) {
                const char* nl = (const char*)memchr(text + q, '\n', (size_t)(hi - q));
                if (!nl) break;
                if (!count_only) part[(size_t)r].push_back(nl - text);
                ++cnt;
                q = (nl - text) + 1;
            }
            part_count[(size_t)r] = cnt;
        }
    }, 2);
    std::vector<int64_t> prefix((size_t)PN + 1, 0);
    for (int r = 0; r < PN; ++r) prefix[(size_t)r + 1] = prefix[(size_t)r] + part_count[(size_t)r];
    const int64_t n_nl = prefix[(size_t)PN];
    const int64_t n = eff_end > 0 ? n_nl + 1 : 0;
    *n_records = n;
    if (count_only) return DSP_OK;
    DSP_REQUIRE(n <= max_records, DSP_ERR_NOMEM, "dsp_parse_calls: %lld records, buffers hold %lld", (long long)n, (long long)max_records);
    DSP_REQUIRE(chrom_code && pos && strand && pos_in_strand && p0 && p1 && label && kmer && names && names_bytes && n_names,
                DSP_ERR_INVALID, "dsp_parse_calls: null argument");
    // monotone cursor over the per-range newline lists
    struct Cursor {
        const std::vector<std::vector<int64_t>>& part; const std::vector<int64_t>& prefix; size_t r = 0;
        int64_t at(int64_t i) {                                      // i ascending between calls
            while (i >= prefix[r + 1]) ++r;
            return part[r][(size_t)(i - prefix[r])];
        }
    };
    const int P = n >= 4096 && nthreads > 1 ? nthreads : 1;
    struct Local {
        std::unordered_map<std::string_view, int32_t> ids; std::vector<std::string_view> names; int64_t a = 0, b = 0;
        // names of up to 8 bytes (chr1 ... chrUn_xyz are longer and take the map): 256-slot table keyed by the bytes + length
        uint64_t skey[256]; int32_t scode[256]; int sused = 0;
        Local() { for (int i = 0; i < 256; ++i) { skey[i] = ~0ull; scode[i] = -1; } }
    };
    std::vector<Local> locals((size_t)P);
    std::atomic<int64_t> bad_line{-1};
    std::atomic<int> bad_kind{0};                  // 1 malformed, 2 field too wide
    parallel_for(P, P, [&](int64_t ra, int64_t rb) {
        for (int64_t r = ra; r < rb; ++r) {
            Local& L = locals[(size_t)r];
            L.a = n * r / P; L.b = n * (r + 1) / P;
            std::string_view last_name;
            int32_t last_code = -1;
            Cursor cur{part, prefix};
            if (L.a < L.b && L.a > 0) { size_t r0 = 0; while (L.a - 1 >= prefix[r0 + 1]) ++r0; cur.r = r0; }
            int64_t prev_nl = (L.a > 0 && L.a < L.b) ? cur.at(L.a - 1) : -1;
            for (int64_t i = L.a; i < L.b; ++i) {
                const int64_t this_nl = i < n_nl ? cur.at(i) : eff_end;
                int64_t lb = prev_nl + 1, le = this_nl;
                prev_nl = this_nl;
                while (lb < le && is_space(text[lb])) ++lb;                  // line.strip()
                while (le > lb && is_space(text[le - 1])) --le;
                const char* b = text + lb;
                const char* e = text + le;
                const char* f[12];
                const int nf = split_tabs(b, e, f);
                int kind = 0;
                if (nf < 10) kind = 1;                                       // also an empty line inside the file
                else {
                    const char* fe[10];
                    for (int c = 0; c < 10; ++c) fe[c] = (c + 1 < nf) ? f[c + 1] - 1 : e;
                    long long v;
                    if (!parse_int_fast(f[1], fe[1], &v)) kind = 1; else pos[i] = v;
                    if (!kind && !parse_int_fast(f[3], fe[3], &v)) kind = 1; else if (!kind) pos_in_strand[i] = v;
                    if (!kind && !parse_prob_fast(f[6], fe[6], &p0[i])) kind = 1;
                    if (!kind && !parse_prob_fast(f[7], fe[7], &p1[i])) kind = 1;
                    if (!kind && !parse_int_fast(f[8], fe[8], &v)) kind = 1; else if (!kind) label[i] = (int32_t)v;
                    const int64_t sl = fe[2] - f[2], kl = fe[9] - f[9];
                    if (!kind && (sl > STRAND_W || kl > KMER_W)) kind = 2;
                    if (!kind) {
                        uint32_t sw = 0;                                         // zero-padded cells, built in registers
                        if (sl == 1) sw = (uint8_t)f[2][0]; else memcpy(&sw, f[2], (size_t)sl);
                        memcpy(strand + i * STRAND_W, &sw, STRAND_W);
                        uint64_t kw[3] = {0, 0, 0};
                        if (kl <= 8 && f[9] + 8 <= text + nbytes) {             // the usual 5-mer: one masked 8-byte load
                            uint64_t w;
                            memcpy(&w, f[9], 8);
                            kw[0] = kl == 8 ? w : (w & ((1ull << (8 * kl)) - 1ull));
                        } else memcpy(kw, f[9], (size_t)kl);
                        memcpy(kmer + i * KMER_W, kw, KMER_W);
                        const std::string_view name(f[0], (size_t)(fe[0] - f[0]));
                        if (name != last_name) {                                 // calls come in runs of one chromosome
                            int32_t code = -1;
                            uint64_t key = 0;
                            const bool small = name.size() <= 7 && L.sused < 192;
                            uint32_t slot = 0;
                            if (small) {
                                memcpy(&key, name.data(), name.size());
                                key |= (uint64_t)name.size() << 56;
                                slot = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 56);
                                while (L.skey[slot] != ~0ull && L.skey[slot] != key) slot = (slot + 1) & 255u;
                                if (L.skey[slot] == key) code = L.scode[slot];
                            }
                            if (code < 0) {
                                auto it = L.ids.find(name);
                                if (it == L.ids.end()) { it = L.ids.emplace(name, (int32_t)L.names.size()).first; L.names.push_back(name); }
                                code = it->second;
                                if (small) { L.skey[slot] = key; L.scode[slot] = code; ++L.sused; }
                            }
                            last_name = name; last_code = code;
                        }
                        chrom_code[i] = last_code;
                    }
                }
                if (kind) {
                    int64_t expect = -1;
                    if (bad_line.compare_exchange_strong(expect, i)) bad_kind = kind;
                    return;
                }
            }
        }
    }, 2);
    if (bad_line >= 0) {
        if (bad_kind == 2) { set_error("call_mods file: line %lld has a strand or k-mer column wider than the fixed cells", (long long)bad_line.load() + 1); return 5; }
        set_error("call_mods file: line %lld is malformed (ten tab-separated columns: ints in 2, 4, 9, floats in 7, 8)", (long long)bad_line.load() + 1);
        return DSP_ERR_INVALID;
    }
    // merge the workers' name tables (worker order, first appearance inside a worker) and renumber their codes
    std::unordered_map<std::string_view, int32_t> global;
    std::vector<std::string_view> gnames;
    std::vector<std::vector<int32_t>> remap((size_t)P);
    for (int r = 0; r < P; ++r) {
        for (const std::string_view& nm : locals[(size_t)r].names) {
            auto it = global.find(nm);
            if (it == global.end()) { it = global.emplace(nm, (int32_t)gnames.size()).first; gnames.push_back(nm); }
            remap[(size_t)r].push_back(it->second);
        }
    }
    parallel_for(P, P, [&](int64_t ra, int64_t rb) {
        for (int64_t r = ra; r < rb; ++r) {
            const std::vector<int32_t>& m = remap[(size_t)r];
            for (int64_t i = locals[(size_t)r].a; i < locals[(size_t)r].b; ++i) chrom_code[i] = m[(size_t)chrom_code[i]];
        }
    }, 2);
    int64_t used = 0;
    for (const std::string_view& nm : gnames) used += (int64_t)nm.size() + 1;
    *names_bytes = used;
    *n_names = (int32_t)gnames.size();
    DSP_REQUIRE(used <= names_cap, DSP_ERR_NOMEM, "dsp_parse_calls: chromosome names need %lld bytes, buffer has %lld", (long long)used, (long long)names_cap);
    char* o = names;
    for (const std::string_view& nm : gnames) { memcpy(o, nm.data(), nm.size()); o += nm.size(); *o++ = '\n'; }
    return DSP_OK;
}

// write_sitekey2stats (call_mods_freq.py:87-120) for n sites in output order, HOST pointers.  chrom / strand / kmer:
// the n strings of each column joined by '\n' (no trailing newline needed).  Rows with coverage 0 are skipped (:104).
//   tsv: "%s\t%d\t%s\t%d\t%.3f\t%.3f\t%d\t%d\t%d\t%.4f\t%s\n"  (:112-118), rmet = float(met) / coverage
//   bed: chrom, pos, pos+1, ".", cov, strand, pos, pos+1, "0,0,0", cov, int(round(rmet * 100 + 0.001, 0))  (:106-110)
int dsp_format_freq(const char* chrom_text, const char* strand_text, const char* kmer_text,
                    const int64_t* pos, const int64_t* pos_in_strand, const double* prob_0, const double* prob_1,
                    const int32_t* met, const int32_t* unmet, const int32_t* coverage, int64_t n, int32_t is_bed,
                    char* out, int64_t out_cap, int64_t* out_bytes, int32_t nthreads) {
    DSP_REQUIRE(n >= 0 && out_bytes, DSP_ERR_INVALID, "dsp_format_freq: bad argument");
    *out_bytes = 0;
    if (n == 0) return DSP_OK;
    DSP_REQUIRE(chrom_text && strand_text && kmer_text && pos && pos_in_strand && prob_0 && prob_1 && met && unmet && coverage && out,
                DSP_ERR_INVALID, "dsp_format_freq: null argument");
    auto split = [n](const char* t, std::vector<const char*>& b, std::vector<int32_t>& l) -> bool {
        b.resize((size_t)n); l.resize((size_t)n);
        const char* p = t;
        for (int64_t i = 0; i < n; ++i) {
            const char* e = (i + 1 < n) ? strchr(p, '\n') : p + strlen(p);
            if (e == nullptr) return false;
            if (i + 1 == n && e > p && e[-1] == '\n') --e;
            b[(size_t)i] = p; l[(size_t)i] = (int32_t)(e - p);
            p = e + 1;
        }
        return true;
    };
    std::vector<const char*> cb, sb, kb;
    std::vector<int32_t> cl, sl, kl;
    DSP_REQUIRE(split(chrom_text, cb, cl) && split(strand_text, sb, sl) && split(kmer_text, kb, kl), DSP_ERR_INVALID,
                "dsp_format_freq: a text column has fewer than n entries");
    struct Part { int64_t a = 0, b = 0; std::vector<char> text; };
    const int64_t nparts = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)std::max(1, nthreads) * 4));
    std::vector<Part> parts((size_t)nparts);
    for (int64_t q = 0; q < nparts; ++q) { parts[q].a = n * q / nparts; parts[q].b = n * (q + 1) / nparts; }
    parallel_for(nparts, nthreads, [&](int64_t qa, int64_t qb) {
        for (int64_t q = qa; q < qb; ++q) {
            Part& pt = parts[q];
            size_t used = 0;
            for (int64_t i = pt.a; i < pt.b; ++i) {
                if (coverage[i] <= 0) continue;
                const size_t need = (size_t)cl[i] + sl[i] + kl[i] + 1024;       // %.3f of a double can be long
                if (pt.text.size() - used < need) pt.text.resize(pt.text.size() * 2 + need + 4096);
                char* o = pt.text.data() + used;
                const double rmet = (double)met[i] / (double)coverage[i];
                memcpy(o, cb[i], (size_t)cl[i]); o += cl[i]; *o++ = '\t';
                if (is_bed) {
                    o = std::to_chars(o, o + 24, (long long)pos[i]).ptr; *o++ = '\t';
                    o = std::to_chars(o, o + 24, (long long)pos[i] + 1).ptr; *o++ = '\t';
                    *o++ = '.'; *o++ = '\t';
                    o = std::to_chars(o, o + 24, coverage[i]).ptr; *o++ = '\t';
                    memcpy(o, sb[i], (size_t)sl[i]); o += sl[i]; *o++ = '\t';
                    o = std::to_chars(o, o + 24, (long long)pos[i]).ptr; *o++ = '\t';
                    o = std::to_chars(o, o + 24, (long long)pos[i] + 1).ptr; *o++ = '\t';
                    memcpy(o, "0,0,0\t", 6); o += 6;
                    o = std::to_chars(o, o + 24, coverage[i]).ptr; *o++ = '\t';
                    o = std::to_chars(o, o + 24, (long long)std::nearbyint(rmet * 100 + 0.001)).ptr;
                } else {
                    o = std::to_chars(o, o + 24, (long long)pos[i]).ptr; *o++ = '\t';
                    memcpy(o, sb[i], (size_t)sl[i]); o += sl[i]; *o++ = '\t';
                    o = std::to_chars(o, o + 24, (long long)pos_in_strand[i]).ptr; *o++ = '\t';
                    // "%.3f" / "%.4f": std::to_chars with a precision prints what printf does in the C locale
                    // (correctly rounded), without depending on the process locale
                    o = std::to_chars(o, o + 400, prob_0[i], std::chars_format::fixed, 3).ptr; *o++ = '\t';
                    o = std::to_chars(o, o + 400, prob_1[i], std::chars_format::fixed, 3).ptr; *o++ = '\t';
                    o = std::to_chars(o, o + 24, met[i]).ptr; *o++ = '\t';
                    o = std::to_chars(o, o + 24, unmet[i]).ptr; *o++ = '\t';
                    o = std::to_chars(o, o + 24, coverage[i]).ptr; *o++ = '\t';
                    o = std::to_chars(o, o + 400, rmet, std::chars_format::fixed, 4).ptr; *o++ = '\t';
                    memcpy(o, kb[i], (size_t)kl[i]); o += kl[i];
                }
                *o++ = '\n';
                used = (size_t)(o - pt.text.data());
            }
            pt.text.resize(used);
        }
    }, 2);
    std::vector<int64_t> off((size_t)nparts + 1);
    off[0] = 0;
    for (int64_t q = 0; q < nparts; ++q) off[q + 1] = off[q] + (int64_t)parts[q].text.size();
    *out_bytes = off[nparts];
    DSP_REQUIRE(off[nparts] <= out_cap, DSP_ERR_NOMEM, "dsp_format_freq: output needs %lld bytes, buffer has %lld",
                (long long)off[nparts], (long long)out_cap);
    parallel_for(nparts, nthreads, [&](int64_t qa, int64_t qb) {
        for (int64_t q = qa; q < qb; ++q) memcpy(out + off[q], parts[q].text.data(), parts[q].text.size());
    }, 2);
    return DSP_OK;
}

}  // extern "C"
