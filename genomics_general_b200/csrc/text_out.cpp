// Host-side row formatter for freq.py's output (multi-threaded, no CUDA): the per-site rows are text, 10^7-10^8 of them,
// and writing them from Python costs far more than computing them on the GPU.
//
// Replaces the row assembly of freqs_wrapper (freq.py:100-111): "<scaffold>\t<position>\t<col>\t<col>...\n" where a column is
//   mode 0: "cA,cC,cG,cT" of a population (np.column_stack of ",".join(row) strings, freq.py:52-58)
//   mode 1: a float64 printed the way numpy's astype(str) prints it (shortest round-trip repr: "0.25", "1.0", "nan")
//   mode 2: an integer (--asCounts)
#include <stdint.h>
#include <string.h>

#include <charconv>
#include <cmath>
#include <thread>
#include <vector>

#include "../../include/pgwin.h"

void pg_set_error(const char* fmt, ...);

namespace {

inline char* put_uint(char* p, uint64_t v) {
    char tmp[24];
    int n = 0;
    do {
        tmp[n++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}
inline char* put_int(char* p, int64_t v) {
    if (v < 0) {
        *p++ = '-';
        return put_uint(p, (uint64_t)(-(v + 1)) + 1u);
    }
    return put_uint(p, (uint64_t)v);
}
// Python's repr(float) / numpy's float64 -> str: shortest digits that round-trip; fixed notation for 1e-4 <= |x| < 1e16
inline char* put_double(char* p, double v) {
    if (std::isnan(v)) {
        memcpy(p, "nan", 3);
        return p + 3;
    }
    if (std::isinf(v)) {
        if (v < 0) *p++ = '-';
        memcpy(p, "inf", 3);
        return p + 3;
    }
    if (v == 0.0) {
        if (std::signbit(v)) *p++ = '-';
        memcpy(p, "0.0", 3);
        return p + 3;
    }
    const double a = std::fabs(v);
    if (a >= 1e-4 && a < 1e16) {
        char* e = std::to_chars(p, p + 40, v, std::chars_format::fixed).ptr;
        bool dot = false;
        for (char* q = p; q < e; ++q) dot |= (*q == '.');
        if (!dot) {
            *e++ = '.';
            *e++ = '0';
        }
        return e;
    }
    return std::to_chars(p, p + 40, v, std::chars_format::scientific).ptr;     // "1e-05", as Python prints it
}

}  // namespace

extern "C" int pg_format_freq_rows(int32_t mode, const void* data, int64_t n, int32_t P, const int32_t* pos,
                                   const int32_t* scaf_id, const char* const* scaf_names, const uint8_t* keep, char* out,
                                   size_t seg_cap, int32_t n_threads, size_t* seg_len) {
    if (!data || !pos || !scaf_id || !scaf_names || !out || !seg_len || mode < 0 || mode > 2 || P < 1 || n < 0) {
        pg_set_error("pg_format_freq_rows: bad argument");
        return 1;
    }
    if (n_threads < 1) n_threads = 1;
    std::vector<int> overflow((size_t)n_threads, 0);
    auto work = [&](int t) {
        const int64_t a = n * t / n_threads, b = n * (t + 1) / n_threads;
        char* base = out + (size_t)t * seg_cap;
        char* p = base;
        char* lim = base + seg_cap;
        for (int64_t i = a; i < b; ++i) {
            if (keep && !keep[i]) continue;
            const char* sc = scaf_names[scaf_id[i]];
            const size_t sl = strlen(sc);
            if ((size_t)(lim - p) < sl + 16 + (size_t)P * 48) {
                overflow[(size_t)t] = 1;
                break;
            }
            memcpy(p, sc, sl);
            p += sl;
            *p++ = '\t';
            p = put_int(p, pos[i]);
            for (int x = 0; x < P; ++x) {
                *p++ = '\t';
                if (mode == 0) {
                    const uint16_t* c = (const uint16_t*)data + ((size_t)i * P + x) * 4;
                    p = put_uint(p, c[0]);
                    *p++ = ',';
                    p = put_uint(p, c[1]);
                    *p++ = ',';
                    p = put_uint(p, c[2]);
                    *p++ = ',';
                    p = put_uint(p, c[3]);
                } else if (mode == 1) {
                    p = put_double(p, ((const double*)data)[(size_t)i * P + x]);
                } else {
                    p = put_int(p, (int64_t)((const double*)data)[(size_t)i * P + x]);
                }
            }
            *p++ = '\n';
        }
        seg_len[t] = (size_t)(p - base);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    for (int v : overflow)
        if (v) {
            pg_set_error("pg_format_freq_rows: segment buffer too small");
            return 1;
        }
    return 0;
}

// Rows of a float64 matrix as text: "<prefix[r]><v0><sep><v1>...\n", numbers printed like numpy's float64 -> str.
// Replaces the " ".join(row) over ndarray.round(roundTo).astype(str) of makeDistMat*String (genomics.py:2288-2306) for
// the n x n matrices distMat.py writes per window (the caller rounds; this prints).
extern "C" int pg_format_matrix_rows(const double* v, int64_t rows, int32_t cols, int32_t sep, const char* const* prefix,
                                     char* out, size_t seg_cap, int32_t n_threads, size_t* seg_len) {
    if (!v || !out || !seg_len || rows < 0 || cols < 1) {
        pg_set_error("pg_format_matrix_rows: bad argument");
        return 1;
    }
    if (n_threads < 1) n_threads = 1;
    std::vector<int> overflow((size_t)n_threads, 0);
    auto work = [&](int t) {
        const int64_t a = rows * t / n_threads, b = rows * (t + 1) / n_threads;
        char* base = out + (size_t)t * seg_cap;
        char* p = base;
        char* lim = base + seg_cap;
        for (int64_t r = a; r < b; ++r) {
            const size_t pl = prefix ? strlen(prefix[r]) : 0;
            if ((size_t)(lim - p) < pl + (size_t)cols * 42 + 2) {
                overflow[(size_t)t] = 1;
                break;
            }
            if (pl) {
                memcpy(p, prefix[r], pl);
                p += pl;
            }
            for (int c = 0; c < cols; ++c) {
                if (c) *p++ = (char)sep;
                p = put_double(p, v[(size_t)r * cols + c]);
            }
            *p++ = '\n';
        }
        seg_len[t] = (size_t)(p - base);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    for (int f : overflow)
        if (f) {
            pg_set_error("pg_format_matrix_rows: segment buffer too small");
            return 1;
        }
    return 0;
}
