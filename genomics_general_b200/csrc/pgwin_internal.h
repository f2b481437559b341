// Internal declarations shared by the translation units of libpgwin.so (not part of the C-ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/pgwin.h"

#define PG_OK 0
#define PG_ERR 1

void pg_set_error(const char* fmt, ...);

#define PG_CUDA(call)                                                                                  \
    do {                                                                                               \
        cudaError_t _e = (call);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            pg_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));  \
            return PG_ERR;                                                                             \
        }                                                                                              \
    } while (0)

#define PG_CHECK(cond, ...)                 \
    do {                                    \
        if (!(cond)) {                      \
            pg_set_error(__VA_ARGS__);      \
            return PG_ERR;                  \
        }                                   \
    } while (0)

#define PG_TRY(expr)                 \
    do {                             \
        int _r = (expr);             \
        if (_r != PG_OK) return _r;  \
    } while (0)

struct PgTiming {
    char name[32];
    cudaEvent_t start, stop;
    int launches;
};

// grow-only device scratch buffer
struct PgBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
};

static constexpr int PG_MAX_K1_POPS = 8;     // K1 keeps per-pop counts in registers: P <= 8
static constexpr int PG_MAX_POPS = 64;       // K2 block epilogue

// K1 launch geometry for one (S, H): see DESIGN.md "K1 tiling".
struct K1Plan {
    int pitch;        // bytes per site row on the device = 16 * chunks, chunks odd (bank-conflict-free LDS.128)
    int chunks;       // 16-byte chunks per row
    int G;            // lanes cooperating on one site (power of two, <= 32)
    int I;            // sites per lane per tile
    int wpt;          // consumer warps per tile (a "team"); nw / wpt teams work on different tiles
    int nw;           // consumer warps per CTA (8 or 12)
    int T;            // sites per tile = (32 * wpt / G) * I
    int stages;       // TMA ring depth
    int tile_bytes;   // T * pitch
    int smem_bytes;   // dynamic shared memory of the kernel
    int64_t num_tiles;
    int ctas;         // persistent CTAs, each owning a contiguous tile range
};
K1Plan pg_make_k1_plan(int64_t S, int H, int sm_count, int table_bytes, int nw = 8, int force_G = 0);
int pg_pitch_for(int H);

struct pg_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    // genotype matrix
    int8_t* d_geno = nullptr;
    int32_t* d_pos = nullptr;
    int64_t S = 0;
    int32_t H = 0;
    int32_t pitch = 0;
    size_t geno_cap = 0, pos_cap = 0;
    // populations
    int32_t P = 0;
    std::vector<int32_t> hap_pop;
    // windows (host copies) + segments
    int64_t W = 0;
    std::vector<int64_t> win_lo, win_hi;
    std::vector<int64_t> brk;                 // segment breakpoints, brk[0]=0 .. brk[nseg]=S
    std::vector<int32_t> win_seg_lo, win_seg_hi;
    // timings of the last statistics call
    std::vector<PgTiming> timings;
    std::vector<cudaEvent_t> event_pool;
    size_t events_used = 0;
    int64_t launches = 0;
    // scratch
    PgBuf tables, part, segmeta, winmeta, out_d, out_i, planes, planes2, pairs, misc, misc2, misc3, misc4, misc5;
    PgBuf text, starts, meta;                 // device-side text ingest (ingest.cu)
    int64_t ingest_sites = -1;
    void* h_text[2] = {nullptr, nullptr};     // pinned staging of the text
    cudaEvent_t h_text_free[2] = {nullptr, nullptr};
    // upload pipeline: copy stream + two staging buffers
    cudaStream_t copy_stream = nullptr;
    PgBuf stage[2];
    cudaEvent_t stage_full[2] = {nullptr, nullptr}, stage_free[2] = {nullptr, nullptr};
    bool want_freq = false;                   // carry the popFreq counters in the popgen site pass
    uint64_t epoch = 1;                       // bumped by every change of data shape / populations / windows
    void* k1_cache[3] = {nullptr, nullptr, nullptr};   // cached launch state (popgen, abba, fourpop) — owned by k1.cu
    std::vector<unsigned long long> h_rec;    // host copy of the per-window records
    // native NCCL gather (nccl_gather.cu)
    void* nccl_comm = nullptr;
    int nccl_ranks = 1, nccl_rank = 0;
    PgBuf gather;
    size_t gather_words = 0;
    // pipelined gather (pg_popgen_gather_begin / _end): two record tables, exchange + read-back on a side stream
    cudaStream_t gather_stream = nullptr;
    cudaEvent_t g_rec[2] = {nullptr, nullptr}, g_done[2] = {nullptr, nullptr};
    PgBuf gslot[2];
    void* gslot_host[2] = {nullptr, nullptr};
    size_t gslot_host_cap[2] = {0, 0}, gslot_words[2] = {0, 0};
    int64_t gslot_wmax[2] = {0, 0};
    int32_t gslot_min_sites[2] = {0, 0};
    double gslot_min_data[2] = {0, 0};
    void* h_pinned = nullptr;                 // small pinned staging for result read-back
    size_t h_pinned_cap = 0;
};

// timing helpers: every kernel launch is bracketed by events on ctx->stream
void pg_timings_reset(pg_ctx* ctx);
int pg_time_begin(pg_ctx* ctx, const char* name);   // returns timing index
void pg_time_end(pg_ctx* ctx, int idx);
int pg_pinned(pg_ctx* ctx, size_t bytes, void** out);
int pg_d2h_staged(pg_ctx* ctx, void* dst, const void* src, size_t bytes);   // large device -> pageable host copy
int pg_build_segments(pg_ctx* ctx);

// tensor-core pairwise path (k2t.cu): bit-packed operand planes of one site span
struct K2TPlanes {
    int Hk = 0, R = 0;                 // plane rows (haplotypes in `order`), rows allocated (multiple of 16, pad rows zero)
    int64_t site_base = 0;             // first site of valid-plane chunk 0 (multiple of 64)
    int64_t nchunk_v = 0;
    const uint64_t* vplane = nullptr;  // [nchunk_v][R]: bit b of word (c, r) = haplotype r non-missing at site site_base + 64 c + b
    const int32_t* cps = nullptr;      // [64 nchunk_v + 1]: pseudo-sites before each site of the span
    int64_t npseudo = 0;
    const uint64_t* pq = nullptr;      // [ceil(npseudo / 64)][2][R]: P and Q planes of the pseudo-sites
    const int32_t* d_iota = nullptr;   // [Hk] 0..Hk-1
    // n_ij is computed over "mask rows": one per plane row, or — when rows 2k and 2k+1 have identical valid planes (the two
    // haplotypes of a sample with per-genotype missingness) — one per pair of rows
    int Hm = 0, R2 = 0;
    const uint64_t* vpair = nullptr;   // [nchunk_v][R2] valid words of the even rows (nullptr: every row is its own mask row)
    const int32_t* d_mid = nullptr;    // [Hk] mask row of each plane row
};
bool pg_k2_use_tensor();               // false when PG_K2_POPC is set (the bit-plane POPC kernels, kept as a checker)
int pg_k2t_build(pg_ctx* ctx, const std::vector<int32_t>& order, int64_t lo, int64_t hi, K2TPlanes& ps);
int pg_k2t_pairs(pg_ctx* ctx, const K2TPlanes& ps, const int64_t* d_lo, const int64_t* d_hi, int nb, int32_t* d_diff,
                 int32_t* d_n);
int pg_k2t_seq_nonnan(pg_ctx* ctx, const K2TPlanes& ps, const int64_t* d_lo, const int64_t* d_hi, int nb, long long* d_out);
int pg_k2t_het(pg_ctx* ctx, const K2TPlanes& ps, const int64_t* d_lo, const int64_t* d_hi, int nb, const int32_t* d_ind_start,
               int n_ind, int min_sites, double* d_out);

// implemented in k1.cu / k2.cu
// pairwise statistics for the listed windows, written into the DEVICE record table (stride RC words)
int pg_k2_popgen_windows(pg_ctx* ctx, const std::vector<int64_t>& wins, int32_t min_sites, double min_data,
                         void* d_rec, int RC);
void pg_k1_cache_free(pg_ctx* ctx);
int pg_nccl_allreduce_i64(pg_ctx* ctx, void* d_buf, size_t count);   // nccl_gather.cu
int pg_popgen_enqueue(pg_ctx* ctx, int32_t min_sites, double min_data, int32_t force_path, void* d_rec, int** h_count);
int pg_popgen_resolve(pg_ctx* ctx, int32_t min_sites, double min_data, void* d_rec, int nk2);
int pg_abba_enqueue(pg_ctx* ctx, const int* sel, double min_data, void* d_rec);                 // records [W x 8]
int pg_fourpop_enqueue(pg_ctx* ctx, const int* sel, double min_data, int mode, void* d_rec);    // records [W x 17]
