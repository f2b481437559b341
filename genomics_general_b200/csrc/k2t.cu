// K2T — the pairwise path on the 5th-generation tensor cores (tcgen05.mma kind::i8, int32 accumulators in TMEM).
//
// Reference semantics (genomics.py:903-916, 1042-1047, 1219-1221): for every haplotype pair of a window
//   n_ij    = #sites where both are non-missing                 = (V V^T)_ij          V  = 0/1 valid indicator
//   diff_ij = #sites where both are non-missing and different   = (P Q^T + Q P^T)_ij  over "pseudo-sites"
// Both are Gram matrices of 0/1 operands, so uint8 x uint8 -> int32 MMAs are bit-exact.
//
// Exact work reduction for diff: a site where fewer than two alleles are present among the selected haplotypes
// cannot contribute to any diff_ij.  A site with alleles a_0 < a_1 < ... < a_{m-1} present is split into m-1
// pseudo-sites k = 0..m-2 with P = [allele == a_k], Q = [allele in {a_{k+1}, ...}]; then
//   sum_k (P_i Q_j + Q_i P_j) = [both valid and different]          (each unordered allele pair is counted once).
// A biallelic site is ONE pseudo-site; monomorphic sites vanish.  Pseudo-sites are compacted (exclusive scan), so the
// diff Gram runs over ~(variable sites) columns instead of every site.
//
// Data flow (all operands stay bit-packed in HBM, 1 bit per genotype):
//   k2t_valid_class : resident one-hot bytes [S x pitch] -> valid plane (64-site chunks, chunk-major) + per-site allele
//                     presence nibble + pseudo-site count per chunk                       (one pass, HBM-bound)
//   k2t_scan / k2t_inv : exclusive scan -> cps[site] (pseudo-site prefix) + inverse map pseudo-site -> (site, P bit, Q mask)
//   k2t_build_pq    : gathers the variable sites' rows -> P / Q planes (64 pseudo-site chunks)
//   k2t_gram<NPL>   : one CTA per (tile group, window): 4 producer warps expand plane words to 0/1 bytes in the
//                     K-major no-swizzle core-matrix layout in shared memory, one thread issues tcgen05.mma (M=128,
//                     N<=256, K=32 per instruction) into TMEM, tcgen05.commit releases the stage; the same 4 warps
//                     read the accumulators back with tcgen05.ld and write the symmetric int32 matrix.
#include <stdlib.h>

#include <algorithm>

#include "pgwin_internal.h"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// ------------------------------------------------------------------------------------------------
// pass 1: valid plane + allele presence per site
// ------------------------------------------------------------------------------------------------
struct VcParams {
    const uint32_t* geno32;     // resident matrix as words, pw words per site row
    int pw, pitch;
    int64_t S;                  // sites in the matrix
    int64_t site_base;          // first site of chunk 0 (multiple of 64)
    int64_t nchunk;
    const int32_t* c2r;         // [pitch] column -> plane row (-1: unused)
    const uint32_t* cmask;      // [pw] 0xff in the bytes of used columns
    uint64_t* vplane;           // [nchunk][R]
    int R, Hk;
    uint8_t* cls;               // [nchunk*64] presence nibble (bit a: allele a present among the used haplotypes)
    int32_t* chunk_tot;         // [nchunk] pseudo-sites of the chunk
};

// one-hot bytes (bits 0,2,4,6) of two sites -> per byte: bit0 = site0 valid, bit1 = site1 valid
__device__ __forceinline__ uint32_t valid2(uint32_t w0, uint32_t w1) {
    const uint32_t z = w0 | (w1 << 1);
    const uint32_t t = z | (z >> 4);
    return (t | (t >> 2)) & 0x03030303u;
}

__global__ void __launch_bounds__(256) k2t_valid_class(const __grid_constant__ VcParams p) {
    extern __shared__ __align__(16) uint32_t vc_st[];      // [8 octets][pw]: byte (o, c) = valid bits of 8 sites of column c
    __shared__ int s_wtot[8];
    const int tid = threadIdx.x, lane = tid & 31, o = tid >> 5;
    for (int64_t chunk = blockIdx.x; chunk < p.nchunk; chunk += gridDim.x) {
        const int64_t site0 = p.site_base + chunk * 64 + o * 8;
        uint32_t pres[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) pres[k] = 0u;
        // one lane owns 4 column words (16 haplotype bytes) of the warp's 8 sites: 8 x LDG.128 in flight per lane
        const uint4* g4 = reinterpret_cast<const uint4*>(p.geno32);
        const int pw4 = p.pw >> 2;
        for (int q = lane; q < pw4; q += 32) {
            const uint4 cm = reinterpret_cast<const uint4*>(p.cmask)[q];
            uint4 w[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int64_t s = site0 + k;
                w[k] = (s < p.S) ? __ldg(g4 + s * pw4 + q) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) pres[k] |= (w[k].x & cm.x) | (w[k].y & cm.y) | (w[k].z & cm.z) | (w[k].w & cm.w);
            uint4 o4;
#define VC_WORD(C) (valid2(w[0].C, w[1].C) | (valid2(w[2].C, w[3].C) << 2) | (valid2(w[4].C, w[5].C) << 4) | (valid2(w[6].C, w[7].C) << 6))
            o4.x = VC_WORD(x);
            o4.y = VC_WORD(y);
            o4.z = VC_WORD(z);
            o4.w = VC_WORD(w);
#undef VC_WORD
            reinterpret_cast<uint4*>(vc_st + o * p.pw)[q] = o4;
        }
        uint32_t mine = 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t v = pres[k];
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, d);
            if (lane == k) mine = v;
        }
        int ps = 0;
        if (lane < 8) {
            uint32_t b = mine;
            b |= b >> 16;
            b |= b >> 8;
            const uint32_t nib = (b & 1u) | ((b >> 1) & 2u) | ((b >> 2) & 4u) | ((b >> 3) & 8u);
            p.cls[chunk * 64 + o * 8 + lane] = (uint8_t)nib;
            const int cnt = __popc(nib);
            ps = cnt > 1 ? cnt - 1 : 0;
        }
#pragma unroll
        for (int d = 4; d >= 1; d >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, d);
        if (lane == 0) s_wtot[o] = ps;
        __syncthreads();
        const uint8_t* st8 = reinterpret_cast<const uint8_t*>(vc_st);
        for (int c = tid; c < p.pitch; c += 256) {
            const int r = p.c2r[c];
            if (r < 0) continue;
            uint64_t v = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) v |= (uint64_t)st8[(size_t)q * p.pitch + c] << (8 * q);
            p.vplane[chunk * p.R + r] = v;
        }
        for (int r = p.Hk + tid; r < p.R; r += 256) p.vplane[chunk * p.R + r] = 0ull;
        if (tid == 0) {
            int t = 0;
            for (int q = 0; q < 8; ++q) t += s_wtot[q];
            p.chunk_tot[chunk] = t;
        }
        __syncthreads();
    }
}

// exclusive scan of the chunk totals (single CTA, 4 elements per thread and iteration; 1e8 sites = 1.6 M chunks = 400 iterations)
__global__ void __launch_bounds__(1024) k2t_scan(const int32_t* __restrict__ tot, int32_t* __restrict__ off, int64_t n) {
    __shared__ int wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int carry = 0;
    for (int64_t base = 0; base < n; base += 4096) {
        const int64_t i = base + 4 * tid;
        int4 v = make_int4(0, 0, 0, 0);
        if (i + 3 < n) v = *reinterpret_cast<const int4*>(tot + i);
        else {
            if (i < n) v.x = tot[i];
            if (i + 1 < n) v.y = tot[i + 1];
            if (i + 2 < n) v.z = tot[i + 2];
        }
        const int mine = v.x + v.y + v.z + v.w;
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int x = wsum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += t;
            }
            wsum[lane] = x;
        }
        __syncthreads();
        const int e0 = carry + (warp ? wsum[warp - 1] : 0) + incl - mine;
        if (i + 3 < n) *reinterpret_cast<int4*>(off + i) = make_int4(e0, e0 + v.x, e0 + v.x + v.y, e0 + v.x + v.y + v.z);
        else {
            if (i < n) off[i] = e0;
            if (i + 1 < n) off[i + 1] = e0 + v.x;
            if (i + 2 < n) off[i + 2] = e0 + v.x + v.y;
        }
        carry += wsum[31];
        __syncthreads();
    }
    if (tid == 0) off[n] = carry;
}

// cps[site] = pseudo-sites before the site; inv[pseudo-site] = (site relative to site_base, P shift | Q mask << 8)
__global__ void __launch_bounds__(256) k2t_inv(const uint8_t* __restrict__ cls, const int32_t* __restrict__ chunk_off,
                                               int64_t nchunk, int32_t* __restrict__ cps, uint2* __restrict__ inv) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5, nwarp = ((int64_t)gridDim.x * 256) >> 5;
    for (int64_t chunk = wid; chunk < nchunk; chunk += nwarp) {
        const uint32_t n0 = cls[chunk * 64 + 2 * lane], n1 = cls[chunk * 64 + 2 * lane + 1];
        const int c0 = __popc(n0) > 1 ? __popc(n0) - 1 : 0, c1 = __popc(n1) > 1 ? __popc(n1) - 1 : 0;
        const int v = c0 + c1;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const int base = chunk_off[chunk] + incl - v;
        cps[chunk * 64 + 2 * lane] = base;
        cps[chunk * 64 + 2 * lane + 1] = base + c0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t nib = h ? n1 : n0;
            const int cnt = h ? c1 : c0;
            int j = base + (h ? c0 : 0);
            const uint32_t site_rel = (uint32_t)(chunk * 64 + 2 * lane + h);
            for (int k = 0; k < cnt; ++k, ++j) {
                const int a = __ffs(nib) - 1;          // lowest remaining allele
                nib &= nib - 1;
                uint32_t qm = 0;                        // one-hot byte mask of the remaining (higher) alleles
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (nib & (1u << b)) qm |= 1u << (2 * b);
                inv[j] = make_uint2(site_rel, (uint32_t)(2 * a) | (qm << 8));
            }
        }
        if (chunk == nchunk - 1 && lane == 31) cps[nchunk * 64] = chunk_off[nchunk];
    }
}

// ------------------------------------------------------------------------------------------------
// pass 2: P / Q planes of the pseudo-sites
// ------------------------------------------------------------------------------------------------
struct PqParams {
    const uint32_t* geno32;
    int pw, pitch;
    int64_t site_base;
    int64_t total;              // pseudo-sites
    int64_t nchunk;             // ceil(total / 64)
    const uint2* inv;
    const int32_t* c2r;
    uint64_t* pq;               // [nchunk][2][R]
    int R, Hk;
};

__global__ void __launch_bounds__(256) k2t_build_pq(const __grid_constant__ PqParams p) {
    extern __shared__ __align__(16) uint32_t pq_st[];      // [2][8][pw]
    const int tid = threadIdx.x, lane = tid & 31, o = tid >> 5;
    for (int64_t chunk = blockIdx.x; chunk < p.nchunk; chunk += gridDim.x) {
        const int64_t j0 = chunk * 64 + o * 8;
        const uint32_t* row[8];
        uint32_t psh[8], qm[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const bool ok = j0 + k < p.total;
            const uint2 e = ok ? __ldg(p.inv + j0 + k) : make_uint2(0u, 0u);
            row[k] = p.geno32 + (p.site_base + (int64_t)e.x) * p.pw;
            psh[k] = ok ? (e.y & 0xffu) : 0u;
            qm[k] = ok ? ((e.y >> 8) & 0xffu) * 0x01010101u : 0u;
            if (!ok) row[k] = nullptr;
        }
        const int pw4 = p.pw >> 2;
        for (int q = lane; q < pw4; q += 32) {
            uint4 w[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                w[k] = row[k] ? __ldg(reinterpret_cast<const uint4*>(row[k]) + q) : make_uint4(0u, 0u, 0u, 0u);
            uint4 op = make_uint4(0u, 0u, 0u, 0u), oq = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
#define PQ_WORD(C)                                         \
    {                                                      \
        op.C |= ((w[k].C >> psh[k]) & 0x01010101u) << k;   \
        uint32_t t = w[k].C & qm[k];                       \
        t |= t >> 4;                                       \
        t |= t >> 2;                                       \
        oq.C |= (t & 0x01010101u) << k;                    \
    }
                PQ_WORD(x)
                PQ_WORD(y)
                PQ_WORD(z)
                PQ_WORD(w)
#undef PQ_WORD
            }
            reinterpret_cast<uint4*>(pq_st + (0 * 8 + o) * p.pw)[q] = op;
            reinterpret_cast<uint4*>(pq_st + (1 * 8 + o) * p.pw)[q] = oq;
        }
        __syncthreads();
        const uint8_t* st8 = reinterpret_cast<const uint8_t*>(pq_st);
        for (int c = tid; c < p.pitch; c += 256) {
            const int r = p.c2r[c];
            if (r < 0) continue;
            uint64_t vp = 0, vq = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                vp |= (uint64_t)st8[(size_t)q * p.pitch + c] << (8 * q);
                vq |= (uint64_t)st8[(size_t)(8 + q) * p.pitch + c] << (8 * q);
            }
            p.pq[(chunk * 2 + 0) * p.R + r] = vp;
            p.pq[(chunk * 2 + 1) * p.R + r] = vq;
        }
        for (int r = p.Hk + tid; r < p.R; r += 256) {
            p.pq[(chunk * 2 + 0) * p.R + r] = 0ull;
            p.pq[(chunk * 2 + 1) * p.R + r] = 0ull;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Gram kernel
// ------------------------------------------------------------------------------------------------
struct GramGroup {
    int a_row0;       // first row of the 128-row A tile
    int b_row0;       // first row of the B range
    int nb_rows;      // rows of the B range: multiple of 16, <= 512
    int pad;
};
struct GramParams {
    const uint64_t* plane;      // NPL == 1: [chunk][R];  NPL == 2: [chunk][2][R]
    int R, Hk;
    int64_t site_base;          // plane coordinate of an absolute site = site - site_base
    const int64_t* win_lo;      // [nb] absolute site ranges
    const int64_t* win_hi;
    const int32_t* cps;         // NPL == 2: plane coordinate = cps[site - site_base]
    const GramGroup* groups;
    int nbmax;                  // max nb_rows over the groups (shared-memory geometry)
    int nstages;                // operand ring depth
    int nraw;                   // raw plane-word ring depth
    int32_t* out;               // [nb][Hk][Hk]
};

constexpr int GRAM_PRODUCERS = 384;        // 12 expanding warps: three per scheduler, so that dependent ALU chains interleave
constexpr int GRAM_NPW = GRAM_PRODUCERS / 32;
constexpr int GRAM_THREADS = GRAM_PRODUCERS + 64;   // warps 0..NPW-1: expand, then (0-7) epilogue; then TMEM allocation + MMA issue; then TMA
constexpr int GRAM_MAX_STAGES = 4;
constexpr int GRAM_MAX_RAW = 8;            // depth of the raw plane-word ring (TMA runs this many chunks ahead)
constexpr int GRAM_MAX_ITEMS = ((128 + 512) * 2 + GRAM_PRODUCERS - 1) / GRAM_PRODUCERS;    // plane words per thread and stage

// 16 bits -> 16 bytes of 0/1 (byte k = bit k)
__device__ __forceinline__ uint4 expand16(uint32_t x) {
    uint4 r;
    r.x = ((x & 0xfu) * 0x00204081u) & 0x01010101u;
    r.y = (((x >> 4) & 0xfu) * 0x00204081u) & 0x01010101u;
    r.z = (((x >> 8) & 0xfu) * 0x00204081u) & 0x01010101u;
    r.w = (((x >> 12) & 0xfu) * 0x00204081u) & 0x01010101u;
    return r;
}

// shared-memory matrix descriptor: K-major, no swizzle; core matrix = 8 rows x 16 bytes stored as 128 contiguous bytes;
// LBO = byte distance between the two K cores of a K=32 slab (128), SBO = byte distance between 8-row groups (256)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) |
           (1ull << 46);
}
// instruction descriptor: D = S32, A = B = UINT8, both K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc(int N) {
    return (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// Shared memory: [raw ring: nraw slots of NPL x (128 + nbmax) plane words, filled by 1-D TMA bulk copies]
//                [operand ring: nstages stages of 2 K steps x NPL planes x (128 + nbmax) rows x 32 bytes]
template <int NPL>
__global__ void __launch_bounds__(GRAM_THREADS, 1) k2t_gram(const __grid_constant__ GramParams gp) {
    extern __shared__ __align__(128) uint8_t gsm[];
    __shared__ __align__(8) uint64_t full[GRAM_MAX_STAGES], empty[GRAM_MAX_STAGES], raw_full[GRAM_MAX_RAW],
        raw_empty[GRAM_MAX_RAW], done;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const GramGroup g = gp.groups[blockIdx.x];
    const int wb = blockIdx.y;
    const int NS = gp.nstages, RD = gp.nraw;
    const int RROWS = 128 + gp.nbmax;                   // rows of one plane in a raw slot / operand block
    const int RAW = NPL * RROWS * 8;                    // bytes of one raw slot
    const int BLK = RROWS * 32;                         // one (K step, plane) block: A region 128 rows, then the B region
    const int STAGE = 2 * NPL * BLK;
    uint8_t* const raw_base = gsm;
    uint8_t* const op_base = gsm + (size_t)RD * RAW;

    // window in plane coordinates
    int64_t lo = gp.win_lo[wb] - gp.site_base, hi = gp.win_hi[wb] - gp.site_base;
    if (NPL == 2) {
        lo = gp.cps[lo];
        hi = gp.cps[hi];
    }
    const int64_t c_first = lo >> 6;
    const int nst = (hi > lo) ? (int)(((hi - 1) >> 6) - c_first + 1) : 0;

    uint32_t ncols = 32;
    while ((int)ncols < g.nb_rows) ncols <<= 1;
    const bool a_in_b = (g.a_row0 == g.b_row0);
    const uint32_t a_base = a_in_b ? 4096u : 0u;       // byte offset of the A tile inside a (K step, plane) block
    // rows of the A tile past the last plane row are never copied: they must read as zero
    for (int i = tid; i < RD * RAW / 16; i += GRAM_THREADS) reinterpret_cast<uint4*>(raw_base)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == GRAM_NPW) {
        if (lane == 0) {
            for (int s = 0; s < NS; ++s) {
                mbar_init(&full[s], GRAM_NPW);            // one arrival per expanding warp
                mbar_init(&empty[s], 1);
            }
            for (int s = 0; s < RD; ++s) {
                mbar_init(&raw_full[s], 1);
                mbar_init(&raw_empty[s], GRAM_NPW);
            }
            mbar_init(&done, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the zero fill above precedes the bulk copies
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp == GRAM_NPW + 1) {
        // ---------------- TMA: plane words of chunk c_first + it -> raw slot it % RD ----------------
        if (lane == 0) {
            const int a_rows = min(128, gp.R - g.a_row0);
            const uint32_t bytes_a = (uint32_t)a_rows * 8u, bytes_b = (uint32_t)g.nb_rows * 8u;
            int slot = 0;
            uint32_t ph = 0;
            for (int it = 0; it < nst; ++it) {
                if (it >= RD) mbar_wait(&raw_empty[slot], ph ^ 1u);
                mbar_expect_tx(&raw_full[slot], NPL * ((a_in_b ? 0u : bytes_a) + bytes_b));
                const int64_t chunk = c_first + it;
#pragma unroll
                for (int pl = 0; pl < NPL; ++pl) {
                    const uint64_t* src = gp.plane + (chunk * NPL + pl) * gp.R;
                    uint8_t* dst = raw_base + (size_t)slot * RAW + (size_t)pl * RROWS * 8;
                    if (!a_in_b) bulk_g2s(dst, src + g.a_row0, bytes_a, &raw_full[slot]);
                    bulk_g2s(dst + 128 * 8, src + g.b_row0, bytes_b, &raw_full[slot]);
                }
                if (++slot == RD) {
                    slot = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp < GRAM_NPW) {
        // ---------------- producers: plane words -> 0/1 bytes in the core-matrix layout ----------------
        // a diagonal group (A tile = first rows of the B range) expands the B rows only: the A descriptor points into them
        const int skip_a = a_in_b ? 128 : 0;
        const int rows_tot = 128 + g.nb_rows - skip_a;
        const int nitems = rows_tot * NPL;
        int r_idx[GRAM_MAX_ITEMS], d_off[GRAM_MAX_ITEMS];          // raw word index / byte offset in a K-step block, -1: none
#pragma unroll
        for (int q = 0; q < GRAM_MAX_ITEMS; ++q) {
            const int item = tid + q * GRAM_PRODUCERS;
            r_idx[q] = -1;
            d_off[q] = 0;
            if (item < nitems) {
                const int pl = (NPL == 2 && item >= rows_tot) ? 1 : 0;
                const int rr = item - pl * rows_tot + skip_a;
                const int x = (rr < 128) ? rr : rr - 128;
                r_idx[q] = pl * RROWS + rr;
                d_off[q] = pl * BLK + ((rr < 128) ? 0 : 4096) + (x >> 3) * 256 + (x & 7) * 16;
            }
        }
        int s = 0, slot = 0;
        uint32_t ph_s = 0, ph_r = 0;
        for (int it = 0; it < nst; ++it) {
            const int64_t chunk = c_first + it;
            uint64_t mask = ~0ull;
            {
                const int64_t b0 = chunk << 6;
                if (lo > b0) mask &= ~0ull << (int)(lo - b0);
                if (hi < b0 + 64) mask &= ~0ull >> (int)(b0 + 64 - hi);
            }
            mbar_wait(&raw_full[slot], ph_r);
            const uint64_t* rw = reinterpret_cast<const uint64_t*>(raw_base + (size_t)slot * RAW);
            uint64_t v[GRAM_MAX_ITEMS];
#pragma unroll
            for (int q = 0; q < GRAM_MAX_ITEMS; ++q) v[q] = (r_idx[q] >= 0) ? (rw[r_idx[q]] & mask) : 0ull;
            if (it >= NS) mbar_wait(&empty[s], ph_s ^ 1u);
            uint8_t* sb = op_base + (size_t)s * STAGE;
#pragma unroll
            for (int q = 0; q < GRAM_MAX_ITEMS; ++q) {
                if (r_idx[q] >= 0) {
                    const uint32_t wlo = (uint32_t)v[q], whi = (uint32_t)(v[q] >> 32);
                    uint8_t* d0 = sb + d_off[q];
                    uint8_t* d1 = d0 + NPL * BLK;
                    *reinterpret_cast<uint4*>(d0) = expand16(wlo & 0xffffu);
                    *reinterpret_cast<uint4*>(d0 + 128) = expand16(wlo >> 16);
                    *reinterpret_cast<uint4*>(d1) = expand16(whi & 0xffffu);
                    *reinterpret_cast<uint4*>(d1 + 128) = expand16(whi >> 16);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&full[s]);
                // released only now: the stores above consumed the words, so the loads from the slot have completed before
                // the TMA (async proxy) may overwrite it — an arrive right after issuing the loads raced with the refill
                mbar_arrive(&raw_empty[slot]);
            }
            if (++s == NS) {
                s = 0;
                ph_s ^= 1u;
            }
            if (++slot == RD) {
                slot = 0;
                ph_r ^= 1u;
            }
        }
    } else if (lane == 0) {
        // ---------------- MMA issue (one thread) ----------------
        const uint32_t sbase = smem_u32(op_base);
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < nst; ++it) {
            mbar_wait(&full[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t st = sbase + (uint32_t)s * STAGE;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const uint32_t acc0 = (it > 0 || ks > 0) ? 1u : 0u;
                for (int n0 = 0; n0 < g.nb_rows; n0 += 256) {
                    const int nn = min(256, g.nb_rows - n0);
                    const uint32_t idesc = umma_idesc(nn);
                    if (NPL == 1) {
                        const uint32_t blk = st + ks * BLK;
                        umma_i8(tmem + n0, umma_desc(blk + a_base), umma_desc(blk + 4096 + n0 * 32), idesc, acc0);
                    } else {
                        const uint32_t bp = st + (ks * 2 + 0) * BLK, bq = st + (ks * 2 + 1) * BLK;
                        umma_i8(tmem + n0, umma_desc(bp + a_base), umma_desc(bq + 4096 + n0 * 32), idesc, acc0);
                        umma_i8(tmem + n0, umma_desc(bq + a_base), umma_desc(bp + 4096 + n0 * 32), idesc, 1u);
                    }
                }
            }
            umma_commit(&empty[s]);       // arrives when the MMAs above have read the stage
            if (++s == NS) {
                s = 0;
                ph ^= 1u;
            }
        }
        umma_commit(&done);
    }

    if (warp < 8) {
        // ---------------- epilogue: TMEM -> registers -> symmetric int32 matrix ----------------
        if (nst > 0) {
            mbar_wait(&done, 0u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        // Lane = matrix row in TMEM (warp w may touch lanes 32 (w % 4) ..): warps w and w + 4 take alternate 32-column
        // blocks of the same rows.  The mirror element [j][i] is written straight from the registers (lanes run along i:
        // coalesced); the direct element [i][j] goes through a 32 x 32 transpose in shared memory (the operand ring is idle
        // now) so that lanes run along j as well.
        const int qd = warp & 3;
        const int i0 = g.a_row0 + qd * 32;
        int32_t* o = gp.out + (size_t)wb * gp.Hk * gp.Hk;
        uint32_t* tr = reinterpret_cast<uint32_t*>(op_base) + warp * (32 * 33);
        for (int c0 = (warp >> 2) * 32; c0 < g.nb_rows; c0 += 64) {
            uint32_t v[32];
            if (nst > 0) {
                const uint32_t taddr = tmem + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
                    "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                      "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                      "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = 0u;
            }
            const int i = i0 + lane;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const int j = g.b_row0 + c0 + e;
                if (i < gp.Hk && j < gp.Hk) o[(size_t)j * gp.Hk + i] = (int32_t)v[e];
                tr[lane * 33 + e] = v[e];
            }
            __syncwarp();
            const int j = g.b_row0 + c0 + lane;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                const uint32_t x = tr[r * 33 + lane];
                if (i0 + r < gp.Hk && j < gp.Hk) o[(size_t)(i0 + r) * gp.Hk + j] = (int32_t)x;
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == GRAM_NPW) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols) : "memory");
    }
}

// ---- small consumers of the planes ---------------------------------------------------------------------
__device__ __forceinline__ uint64_t chunk_mask(int64_t chunk, int64_t lo, int64_t hi) {
    uint64_t m = ~0ull;
    const int64_t b0 = chunk << 6;
    if (lo > b0) m &= ~0ull << (int)(lo - b0);
    if (hi < b0 + 64) m &= ~0ull >> (int)(b0 + 64 - hi);
    return m;
}

// Alignment.seqNonNan (genomics.py:1038-1040): thread = plane row, one CTA column per window
__global__ void __launch_bounds__(128) k2t_seq_nonnan(const uint64_t* __restrict__ vplane, int R, int Hk, int64_t site_base,
                                                      const int64_t* __restrict__ win_lo, const int64_t* __restrict__ win_hi,
                                                      long long* __restrict__ out) {
    const int r = blockIdx.x * 128 + threadIdx.x, wb = blockIdx.y;
    if (r >= Hk) return;
    const int64_t lo = win_lo[wb] - site_base, hi = win_hi[wb] - site_base;
    long long n = 0;
    if (hi > lo)
        for (int64_t c = lo >> 6; c <= (hi - 1) >> 6; ++c) n += __popcll(vplane[c * R + r] & chunk_mask(c, lo, hi));
    out[(size_t)wb * Hk + r] = n;
}

// Alignment.sampleHet (genomics.py:918-929): thread = individual (rows ind_start[a], +1)
__global__ void __launch_bounds__(128) k2t_het(const uint64_t* __restrict__ vplane, const uint64_t* __restrict__ pq,
                                               const int32_t* __restrict__ cps, int R, int64_t site_base,
                                               const int64_t* __restrict__ win_lo, const int64_t* __restrict__ win_hi,
                                               const int32_t* __restrict__ ind_start, int n_ind, int min_sites,
                                               double* __restrict__ out) {
    const int a = blockIdx.x * 128 + threadIdx.x, wb = blockIdx.y;
    if (a >= n_ind) return;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    const int r0 = ind_start[a], r1 = ind_start[a + 1];
    double v = nan;
    if (r1 - r0 == 2) {             // len(x) == 2 is required (the reference raises IndexError for len(x) == 1)
        const int64_t lo = win_lo[wb] - site_base, hi = win_hi[wb] - site_base;
        long long n = 0, diff = 0;
        if (hi > lo) {
            for (int64_t c = lo >> 6; c <= (hi - 1) >> 6; ++c)
                n += __popcll(vplane[c * R + r0] & vplane[c * R + r0 + 1] & chunk_mask(c, lo, hi));
            const int64_t plo = cps[lo], phi = cps[hi];
            if (phi > plo)
                for (int64_t c = plo >> 6; c <= (phi - 1) >> 6; ++c) {
                    const uint64_t p0 = pq[(c * 2) * R + r0], p1 = pq[(c * 2) * R + r0 + 1];
                    const uint64_t q0 = pq[(c * 2 + 1) * R + r0], q1 = pq[(c * 2 + 1) * R + r0 + 1];
                    diff += __popcll(((p0 & q1) | (q0 & p1)) & chunk_mask(c, plo, phi));
                }
        }
        // `len(x)==2 & np.sum(mask) >= 1` parses as len(x) == (2 & n) >= 1: bit 1 of n must be set (924, 927)
        if ((n & 2) == 2 && !(min_sites > 0 && n < min_sites)) v = (double)diff / (double)n;
    }
    out[(size_t)wb * n_ind + a] = v;
}

__global__ void k2t_iota(int32_t* p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
bool pg_k2_use_tensor() { return getenv("PG_K2_POPC") == nullptr; }

// Planes for sites [lo, hi) of the haplotype columns in `order` (plane row r = column order[r]).
int pg_k2t_build(pg_ctx* ctx, const std::vector<int32_t>& order, int64_t lo, int64_t hi, K2TPlanes& ps) {
    const int Hk = (int)order.size();
    PG_CHECK(Hk >= 1, "pairwise path: no haplotypes selected");
    const int R = (Hk + 15) / 16 * 16;
    const int64_t sb = lo & ~(int64_t)63;
    const int64_t nchunk = (hi - sb + 63) / 64;
    PG_CHECK(nchunk * 64 < (int64_t)1 << 31, "pairwise path: site span too large for one call");
    const int pitch = ctx->pitch, pw = pitch / 4;
    PG_CHECK((size_t)16 * pw * 4 <= 96 * 1024, "pairwise path: %d haplotype columns are too many for the plane builders", pitch);
    {
        static bool attr_dev[64] = {};
        if (!attr_dev[ctx->device & 63]) {
            PG_CUDA(cudaFuncSetAttribute(k2t_build_pq, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            PG_CUDA(cudaFuncSetAttribute(k2t_valid_class, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            attr_dev[ctx->device & 63] = true;
        }
    }
    // column tables
    std::vector<int32_t> c2r(pitch, -1);
    for (int r = 0; r < Hk; ++r) c2r[order[r]] = r;
    std::vector<uint32_t> cmask(pw, 0u);
    for (int c = 0; c < pitch; ++c)
        if (c2r[c] >= 0) cmask[c / 4] |= 0xffu << (8 * (c % 4));
    PG_TRY(ctx->misc2.ensure((size_t)pitch * 4 + (size_t)pw * 4 + (size_t)Hk * 4 + 256));
    int32_t* d_c2r = (int32_t*)ctx->misc2.p;
    uint32_t* d_cmask = (uint32_t*)(d_c2r + pitch);
    int32_t* d_iota = (int32_t*)(d_cmask + pw);
    PG_CUDA(cudaMemcpyAsync(d_c2r, c2r.data(), (size_t)pitch * 4, cudaMemcpyHostToDevice, ctx->stream));
    PG_CUDA(cudaMemcpyAsync(d_cmask, cmask.data(), (size_t)pw * 4, cudaMemcpyHostToDevice, ctx->stream));
    k2t_iota<<<(Hk + 255) / 256, 256, 0, ctx->stream>>>(d_iota, Hk);
    // plane memory: vplane | cls | chunk_tot | chunk_off | cps
    const size_t span = (size_t)nchunk * 64;
    size_t off = 0;
    auto carve = [&](size_t bytes) {
        const size_t o = off;
        off += (bytes + 255) / 256 * 256;
        return o;
    };
    const size_t o_v = carve((size_t)nchunk * R * 8), o_cls = carve(span), o_tot = carve((size_t)nchunk * 4),
                 o_off = carve((size_t)(nchunk + 1) * 4), o_cps = carve((span + 1) * 4);
    PG_TRY(ctx->planes.ensure(off));
    uint8_t* base = (uint8_t*)ctx->planes.p;
    VcParams vp;
    vp.geno32 = (const uint32_t*)ctx->d_geno;
    vp.pw = pw;
    vp.pitch = pitch;
    vp.S = ctx->S;
    vp.site_base = sb;
    vp.nchunk = nchunk;
    vp.c2r = d_c2r;
    vp.cmask = d_cmask;
    vp.vplane = (uint64_t*)(base + o_v);
    vp.R = R;
    vp.Hk = Hk;
    vp.cls = base + o_cls;
    vp.chunk_tot = (int32_t*)(base + o_tot);
    int32_t* d_off = (int32_t*)(base + o_off);
    int32_t* d_cps = (int32_t*)(base + o_cps);
    const int grid1 = (int)std::min<int64_t>(nchunk, (int64_t)ctx->sm_count * 8);
    {
        const int ti = pg_time_begin(ctx, "k2t_valid_class");
        k2t_valid_class<<<grid1, 256, (size_t)8 * pw * 4, ctx->stream>>>(vp);
        pg_time_end(ctx, ti);
        PG_CUDA(cudaGetLastError());
    }
    {
        const int ti = pg_time_begin(ctx, "k2t_scan");
        k2t_scan<<<1, 1024, 0, ctx->stream>>>(vp.chunk_tot, d_off, nchunk);
        pg_time_end(ctx, ti);
        PG_CUDA(cudaGetLastError());
    }
    int32_t total = 0;
    PG_CUDA(cudaMemcpyAsync(&total, d_off + nchunk, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    const int64_t nchunk_d = ((int64_t)total + 63) / 64;
    off = 0;
    const size_t o_inv = carve((size_t)std::max<int64_t>(total, 1) * 8), o_pq = carve((size_t)std::max<int64_t>(nchunk_d, 1) * 2 * R * 8);
    PG_TRY(ctx->planes2.ensure(off));
    uint8_t* base2 = (uint8_t*)ctx->planes2.p;
    {
        const int ti = pg_time_begin(ctx, "k2t_inv");
        const int gridi = (int)std::min<int64_t>((nchunk + 7) / 8, (int64_t)ctx->sm_count * 16);
        k2t_inv<<<gridi, 256, 0, ctx->stream>>>(vp.cls, d_off, nchunk, d_cps, (uint2*)(base2 + o_inv));
        pg_time_end(ctx, ti);
        PG_CUDA(cudaGetLastError());
    }
    if (nchunk_d > 0) {
        PqParams pp;
        pp.geno32 = vp.geno32;
        pp.pw = pw;
        pp.pitch = pitch;
        pp.site_base = sb;
        pp.total = total;
        pp.nchunk = nchunk_d;
        pp.inv = (const uint2*)(base2 + o_inv);
        pp.c2r = d_c2r;
        pp.pq = (uint64_t*)(base2 + o_pq);
        pp.R = R;
        pp.Hk = Hk;
        const size_t smem = (size_t)16 * pw * 4;
        const int grid2 = (int)std::min<int64_t>(nchunk_d, (int64_t)ctx->sm_count * 8);
        const int ti = pg_time_begin(ctx, "k2t_build_pq");
        k2t_build_pq<<<grid2, 256, smem, ctx->stream>>>(pp);
        pg_time_end(ctx, ti);
        PG_CUDA(cudaGetLastError());
    }
    ps.Hk = Hk;
    ps.R = R;
    ps.site_base = sb;
    ps.nchunk_v = nchunk;
    ps.vplane = vp.vplane;
    ps.cps = d_cps;
    ps.npseudo = total;
    ps.pq = (uint64_t*)(base2 + o_pq);
    ps.d_iota = d_iota;
    return PG_OK;
}

// diff [nb][Hk^2] and n [nb][Hk^2] for nb windows (absolute site ranges on the device)
int pg_k2t_pairs(pg_ctx* ctx, const K2TPlanes& ps, const int64_t* d_lo, const int64_t* d_hi, int nb, int32_t* d_diff,
                 int32_t* d_n) {
    // tile groups: one 128-row A tile x up to 512 B rows (TMEM has 512 int32 columns per SM)
    std::vector<GramGroup> groups;
    int nbmax = 16;
    for (int a0 = 0; a0 < ps.R; a0 += 128)
        for (int c = a0; c < ps.R; c += 512) {
            GramGroup g;
            g.a_row0 = a0;
            g.b_row0 = c;
            g.nb_rows = std::min(512, ps.R - c);
            g.pad = 0;
            nbmax = std::max(nbmax, g.nb_rows);
            groups.push_back(g);
        }
    PG_TRY(ctx->misc4.ensure(groups.size() * sizeof(GramGroup) + 64));
    PG_CUDA(cudaMemcpyAsync(ctx->misc4.p, groups.data(), groups.size() * sizeof(GramGroup), cudaMemcpyHostToDevice, ctx->stream));
    GramParams gp;
    gp.R = ps.R;
    gp.Hk = ps.Hk;
    gp.site_base = ps.site_base;
    gp.win_lo = d_lo;
    gp.win_hi = d_hi;
    gp.groups = (const GramGroup*)ctx->misc4.p;
    gp.nbmax = nbmax;
    static bool attr_dev[64] = {};
    if (!attr_dev[ctx->device & 63]) {
        PG_CUDA(cudaFuncSetAttribute(k2t_gram<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        PG_CUDA(cudaFuncSetAttribute(k2t_gram<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_dev[ctx->device & 63] = true;
    }
    const dim3 grid((unsigned)groups.size(), (unsigned)nb);
    // shared memory: an operand ring of >= 2 stages, the rest (up to 8 slots) for the raw plane words
    const int budget = 206 * 1024;
    auto geometry = [&](int npl, int& nstages, int& nraw) {
        const int stage = 2 * npl * (128 + nbmax) * 32, raw = npl * (128 + nbmax) * 8;
        nstages = std::max(2, std::min(3, (budget - 4 * raw) / stage));
        nraw = std::max(2, std::min(GRAM_MAX_RAW, (budget - nstages * stage) / raw));
        if (const char* e = getenv("PG_K2T_NRAW")) nraw = std::max(1, std::min(nraw, atoi(e)));
        if (const char* e = getenv("PG_K2T_NSTAGES")) nstages = std::max(1, std::min(nstages, atoi(e)));
        // the epilogue's 8 transpose tiles (32 x 33 words each) reuse the operand ring
        // (+ 4 KB: the 128-row A tile of a diagonal group may reach past a B range of fewer than 128 rows)
        return std::max((size_t)nstages * stage, (size_t)8 * 32 * 33 * 4) + (size_t)nraw * raw + 4096;
    };
    {
        const size_t smem = geometry(1, gp.nstages, gp.nraw);
        gp.plane = ps.vplane;
        gp.cps = nullptr;
        gp.out = d_n;
        const int ti = pg_time_begin(ctx, "k2t_gram_n");
        k2t_gram<1><<<grid, GRAM_THREADS, smem, ctx->stream>>>(gp);
        pg_time_end(ctx, ti);
        PG_CUDA(cudaGetLastError());
    }
    {
        const size_t smem = geometry(2, gp.nstages, gp.nraw);
        gp.plane = ps.pq;
        gp.cps = ps.cps;
        gp.out = d_diff;
        const int ti = pg_time_begin(ctx, "k2t_gram_diff");
        k2t_gram<2><<<grid, GRAM_THREADS, smem, ctx->stream>>>(gp);
        pg_time_end(ctx, ti);
        PG_CUDA(cudaGetLastError());
    }
    return PG_OK;
}

int pg_k2t_seq_nonnan(pg_ctx* ctx, const K2TPlanes& ps, const int64_t* d_lo, const int64_t* d_hi, int nb, long long* d_out) {
    const int ti = pg_time_begin(ctx, "k2_seq_nonnan");
    k2t_seq_nonnan<<<dim3((unsigned)((ps.Hk + 127) / 128), (unsigned)nb), 128, 0, ctx->stream>>>(ps.vplane, ps.R, ps.Hk,
                                                                                                ps.site_base, d_lo, d_hi, d_out);
    pg_time_end(ctx, ti);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

int pg_k2t_het(pg_ctx* ctx, const K2TPlanes& ps, const int64_t* d_lo, const int64_t* d_hi, int nb, const int32_t* d_ind_start,
               int n_ind, int min_sites, double* d_out) {
    const int ti = pg_time_begin(ctx, "k2_het");
    k2t_het<<<dim3((unsigned)((n_ind + 127) / 128), (unsigned)nb), 128, 0, ctx->stream>>>(
        ps.vplane, ps.pq, ps.cps, ps.R, ps.site_base, d_lo, d_hi, d_ind_start, n_ind, min_sites, d_out);
    pg_time_end(ctx, ti);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
