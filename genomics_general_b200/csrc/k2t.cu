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
//   k2t_gram<NPL>   : persistent CTAs (one per SM) over (window, tile group) items: TMA warps bring plane words into a raw
//                     ring, three groups of warps expand them to 0/1 bytes in the K-major no-swizzle core-matrix layout
//                     in shared memory, one warp issues tcgen05.mma (M=128, N<=256, K=32 per instruction) into TMEM,
//                     tcgen05.commit releases the stage; epilogue warps read the accumulators back with tcgen05.ld and
//                     write the upper triangle of the symmetric int32 matrix.
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
    uint64_t* vpair;            // [nchunk][R2] valid words of the even rows (nullptr: not wanted)
    int R2;
    int32_t* pair_flag;         // set to 1 when some row 2k and 2k+1 differ in a valid word
};

// one-hot bytes (bits 0,2,4,6) of two sites -> per byte: bit0 = site0 valid, bit1 = site1 valid
__device__ __forceinline__ uint32_t valid2(uint32_t w0, uint32_t w1) {
    const uint32_t z = w0 | (w1 << 1);
    const uint32_t t = z | (z >> 4);
    return (t | (t >> 2)) & 0x03030303u;
}

// eight words [4 columns x 1 byte] (one per site octet) -> per column the 64-site word (lo, hi): two 4x4 byte transposes
__device__ __forceinline__ void octets_to_words(const uint32_t (&a)[8], uint32_t (&lo)[4], uint32_t (&hi)[4]) {
    {
        const uint32_t t0 = __byte_perm(a[0], a[1], 0x5140), t1 = __byte_perm(a[2], a[3], 0x5140);
        const uint32_t t2 = __byte_perm(a[0], a[1], 0x7362), t3 = __byte_perm(a[2], a[3], 0x7362);
        lo[0] = __byte_perm(t0, t1, 0x5410);
        lo[1] = __byte_perm(t0, t1, 0x7632);
        lo[2] = __byte_perm(t2, t3, 0x5410);
        lo[3] = __byte_perm(t2, t3, 0x7632);
    }
    {
        const uint32_t t0 = __byte_perm(a[4], a[5], 0x5140), t1 = __byte_perm(a[6], a[7], 0x5140);
        const uint32_t t2 = __byte_perm(a[4], a[5], 0x7362), t3 = __byte_perm(a[6], a[7], 0x7362);
        hi[0] = __byte_perm(t0, t1, 0x5410);
        hi[1] = __byte_perm(t0, t1, 0x7632);
        hi[2] = __byte_perm(t2, t3, 0x5410);
        hi[3] = __byte_perm(t2, t3, 0x7632);
    }
}

// ALL: every column of the row is a selected haplotype or padding (padding bytes are 0 = missing): no column mask needed
template <bool ALL>
__global__ void __launch_bounds__(256, 4) k2t_valid_class(const __grid_constant__ VcParams p) {
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
            uint4 cm = make_uint4(~0u, ~0u, ~0u, ~0u);
            if (!ALL) cm = reinterpret_cast<const uint4*>(p.cmask)[q];
            uint4 w[8];
            const uint4* rp = g4 + site0 * pw4 + q;
            if (site0 + 8 <= p.S) {      // (warp-uniform) every chunk but the last: no per-row guards
#pragma unroll
                for (int k = 0; k < 8; ++k) w[k] = __ldg(rp + (int64_t)k * pw4);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) w[k] = (site0 + k < p.S) ? __ldg(rp + (int64_t)k * pw4) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
                pres[k] |= ALL ? (w[k].x | w[k].y | w[k].z | w[k].w)
                               : ((w[k].x & cm.x) | (w[k].y & cm.y) | (w[k].z & cm.z) | (w[k].w & cm.w));
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
            const uint32_t v = __reduce_or_sync(0xffffffffu, pres[k]);      // REDUX.OR: one instruction per site
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
        uint64_t* s_v = reinterpret_cast<uint64_t*>(vc_st + 8 * p.pw);      // [R] this chunk's words by plane row
        // a thread turns four columns: 8 octet words [4 columns x 1 byte] -> 4 words of 64 sites (two 4x4 byte transposes)
        for (int cw = tid; cw < p.pw; cw += 256) {
            uint32_t a[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] = vc_st[q * p.pw + cw];
            uint32_t lo[4], hi[4];
            octets_to_words(a, lo, hi);
            const int4 r4 = *reinterpret_cast<const int4*>(p.c2r + 4 * cw);
            const int rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (rr[j] < 0) continue;
                const uint64_t v = (uint64_t)lo[j] | ((uint64_t)hi[j] << 32);
                p.vplane[chunk * p.R + rr[j]] = v;
                s_v[rr[j]] = v;
            }
        }
        for (int r = p.Hk + tid; r < p.R; r += 256) p.vplane[chunk * p.R + r] = 0ull;
        if (p.vpair) {
            // Missingness is usually per genotype: the two haplotypes of a sample then share their valid words, n_ij needs
            // one row per sample, and the co-valid Gram shrinks 4x.  Checked here for every word; the flag decides later.
            __syncthreads();
            bool bad = false;
            for (int k2 = tid; k2 < p.R2; k2 += 256) {
                uint64_t a = 0;
                if (2 * k2 + 1 < p.Hk) {
                    a = s_v[2 * k2];
                    bad |= (a != s_v[2 * k2 + 1]);
                }
                p.vpair[chunk * p.R2 + k2] = a;
            }
            if (bad) atomicOr(p.pair_flag, 1);
        }
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

__global__ void __launch_bounds__(256, 4) k2t_build_pq(const __grid_constant__ PqParams p) {
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
            if (j0 + 8 <= p.total) {       // (warp-uniform) all eight pseudo-sites exist
#pragma unroll
                for (int k = 0; k < 8; ++k) w[k] = __ldg(reinterpret_cast<const uint4*>(row[k]) + q);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    w[k] = row[k] ? __ldg(reinterpret_cast<const uint4*>(row[k]) + q) : make_uint4(0u, 0u, 0u, 0u);
            }
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
        for (int cw = tid; cw < p.pw; cw += 256) {      // a thread turns four columns of both planes
            uint32_t a[8], plo[4], phi[4], qlo[4], qhi[4];
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] = pq_st[q * p.pw + cw];
            octets_to_words(a, plo, phi);
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] = pq_st[(8 + q) * p.pw + cw];
            octets_to_words(a, qlo, qhi);
            const int4 r4 = *reinterpret_cast<const int4*>(p.c2r + 4 * cw);
            const int rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (rr[j] < 0) continue;
                p.pq[(chunk * 2 + 0) * p.R + rr[j]] = (uint64_t)plo[j] | ((uint64_t)phi[j] << 32);
                p.pq[(chunk * 2 + 1) * p.R + rr[j]] = (uint64_t)qlo[j] | ((uint64_t)qhi[j] << 32);
            }
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
    int ngroups;
    int nb;                     // windows
    int nbmax;                  // max nb_rows over the groups (shared-memory geometry)
    int a_sep;                  // some group's A tile lies outside its B range: blocks carry a separate 128-row A region
    int nstages;                // operand ring depth (a multiple of xg)
    int nraw;                   // raw plane-word ring depth (a multiple of xg)
    int xg;                     // expanding groups at work (3, or 2 when only two operand stages fit)
    int64_t nchunks;            // chunks of the plane (a stage of CH chunks may reach past the last one: clamped, masked to 0)
    int32_t* out;               // [nb][Hk][Hk], upper triangle (i <= j) only
};

// Warp roles of the persistent CTA (one per SM), geometry <GW, EW>:
//   3 GW warps  expand: three groups of GW warps (GW / 4 per scheduler each); group k owns the stages k, k+3, k+6, ... so
//               that while one group waits (shared-memory loads, the proxy fence) the other two keep the ALUs busy
//   1 warp      TMEM allocation + MMA issue (warp-uniform loop, an elected lane issues)
//   3 warps     TMA: warp t brings the plane words of group t's stages into group t's raw slots
//   EW warps    epilogue: TMEM -> registers -> global (EW = 4 or 8; with 8, warps w and w+4 share TMEM lane quarter w % 4)
// <8, 4> = 32 warps (the default: 64 registers), <4, 8> = 24 warps (80 registers; PG_K2T_GW=4).
constexpr int GRAM_XGROUPS = 3;
constexpr int gram_threads(int GW, int EW) { return (GW * GRAM_XGROUPS + 1 + GRAM_XGROUPS + EW) * 32; }
constexpr int GRAM_MAX_STAGES = 9;
constexpr int GRAM_MAX_RAW = 48;           // depth of the raw plane-word ring (TMA runs this many chunks ahead)

// 16 bits -> 16 bytes of 0/1 (byte k = bit k): 4 bits -> 4 bytes is one IMAD + LOP3.  (A 256-entry shared-memory table,
// 8 bits -> 8 bytes per LDS.64, was measured and dropped: the kernel is short of shared-memory bandwidth — the SS-mode MMAs
// read (128 + N) x 32 bytes per instruction — not of integer issue slots; gram_diff went from 1.07 to 1.35 ms with it.)
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
// one lane of a converged warp (the MMA warp runs its loop with all 32 lanes so that addresses and descriptors stay in
// uniform registers; only the elected lane issues)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// position in a ring of `n` slots + the parity of the current pass; advances by small steps
struct RingPos {
    int s;
    uint32_t ph;
    __device__ __forceinline__ void advance(int delta, int n) {
        s += delta;
        while (s >= n) {
            s -= n;
            ph ^= 1u;
        }
    }
};

// work item j of this launch -> (window, group) and the chunk range of the window in plane coordinates
struct GramItem {
    GramGroup g;
    int wb;
    int64_t lo, hi, c_first;
    int nst;
};
template <int NPL, int CH>
__device__ __forceinline__ GramItem gram_item(const GramParams& gp, int64_t j) {
    constexpr int SH = 6 + (CH == 4 ? 2 : (CH == 2 ? 1 : 0));   // a stage covers CH chunks of 64 (pseudo-)sites
    GramItem it;
    it.wb = (int)(j / gp.ngroups);
    it.g = gp.groups[j - (int64_t)it.wb * gp.ngroups];
    it.lo = gp.win_lo[it.wb] - gp.site_base;
    it.hi = gp.win_hi[it.wb] - gp.site_base;
    if (NPL == 2) {
        it.lo = gp.cps[it.lo];
        it.hi = gp.cps[it.hi];
    }
    it.c_first = it.lo >> SH;                           // first STAGE of the window
    it.nst = (it.hi > it.lo) ? (int)(((it.hi - 1) >> SH) - it.c_first + 1) : 0;
    return it;
}

// Shared memory: [raw ring: nraw slots of NPL x RROWS plane words, filled by 1-D TMA bulk copies]
//                [operand ring: nstages stages of 2 K steps x NPL planes x RROWS rows x 32 bytes] [+ slack]
// RROWS = (128 rows of a separate A tile, only when some group needs one) + nbmax rows of the B range.
// CH = chunks per stage (1, 2 or 4): every per-stage hand-over (TMA wait, proxy fence, MMA issue, commit) serves CH x 64 sites.
template <int NPL, int GW, int EW, int CH = 1>
__global__ void __launch_bounds__(gram_threads(GW, EW), 1) k2t_gram(const __grid_constant__ GramParams gp) {
    static_assert(CH == 1 || CH == 2 || CH == 4, "chunks per stage");
    constexpr int GRAM_XWARPS = GW * GRAM_XGROUPS, GTHREADS = GW * 32;
    constexpr int GRAM_WARP_MMA = GRAM_XWARPS, GRAM_WARP_TMA = GRAM_XWARPS + 1, GRAM_WARP_EPI = GRAM_WARP_TMA + GRAM_XGROUPS;
    constexpr int GRAM_EPI_WARPS = EW;
    constexpr int GRAM_MAX_ITEMS = ((128 + 512) * NPL + GTHREADS - 1) / GTHREADS;    // plane rows per expanding thread and stage
    static_assert(GW % 4 == 0 && (EW == 4 || EW == 8) , "warp roles");
    extern __shared__ __align__(128) uint8_t gsm[];
    __shared__ __align__(8) uint64_t full[GRAM_MAX_STAGES], empty[GRAM_MAX_STAGES], raw_full[GRAM_MAX_RAW],
        raw_empty[GRAM_MAX_RAW], tmem_full, tmem_empty;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NS = gp.nstages, RD = gp.nraw;
    const int AOFF = gp.a_sep ? 128 : 0;                // rows of the separate A region in front of the B rows
    const int RROWS = AOFF + gp.nbmax;                  // rows of one plane in a raw slot / operand block
    const int RAW1 = NPL * RROWS * 8;                   // plane words of ONE chunk in a raw slot
    const int RAW = CH * RAW1;                          // bytes of one raw slot
    const int BLK = RROWS * 32;                         // one (K step, plane) operand block
    const int STAGE = CH * 2 * NPL * BLK;
    uint8_t* const raw_base = gsm;
    uint8_t* const op_base = gsm + (size_t)RD * RAW;

    // contiguous range of work items (window-major, groups of a window adjacent) of this CTA
    const int64_t n_items = (int64_t)gp.nb * gp.ngroups;
    const int64_t j0 = n_items * blockIdx.x / gridDim.x, j1 = n_items * (blockIdx.x + 1) / gridDim.x;

    if (warp == GRAM_WARP_MMA) {
        if (lane == 0) {
            for (int s = 0; s < NS; ++s) {
                mbar_init(&full[s], GW);                // the warps of the expanding group that owns the slot
                mbar_init(&empty[s], 1);
            }
            for (int s = 0; s < RD; ++s) {
                mbar_init(&raw_full[s], 1);
                mbar_init(&raw_empty[s], GW);
            }
            mbar_init(&tmem_full, 1);
            mbar_init(&tmem_empty, GRAM_EPI_WARPS);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp >= GRAM_WARP_TMA && warp < GRAM_WARP_EPI) {
        // ---------------- TMA: plane words of every chunk of every item -> raw ring ----------------
        // One warp per expanding group: warp t loads exactly the stages group t expands (gs % XG == t) into that group's raw
        // slots, walking the same slot / parity sequence as the group.  (ONE thread issuing every stage's copies — ~60
        // dependent instructions per stage — was what paced the whole kernel; the loop runs warp-uniform and an elected lane
        // issues, so addresses stay in uniform registers.)
        const int XG = gp.xg;
        const int t = warp - GRAM_WARP_TMA;
        if (t < XG) {
            const int RM = RD / XG;
            int n_done = 0, rm = 0;
            uint32_t rph = 0;
            int64_t gbase = 0;
            for (int64_t j = j0; j < j1; ++j) {
                const GramItem im = gram_item<NPL, CH>(gp, j);
                const bool a_in_b = (im.g.a_row0 == im.g.b_row0);
                const int a_rows = a_in_b ? 0 : min(128, gp.R - im.g.a_row0);
                const uint32_t bytes_a = (uint32_t)a_rows * 8u, bytes_b = (uint32_t)im.g.nb_rows * 8u;
                int it = (int)(((int64_t)t - gbase % XG + XG) % XG);
                for (; it < im.nst; it += XG) {
                    const int rslot = t + XG * rm;
                    if (n_done >= RM) mbar_wait(&raw_empty[rslot], rph ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(&raw_full[rslot], CH * NPL * (bytes_a + bytes_b));
#pragma unroll
                        for (int h = 0; h < CH; ++h) {
                            // a chunk past the end of the plane is read from the last one; its mask is 0
                            const int64_t chunk = min((im.c_first + it) * CH + h, gp.nchunks - 1);
#pragma unroll
                            for (int pl = 0; pl < NPL; ++pl) {
                                const uint64_t* src = gp.plane + (chunk * NPL + pl) * gp.R;
                                uint8_t* dst = raw_base + (size_t)rslot * RAW + (size_t)h * RAW1 + (size_t)pl * RROWS * 8;
                                if (bytes_a) bulk_g2s(dst, src + im.g.a_row0, bytes_a, &raw_full[rslot]);
                                bulk_g2s(dst + AOFF * 8, src + im.g.b_row0, bytes_b, &raw_full[rslot]);
                            }
                        }
                    }
                    __syncwarp();
                    ++n_done;
                    if (++rm == RM) {
                        rm = 0;
                        rph ^= 1u;
                    }
                }
                gbase += im.nst;
            }
        }
    } else if (warp < GRAM_XWARPS) {
        // ---------------- expand: plane words -> 0/1 bytes in the core-matrix layout ----------------
        // Group xg owns the global stages gs with gs % XG == xg.  NS and RD are multiples of XG, so a group always cycles
        // through the same operand slots (xg, xg + XG, ...) and the same raw slots in the same order: every barrier it waits on
        // is one whose previous phase the same group consumed, and a parity wait can never alias with an older phase.
        // Several slots per group matter: a slot comes back only after the MMAs that read it have completed (the tensor
        // pipe's latency), and with one slot per group that latency sat in every group's critical path.
        const int XG = gp.xg;
        const int xg = warp / GW, xt = tid % GTHREADS;     // expanding group, thread inside the group
        const int SM_ = NS / XG, RM = RD / XG;             // operand / raw slots of this group
        int n_done = 0;                                    // stages this group has processed: stage n is gs = xg + n XG
        int sm = 0, rm = 0;                                // n_done % SM_, n_done % RM
        uint32_t sph = 0, rph = 0;                         // (n_done / SM_) & 1, (n_done / RM) & 1
        int64_t gbase = 0;                                 // global stage index of the item's first stage
        for (int64_t j = j0; j < j1 && xg < XG; ++j) {
            const GramItem im = gram_item<NPL, CH>(gp, j);
            const bool a_in_b = (im.g.a_row0 == im.g.b_row0);
            // rows expanded per plane: [separate A tile (128 rows)] + B range
            const int skip_a = a_in_b ? 128 : 0;
            const int rows_tot = 128 + im.g.nb_rows - skip_a;
            const int nitems = rows_tot * NPL;
            int r_idx[GRAM_MAX_ITEMS], d_off[GRAM_MAX_ITEMS];      // raw word index (-1 none, -2 zero row) / byte offset in a block
#pragma unroll
            for (int q = 0; q < GRAM_MAX_ITEMS; ++q) {
                const int item = xt + q * GTHREADS;
                r_idx[q] = -1;
                d_off[q] = 0;
                if (item < nitems) {
                    const int pl = (NPL == 2 && item >= rows_tot) ? 1 : 0;
                    const int rr = item - pl * rows_tot + skip_a;          // < 128: row of the separate A tile
                    const int x = (rr < 128) ? rr : rr - 128;
                    const int row = (rr < 128) ? x : AOFF + x;             // row inside the block / raw slot
                    r_idx[q] = (rr < 128 && im.g.a_row0 + x >= gp.R) ? -2 : pl * RROWS + row;
                    d_off[q] = pl * BLK + (row >> 3) * 256 + (row & 7) * 16;
                }
            }
            // first local stage of this group: (gbase + it) % XG == xg
            int it = (int)(((int64_t)xg - gbase % XG + XG) % XG);
            for (; it < im.nst; it += XG) {
                const int rslot = xg + XG * rm, sslot = xg + XG * sm;
                uint64_t mask[CH];
#pragma unroll
                for (int h = 0; h < CH; ++h) {
                    const int64_t b0 = ((im.c_first + it) * CH + h) << 6;
                    uint64_t m = 0ull;                                  // a chunk outside the window (CH > 1 only)
                    if (b0 < im.hi && b0 + 64 > im.lo) {
                        m = ~0ull;
                        if (im.lo > b0) m &= ~0ull << (int)(im.lo - b0);
                        if (im.hi < b0 + 64) m &= ~0ull >> (int)(b0 + 64 - im.hi);
                    }
                    mask[h] = m;
                }
                mbar_wait(&raw_full[rslot], rph);
                const uint64_t* rw = reinterpret_cast<const uint64_t*>(raw_base + (size_t)rslot * RAW);
                uint64_t v[GRAM_MAX_ITEMS][CH];
#pragma unroll
                for (int q = 0; q < GRAM_MAX_ITEMS; ++q)
#pragma unroll
                    for (int h = 0; h < CH; ++h) v[q][h] = (r_idx[q] >= 0) ? (rw[h * (RAW1 / 8) + r_idx[q]] & mask[h]) : 0ull;
                if (n_done >= SM_) mbar_wait(&empty[sslot], sph ^ 1u);
                uint8_t* sb = op_base + (size_t)sslot * STAGE;
#pragma unroll
                for (int q = 0; q < GRAM_MAX_ITEMS; ++q) {
                    if (q * GTHREADS < nitems) {            // warp-uniform: no instructions for item slots nobody uses
                        if (r_idx[q] != -1) {
#pragma unroll
                            for (int h = 0; h < CH; ++h) {
                                const uint32_t wlo = (uint32_t)v[q][h], whi = (uint32_t)(v[q][h] >> 32);
                                uint8_t* d0 = sb + (size_t)h * 2 * NPL * BLK + d_off[q];      // K steps 2h, 2h + 1
                                uint8_t* d1 = d0 + NPL * BLK;
                                *reinterpret_cast<uint4*>(d0) = expand16(wlo & 0xffffu);
                                *reinterpret_cast<uint4*>(d0 + 128) = expand16(wlo >> 16);
                                *reinterpret_cast<uint4*>(d1) = expand16(whi & 0xffffu);
                                *reinterpret_cast<uint4*>(d1 + 128) = expand16(whi >> 16);
                            }
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&full[sslot]);
                    // released only now: the stores above consumed the words, so the loads from the slot have completed
                    // before the TMA (async proxy) may overwrite it — an arrive right after issuing the loads raced with
                    // the refill
                    mbar_arrive(&raw_empty[rslot]);
                }
                ++n_done;
                if (++rm == RM) {
                    rm = 0;
                    rph ^= 1u;
                }
                if (++sm == SM_) {
                    sm = 0;
                    sph ^= 1u;
                }
            }
            gbase += im.nst;
        }
    } else if (warp == GRAM_WARP_MMA) {
        // ---------------- MMA issue ----------------
        // The whole warp runs the loop (waits included) so that everything stays warp-uniform — the compiler keeps ring
        // positions, shared-memory addresses and descriptors in uniform registers, which is what UTCIMMA takes; one elected lane
        // issues.  Everything that does not change per stage is hoisted and a descriptor is one 32-bit add (the shared-memory
        // address field sits in the low word).  (Earlier versions ran the loop in lane 0 alone: ~450, then ~110 dependent
        // instructions per stage through R2UR moves — the issuing thread, not the tensor pipe, set the pace.)
        {
            const uint32_t sbase16 = smem_u32(op_base) >> 4;
            const uint32_t DLO = (128u >> 4) << 16;                       // LBO
            const uint32_t DHI = (256u >> 4) | (1u << 14);                // SBO | descriptor version 1
            const uint32_t stage16 = (uint32_t)STAGE >> 4, blk16 = (uint32_t)BLK >> 4;
            auto D = [&](uint32_t a16) { return ((uint64_t)DHI << 32) | (uint64_t)(DLO + a16); };
            RingPos sp = {0, 0u};
            int64_t k = 0;                                  // items done by this CTA
            for (int64_t j = j0; j < j1; ++j, ++k) {
                const GramItem im = gram_item<NPL, CH>(gp, j);
                // A tile inside the B rows for a diagonal group, else the separate A region in front of them
                const uint32_t a16 = (im.g.a_row0 == im.g.b_row0) ? (uint32_t)AOFF * 2u : 0u;
                const uint32_t b16 = (uint32_t)AOFF * 2u;                 // 32 bytes per row = 2 units of 16 bytes
                const int n_a = min(256, im.g.nb_rows), n_b = im.g.nb_rows - n_a;
                const uint32_t id_a = umma_idesc(n_a), id_b = umma_idesc(n_b > 0 ? n_b : 16);
                if (k > 0) {                                // the epilogue must have drained the previous accumulators
                    mbar_wait(&tmem_empty, (uint32_t)((k - 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int it = 0; it < im.nst; ++it) {
                    mbar_wait(&full[sp.s], sp.ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st16 = sbase16 + (uint32_t)sp.s * stage16;
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < 2 * CH; ++ks) {
                            const uint32_t acc0 = (it > 0 || ks > 0) ? 1u : 0u;
                            if (NPL == 1) {
                                const uint32_t blk = st16 + ks * blk16;
                                umma_i8(tmem, D(blk + a16), D(blk + b16), id_a, acc0);
                                if (n_b > 0) umma_i8(tmem + 256, D(blk + a16), D(blk + b16 + 512), id_b, acc0);
                            } else {
                                const uint32_t bp = st16 + (ks * 2) * blk16, bq = bp + blk16;
                                umma_i8(tmem, D(bp + a16), D(bq + b16), id_a, acc0);
                                umma_i8(tmem, D(bq + a16), D(bp + b16), id_a, 1u);
                                if (n_b > 0) {
                                    umma_i8(tmem + 256, D(bp + a16), D(bq + b16 + 512), id_b, acc0);
                                    umma_i8(tmem + 256, D(bq + a16), D(bp + b16 + 512), id_b, 1u);
                                }
                            }
                        }
                        umma_commit(&empty[sp.s]);  // arrives when the MMAs above have read the stage
                    }
                    __syncwarp();
                    sp.advance(1, NS);
                }
                if (elect_one()) umma_commit(&tmem_full);   // arrives when every MMA of the item has completed
                __syncwarp();
            }
        }
    } else {
        // ---------------- epilogue: TMEM -> registers -> symmetric int32 matrix ----------------
        // Lane = matrix row in TMEM (a warp may touch lanes 32 (warp % 4) ..): the two warps of a quarter take alternate
        // 32-column blocks; a lane stores its 32 consecutive columns (128 contiguous bytes of its row) with 16-byte stores.
        // Only [i][j] with i in the A tile and j in the B range is written — the upper triangle of the symmetric matrix
        // (readers index it through (min, max)).
        const int ew = warp - GRAM_WARP_EPI;
        const int qd = warp & 3;
        const bool vec_ok = (gp.Hk & 3) == 0;              // rows are 16-byte aligned: a lane stores its 32 columns as 8 x 16 bytes
        int64_t k = 0;
        for (int64_t j = j0; j < j1; ++j, ++k) {
            const GramItem im = gram_item<NPL, CH>(gp, j);
            mbar_wait(&tmem_full, (uint32_t)(k & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int i = im.g.a_row0 + qd * 32 + lane;     // this lane's matrix row
            int32_t* orow = gp.out + (size_t)im.wb * gp.Hk * gp.Hk + (size_t)i * gp.Hk;
            const int cfirst = (ew >> 2) * 32;
            int c_last = cfirst;                            // last block this warp reads
            constexpr int CSTEP = (EW / 4) * 32;          // the warps of a lane quarter take alternate 32-column blocks
            while (c_last + CSTEP < im.g.nb_rows) c_last += CSTEP;
            if (cfirst >= im.g.nb_rows) {                   // nothing to read: release the accumulators right away
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty);
            }
            for (int c0 = cfirst; c0 < im.g.nb_rows; c0 += CSTEP) {
                uint32_t v[32];
                if (im.nst > 0) {
                    const uint32_t taddr = tmem + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,"
                        "%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
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
                if (c0 == c_last) {                         // last read of the accumulators: the next item's MMAs may start
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty);
                }
                if (i < gp.Hk) {
                    const int jb = im.g.b_row0 + c0;        // first column of the block (multiple of 16)
                    if (vec_ok) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4)
                            if (jb + e + 4 <= gp.Hk)
                                *reinterpret_cast<uint4*>(orow + jb + e) = make_uint4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (jb + e < gp.Hk) orow[jb + e] = (int32_t)v[e];
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == GRAM_WARP_MMA) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
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

__global__ void k2t_half(int32_t* p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i >> 1;
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
    PG_CHECK((size_t)16 * pw * 4 + (size_t)R * 8 <= 96 * 1024, "pairwise path: %d haplotype columns are too many for the plane builders", pitch);
    {
        static bool attr_dev[64] = {};
        if (!attr_dev[ctx->device & 63]) {
            PG_CUDA(cudaFuncSetAttribute(k2t_build_pq, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            PG_CUDA(cudaFuncSetAttribute(k2t_valid_class<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            PG_CUDA(cudaFuncSetAttribute(k2t_valid_class<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
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
    const bool want_pairs = (Hk % 2 == 0) && !getenv("PG_K2T_NO_PAIRS");
    const int R2 = (Hk / 2 + 15) / 16 * 16;
    const size_t o_v = carve((size_t)nchunk * R * 8), o_cls = carve(span), o_tot = carve((size_t)nchunk * 4),
                 o_off = carve((size_t)(nchunk + 2) * 4), o_cps = carve((span + 1) * 4),
                 o_vp = carve(want_pairs ? (size_t)nchunk * R2 * 8 : 0), o_mid = carve((size_t)Hk * 4);
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
    vp.vpair = want_pairs ? (uint64_t*)(base + o_vp) : nullptr;
    vp.R2 = R2;
    vp.pair_flag = d_off + nchunk + 1;
    PG_CUDA(cudaMemsetAsync(vp.pair_flag, 0, 4, ctx->stream));
    const int grid1 = (int)std::min<int64_t>(nchunk, (int64_t)ctx->sm_count * 8);
    {
        const int ti = pg_time_begin(ctx, "k2t_valid_class");
        bool all_used = true;                    // unselected real columns? (padding columns hold 0 = missing and never count)
        for (int c = 0; c < ctx->H; ++c) all_used = all_used && c2r[c] >= 0;
        if (all_used) k2t_valid_class<true><<<grid1, 256, (size_t)8 * pw * 4 + (size_t)R * 8, ctx->stream>>>(vp);
        else k2t_valid_class<false><<<grid1, 256, (size_t)8 * pw * 4 + (size_t)R * 8, ctx->stream>>>(vp);
        pg_time_end(ctx, ti);
        PG_CUDA(cudaGetLastError());
    }
    {
        const int ti = pg_time_begin(ctx, "k2t_scan");
        k2t_scan<<<1, 1024, 0, ctx->stream>>>(vp.chunk_tot, d_off, nchunk);
        pg_time_end(ctx, ti);
        PG_CUDA(cudaGetLastError());
    }
    int32_t tf[2] = {0, 0};                       // pseudo-sites, "some sample's haplotypes differ in missingness"
    PG_CUDA(cudaMemcpyAsync(tf, d_off + nchunk, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    const int32_t total = tf[0];
    const bool pairs_ok = want_pairs && tf[1] == 0;
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
    // mask ids for the epilogues: row r -> r / 2 when the valid words are shared by consecutive rows
    ps.Hm = pairs_ok ? Hk / 2 : Hk;
    ps.R2 = R2;
    ps.vpair = pairs_ok ? vp.vpair : nullptr;
    if (pairs_ok) {
        int32_t* d_mid = (int32_t*)(base + o_mid);
        k2t_half<<<(Hk + 255) / 256, 256, 0, ctx->stream>>>(d_mid, Hk);
        ps.d_mid = d_mid;
    } else {
        ps.d_mid = d_iota;
    }
    return PG_OK;
}

// diff [nb][Hk^2] and n [nb][Hk^2] for nb windows (absolute site ranges on the device)
namespace {
// tile groups of an R-row Gram: one 128-row A tile x up to 512 B rows (TMEM has 512 int32 columns per SM)
void gram_groups(int R, std::vector<GramGroup>& groups, int& nbmax, int& a_sep) {
    groups.clear();
    nbmax = 16;
    a_sep = 0;
    for (int a0 = 0; a0 < R; a0 += 128)
        for (int c = a0; c < R; c += 512) {
            GramGroup g;
            g.a_row0 = a0;
            g.b_row0 = c;
            g.nb_rows = std::min(512, R - c);
            g.pad = 0;
            nbmax = std::max(nbmax, g.nb_rows);
            if (c != a0) a_sep = 1;
            groups.push_back(g);
        }
}
}  // namespace

int pg_k2t_pairs(pg_ctx* ctx, const K2TPlanes& ps, const int64_t* d_lo, const int64_t* d_hi, int nb, int32_t* d_diff,
                 int32_t* d_n) {
    static bool attr_dev[64] = {};
    if (!attr_dev[ctx->device & 63]) {
        PG_CUDA(cudaFuncSetAttribute(k2t_gram<1, 4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        PG_CUDA(cudaFuncSetAttribute(k2t_gram<2, 4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        PG_CUDA(cudaFuncSetAttribute(k2t_gram<1, 8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        PG_CUDA(cudaFuncSetAttribute(k2t_gram<2, 8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        PG_CUDA(cudaFuncSetAttribute(k2t_gram<1, 8, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        PG_CUDA(cudaFuncSetAttribute(k2t_gram<1, 8, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        attr_dev[ctx->device & 63] = true;
    }
    // expanding groups of 8 warps + 4 epilogue warps (30 warps) or 4 + 8 (22 warps): PG_K2T_GW = 4 | 8, per kernel
    // PG_K2T_GW_N / PG_K2T_GW_D
    auto gw_of = [](const char* name, int dflt) {
        const char* e = getenv(name);
        if (!e) e = getenv("PG_K2T_GW");
        return e ? atoi(e) : dflt;
    };
    const bool wide_n = gw_of("PG_K2T_GW_N", 8) == 8, wide_d = gw_of("PG_K2T_GW_D", 8) == 8;
    // n_ij over the mask rows (one per sample when the haplotypes of a sample share their missingness), diff_ij over all rows
    const int Rn = ps.vpair ? ps.R2 : ps.R;
    std::vector<GramGroup> gn, gd;
    int nbmax_n, asep_n, nbmax_d, asep_d;
    gram_groups(Rn, gn, nbmax_n, asep_n);
    gram_groups(ps.R, gd, nbmax_d, asep_d);
    PG_TRY(ctx->misc4.ensure((gn.size() + gd.size()) * sizeof(GramGroup) + 64));
    GramGroup* d_gn = (GramGroup*)ctx->misc4.p;
    GramGroup* d_gd = d_gn + gn.size();
    PG_CUDA(cudaMemcpyAsync(d_gn, gn.data(), gn.size() * sizeof(GramGroup), cudaMemcpyHostToDevice, ctx->stream));
    PG_CUDA(cudaMemcpyAsync(d_gd, gd.data(), gd.size() * sizeof(GramGroup), cudaMemcpyHostToDevice, ctx->stream));
    GramParams gp;
    gp.site_base = ps.site_base;
    gp.win_lo = d_lo;
    gp.win_hi = d_hi;
    gp.nb = nb;
    // shared memory: an operand ring of (ideally) one stage per expanding group + the raw plane-word ring + 4 KB slack (the
    // 128-row A tile of a diagonal group may reach past a short B range)
    const int budget = 222 * 1024;            // + 2.3 KB of static shared memory (barriers, the expansion table) <= 227 KB
    const int fixed = 4096;
    auto geometry = [&](int npl, int nbmax, int a_sep, int& nstages, int& nraw, int& xg, int ch = 1) {
        const int rrows = (a_sep ? 128 : 0) + nbmax;
        const int stage = ch * 2 * npl * rrows * 32, raw = ch * npl * rrows * 8;
        // operand stages: 9, 6 or 3 (three expanding groups with 3 / 2 / 1 slots each), else 2 (two groups); raw slots a
        // multiple of the group count too, so that every slot has ONE consumer group
        const int avail = budget - fixed;
        nstages = 2;
        for (int cand : {9, 6, 3})
            if (cand * stage + 3 * raw <= avail) {
                nstages = cand;
                break;
            }
        if (const char* e = getenv("PG_K2T_NSTAGES")) nstages = std::max(1, std::min(nstages, atoi(e)));
        xg = (nstages % 3 == 0) ? 3 : (nstages % 2 == 0 ? 2 : 1);
        int mult = std::min(GRAM_MAX_RAW / xg, (avail - nstages * stage) / (raw * xg));
        if (const char* e = getenv("PG_K2T_NRAW")) mult = std::min(mult, atoi(e));
        nraw = std::max(1, mult) * xg;
        return (size_t)nstages * stage + (size_t)nraw * raw + fixed;
    };
    {   // persistent CTAs: one per SM, each works through a contiguous range of (window, group) items
        gp.R = Rn;
        gp.Hk = ps.Hm;
        gp.groups = d_gn;
        gp.ngroups = (int)gn.size();
        gp.nbmax = nbmax_n;
        gp.a_sep = asep_n;
        // 256-site stages (128 / 64 where three of the larger ones do not fit): every site counts for n_ij, so its K is the
        // long one and the per-stage hand-overs — TMA wait, proxy fence, MMA issue, commit — are what paces the kernel:
        // 0.70 (64) -> 0.55 (128) -> 0.50 ms (256 sites per stage) on the C2 shape, bit-identical sums.  PG_K2T_CH = 1 | 2 | 4.
        int ch_n = wide_n ? 4 : 1;
        if (const char* e = getenv("PG_K2T_CH")) {
            const int v = atoi(e);
            ch_n = (v == 4 && wide_n) ? 4 : ((v == 2 && wide_n) ? 2 : 1);
        }
        size_t smem = geometry(1, nbmax_n, asep_n, gp.nstages, gp.nraw, gp.xg, ch_n);
        while (ch_n > 1 && gp.nstages < 3) {
            ch_n /= 2;
            smem = geometry(1, nbmax_n, asep_n, gp.nstages, gp.nraw, gp.xg, ch_n);
        }
        gp.plane = ps.vpair ? ps.vpair : ps.vplane;
        gp.nchunks = ps.nchunk_v;
        gp.cps = nullptr;
        gp.out = d_n;
        const unsigned grid = (unsigned)std::min<int64_t>((int64_t)nb * gp.ngroups, ctx->sm_count);
        const int ti = pg_time_begin(ctx, "k2t_gram_n");
        if (ch_n == 4) k2t_gram<1, 8, 4, 4><<<grid, gram_threads(8, 4), smem, ctx->stream>>>(gp);
        else if (ch_n == 2) k2t_gram<1, 8, 4, 2><<<grid, gram_threads(8, 4), smem, ctx->stream>>>(gp);
        else if (wide_n) k2t_gram<1, 8, 4><<<grid, gram_threads(8, 4), smem, ctx->stream>>>(gp);
        else k2t_gram<1, 4, 8><<<grid, gram_threads(4, 8), smem, ctx->stream>>>(gp);
        pg_time_end(ctx, ti);
        PG_CUDA(cudaGetLastError());
    }
    {
        gp.R = ps.R;
        gp.Hk = ps.Hk;
        gp.groups = d_gd;
        gp.ngroups = (int)gd.size();
        gp.nbmax = nbmax_d;
        gp.a_sep = asep_d;
        // (128-pseudo-site stages do not fit here: two planes per stage, 3 x 102 KB for 400 rows)
        const size_t smem = geometry(2, nbmax_d, asep_d, gp.nstages, gp.nraw, gp.xg);
        gp.plane = ps.pq;
        gp.nchunks = (ps.npseudo + 63) / 64;
        gp.cps = ps.cps;
        gp.out = d_diff;
        const unsigned grid = (unsigned)std::min<int64_t>((int64_t)nb * gp.ngroups, ctx->sm_count);
        const int ti = pg_time_begin(ctx, "k2t_gram_diff");
        if (wide_d) k2t_gram<2, 8, 4><<<grid, gram_threads(8, 4), smem, ctx->stream>>>(gp);
        else k2t_gram<2, 4, 8><<<grid, gram_threads(4, 8), smem, ctx->stream>>>(gp);
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
