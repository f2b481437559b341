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
        for (int cw = lane; cw < p.pw; cw += 32) {
            const uint32_t cm = p.cmask[cw];
            uint32_t w[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int64_t s = site0 + k;
                w[k] = (s < p.S) ? __ldg(p.geno32 + s * p.pw + cw) : 0u;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) pres[k] |= w[k] & cm;
            vc_st[o * p.pw + cw] = valid2(w[0], w[1]) | (valid2(w[2], w[3]) << 2) | (valid2(w[4], w[5]) << 4) |
                                   (valid2(w[6], w[7]) << 6);
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

// exclusive scan of the chunk totals (single CTA; 1e8 sites = 1.6 M chunks = 1.6 k iterations)
__global__ void __launch_bounds__(1024) k2t_scan(const int32_t* __restrict__ tot, int32_t* __restrict__ off, int64_t n) {
    __shared__ int wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int carry = 0;
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + tid;
        const int v = (i < n) ? tot[i] : 0;
        int incl = v;
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
        const int wbase = warp ? wsum[warp - 1] : 0;
        if (i < n) off[i] = carry + wbase + incl - v;
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
        for (int cw = lane; cw < p.pw; cw += 32) {
            uint32_t outp = 0u, outq = 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t w = row[k] ? __ldg(row[k] + cw) : 0u;
                outp |= ((w >> psh[k]) & 0x01010101u) << k;
                uint32_t t = w & qm[k];
                t |= t >> 4;
                t |= t >> 2;
                outq |= (t & 0x01010101u) << k;
            }
            pq_st[(0 * 8 + o) * p.pw + cw] = outp;
            pq_st[(1 * 8 + o) * p.pw + cw] = outq;
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
    int nstages;
    int32_t* out;               // [nb][Hk][Hk]
};

constexpr int GRAM_PRODUCERS = 128;
constexpr int GRAM_THREADS = 160;          // warps 0-3: producers, then epilogue; warp 4: TMEM allocation + MMA issue
constexpr int GRAM_MAX_STAGES = 4;
constexpr int GRAM_MAX_ITEMS = (128 + 512) * 2 / GRAM_PRODUCERS;    // plane words per producer thread and stage

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

template <int NPL>
__global__ void __launch_bounds__(GRAM_THREADS, 1) k2t_gram(const __grid_constant__ GramParams gp) {
    extern __shared__ __align__(128) uint8_t gsm[];
    __shared__ __align__(8) uint64_t full[GRAM_MAX_STAGES], empty[GRAM_MAX_STAGES], done;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const GramGroup g = gp.groups[blockIdx.x];
    const int wb = blockIdx.y;
    const int NS = gp.nstages;
    const int BLK = (128 + gp.nbmax) * 32;              // one (K step, plane) block: A region 128 rows, then the B region
    const int STAGE = 2 * NPL * BLK;

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
    if (warp == 4) {
        if (lane == 0) {
            for (int s = 0; s < NS; ++s) {
                mbar_init(&full[s], GRAM_PRODUCERS);
                mbar_init(&empty[s], 1);
            }
            mbar_init(&done, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp < 4) {
        // ---------------- producers: plane words -> 0/1 bytes in the core-matrix layout ----------------
        const int rows_tot = 128 + g.nb_rows;
        const int nitems = rows_tot * NPL;
        for (int it = 0; it < nst; ++it) {
            const int s = it % NS;
            if (it >= NS) mbar_wait(&empty[s], (uint32_t)(((it / NS) - 1) & 1));
            const int64_t chunk = c_first + it;
            uint64_t mask = ~0ull;
            {
                const int64_t b0 = chunk << 6;
                if (lo > b0) mask &= ~0ull << (int)(lo - b0);
                if (hi < b0 + 64) mask &= ~0ull >> (int)(b0 + 64 - hi);
            }
            uint64_t v[GRAM_MAX_ITEMS];
#pragma unroll
            for (int q = 0; q < GRAM_MAX_ITEMS; ++q) {
                const int item = tid + q * GRAM_PRODUCERS;
                v[q] = 0ull;
                if (item < nitems) {
                    const int pl = (NPL == 2 && item >= rows_tot) ? 1 : 0;
                    const int rr = item - pl * rows_tot;
                    const int row = (rr < 128) ? g.a_row0 + rr : g.b_row0 + rr - 128;
                    if (row < gp.R) v[q] = __ldg(gp.plane + (chunk * NPL + pl) * gp.R + row) & mask;
                }
            }
            uint8_t* sb = gsm + (size_t)s * STAGE;
#pragma unroll
            for (int q = 0; q < GRAM_MAX_ITEMS; ++q) {
                const int item = tid + q * GRAM_PRODUCERS;
                if (item < nitems) {
                    const int pl = (NPL == 2 && item >= rows_tot) ? 1 : 0;
                    const int rr = item - pl * rows_tot;
                    const int x = (rr < 128) ? rr : rr - 128;
                    const int ro = ((rr < 128) ? 0 : 4096) + (x >> 3) * 256 + (x & 7) * 16;
                    const uint32_t wlo = (uint32_t)v[q], whi = (uint32_t)(v[q] >> 32);
                    uint8_t* d0 = sb + (0 * NPL + pl) * BLK + ro;
                    uint8_t* d1 = sb + (1 * NPL + pl) * BLK + ro;
                    *reinterpret_cast<uint4*>(d0) = expand16(wlo & 0xffffu);
                    *reinterpret_cast<uint4*>(d0 + 128) = expand16(wlo >> 16);
                    *reinterpret_cast<uint4*>(d1) = expand16(whi & 0xffffu);
                    *reinterpret_cast<uint4*>(d1 + 128) = expand16(whi >> 16);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
            mbar_arrive(&full[s]);
        }
    } else if (lane == 0) {
        // ---------------- MMA issue (one thread) ----------------
        const uint32_t sbase = smem_u32(gsm);
        for (int it = 0; it < nst; ++it) {
            const int s = it % NS;
            mbar_wait(&full[s], (uint32_t)((it / NS) & 1));
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
                        umma_i8(tmem + n0, umma_desc(blk), umma_desc(blk + 4096 + n0 * 32), idesc, acc0);
                    } else {
                        const uint32_t bp = st + (ks * 2 + 0) * BLK, bq = st + (ks * 2 + 1) * BLK;
                        umma_i8(tmem + n0, umma_desc(bp), umma_desc(bq + 4096 + n0 * 32), idesc, acc0);
                        umma_i8(tmem + n0, umma_desc(bq), umma_desc(bp + 4096 + n0 * 32), idesc, 1u);
                    }
                }
            }
            umma_commit(&empty[s]);       // arrives when the MMAs above have read the stage
        }
        umma_commit(&done);
    }

    if (warp < 4) {
        // ---------------- epilogue: TMEM -> registers -> symmetric int32 matrix ----------------
        if (nst > 0) {
            mbar_wait(&done, 0u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const int i = g.a_row0 + warp * 32 + lane;
        int32_t* o = gp.out + (size_t)wb * gp.Hk * gp.Hk;
        for (int c0 = 0; c0 < g.nb_rows; c0 += 16) {
            uint32_t v[16];
            if (nst > 0) {
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = 0u;
            }
            if (i < gp.Hk) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int j = g.b_row0 + c0 + e;
                    if (j < gp.Hk) {
                        o[(size_t)i * gp.Hk + j] = (int32_t)v[e];
                        o[(size_t)j * gp.Hk + i] = (int32_t)v[e];
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
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
        PG_CUDA(cudaFuncSetAttribute(k2t_gram<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        PG_CUDA(cudaFuncSetAttribute(k2t_gram<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_dev[ctx->device & 63] = true;
    }
    const dim3 grid((unsigned)groups.size(), (unsigned)nb);
    {
        const int stage = 2 * 1 * (128 + nbmax) * 32;
        gp.nstages = std::max(2, std::min(GRAM_MAX_STAGES, (200 * 1024) / stage));
        gp.plane = ps.vplane;
        gp.cps = nullptr;
        gp.out = d_n;
        const int ti = pg_time_begin(ctx, "k2t_gram_n");
        k2t_gram<1><<<grid, GRAM_THREADS, (size_t)gp.nstages * stage, ctx->stream>>>(gp);
        pg_time_end(ctx, ti);
        PG_CUDA(cudaGetLastError());
    }
    {
        const int stage = 2 * 2 * (128 + nbmax) * 32;
        gp.nstages = std::max(2, std::min(GRAM_MAX_STAGES, (200 * 1024) / stage));
        gp.plane = ps.pq;
        gp.cps = ps.cps;
        gp.out = d_diff;
        const int ti = pg_time_begin(ctx, "k2t_gram_diff");
        k2t_gram<2><<<grid, GRAM_THREADS, (size_t)gp.nstages * stage, ctx->stream>>>(gp);
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
