// K1 — the HBM-bound "site pass": one read of the int8 genotype matrix produces per-population allele
// counts per site and reduces them into per-window sums.
//
//   mode POPGEN : closed-form pi / dxy / Fst sums (exact when a window has no partially-missing site;
//                 otherwise the window is flagged "ragged" and routed to K2)   -- genomics.py:956-995
//   mode ABBA   : ABBA / BABA / D / fd / fdM sums                             -- genomics.py:1647-1695
//   mode COUNTS : per-site per-population A,C,G,T counts                      -- genomics.py:1049-1052
//
// Data path: a persistent CTA per SM owns a contiguous range of tiles; a tile is T
// consecutive sites = one contiguous T*pitch byte range, brought into shared memory by 1-D TMA bulk copies
// (cp.async.bulk ... mbarrier::complete_tx) through a `stages`-deep ring.  One lane (or G lanes) owns one
// site row and walks it with conflict-free LDS.128; alleles are counted with SWAR byte-lane accumulators
// (8 integer ops per 4 genotypes).  Per-lane running sums are flushed per (warp, segment) into private
// slots — no atomics, deterministic — and a finalize kernel folds slots -> segments -> windows -> statistics.
#include <stdlib.h>

#include <algorithm>
#include <cmath>

#include "pgwin_internal.h"

namespace {

// consumer warps per CTA (+ one TMA producer warp): 8, or 12 where the register budget allows (65536 / 13 / 32 = 157)
constexpr int K1_MAX_WARPS = 12;

enum { MODE_POPGEN = 0, MODE_ABBA = 1, MODE_COUNTS = 2, MODE_POPGEN_FREQ = 3, MODE_FOURPOP = 4,   // FREQ = POPGEN + popFreq counters
       MODE_FOURPOP_Q = 5 };   // FOURPOP with the informative sites of a warp queued and evaluated 32 at a time (site pass only;
                               // experimental, PG_K1_FOURPOP_QUEUE: measured slower, see the comment in k1_site_pass)

struct K1Params {
    const uint8_t* geno;
    const int32_t* pos;
    int64_t site_begin, site_end;     // sites processed by this launch
    int64_t num_tiles;
    int pitch, G, I, T, wpt, stages, tile_bytes, nw;
    // hap -> pop tables (shared-memory copies are made at kernel start)
    const int32_t* ent_chunk;
    const uint4* ent_mask;
    int n_ent;
    int ent_lo[PG_MAX_K1_POPS], ent_hi[PG_MAX_K1_POPS], full_lo[PG_MAX_K1_POPS], full_hi[PG_MAX_K1_POPS];
    int popN[PG_MAX_K1_POPS];
    // segments / slots
    const int64_t* brk;
    int nseg;
    const int32_t* cta_seg_first;
    const int64_t* cta_slot_off;
    unsigned long long* part;
    // ABBA / FOURPOP: minimum non-missing count per population (exact integer form of n/N >= minData)
    int thr[4];
    int acc_limit;           // POPGEN: sites a lane may add to its 32-bit sums between flushes
    int lanepop;             // POPGEN: launch the lane-per-population variant (G == P lanes per site)
    int bytes;               // POPGEN: every population has <= 255 haplotypes -> byte-packed counts (IDP.4A statistics)
    int variant;             // FOURPOP allele choice: 0 = third of argsort (the rarer allele), 1 = polarize, 2 = fixed
    // COUNTS
    uint16_t* counts_out;
    int64_t counts_stride;   // uint16 elements per site
    int counts_pops;         // populations actually written (<= P)
};

// ---- PTX helpers: mbarrier + 1-D TMA bulk copy -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
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

// ---- allele counting on the resident one-hot code (A 0x01, C 0x04, G 0x10, T 0x40, missing 0x00) ----------
// level 1: three words are added -> 2-bit fields hold 0..3
// level 2: split into 4-bit fields (lo = A | G<<4, hi = C | T<<4 per byte), up to 5 level-1 sums
// level 3: split into byte lanes, up to 17 level-2 flushes, then __dp4a folds the 4 byte lanes
struct Tally {
    uint32_t nlo, nhi;            // nibble fields
    uint32_t bA, bC, bG, bT;      // byte lanes
    uint32_t tA, tC, tG, tT;      // totals
    int load;                     // upper bound of what one byte lane holds
};
__device__ __forceinline__ void tally_init(Tally& t) {
    t.nlo = t.nhi = t.bA = t.bC = t.bG = t.bT = t.tA = t.tC = t.tG = t.tT = 0;
    t.load = 0;
}
__device__ __forceinline__ void add3(Tally& t, uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t s = a + b + c;
    t.nlo += s & 0x33333333u;
    t.nhi += (s >> 2) & 0x33333333u;
}
__device__ __forceinline__ void nib_flush(Tally& t) {
    t.bA += t.nlo & 0x0f0f0f0fu;
    t.bG += (t.nlo >> 4) & 0x0f0f0f0fu;
    t.bC += t.nhi & 0x0f0f0f0fu;
    t.bT += (t.nhi >> 4) & 0x0f0f0f0fu;
    t.nlo = t.nhi = 0;
}
__device__ __forceinline__ void byte_flush(Tally& t) {
    t.tA = __dp4a(t.bA, 0x01010101u, t.tA);
    t.tC = __dp4a(t.bC, 0x01010101u, t.tC);
    t.tG = __dp4a(t.bG, 0x01010101u, t.tG);
    t.tT = __dp4a(t.bT, 0x01010101u, t.tT);
    t.bA = t.bC = t.bG = t.bT = 0;
    t.load = 0;
}
__device__ __forceinline__ uint4 and4(uint4 w, uint4 m) { return make_uint4(w.x & m.x, w.y & m.y, w.z & m.z, w.w & m.w); }
// 3 chunks = 12 words -> 4 level-1 sums (<= 12 per nibble), one nibble flush
__device__ __forceinline__ void add_chunks3(Tally& t, uint4 x, uint4 y, uint4 z) {
    add3(t, x.x, y.x, z.x);
    add3(t, x.y, y.y, z.y);
    add3(t, x.z, y.z, z.z);
    add3(t, x.w, y.w, z.w);
    nib_flush(t);
    t.load += 12;
    if (t.load > 240) byte_flush(t);
}
__device__ __forceinline__ void add_chunks2(Tally& t, uint4 x, uint4 y) {
    add3(t, x.x, x.y, x.z);
    add3(t, x.w, y.x, y.y);
    add3(t, y.z, y.w, 0u);
    nib_flush(t);
    t.load += 8;
    if (t.load > 240) byte_flush(t);
}
__device__ __forceinline__ void add_chunks1(Tally& t, uint4 x) {
    add3(t, x.x, x.y, x.z);
    add3(t, x.w, 0u, 0u);
    nib_flush(t);
    t.load += 4;
    if (t.load > 240) byte_flush(t);
}

// Per-lane running sums: 64-bit integers, 32-bit integers (flushed before they can overflow: K1Params::acc_limit),
// doubles.  A slot in global memory holds them as QI + QU + QD 8-byte words in that order.
template <int QI, int QU, int QD>
struct Acc {
    long long i[QI > 0 ? QI : 1];
    uint32_t u[QU > 0 ? QU : 1];
    double d[QD > 0 ? QD : 1];
};

// Warp-cooperative flush of the per-lane running sums into this warp's private slots.
template <int QI, int QU, int QD>
__device__ __forceinline__ void warp_flush(Acc<QI, QU, QD>& acc, int cur_seg, unsigned long long* part, int64_t slot_base,
                                           int seg_first, int warp, int lane, int nw) {
    constexpr int Q = QI + QU + QD;
    unsigned pending = __ballot_sync(0xffffffffu, cur_seg >= 0);
    while (pending) {
        const int leader = __ffs(pending) - 1;
        const int g = __shfl_sync(0xffffffffu, cur_seg, leader);
        const bool mine = (cur_seg == g);
        unsigned long long* dst = part + slot_base + ((int64_t)(g - seg_first) * nw + warp) * Q;
#pragma unroll
        for (int q = 0; q < QI; ++q) {
            long long v = mine ? acc.i[q] : 0ll;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if (lane == 0) dst[q] = (unsigned long long)((long long)dst[q] + v);
            if (mine) acc.i[q] = 0;
        }
#pragma unroll
        for (int q = 0; q < QU; ++q) {
            // 32-bit lane sums: the first butterfly step stays in 32 bits when two lanes cannot overflow... they can,
            // so widen first (33 bits after one step) — 5 steps of 64-bit adds on a once-per-segment path
            long long v = mine ? (long long)acc.u[q] : 0ll;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if (lane == 0) dst[QI + q] = (unsigned long long)((long long)dst[QI + q] + v);
            if (mine) acc.u[q] = 0u;
        }
#pragma unroll
        for (int q = 0; q < QD; ++q) {
            double v = mine ? acc.d[q] : 0.0;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);   // fixed butterfly order
            if (lane == 0)
                dst[QI + QU + q] =
                    (unsigned long long)__double_as_longlong(__longlong_as_double((long long)dst[QI + QU + q]) + v);
            if (mine) acc.d[q] = 0.0;
        }
        pending &= ~__ballot_sync(0xffffffffu, mine);
    }
}

__device__ __forceinline__ int find_seg(const int64_t* __restrict__ brk, int nseg, int from, int64_t site) {
    int lo = from, hi = nseg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(brk + mid) <= site) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <int MODE, int P>
struct ModeTraits;
template <int P>
struct ModeTraits<MODE_POPGEN, P> {
    static constexpr int QI = 3, QU = P + P * (P - 1) / 2, QD = 0;
};
template <int P>
struct ModeTraits<MODE_POPGEN_FREQ, P> {
    static constexpr int QI = 3, QU = P + P * (P - 1) / 2 + P, QD = 0;   // + segregating-site counts
};
template <int P>
struct ModeTraits<MODE_ABBA, P> {
    static constexpr int QI = 3, QU = 0, QD = 6;
};
template <int P>
struct ModeTraits<MODE_COUNTS, P> {
    static constexpr int QI = 0, QU = 0, QD = 0;
};
template <int P>
struct ModeTraits<MODE_FOURPOP, P> {
    static constexpr int QI = 3, QU = 0, QD = 16;
};
template <int P>
struct ModeTraits<MODE_FOURPOP_Q, P> {
    static constexpr int QI = 3, QU = 0, QD = 16;
};

// genomics.py:1409-1418, operation order of the reference's numpy expressions
__device__ __forceinline__ double f4_dev(double p1, double p2, double p3, double p4) {
    return (1 - p1) * p2 * p3 * (1 - p4) - p1 * (1 - p2) * p3 * (1 - p4);
}
__device__ __forceinline__ double f4c_dev(double p1, double p2, double p3, double p4) {
    return f4_dev(p1, p2, p3, p4) + f4_dev(1 - p1, 1 - p2, 1 - p3, 1 - p4);
}
// np.amax propagates nan
__device__ __forceinline__ double nmax(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }

// The 16 running sums of genomics.fourPop (genomics.py:1617-1643) for one informative site: counts of the chosen allele
// k_X and non-missing haplotypes n_X of P1..P4, packed as k | n << 16.
template <class ACC>
__device__ __forceinline__ void fourpop_add(ACC& acc, uint32_t e1, uint32_t e2, uint32_t e3, uint32_t e4) {
    const double p1 = (double)(e1 & 0xffffu) / (double)(e1 >> 16);      // 0/0 = nan, as in the reference (genomics.py:597)
    const double p2 = (double)(e2 & 0xffffu) / (double)(e2 >> 16);
    const double p3 = (double)(e3 & 0xffffu) / (double)(e3 >> 16);
    const double p4 = (double)(e4 & 0xffffu) / (double)(e4 >> 16);
    const double abba = (1 - p1) * p2 * p3 * (1 - p4);
    const double baba = p1 * (1 - p2) * p3 * (1 - p4);
    const double pd = p2 * (p2 > p3 ? 1.0 : 0.0) + p3 * (p3 >= p2 ? 1.0 : 0.0);
    const bool A = p3 > p1, Bq = p3 > p2, Xq = p1 > p2, Yq = !Xq;
    const double xa = (Xq && A) ? 1.0 : 0.0, yb = (Yq && Bq) ? 1.0 : 0.0;
    const double xna = (Xq && !A) ? 1.0 : 0.0, ynb = (Yq && !Bq) ? 1.0 : 0.0;
    const double pdm1 = p3 * xa + p1 * (1.0 - xa);
    const double pdm2 = p3 * yb + p2 * (1.0 - yb);
    const double pdm3 = -p3 * xa + p3 * yb - p1 * xna + p2 * ynb;
    const double t11 = f4c_dev(p1, p3, p3, p4), t12 = f4c_dev(p4, p2, p3, p4);
    const double t21 = f4c_dev(p3, p2, p3, p4), t22 = f4c_dev(p1, p4, p3, p4);
    const double t31 = f4c_dev(p1, p2, p2, p4), t32 = f4c_dev(p1, p2, p3, p1);
    const double t41 = f4c_dev(p1, p2, p1, p4), t42 = f4c_dev(p1, p2, p3, p2);
    const double m4 = nmax(nmax(t11, t12), nmax(t21, t22));
    const double m8 = nmax(nmax(m4, nmax(t31, t32)), nmax(t41, t42));
    const double u1 = fabs(p1 - p2), u2 = fabs(p3 - p4);
    const double um = u1 * (u1 > u2 ? 1.0 : 0.0) + u2 * (u2 >= u1 ? 1.0 : 0.0);
    acc.d[0] += f4_dev(p1, p2, p3, p4);
    acc.d[1] += f4_dev(p1, p3, p3, p4);
    acc.d[2] += f4c_dev(p1, p2, p3, p4);
    acc.d[3] += t11;
    acc.d[4] += abba + baba;
    acc.d[5] += f4_dev(p1, pd, pd, p4);
    acc.d[6] += f4c_dev(p1, pd, pd, p4);
    acc.d[7] += f4_dev(pdm1, pdm2, pdm3, p4);
    acc.d[8] += f4c_dev(pdm1, pdm2, pdm3, p4);
    acc.d[9] += m4;
    acc.d[10] += m8;
    acc.d[11] += um * um;
    acc.d[12] += abba;
    acc.d[13] += baba;
    acc.d[14] += (1 - p1) * p2 * (1 - p3) * (1 - p4);
    acc.d[15] += p1 * (1 - p2) * (1 - p3) * (1 - p4);
}

// TMA producer (one elected lane of the producer warp): keeps the ring of tiles full.
template <int MODE>
__device__ __forceinline__ void k1_producer(const K1Params& prm, uint8_t* tiles, uint64_t* full, uint64_t* empty,
                                            volatile int* s_issued, int ntiles, int64_t t0) {
    for (int it = 0; it < ntiles; ++it) {
        const int stage = it % prm.stages;
        if (it >= prm.stages) mbar_wait(&empty[stage], (uint32_t)(((it / prm.stages) - 1) & 1));
        const int64_t s_lo = prm.site_begin + (t0 + it) * prm.T;
        int64_t rows = prm.site_end - s_lo;
        if (rows > prm.T) rows = prm.T;
        const uint32_t bytes = (uint32_t)(rows * prm.pitch);
        // positions of the tile ride along (rounded up to 16 bytes; the array has zeroed slack)
        const uint32_t pbytes = (MODE == MODE_COUNTS) ? 0u : (uint32_t)(((rows * 4 + 15) / 16) * 16);
        mbar_expect_tx(&full[stage], bytes + pbytes);
        const uint8_t* src = prm.geno + s_lo * prm.pitch;
        uint8_t* dst = tiles + (size_t)stage * prm.tile_bytes;
        for (uint32_t off = 0; off < bytes; off += 32768u) {
            const uint32_t n = (bytes - off) < 32768u ? (bytes - off) : 32768u;
            bulk_g2s(dst + off, src + off, n, &full[stage]);
        }
        if (pbytes) bulk_g2s(dst + (size_t)prm.T * prm.pitch, prm.pos + s_lo, pbytes, &full[stage]);
        // publish "tile `it` is armed": a consumer must not test a phase parity before its phase has
        // been armed, or try_wait.parity would alias it with the previous (already complete) phase
        __threadfence_block();
        atomicExch(const_cast<int*>(s_issued), it + 1);      // an atomic, so that racecheck sees the flag as synchronisation
    }
}

// BYTES (POPGEN modes, every population <= 255 haplotypes): the four allele counts of a population travel as the
// bytes of one word, so that sum c^2 and sum c_X c_Y are ONE IDP.4A each, accumulate included.
template <int MODE, int P, int NW, bool BYTES = false>
__global__ void __launch_bounds__((NW + 1) * 32, 1) k1_site_pass(const __grid_constant__ K1Params prm) {
    constexpr int K1_THREADS = (NW + 1) * 32;
    constexpr int QI = ModeTraits<MODE, P>::QI, QU = ModeTraits<MODE, P>::QU, QD = ModeTraits<MODE, P>::QD;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* tiles = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)prm.stages * prm.tile_bytes);   // [stages]
    uint64_t* empty = full + 8;                                                                  // [stages]
    volatile int* s_issued = reinterpret_cast<volatile int*>(empty + 8);   // tiles armed by the producer so far
    uint4* s_ent_mask = reinterpret_cast<uint4*>(smem + (size_t)prm.stages * prm.tile_bytes + 256);
    int32_t* s_ent_chunk = reinterpret_cast<int32_t*>(s_ent_mask + prm.n_ent);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, B = gridDim.x;
    const int64_t t0 = (int64_t)b * prm.num_tiles / B, t1 = (int64_t)(b + 1) * prm.num_tiles / B;
    const int ntiles = (int)(t1 - t0);

    for (int e = tid; e < prm.n_ent; e += K1_THREADS) {
        s_ent_mask[e] = prm.ent_mask[e];
        s_ent_chunk[e] = prm.ent_chunk[e];
    }
    if (tid == 0) {
        for (int s = 0; s < prm.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], (uint32_t)prm.wpt);
        }
        *s_issued = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == NW) {
        if (lane == 0) k1_producer<MODE>(prm, tiles, full, empty, s_issued, ntiles, t0);
        return;
    }

    // ---------------- consumers: team m = warp / wpt owns tiles m, m + nteams, ... ----------------
    const int G = prm.G;
    const int spw = 32 / G;              // sites per warp per iteration
    const int gsub = lane / spw;         // which part of the row this lane walks
    const int sl = lane % spw;
    const int nteams = NW / prm.wpt;
    const int team = warp / prm.wpt, lw = warp % prm.wpt;
    const int sites_per_iter = prm.wpt * spw;

    Acc<QI, QU, QD> acc;
#pragma unroll
    for (int q = 0; q < QI; ++q) acc.i[q] = 0;
#pragma unroll
    for (int q = 0; q < QU; ++q) acc.u[q] = 0u;
    int since_flush = 0;     // sites added to the 32-bit sums since they were last flushed (warp-uniform)
#pragma unroll
    for (int q = 0; q < QD; ++q) acc.d[q] = 0.0;
    int cur_seg = -1;
    int64_t seg_end = -1;
    const int seg_first = (MODE == MODE_COUNTS) ? 0 : prm.cta_seg_first[b];
    const int64_t slot_base = (MODE == MODE_COUNTS) ? 0 : prm.cta_slot_off[b];
    // MODE_FOURPOP_Q: one queued informative site per lane (k | n << 16 of P1..P4).  Sites are queued only while every lane
    // of the warp is in the same segment, and the queue is emptied before any lane changes segment, so whichever lane
    // evaluates a queued site adds it to the sums of the right segment.
    uint32_t q1 = 0u, q2 = 0u, q3 = 0u, q4 = 0u;
    bool qpend = false;

    for (int it = team; it < ntiles; it += nteams) {
        const int stage = it % prm.stages;
        if (lane == 0)
            while (atomicAdd(const_cast<int*>(s_issued), 0) <= it) __nanosleep(20);
        __syncwarp();
        mbar_wait(&full[stage], (uint32_t)((it / prm.stages) & 1));
        const uint8_t* tile = tiles + (size_t)stage * prm.tile_bytes;
        const int64_t tile_site0 = prm.site_begin + (t0 + it) * prm.T;

        for (int i = 0; i < prm.I; ++i) {
            const int slot = i * sites_per_iter + lw * spw + sl;
            const int64_t site = tile_site0 + slot;
            const bool valid = site < prm.site_end;
            const bool owner = valid && (gsub == 0);
            const uint4* row = reinterpret_cast<const uint4*>(tile + (size_t)(valid ? slot : 0) * prm.pitch);
            // the tile's positions were staged behind its genotype rows by the producer
            int posv = 0;
            if (MODE != MODE_COUNTS)
                posv = owner ? reinterpret_cast<const int32_t*>(tile + (size_t)prm.T * prm.pitch)[slot] : 0;

            uint32_t n[P], c[BYTES ? 1 : P][4], cb[BYTES ? P : 1];
#pragma unroll
            for (int X = 0; X < P; ++X) {
                Tally t;
                tally_init(t);
                {   // the population's main run of fully-owned 16-byte chunks: no masks
                    int ch = prm.full_lo[X] + gsub;
                    const int hi = prm.full_hi[X];
                    for (; ch + 2 * G < hi; ch += 3 * G) add_chunks3(t, row[ch], row[ch + G], row[ch + 2 * G]);
                    if (ch + G < hi) add_chunks2(t, row[ch], row[ch + G]);
                    else if (ch < hi) add_chunks1(t, row[ch]);
                }
                {   // chunks shared with other populations / unused haplotypes / row padding: masked
                    int e = prm.ent_lo[X] + gsub;
                    const int hi = prm.ent_hi[X];
                    for (; e + G < hi; e += 2 * G)
                        add_chunks2(t, and4(row[s_ent_chunk[e]], s_ent_mask[e]), and4(row[s_ent_chunk[e + G]], s_ent_mask[e + G]));
                    if (e < hi) add_chunks1(t, and4(row[s_ent_chunk[e]], s_ent_mask[e]));
                }
                byte_flush(t);
                if (BYTES) {
                    // counts <= 255: one word per population, the G lanes of the site add their packed bytes
                    uint32_t pk = t.tA | (t.tC << 8) | (t.tG << 16) | (t.tT << 24);
                    for (int d = spw; d < 32; d <<= 1) pk += __shfl_xor_sync(0xffffffffu, pk, d);
                    cb[X] = pk;
                    n[X] = __dp4a(pk, 0x01010101u, 0u);
                } else {
                    // combine the G lanes of this site (16-bit fields: counts < 65536)
                    uint32_t p0 = t.tA | (t.tC << 16), p1 = t.tG | (t.tT << 16);
                    for (int d = spw; d < 32; d <<= 1) {
                        p0 += __shfl_xor_sync(0xffffffffu, p0, d);
                        p1 += __shfl_xor_sync(0xffffffffu, p1, d);
                    }
                    c[X][0] = p0 & 0xffffu;
                    c[X][1] = p0 >> 16;
                    c[X][2] = p1 & 0xffffu;
                    c[X][3] = p1 >> 16;
                    n[X] = c[X][0] + c[X][1] + c[X][2] + c[X][3];
                }
            }

            if (MODE == MODE_COUNTS) {
                if (owner) {
                    uint16_t* o = prm.counts_out + (site - prm.site_begin) * prm.counts_stride;
#pragma unroll
                    for (int X = 0; X < P; ++X)
                        if (X < prm.counts_pops) {
                            ushort4 v = make_ushort4((unsigned short)c[X][0], (unsigned short)c[X][1],
                                                     (unsigned short)c[X][2], (unsigned short)c[X][3]);
                            *reinterpret_cast<ushort4*>(o + X * 4) = v;
                        }
                }
                continue;
            }

            // ---- segment bookkeeping (warp-uniform control flow) ----
            int sg = cur_seg;
            if (owner && site >= seg_end) sg = find_seg(prm.brk, prm.nseg, cur_seg + 1, site);
            if (__any_sync(0xffffffffu, sg != cur_seg)) {
                if (MODE == MODE_FOURPOP_Q) {
                    if (qpend) fourpop_add(acc, q1, q2, q3, q4);
                    qpend = false;
                }
                warp_flush<QI, QU, QD>(acc, cur_seg, prm.part, slot_base, seg_first, warp, lane, NW);
                since_flush = 0;
                if (sg != cur_seg) {
                    cur_seg = sg;
                    seg_end = __ldg(prm.brk + sg + 1);
                }
            }

            if (MODE == MODE_POPGEN || MODE == MODE_POPGEN_FREQ) {
                bool allpres = true, allmiss = true;
#pragma unroll
                for (int X = 0; X < P; ++X) {
                    allpres = allpres && (n[X] == (uint32_t)prm.popN[X]);
                    allmiss = allmiss && (n[X] == 0u);
                }
                const bool pres = owner && allpres;
                const bool ragged = owner && !allpres && !allmiss;
                if (++since_flush > prm.acc_limit) {      // the 32-bit sums must not overflow (never taken for N < ~900)
                    warp_flush<QI, QU, QD>(acc, cur_seg, prm.part, slot_base, seg_first, warp, lane, NW);
                    since_flush = 1;
                }
                acc.i[0] += pres ? 1 : 0;
                acc.i[1] += ragged ? 1 : 0;
                acc.i[2] += (long long)posv;
                if (BYTES) {
                    uint32_t cf[P];      // counts of a site that does not count are zeroed once, instead of every product
#pragma unroll
                    for (int X = 0; X < P; ++X) cf[X] = pres ? cb[X] : 0u;
#pragma unroll
                    for (int X = 0; X < P; ++X) {
                        if (MODE == MODE_POPGEN_FREQ) {
                            const uint32_t sq = __dp4a(cf[X], cb[X], 0u);
                            acc.u[X] += sq;
                            acc.u[P + P * (P - 1) / 2 + X] += (pres && sq != n[X] * n[X]) ? 1u : 0u;
                        } else {
                            acc.u[X] = __dp4a(cf[X], cb[X], acc.u[X]);
                        }
                    }
                    int kb = 0;
#pragma unroll
                    for (int X = 0; X < P; ++X)
#pragma unroll
                        for (int Y = X + 1; Y < P; ++Y) {
                            acc.u[P + kb] = __dp4a(cf[X], cb[Y], acc.u[P + kb]);
                            ++kb;
                        }
                }
                const uint32_t f = pres ? 1u : 0u;
#pragma unroll
                for (int X = 0; X < (BYTES ? 0 : P); ++X) {
                    const uint32_t sq = c[X][0] * c[X][0] + c[X][1] * c[X][1] + c[X][2] * c[X][2] + c[X][3] * c[X][3];
                    acc.u[X] += sq * f;
                    // groupFreqStats (genomics.py:1002-1028): a complete site is segregating in X iff sum c^2 < N^2
                    // (two populations share one 64-bit accumulator: 32-bit fields)
                    if (MODE == MODE_POPGEN_FREQ)
                        acc.u[P + P * (P - 1) / 2 + X] += (pres && sq != n[X] * n[X]) ? 1u : 0u;
                }
                int k = 0;
#pragma unroll
                for (int X = 0; X < (BYTES ? 0 : P); ++X)
#pragma unroll
                    for (int Y = X + 1; Y < P; ++Y) {
                        const uint32_t cr = c[X][0] * c[Y][0] + c[X][1] * c[Y][1] + c[X][2] * c[Y][2] + c[X][3] * c[Y][3];
                        acc.u[P + k] += cr * f;
                        ++k;
                    }
            }

            if (MODE == MODE_ABBA) {
                // genomics.py:1655-1662: biallelic over P1+P2+P3+O and enough data in each population
                uint32_t tot[4];
                int nall = 0;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    tot[a] = c[0][a] + c[1][a] + c[2][a] + c[3][a];
                    nall += tot[a] > 0 ? 1 : 0;
                }
                bool good = owner && (nall == 2);
#pragma unroll
                for (int X = 0; X < 4; ++X) good = good && ((int)n[X] >= prm.thr[X]);
                acc.i[1] += good ? 1 : 0;
                acc.i[2] += (long long)posv;
                // derived allele: present overall, absent in the outgroup (1672). With two alleles overall and a
                // non-empty outgroup at most one allele qualifies, so one body serves the whole warp.
                int da = -1;
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    if (tot[a] > 0 && c[3][a] == 0) da = a;
                const bool hit = good && n[3] > 0 && da >= 0;
                if (hit) {
                    const uint32_t k1 = da == 0 ? c[0][0] : (da == 1 ? c[0][1] : (da == 2 ? c[0][2] : c[0][3]));
                    const uint32_t k2 = da == 0 ? c[1][0] : (da == 1 ? c[1][1] : (da == 2 ? c[1][2] : c[1][3]));
                    const uint32_t k3 = da == 0 ? c[2][0] : (da == 1 ? c[2][1] : (da == 2 ? c[2][2] : c[2][3]));
                    // the derived allele is absent from the outgroup: p4 = 0/n4 = 0 and every (1 - p4) factor of the
                    // reference's formulas is exactly 1.0 — dropping those factors leaves the values bit-identical
                    const double p1 = (double)k1 / (double)n[0];
                    const double p2 = (double)k2 / (double)n[1];
                    const double p3 = (double)k3 / (double)n[2];
                    const double abba = (1 - p1) * p2 * p3;
                    const double baba = p1 * (1 - p2) * p3;
                    const double pd = p2 * (p2 > p3 ? 1.0 : 0.0) + p3 * (p3 >= p2 ? 1.0 : 0.0);
                    const double fd_den = (1 - p1) * pd * pd - p1 * (1 - pd) * pd;
                    const bool A = p3 > p1, Bq = p3 > p2, Xq = p1 > p2, Yq = !Xq;
                    const double xa = (Xq && A) ? 1.0 : 0.0, yb = (Yq && Bq) ? 1.0 : 0.0;
                    const double xna = (Xq && !A) ? 1.0 : 0.0, ynb = (Yq && !Bq) ? 1.0 : 0.0;
                    const double pdm1 = p3 * xa + p1 * (1.0 - xa);
                    const double pdm2 = p3 * yb + p2 * (1.0 - yb);
                    const double pdm3 = -p3 * xa + p3 * yb - p1 * xna + p2 * ynb;
                    const double fdm_den = (1 - pdm1) * pdm2 * pdm3 - pdm1 * (1 - pdm2) * pdm3;
                    acc.i[0] += 1;
                    acc.d[0] += abba;
                    acc.d[1] += baba;
                    acc.d[2] += abba - baba;
                    acc.d[3] += abba + baba;
                    acc.d[4] += fd_den;
                    acc.d[5] += fdm_den;
                }
            }

            if (MODE == MODE_FOURPOP || MODE == MODE_FOURPOP_Q) {
                // genomics.py:1595-1603: biallelic over P1+P2+P3+P4 and enough data in each population
                uint32_t tot[4];
                int nall = 0;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    tot[a] = c[0][a] + c[1][a] + c[2][a] + c[3][a];
                    nall += tot[a] > 0 ? 1 : 0;
                }
                bool good = owner && (nall == 2);
#pragma unroll
                for (int X = 0; X < 4; ++X) good = good && ((int)n[X] >= prm.thr[X]);
                acc.i[1] += good ? 1 : 0;
                acc.i[2] += (long long)posv;
                int da = -1;
                if (prm.variant == 0) {
                    // np.argsort(all4freqs)[:,2] (1615): of the two alleles present, the rarer one; an exact tie is
                    // resolved by numpy's sort implementation in the reference — here the lower allele index
                    uint32_t best = 0xffffffffu;
#pragma unroll
                    for (int a = 3; a >= 0; --a)
                        if (tot[a] > 0 && tot[a] <= best) {
                            best = tot[a];
                            da = a;
                        }
                } else {
                    // polarize (1610): present overall, absent in P4 (needs data in P4: nan == 0 is False)
#pragma unroll
                    for (int a = 0; a < 4; ++a)
                        if (tot[a] > 0 && c[3][a] == 0) da = a;
                    if (n[3] == 0) da = -1;
                }
                bool hit = good && da >= 0;
                const uint32_t k1 = da == 0 ? c[0][0] : (da == 1 ? c[0][1] : (da == 2 ? c[0][2] : c[0][3]));
                const uint32_t k2 = da == 0 ? c[1][0] : (da == 1 ? c[1][1] : (da == 2 ? c[1][2] : c[1][3]));
                const uint32_t k3 = da == 0 ? c[2][0] : (da == 1 ? c[2][1] : (da == 2 ? c[2][2] : c[2][3]));
                const uint32_t k4 = da == 0 ? c[3][0] : (da == 1 ? c[3][1] : (da == 2 ? c[3][2] : c[3][3]));
                if (prm.variant == 2)       // fixed (1611-1614): frequency exactly 0 or 1 in P1, P2, P3 (nan fails both)
                    hit = hit && n[0] > 0 && n[1] > 0 && n[2] > 0 && (k1 == 0 || k1 == n[0]) && (k2 == 0 || k2 == n[1]) &&
                          (k3 == 0 || k3 == n[2]);
                const uint32_t e1 = k1 | (n[0] << 16), e2 = k2 | (n[1] << 16), e3 = k3 | (n[2] << 16), e4 = k4 | (n[3] << 16);
                acc.i[0] += hit ? 1 : 0;
                if (MODE == MODE_FOURPOP) {
                    if (hit) fourpop_add(acc, e1, e2, e3, e4);
                } else {
                    // About a third of the lanes hold an informative site, and the ~1000 instructions of the evaluation
                    // run for the whole warp whenever one does.  Here the informative sites move into free queue slots of
                    // the warp (one per lane) and are evaluated when the next ones no longer fit, i.e. with (nearly) all
                    // 32 lanes at work.  MEASURED on B200 (profiles/r02c_fourpop_queue.jsonl, C2 shape): identical sums,
                    // but 1.15 ms against 0.90 ms for the default / polarize modes (the ballot / rank / __fns / four
                    // shuffles of every iteration cost more than the evaluations they save) and 0.70 against 0.78 ms for
                    // `fixed`, where informative sites are rare — so it stays an experiment (PG_K1_FOURPOP_QUEUE).
                    const unsigned full_m = 0xffffffffu;
                    const int seg0 = __shfl_sync(full_m, cur_seg, 0);
                    if (!__all_sync(full_m, cur_seg == seg0)) {
                        // lanes in two segments (the iteration that crosses a window boundary, the tail of the data): the
                        // queue is empty — a segment change empties it — and every lane evaluates its own site
                        if (hit) fourpop_add(acc, e1, e2, e3, e4);
                    } else {
                        const unsigned hits = __ballot_sync(full_m, hit);
                        if (hits) {
                            unsigned pm = __ballot_sync(full_m, qpend);
                            const int nh = __popc(hits);
                            if (nh > 32 - __popc(pm)) {         // no room for the new sites: evaluate the queued ones
                                if (qpend) fourpop_add(acc, q1, q2, q3, q4);
                                qpend = false;
                                pm = 0u;
                            }
                            const unsigned fr = ~pm;
                            const int r = __popc(fr & ((1u << lane) - 1u));          // rank of this lane among the free ones
                            const bool take = ((fr >> lane) & 1u) && r < nh;
                            const int src = take ? (int)__fns(hits, 0, r + 1) : lane;  // lane of the (r+1)-th new site
                            const uint32_t t1 = __shfl_sync(full_m, e1, src), t2 = __shfl_sync(full_m, e2, src);
                            const uint32_t t3 = __shfl_sync(full_m, e3, src), t4 = __shfl_sync(full_m, e4, src);
                            if (take) {
                                q1 = t1;
                                q2 = t2;
                                q3 = t3;
                                q4 = t4;
                                qpend = true;
                            }
                        }
                    }
                }
            }
        }

        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);   // this warp is done with the stage's bytes
    }
    if (MODE == MODE_FOURPOP_Q && qpend) fourpop_add(acc, q1, q2, q3, q4);
    if (MODE != MODE_COUNTS) warp_flush<QI, QU, QD>(acc, cur_seg, prm.part, slot_base, seg_first, warp, lane, NW);
}


// ---- lane-per-population variant for LONG rows (popgen modes, byte-packed counts) ----------------------------------
// P lanes share one site and lane X walks population X's chunks alone: no per-population cross-lane combine, every
// lane holds ONE packed count word, and the pair products are spread over the lanes — lane X accumulates
// sum c_X^2 and sum c_X c_{X+d} for d = 1 .. P/2 (partner words arrive by shuffle) — so a lane carries 1 + P/2 32-bit
// sums instead of P + P(P-1)/2.  The slot layout in global memory is the same as k1_site_pass's.
template <int QI, int QU>
__device__ __forceinline__ void warp_flush_lp(long long (&ai)[QI], uint32_t (&au)[QU], int cur_seg, unsigned long long* part,
                                              int64_t slot_base, int seg_first, int warp, int lane, int nw, int Q, int spw,
                                              int X, const int* s_q) {
    unsigned pending = __ballot_sync(0xffffffffu, cur_seg >= 0);
    while (pending) {
        const int leader = __ffs(pending) - 1;
        const int g = __shfl_sync(0xffffffffu, cur_seg, leader);
        const bool mine = (cur_seg == g);
        unsigned long long* dst = part + slot_base + ((int64_t)(g - seg_first) * nw + warp) * Q;
        const bool head = (lane % spw) == 0;
#pragma unroll
        for (int q = 0; q < QI; ++q) {      // site bookkeeping lives in the lanes of population 0
            long long v = (mine && X == 0) ? ai[q] : 0ll;
            for (int d = spw >> 1; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if (lane == 0) dst[q] = (unsigned long long)((long long)dst[q] + v);
            if (mine) ai[q] = 0;
        }
#pragma unroll
        for (int q = 0; q < QU; ++q) {      // lanes of the same population are neighbours: butterfly inside the group
            long long v = mine ? (long long)au[q] : 0ll;
            for (int d = spw >> 1; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            const int slot = s_q[X * QU + q];
            if (head && slot >= 0) dst[slot] = (unsigned long long)((long long)dst[slot] + v);
            if (mine) au[q] = 0u;
        }
        pending &= ~__ballot_sync(0xffffffffu, mine);
    }
}

template <int MODE, int P, int NW>
__global__ void __launch_bounds__((NW + 1) * 32, 1) k1_site_pass_lp(const __grid_constant__ K1Params prm) {
    static_assert(MODE == MODE_POPGEN || MODE == MODE_POPGEN_FREQ || MODE == MODE_COUNTS, "lane-per-population: popgen / counts");
    constexpr int K1_THREADS = (NW + 1) * 32;
    constexpr int HP = P / 2;
    constexpr int QU = 1 + HP + (MODE == MODE_POPGEN_FREQ ? 1 : 0);      // sq, cross d = 1..P/2, (segregating sites)
    constexpr int spw = 32 / P;                                            // sites per warp per step
    const int Q = 3 + P + P * (P - 1) / 2 + (MODE == MODE_POPGEN_FREQ ? P : 0);
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* tiles = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)prm.stages * prm.tile_bytes);
    uint64_t* empty = full + 8;
    volatile int* s_issued = reinterpret_cast<volatile int*>(empty + 8);
    uint4* s_ent_mask = reinterpret_cast<uint4*>(smem + (size_t)prm.stages * prm.tile_bytes + 256);
    int32_t* s_ent_chunk = reinterpret_cast<int32_t*>(s_ent_mask + prm.n_ent);
    int* s_pop = s_ent_chunk + prm.n_ent;            // [5][P]: full_lo, full_hi, ent_lo, ent_hi, popN
    int* s_q = s_pop + 5 * P;                        // [P][QU]: slot word of each lane-local sum (-1 = unused)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, B = gridDim.x;
    const int64_t t0 = (int64_t)b * prm.num_tiles / B, t1 = (int64_t)(b + 1) * prm.num_tiles / B;
    const int ntiles = (int)(t1 - t0);

    for (int e = tid; e < prm.n_ent; e += K1_THREADS) {
        s_ent_mask[e] = prm.ent_mask[e];
        s_ent_chunk[e] = prm.ent_chunk[e];
    }
    if (tid < P) {
        s_pop[0 * P + tid] = prm.full_lo[tid];
        s_pop[1 * P + tid] = prm.full_hi[tid];
        s_pop[2 * P + tid] = prm.ent_lo[tid];
        s_pop[3 * P + tid] = prm.ent_hi[tid];
        s_pop[4 * P + tid] = prm.popN[tid];
        // slot words (layout of k1_site_pass): [3 ints][P sq][pairs (x<y) in order][P segregating]
        const int x = tid;
        s_q[x * QU + 0] = 3 + x;
        for (int d = 1; d <= HP; ++d) {
            int slot = -1;
            if (d < HP || x < HP) {
                const int y = (x + d) % P;
                const int lo = x < y ? x : y, hi = x < y ? y : x;
                int kp = 0;
                for (int xx = 0; xx < lo; ++xx) kp += P - 1 - xx;
                kp += hi - lo - 1;
                slot = 3 + P + kp;
            }
            s_q[x * QU + d] = slot;
        }
        if (MODE == MODE_POPGEN_FREQ) s_q[x * QU + 1 + HP] = 3 + P + P * (P - 1) / 2 + x;
    }
    if (tid == 0) {
        for (int st = 0; st < prm.stages; ++st) {
            mbar_init(&full[st], 1);
            mbar_init(&empty[st], (uint32_t)prm.wpt);
        }
        *s_issued = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == NW) {
        if (lane == 0) k1_producer<MODE>(prm, tiles, full, empty, s_issued, ntiles, t0);
        return;
    }

    const int X = lane / spw;                // this lane's population
    const int sl = lane % spw;               // its site within the warp's step
    unsigned site_lanes = 0;                 // the P lanes that share this lane's site
#pragma unroll
    for (int k = 0; k < P; ++k) site_lanes |= 1u << (sl + k * spw);
    const int nteams = NW / prm.wpt;
    const int team = warp / prm.wpt, lw = warp % prm.wpt;
    const int sites_per_iter = prm.wpt * spw;
    const int f_lo = s_pop[0 * P + X], f_hi = s_pop[1 * P + X], e_lo = s_pop[2 * P + X], e_hi = s_pop[3 * P + X];
    const uint32_t myN = (uint32_t)s_pop[4 * P + X];

    long long ai[3] = {0, 0, 0};
    uint32_t au[QU];
#pragma unroll
    for (int q = 0; q < QU; ++q) au[q] = 0u;
    int since_flush = 0;
    int cur_seg = -1;
    int64_t seg_end = -1;
    const int seg_first = (MODE == MODE_COUNTS) ? 0 : prm.cta_seg_first[b];
    const int64_t slot_base = (MODE == MODE_COUNTS) ? 0 : prm.cta_slot_off[b];

    for (int it = team; it < ntiles; it += nteams) {
        const int stage = it % prm.stages;
        if (lane == 0)
            while (atomicAdd(const_cast<int*>(s_issued), 0) <= it) __nanosleep(20);
        __syncwarp();
        mbar_wait(&full[stage], (uint32_t)((it / prm.stages) & 1));
        const uint8_t* tile = tiles + (size_t)stage * prm.tile_bytes;
        const int64_t tile_site0 = prm.site_begin + (t0 + it) * prm.T;

        for (int i = 0; i < prm.I; ++i) {
            const int slot = i * sites_per_iter + lw * spw + sl;
            const int64_t site = tile_site0 + slot;
            const bool valid = site < prm.site_end;
            const bool owner = valid && (X == 0);
            const uint4* row = reinterpret_cast<const uint4*>(tile + (size_t)(valid ? slot : 0) * prm.pitch);
            int posv = 0;
            if (MODE != MODE_COUNTS)
                posv = owner ? reinterpret_cast<const int32_t*>(tile + (size_t)prm.T * prm.pitch)[slot] : 0;

            Tally t;
            tally_init(t);
            {
                int ch = f_lo;
                for (; ch + 2 < f_hi; ch += 3) add_chunks3(t, row[ch], row[ch + 1], row[ch + 2]);
                if (ch + 1 < f_hi) add_chunks2(t, row[ch], row[ch + 1]);
                else if (ch < f_hi) add_chunks1(t, row[ch]);
            }
            {
                int e = e_lo;
                for (; e + 1 < e_hi; e += 2)
                    add_chunks2(t, and4(row[s_ent_chunk[e]], s_ent_mask[e]), and4(row[s_ent_chunk[e + 1]], s_ent_mask[e + 1]));
                if (e < e_hi) add_chunks1(t, and4(row[s_ent_chunk[e]], s_ent_mask[e]));
            }
            byte_flush(t);
            if (MODE == MODE_COUNTS) {       // every lane writes its own population's four counts
                if (valid && X < prm.counts_pops)
                    *reinterpret_cast<ushort4*>(prm.counts_out + (site - prm.site_begin) * prm.counts_stride + X * 4) =
                        make_ushort4((unsigned short)t.tA, (unsigned short)t.tC, (unsigned short)t.tG, (unsigned short)t.tT);
                continue;
            }
            const uint32_t cb = t.tA | (t.tC << 8) | (t.tG << 16) | (t.tT << 24);
            const uint32_t n = __dp4a(cb, 0x01010101u, 0u);

            // ---- segment bookkeeping: the site's owner lane looks it up, its P lanes share it ----
            int sg = cur_seg;
            if (owner && site >= seg_end) sg = find_seg(prm.brk, prm.nseg, cur_seg + 1, site);
            sg = __shfl_sync(0xffffffffu, sg, sl);                   // lane sl is population 0 of this site
            if (!valid) sg = cur_seg;
            if (__any_sync(0xffffffffu, sg != cur_seg)) {
                warp_flush_lp<3, QU>(ai, au, cur_seg, prm.part, slot_base, seg_first, warp, lane, NW, Q, spw, X, s_q);
                since_flush = 0;
                if (sg != cur_seg) {
                    cur_seg = sg;
                    seg_end = __ldg(prm.brk + sg + 1);
                }
            }
            if (++since_flush > prm.acc_limit) {
                warp_flush_lp<3, QU>(ai, au, cur_seg, prm.part, slot_base, seg_first, warp, lane, NW, Q, spw, X, s_q);
                since_flush = 1;
            }

            const unsigned bf = __ballot_sync(0xffffffffu, valid && n == myN);
            const unsigned bz = __ballot_sync(0xffffffffu, valid && n == 0u);
            const bool allpres = (bf & site_lanes) == site_lanes;
            const bool allmiss = (bz & site_lanes) == site_lanes;
            const bool pres = valid && allpres;
            ai[0] += (owner && allpres) ? 1 : 0;
            ai[1] += (owner && !allpres && !allmiss) ? 1 : 0;
            ai[2] += (long long)posv;
            const uint32_t cf = pres ? cb : 0u;
            if (MODE == MODE_POPGEN_FREQ) {
                const uint32_t sq = __dp4a(cf, cb, 0u);
                au[0] += sq;
                au[1 + HP] += (pres && sq != n * n) ? 1u : 0u;
            } else {
                au[0] = __dp4a(cf, cb, au[0]);
            }
#pragma unroll
            for (int d = 1; d <= HP; ++d) {
                int px = X + d;
                if (px >= P) px -= P;
                const uint32_t cbp = __shfl_sync(0xffffffffu, cb, px * spw + sl);
                if (d < HP || X < HP) au[d] = __dp4a(cf, cbp, au[d]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
    }
    if (MODE != MODE_COUNTS) warp_flush_lp<3, QU>(ai, au, cur_seg, prm.part, slot_base, seg_first, warp, lane, NW, Q, spw, X, s_q);
}

// ---- finalize: slots -> segments -> windows -> statistics ------------------------------------------
struct FinParams {
    const unsigned long long* part;
    const int32_t* seg_cta_lo;     // [nseg] first CTA touching the segment
    const int32_t* seg_cta_hi;     // [nseg] last CTA (inclusive)
    const int32_t* cta_seg_first;  // [ctas]
    const int64_t* cta_slot_off;   // [ctas]
    const int32_t* win_seg_lo;     // [W]
    const int32_t* win_seg_hi;     // [W]
    const int64_t* win_lo;
    const int64_t* win_hi;
    int64_t W;
    int Q, QI, nw;
    int P;                         // real population count
    int Ppad;                      // template P used by the site pass
    int popN[PG_MAX_K1_POPS];
    double harm_a[PG_MAX_K1_POPS], harm_a2[PG_MAX_K1_POPS];
    int min_sites;
    double min_data;
    int force_path;
    int with_freq;                 // the site pass carried the popFreq counters
    int bookkeeping_only;          // P > 8: the site pass saw one collapsed population; statistics come from K2
    // outputs: fixed-width 8-byte records per window
    //   popgen: [sites(i64) pos_sum(i64) path(i64) pi[P] dxy[npairs] fst[npairs]]
    //   abba  : [sites(i64) pos_sum(i64) ABBA BABA D fd fdM sitesUsed]
    unsigned long long* rec;
    int RC;
    int32_t* path;                 // [W] popgen routing, also counted in *n_pairwise
    int* n_pairwise;
};

__device__ __forceinline__ unsigned long long d2u(double x) { return (unsigned long long)__double_as_longlong(x); }

__device__ __forceinline__ double nan_d() { return __longlong_as_double(0x7ff8000000000000ll); }

// mean over the off-diagonal entries of an N x N block whose pairs all have n_ij = Lp
__device__ __forceinline__ double cf_pi(long long N, long long Lp, long long sumsq, bool all_nan, double min_data) {
    if (N <= 0) return nan_d();
    const double size = (double)(N * N);
    const double nan_cnt = all_nan ? size : (double)N;
    if (1.0 - (1.0 * nan_cnt / size) < min_data) return nan_d();      // nanmean_min, genomics.py:88-90
    if (all_nan) return nan_d();
    const double num = (double)(N * N * Lp - sumsq);
    const double den = (double)((N * N - N) * Lp);
    return num / den;
}

template <int MODE>
__global__ void __launch_bounds__(64) k1_finalize(const __grid_constant__ FinParams fp) {
    __shared__ unsigned long long sums[64];
    const int q = threadIdx.x;
    for (int64_t w = blockIdx.x; w < fp.W; w += gridDim.x) {
        __syncthreads();
        if (q < fp.Q) {
            long long si = 0;
            double sd = 0.0;
            for (int g = fp.win_seg_lo[w]; g < fp.win_seg_hi[w]; ++g) {
                for (int b = fp.seg_cta_lo[g]; b <= fp.seg_cta_hi[g]; ++b) {
                    const unsigned long long* src =
                        fp.part + fp.cta_slot_off[b] + (int64_t)(g - fp.cta_seg_first[b]) * fp.nw * fp.Q;
                    for (int wp = 0; wp < fp.nw; ++wp) {
                        const unsigned long long v = src[wp * fp.Q + q];
                        if (q < fp.QI) si += (long long)v; else sd += __longlong_as_double((long long)v);
                    }
                }
            }
            sums[q] = (q < fp.QI) ? (unsigned long long)si : (unsigned long long)__double_as_longlong(sd);
        }
        __syncthreads();
        if (q != 0) continue;
        const long long sites = fp.win_hi[w] - fp.win_lo[w];
        unsigned long long* rec = fp.rec + (size_t)w * fp.RC;
        rec[0] = (unsigned long long)sites;
        rec[1] = sums[2];
        if (MODE == MODE_POPGEN) {
            const int P = fp.P, Pp = fp.Ppad;
            const int npairs = P * (P - 1) / 2;
            double* pi_o = reinterpret_cast<double*>(rec + 3);
            double* dxy_o = pi_o + P;
            double* fst_o = dxy_o + npairs;
            const long long Lp = (long long)sums[0];
            const bool ragged = (long long)sums[1] > 0;
            if (fp.bookkeeping_only) {
                const int path = sites < fp.min_sites ? 0 : 2;
                fp.path[w] = path;
                rec[2] = (unsigned long long)path;
                if (path == 2) atomicAdd(fp.n_pairwise, 1);
                for (int k = 3; k < fp.RC; ++k) rec[k] = d2u(nan_d());
                continue;
            }
            {   // popFreq columns (valid for every window: they only use sites complete in all haplotypes)
                double* fq = fst_o + npairs;           // [l, S[P], thetaPi[P], thetaW[P], TajD[P]]
                fq[0] = fp.with_freq ? (double)Lp : nan_d();
                const int npp = Pp * (Pp - 1) / 2;
                for (int x = 0; x < P; ++x) {
                    double Sx = nan_d(), tpi = nan_d(), tw = nan_d(), tD = nan_d();
                    if (Lp >= 1 && fp.with_freq) {
                        const long long N = fp.popN[x];
                        const long long seg = (long long)sums[3 + Pp + npp + x];
                        const long long pairs = (N * N * Lp - (long long)sums[3 + x]) / 2;   // sum over sites of sum_{a<b} c_a c_b
                        Sx = (double)seg;
                        tpi = (double)pairs / (.5 * (double)N * (double)(N - 1));
                        const double a = fp.harm_a[x], a2 = fp.harm_a2[x];   // sum 1/i, sum 1/i^2 for i < N (host, same order)
                        tw = (double)seg / a;
                        // TajimaD (genomics.py:619-632)
                        const double n_ = (double)N;
                        const double b1 = (n_ + 1.) / (3 * (n_ - 1));
                        const double b2 = (2. * (n_ * n_ + n_ + 3)) / (9 * n_ * (n_ - 1));
                        const double c1 = b1 - (1. / a);
                        const double c2 = b2 - ((n_ + 2) / (a * n_)) + a2 / (a * a);
                        const double e1 = c1 / a;
                        const double e2 = c2 / (a * a + a2);
                        const double d = tpi - tw;
                        tD = d / sqrt(e1 * Sx + e2 * Sx * (Sx - 1));
                    }
                    fq[1 + x] = Sx;
                    fq[1 + P + x] = tpi;
                    fq[1 + 2 * P + x] = tw;
                    fq[1 + 3 * P + x] = tD;
                }
            }
            int path = 1;
            if (sites < fp.min_sites) path = 0;
            else if (ragged || fp.force_path == 2) path = 2;
            fp.path[w] = path;
            rec[2] = (unsigned long long)path;
            if (path == 2) atomicAdd(fp.n_pairwise, 1);
            if (path != 1) {
                for (int x = 0; x < P; ++x) pi_o[x] = nan_d();
                for (int k = 0; k < npairs; ++k) dxy_o[k] = fst_o[k] = nan_d();
                continue;
            }
            const bool all_nan = (Lp == 0) || (fp.min_sites > 0 && Lp < fp.min_sites);
            double piv[PG_MAX_K1_POPS];
            for (int x = 0; x < P; ++x) {
                piv[x] = cf_pi(fp.popN[x], Lp, (long long)sums[3 + x], all_nan, fp.min_data);
                pi_o[x] = piv[x];
            }
            int k = 0;
            for (int x = 0; x < P; ++x)
                for (int y = x + 1; y < P; ++y) {
                    // index of (x,y) in the padded pair enumeration of the site pass
                    int kp = 0;
                    for (int xx = 0; xx < x; ++xx) kp += Pp - 1 - xx;
                    kp += y - x - 1;
                    const long long cross = (long long)sums[3 + Pp + kp];
                    const long long Nx = fp.popN[x], Ny = fp.popN[y];
                    double dxy = nan_d();
                    {
                        const double size = (double)(Nx * Ny);
                        const double nan_cnt = all_nan ? size : 0.0;
                        const bool frac_bad = (1.0 - (1.0 * nan_cnt / size) < fp.min_data);
                        if (!frac_bad && !all_nan) dxy = (double)(Nx * Ny * Lp - cross) / (double)(Nx * Ny * Lp);
                    }
                    const long long sq_t = (long long)sums[3 + x] + (long long)sums[3 + y] + 2 * cross;
                    const double pi_t = cf_pi(Nx + Ny, Lp, sq_t, all_nan, fp.min_data);
                    const double wgt = 1.0 * (double)Nx / (double)(Nx + Ny);
                    const double pi_s = wgt * piv[x] + (1 - wgt) * piv[y];
                    dxy_o[k] = dxy;
                    fst_o[k] = 1 - pi_s / pi_t;
                    ++k;
                }
        } else if (MODE == MODE_FOURPOP) {
            // genomics.py:1623-1643: [fhom, fhom', D, fd, fd', fdm, fdm', fdh, fdh2, fh, ABBA, BABA, ABAA, BAAA, sitesUsed]
            const long long used = (long long)sums[0], n_good = (long long)sums[1];
            double* o = reinterpret_cast<double*>(rec + 2);
            if (n_good < 1) {
                for (int k = 0; k < 14; ++k) o[k] = nan_d();
                o[14] = 0.0;
                continue;
            }
            double d[16];
            for (int k = 0; k < 16; ++k) d[k] = __longlong_as_double((long long)sums[3 + k]);
            o[0] = d[0] * 1. / d[1];
            o[1] = d[2] * 1. / d[3];
            o[2] = d[0] * 1. / d[4];
            o[3] = d[0] * 1. / d[5];
            o[4] = d[2] * 1. / d[6];
            o[5] = d[0] * 1. / d[7];
            o[6] = d[2] * 1. / d[8];
            o[7] = d[2] * 1. / d[9];
            o[8] = d[2] * 1. / d[10];
            o[9] = d[2] * 1. / d[11];
            o[10] = d[12];
            o[11] = d[13];
            o[12] = d[14];
            o[13] = d[15];
            o[14] = (double)used;
        } else {   // MODE_ABBA
            const long long used = (long long)sums[0], n_good = (long long)sums[1];
            double* o = reinterpret_cast<double*>(rec + 2);
            if (n_good < 1) {   // genomics.py:1694-1695: every value nan, sitesUsed included
                for (int k = 0; k < 6; ++k) o[k] = nan_d();
                continue;
            }
            const double s_abba = __longlong_as_double((long long)sums[3]), s_baba = __longlong_as_double((long long)sums[4]);
            const double s_f4 = __longlong_as_double((long long)sums[5]), s_ab = __longlong_as_double((long long)sums[6]);
            const double s_fd = __longlong_as_double((long long)sums[7]), s_fdm = __longlong_as_double((long long)sums[8]);
            o[0] = s_abba;
            o[1] = s_baba;
            o[2] = s_f4 * 1.0 / s_ab;
            o[3] = s_f4 * 1.0 / s_fd;
            o[4] = s_f4 * 1.0 / s_fdm;
            o[5] = (double)used;
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------
struct PopTables {
    std::vector<int32_t> ent_chunk;
    std::vector<uint32_t> ent_mask;   // 4 words per entry
    int ent_lo[PG_MAX_K1_POPS], ent_hi[PG_MAX_K1_POPS], full_lo[PG_MAX_K1_POPS], full_hi[PG_MAX_K1_POPS];
    int popN[PG_MAX_K1_POPS];
};

// hap_pop_local[h] in [0, Ppad) or -1
void build_tables(const std::vector<int32_t>& hap_pop_local, int H, int chunks, int Ppad, PopTables& t) {
    t.ent_chunk.clear();
    t.ent_mask.clear();
    for (int X = 0; X < PG_MAX_K1_POPS; ++X) t.ent_lo[X] = t.ent_hi[X] = t.full_lo[X] = t.full_hi[X] = t.popN[X] = 0;
    for (int X = 0; X < Ppad; ++X) {
        std::vector<uint16_t> cm(chunks, 0);
        int N = 0;
        for (int h = 0; h < H; ++h)
            if (hap_pop_local[h] == X) {
                cm[h / 16] |= (uint16_t)(1u << (h % 16));
                ++N;
            }
        t.popN[X] = N;
        // longest run of completely-owned chunks
        int best_lo = 0, best_hi = 0, run_lo = -1;
        for (int cidx = 0; cidx <= chunks; ++cidx) {
            const bool fullc = cidx < chunks && cm[cidx] == 0xffff;
            if (fullc && run_lo < 0) run_lo = cidx;
            if (!fullc && run_lo >= 0) {
                if (cidx - run_lo > best_hi - best_lo) {
                    best_lo = run_lo;
                    best_hi = cidx;
                }
                run_lo = -1;
            }
        }
        t.full_lo[X] = best_lo;
        t.full_hi[X] = best_hi;
        t.ent_lo[X] = (int)t.ent_chunk.size();
        for (int cidx = 0; cidx < chunks; ++cidx) {
            if (cm[cidx] == 0) continue;
            if (cidx >= best_lo && cidx < best_hi) continue;
            t.ent_chunk.push_back(cidx);
            for (int wd = 0; wd < 4; ++wd) {
                uint32_t m = 0;
                for (int by = 0; by < 4; ++by)
                    if (cm[cidx] & (1u << (wd * 4 + by))) m |= 0xffu << (8 * by);
                t.ent_mask.push_back(m);
            }
        }
        t.ent_hi[X] = (int)t.ent_chunk.size();
    }
}

int check_plan(const K1Plan& pl) {
    const bool pow2G = pl.G >= 1 && pl.G <= 32 && (pl.G & (pl.G - 1)) == 0;
    const bool okw = (pl.wpt == 1 || pl.wpt == 2 || pl.wpt == 4 || pl.wpt == 8) && (pl.nw % pl.wpt) == 0;
    PG_CHECK((pl.T % 4) == 0, "rows of %d bytes are too long for the site-pass kernel", pl.pitch);
    PG_CHECK(pow2G && okw && pl.I >= 1 && pl.stages >= 2 && pl.stages <= 8 && pl.smem_bytes <= 227 * 1024,
             "invalid site-pass geometry G=%d wpt=%d I=%d stages=%d smem=%d", pl.G, pl.wpt, pl.I, pl.stages, pl.smem_bytes);
    return PG_OK;
}

struct K1Launch {
    K1Plan plan;
    K1Params prm;
    std::vector<int32_t> cta_seg_first, seg_cta_lo, seg_cta_hi;
    std::vector<int64_t> cta_slot_off;
    int64_t total_slots = 0;   // 8-byte words
};

int seg_of(const std::vector<int64_t>& brk, int64_t site) {
    // brk[g] <= site < brk[g+1]
    return (int)(std::upper_bound(brk.begin(), brk.end(), site) - brk.begin()) - 1;
}

// device layout of the uploaded tables inside ctx->tables:
//   [ent_mask (16B each)] [ent_chunk] [brk] [cta_seg_first] [cta_slot_off] [seg_cta_lo] [seg_cta_hi]
//   [win_seg_lo] [win_seg_hi] [win_lo] [win_hi]
struct DevTables {
    uint4* ent_mask;
    int32_t* ent_chunk;
    int64_t* brk;
    int32_t* cta_seg_first;
    int64_t* cta_slot_off;
    int32_t* seg_cta_lo;
    int32_t* seg_cta_hi;
    int32_t* win_seg_lo;
    int32_t* win_seg_hi;
    int64_t* win_lo;
    int64_t* win_hi;
};

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <typename T>
int push(pg_ctx* ctx, uint8_t* base, size_t& off, const T* src, size_t n, T** out) {
    off = align_up(off, 16);
    *out = reinterpret_cast<T*>(base + off);
    if (n) PG_CUDA(cudaMemcpyAsync(base + off, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    off += n * sizeof(T);
    return PG_OK;
}

// Everything a windowed launch needs, cached per configuration (data shape, populations, windows): a repeated
// statistics call on the same configuration only clears the slots and launches two kernels.
struct K1Cache {
    bool valid = false;
    bool lanepop = false;
    uint64_t epoch = 0;
    int mode = -1;
    int sel[4] = {-1, -1, -1, -1};
    K1Launch L;
    DevTables dt;
    PopTables pt;
    PgBuf tables;
};

int prepare_windowed(pg_ctx* ctx, K1Cache& c, const std::vector<int32_t>& hap_pop_local, int Ppad, int Q, int nw,
                     int force_G = 0) {
    K1Launch& L = c.L;
    DevTables& dt = c.dt;
    PopTables& pt = c.pt;
    PG_TRY(pg_build_segments(ctx));
    build_tables(hap_pop_local, ctx->H, ctx->pitch / 16, Ppad, pt);
    const int n_ent = (int)pt.ent_chunk.size();
    const int table_bytes = n_ent * 20 + 64 + 512;        // + the per-population tables of the lane-per-population variant
    PG_CHECK(table_bytes <= 48 * 1024, "population layout needs %d bytes of mask tables (limit 48 KiB)", table_bytes);
    L.plan = pg_make_k1_plan(ctx->S, ctx->H, ctx->sm_count, table_bytes, nw, force_G);
    PG_CHECK(L.plan.stages >= 2, "rows of %d haplotypes are too long for the site-pass kernel (pitch %d bytes)", ctx->H,
             L.plan.pitch);
    PG_TRY(check_plan(L.plan));
    for (int X = 0; X < Ppad; ++X) PG_CHECK(pt.popN[X] <= 65535, "a population has more than 65535 haplotypes");
    const K1Plan& pl = L.plan;
    const int B = pl.ctas;
    const int nseg = (int)ctx->brk.size() - 1;
    L.cta_seg_first.assign(B, 0);
    L.cta_slot_off.assign(B, 0);
    L.seg_cta_lo.assign(std::max(nseg, 1), 0);
    L.seg_cta_hi.assign(std::max(nseg, 1), -1);
    std::vector<int> cta_seg_last(B, -1);
    int64_t off = 0;
    for (int b = 0; b < B; ++b) {
        const int64_t t0 = (int64_t)b * pl.num_tiles / B, t1 = (int64_t)(b + 1) * pl.num_tiles / B;
        const int64_t s0 = t0 * pl.T, s1 = std::min<int64_t>(t1 * pl.T, ctx->S);
        L.cta_slot_off[b] = off;
        if (s1 <= s0) continue;
        const int g0 = seg_of(ctx->brk, s0), g1 = seg_of(ctx->brk, s1 - 1);
        L.cta_seg_first[b] = g0;
        cta_seg_last[b] = g1;
        off += (int64_t)(g1 - g0 + 1) * nw * Q;
    }
    L.total_slots = off;
    for (int g = 0; g < nseg; ++g) {
        L.seg_cta_lo[g] = B;
        L.seg_cta_hi[g] = -1;
    }
    for (int b = 0; b < B; ++b) {
        if (cta_seg_last[b] < 0) continue;
        for (int g = L.cta_seg_first[b]; g <= cta_seg_last[b]; ++g) {
            L.seg_cta_lo[g] = std::min(L.seg_cta_lo[g], b);
            L.seg_cta_hi[g] = std::max(L.seg_cta_hi[g], b);
        }
    }
    size_t bytes = 4096 + pt.ent_mask.size() * 4 + pt.ent_chunk.size() * 4 + ctx->brk.size() * 8 + (size_t)B * 12 +
                   (size_t)std::max(nseg, 1) * 8 + (size_t)ctx->W * 24 + 16 * 16;
    PG_TRY(c.tables.ensure(bytes));
    uint8_t* base = (uint8_t*)c.tables.p;
    size_t o = 0;
    uint32_t* d_mask_words = nullptr;
    PG_TRY(push(ctx, base, o, pt.ent_mask.data(), pt.ent_mask.size(), &d_mask_words));
    dt.ent_mask = reinterpret_cast<uint4*>(d_mask_words);
    PG_TRY(push(ctx, base, o, pt.ent_chunk.data(), pt.ent_chunk.size(), &dt.ent_chunk));
    PG_TRY(push(ctx, base, o, ctx->brk.data(), ctx->brk.size(), &dt.brk));
    PG_TRY(push(ctx, base, o, L.cta_seg_first.data(), L.cta_seg_first.size(), &dt.cta_seg_first));
    PG_TRY(push(ctx, base, o, L.cta_slot_off.data(), L.cta_slot_off.size(), &dt.cta_slot_off));
    PG_TRY(push(ctx, base, o, L.seg_cta_lo.data(), L.seg_cta_lo.size(), &dt.seg_cta_lo));
    PG_TRY(push(ctx, base, o, L.seg_cta_hi.data(), L.seg_cta_hi.size(), &dt.seg_cta_hi));
    PG_TRY(push(ctx, base, o, ctx->win_seg_lo.data(), ctx->win_seg_lo.size(), &dt.win_seg_lo));
    PG_TRY(push(ctx, base, o, ctx->win_seg_hi.data(), ctx->win_seg_hi.size(), &dt.win_seg_hi));
    PG_TRY(push(ctx, base, o, ctx->win_lo.data(), ctx->win_lo.size(), &dt.win_lo));
    PG_TRY(push(ctx, base, o, ctx->win_hi.data(), ctx->win_hi.size(), &dt.win_hi));
    PG_CHECK(o <= c.tables.cap, "internal: table buffer overflow");
    // the host vectors must outlive the async copies
    PG_CUDA(cudaStreamSynchronize(ctx->stream));

    K1Params& p = L.prm;
    memset(&p, 0, sizeof(p));
    p.geno = (const uint8_t*)ctx->d_geno;
    p.pos = ctx->d_pos;
    p.site_begin = 0;
    p.site_end = ctx->S;
    p.num_tiles = pl.num_tiles;
    p.pitch = pl.pitch;
    p.G = pl.G;
    p.I = pl.I;
    p.T = pl.T;
    p.wpt = pl.wpt;
    p.nw = nw;
    p.stages = pl.stages;
    p.tile_bytes = pl.tile_bytes;
    p.ent_chunk = dt.ent_chunk;
    p.ent_mask = dt.ent_mask;
    p.n_ent = n_ent;
    for (int X = 0; X < PG_MAX_K1_POPS; ++X) {
        p.ent_lo[X] = pt.ent_lo[X];
        p.ent_hi[X] = pt.ent_hi[X];
        p.full_lo[X] = pt.full_lo[X];
        p.full_hi[X] = pt.full_hi[X];
        p.popN[X] = pt.popN[X];
    }
    p.brk = dt.brk;
    p.nseg = nseg;
    p.cta_seg_first = dt.cta_seg_first;
    p.cta_slot_off = dt.cta_slot_off;
    return PG_OK;
}

// per-call part: zeroed slots
int arm_slots(pg_ctx* ctx, K1Cache& c) {
    const size_t bytes = (size_t)std::max<int64_t>(c.L.total_slots, 1) * 8;
    PG_TRY(ctx->part.ensure(bytes));
    PG_CUDA(cudaMemsetAsync(ctx->part.p, 0, bytes, ctx->stream));
    c.L.prm.part = (unsigned long long*)ctx->part.p;
    c.L.prm.geno = (const uint8_t*)ctx->d_geno;
    c.L.prm.pos = ctx->d_pos;
    return PG_OK;
}

template <int MODE, int P, int NW, bool BYTES>
int launch_site_pass_nw(pg_ctx* ctx, const K1Launch& L, const char* name) {
    auto kern = k1_site_pass<MODE, P, NW, BYTES>;
    static bool attr_set[64] = {};    // per instantiation and per device (the attribute is per device)
    if (!attr_set[ctx->device & 63]) {
        PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[ctx->device & 63] = true;
    }
    const int ti = pg_time_begin(ctx, name);
    kern<<<L.plan.ctas, (NW + 1) * 32, L.plan.smem_bytes, ctx->stream>>>(L.prm);
    pg_time_end(ctx, ti);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

// consumer warps per CTA: 12 (+ the producer = 416 threads, 128 registers each) for rows below 1 KiB; 8 (up to 168
// registers, no spills) for longer rows, where the 8-population instantiations measured 3-10 % faster (tools/k1_sweep2.py)
int k1_env_nw12() {
    const char* e = getenv("PG_K1_NW");
    return (e && *e && atoi(e) == 8) ? 8 : 12;
}
int k1_nw_for(int pitch) {
    const char* e = getenv("PG_K1_NW");
    const int v = (e && *e) ? atoi(e) : (pitch >= 1024 ? 8 : 12);
    return v == 12 ? 12 : 8;
}

template <int MODE, int P, int NW>
int launch_site_pass_lp(pg_ctx* ctx, const K1Launch& L, const char* name) {
    auto kern = k1_site_pass_lp<MODE, P, NW>;
    static bool attr_set[64] = {};
    if (!attr_set[ctx->device & 63]) {
        PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[ctx->device & 63] = true;
    }
    const int ti = pg_time_begin(ctx, name);
    kern<<<L.plan.ctas, (NW + 1) * 32, L.plan.smem_bytes, ctx->stream>>>(L.prm);
    pg_time_end(ctx, ti);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

template <int MODE, int P>
int launch_site_pass(pg_ctx* ctx, const K1Launch& L, const char* name) {
    constexpr bool POPGEN_MODE = (MODE == MODE_POPGEN || MODE == MODE_POPGEN_FREQ);
    if constexpr ((POPGEN_MODE || MODE == MODE_COUNTS) && (P == 4 || P == 8)) {
        if (L.prm.lanepop) {
            if (L.prm.nw == 12) return launch_site_pass_lp<MODE, P, 12>(ctx, L, name);
            return launch_site_pass_lp<MODE, P, 8>(ctx, L, name);
        }
    }
    if constexpr (POPGEN_MODE) {
        if (L.prm.bytes) {
            if (L.prm.nw == 12) return launch_site_pass_nw<MODE, P, 12, true>(ctx, L, name);
            return launch_site_pass_nw<MODE, P, 8, true>(ctx, L, name);
        }
    }
    if (L.prm.nw == 12) return launch_site_pass_nw<MODE, P, 12, false>(ctx, L, name);
    return launch_site_pass_nw<MODE, P, 8, false>(ctx, L, name);
}

int pad_pops(int P) { return P <= 2 ? 2 : (P <= 4 ? 4 : 8); }

K1Cache* cache_of(pg_ctx* ctx, int slot) {
    if (!ctx->k1_cache[slot]) ctx->k1_cache[slot] = new K1Cache();
    return static_cast<K1Cache*>(ctx->k1_cache[slot]);
}

void fill_fin(FinParams& fp, pg_ctx* ctx, const K1Cache& c, int Q, int QI) {
    memset(&fp, 0, sizeof(fp));
    fp.part = (const unsigned long long*)ctx->part.p;
    fp.seg_cta_lo = c.dt.seg_cta_lo;
    fp.seg_cta_hi = c.dt.seg_cta_hi;
    fp.cta_seg_first = c.dt.cta_seg_first;
    fp.cta_slot_off = c.dt.cta_slot_off;
    fp.win_seg_lo = c.dt.win_seg_lo;
    fp.win_seg_hi = c.dt.win_seg_hi;
    fp.win_lo = c.dt.win_lo;
    fp.win_hi = c.dt.win_hi;
    fp.W = ctx->W;
    fp.Q = Q;
    fp.QI = QI;
    fp.nw = c.L.prm.nw;
    for (int X = 0; X < PG_MAX_K1_POPS; ++X) {
        fp.popN[X] = c.pt.popN[X];
        double a = 0.0, a2 = 0.0;                       // TajimaD's python sums (genomics.py:621-623), same order
        for (int i = 1; i < c.pt.popN[X]; ++i) {
            a += 1. / (double)i;
            a2 += 1. / ((double)i * (double)i);
        }
        fp.harm_a[X] = a;
        fp.harm_a2[X] = a2;
    }
}

}  // namespace

void pg_k1_cache_free(pg_ctx* ctx) {
    for (int k = 0; k < 3; ++k)
        if (ctx->k1_cache[k]) {
            K1Cache* c = static_cast<K1Cache*>(ctx->k1_cache[k]);
            c->tables.release();
            delete c;
            ctx->k1_cache[k] = nullptr;
        }
}

// ================================================================================================
// pg_popgen
// ================================================================================================
// Enqueue the site pass + finalize on the ctx stream WITHOUT synchronising; *h_count (pinned) holds the number of
// windows routed to the pairwise path once the stream has been synchronised (nullptr when nothing was launched).
int pg_popgen_enqueue(pg_ctx* ctx, int32_t min_sites, double min_data, int32_t force_path, void* d_rec, int** h_count) {
    PG_CHECK(ctx && d_rec && h_count, "pg_popgen_device: null argument");
    *h_count = nullptr;
    PG_CHECK(ctx->P >= 1, "pg_popgen: call pg_set_pops first");
    PG_CHECK(force_path == 0 || force_path == 2, "pg_popgen: force_path must be 0 or 2");
    PG_CHECK(ctx->P <= PG_MAX_POPS, "pg_popgen: P=%d > %d populations", ctx->P, PG_MAX_POPS);
    PG_CUDA(cudaSetDevice(ctx->device));
    pg_timings_reset(ctx);
    const int P = ctx->P;
    // More populations than the site pass keeps in registers: it still does the bookkeeping (sites, position sums,
    // failed / ragged windows) with all used haplotypes collapsed into one population, and every window that passes
    // minSites goes through the pairwise path, whose epilogue handles up to PG_MAX_POPS populations.
    const bool many = P > PG_MAX_K1_POPS;
    const int npairs = P * (P - 1) / 2;
    const int RC = 3 + P + 2 * npairs + 1 + 4 * P;
    const int64_t W = ctx->W;
    if (W == 0) return PG_OK;
    if (ctx->S == 0) {
        std::vector<unsigned long long> h((size_t)W * RC);
        const double qn = NAN;
        unsigned long long nanbits;
        memcpy(&nanbits, &qn, 8);
        for (int64_t w = 0; w < W; ++w) {
            h[w * RC] = 0;
            h[w * RC + 1] = 0;
            h[w * RC + 2] = (0 < min_sites) ? 0 : 1;
            for (int k = 3; k < RC; ++k) h[w * RC + k] = nanbits;
        }
        PG_CUDA(cudaMemcpy(d_rec, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
        return PG_OK;
    }
    const int Pp = many ? 2 : pad_pops(P);
    const bool wf = ctx->want_freq && !many;
    const int Q = 3 + Pp + Pp * (Pp - 1) / 2 + (wf ? Pp : 0);
    K1Cache& c = *cache_of(ctx, 0);
    if (!c.valid || c.epoch != ctx->epoch) {
        c.valid = false;
        for (int x = 0; x < P; ++x) {
            int N = 0;
            for (int h = 0; h < ctx->H; ++h) N += ctx->hap_pop[h] == x;
            PG_CHECK(N >= 1, "pg_popgen: population %d has no haplotypes", x);
        }
        std::vector<int32_t> collapsed;
        if (many) {
            collapsed.assign(ctx->hap_pop.begin(), ctx->hap_pop.end());
            for (int32_t& v : collapsed) v = v >= 0 ? 0 : -1;
        }
        const std::vector<int32_t>& pop_map = many ? collapsed : ctx->hap_pop;
        // long rows, 4 or 8 real populations of <= 255 haplotypes: one lane per population (k1_site_pass_lp)
        int maxN = 0;
        for (int x = 0; x < P && !many; ++x) {
            int N = 0;
            for (int h = 0; h < ctx->H; ++h) N += ctx->hap_pop[h] == x;
            maxN = std::max(maxN, N);
        }
        bool lp = !many && (Pp == 4 || Pp == 8) && Pp == P && maxN <= 255 && ctx->pitch >= 1024 && !getenv("PG_K1_NO_BYTES");
        if (const char* e = getenv("PG_K1_LANEPOP"))
            lp = atoi(e) != 0 && !many && (Pp == 4 || Pp == 8) && maxN <= 255 && !getenv("PG_K1_NO_BYTES");
        c.lanepop = lp;
        const int nw = lp ? k1_env_nw12() : k1_nw_for(ctx->pitch);
        PG_TRY(prepare_windowed(ctx, c, pop_map, Pp, Q, nw, lp ? Pp : 0));
        c.epoch = ctx->epoch;
        c.valid = true;
    }
    {
        long long maxN = 1;
        for (int X = 0; X < Pp; ++X) maxN = std::max<long long>(maxN, c.pt.popN[X]);
        c.L.prm.acc_limit = (int)std::max<long long>(1, std::min<long long>(0xffffffffll / (maxN * maxN), 1 << 30));
        if (const char* e = getenv("PG_K1_ACC_LIMIT")) c.L.prm.acc_limit = std::max(1, atoi(e));   // test hook: force early flushes
        c.L.prm.bytes = (maxN <= 255 && !getenv("PG_K1_NO_BYTES")) ? 1 : 0;
        c.L.prm.lanepop = c.lanepop ? 1 : 0;
    }
    PG_TRY(arm_slots(ctx, c));
    if (!wf) {
        if (Pp == 2) PG_TRY((launch_site_pass<MODE_POPGEN, 2>(ctx, c.L, "k1_popgen")));
        else if (Pp == 4) PG_TRY((launch_site_pass<MODE_POPGEN, 4>(ctx, c.L, "k1_popgen")));
        else PG_TRY((launch_site_pass<MODE_POPGEN, 8>(ctx, c.L, "k1_popgen")));
    } else {
        if (Pp == 2) PG_TRY((launch_site_pass<MODE_POPGEN_FREQ, 2>(ctx, c.L, "k1_popgen")));
        else if (Pp == 4) PG_TRY((launch_site_pass<MODE_POPGEN_FREQ, 4>(ctx, c.L, "k1_popgen")));
        else PG_TRY((launch_site_pass<MODE_POPGEN_FREQ, 8>(ctx, c.L, "k1_popgen")));
    }

    PG_TRY(ctx->out_i.ensure((size_t)W * 4 + 128));
    int* d_cnt = (int*)ctx->out_i.p;
    int32_t* d_path = (int32_t*)ctx->out_i.p + 16;
    PG_CUDA(cudaMemsetAsync(d_cnt, 0, 4, ctx->stream));
    FinParams fp;
    fill_fin(fp, ctx, c, Q, Q);
    fp.P = P;
    fp.Ppad = Pp;
    fp.min_sites = min_sites;
    fp.min_data = min_data;
    fp.force_path = many ? 2 : force_path;
    fp.bookkeeping_only = many ? 1 : 0;
    fp.with_freq = wf ? 1 : 0;
    fp.rec = (unsigned long long*)d_rec;
    fp.RC = RC;
    fp.path = d_path;
    fp.n_pairwise = d_cnt;
    const int ti = pg_time_begin(ctx, "k1_finalize");
    k1_finalize<MODE_POPGEN><<<(unsigned)std::min<int64_t>(W, 65535), 64, 0, ctx->stream>>>(fp);
    pg_time_end(ctx, ti);
    PG_CUDA(cudaGetLastError());
    void* hp = nullptr;
    PG_TRY(pg_pinned(ctx, (size_t)W * 4 + 256, &hp));
    int* h_cnt = (int*)hp;
    *h_cnt = 0;
    PG_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, 4, cudaMemcpyDeviceToHost, ctx->stream));
    *h_count = h_cnt;
    return PG_OK;
}

// After the stream has been synchronised: run the pairwise path for the windows the finalize kernel routed to it
// (their rows of d_rec are overwritten in place).
int pg_popgen_resolve(pg_ctx* ctx, int32_t min_sites, double min_data, void* d_rec, int nk2) {
    if (nk2 <= 0) return PG_OK;
    const int P = ctx->P;
    const int RC = 3 + P + 2 * (P * (P - 1) / 2) + 1 + 4 * P;
    const int64_t W = ctx->W;
    const int32_t* d_path = (const int32_t*)ctx->out_i.p + 16;
    std::vector<int32_t> h_path((size_t)W);
    PG_CUDA(cudaMemcpyAsync(h_path.data(), d_path, (size_t)W * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<int64_t> k2_windows;
    k2_windows.reserve((size_t)nk2);
    for (int64_t w = 0; w < W; ++w)
        if (h_path[w] == 2) k2_windows.push_back(w);
    return pg_k2_popgen_windows(ctx, k2_windows, min_sites, min_data, d_rec, RC);
}

// Device-record variant: d_rec is a DEVICE buffer of W * (4 + 5P + 2*npairs) 8-byte words.
extern "C" int pg_popgen_device(pg_ctx* ctx, int32_t min_sites, double min_data, int32_t force_path, void* d_rec,
                                int64_t* n_pairwise) {
    if (n_pairwise) *n_pairwise = 0;
    int* h_cnt = nullptr;
    PG_TRY(pg_popgen_enqueue(ctx, min_sites, min_data, force_path, d_rec, &h_cnt));
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    const int nk2 = h_cnt ? *h_cnt : 0;
    if (n_pairwise) *n_pairwise = nk2;
    return pg_popgen_resolve(ctx, min_sites, min_data, d_rec, nk2);
}


extern "C" int pg_popgen(pg_ctx* ctx, int32_t min_sites, double min_data, int32_t force_path, double* pi, double* dxy,
                         double* fst, int64_t* n_sites, int64_t* pos_sum, int32_t* path) {
    PG_CHECK(ctx && pi && dxy && fst && n_sites && pos_sum && path, "pg_popgen: null argument");
    PG_CHECK(ctx->P >= 1, "pg_popgen: call pg_set_pops first");
    PG_CHECK(ctx->P <= PG_MAX_POPS, "pg_popgen: P=%d > %d populations", ctx->P, PG_MAX_POPS);
    const int P = ctx->P;
    const int npairs = P * (P - 1) / 2;
    const int RC = 3 + P + 2 * npairs + 1 + 4 * P;
    const int64_t W = ctx->W;
    if (W == 0) {
        pg_timings_reset(ctx);
        return PG_OK;
    }
    PG_CUDA(cudaSetDevice(ctx->device));
    PG_TRY(ctx->out_d.ensure((size_t)W * RC * 8 + 64));
    // site pass -> finalize -> D2H of the record table into pinned staging, ONE host synchronisation; only when the
    // finalize kernel routed windows to the pairwise path are those rows recomputed and the table read again
    void* hp = nullptr;
    PG_TRY(pg_pinned(ctx, (size_t)W * RC * 8 + 64 + 256, &hp));     // (+ the routed-window counter of the enqueue step)
    int* h_cnt = nullptr;
    PG_TRY(pg_popgen_enqueue(ctx, min_sites, min_data, force_path, ctx->out_d.p, &h_cnt));
    hp = (uint8_t*)hp + 256;
    const unsigned long long* hrec = (const unsigned long long*)hp;
    PG_CUDA(cudaMemcpyAsync(hp, ctx->out_d.p, (size_t)W * RC * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    const int nk2 = h_cnt ? *h_cnt : 0;
    if (nk2 > 0) {
        PG_TRY(pg_popgen_resolve(ctx, min_sites, min_data, ctx->out_d.p, nk2));
        PG_CUDA(cudaMemcpyAsync(hp, ctx->out_d.p, (size_t)W * RC * 8, cudaMemcpyDeviceToHost, ctx->stream));
        PG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (ctx->want_freq) ctx->h_rec.assign(hrec, hrec + (size_t)W * RC);     // kept for pg_popgen_freqstats
    else ctx->h_rec.clear();
    for (int64_t w = 0; w < W; ++w) {
        const unsigned long long* r = hrec + (size_t)w * RC;
        n_sites[w] = (int64_t)r[0];
        pos_sum[w] = (int64_t)r[1];
        path[w] = (int32_t)r[2];
        memcpy(pi + (size_t)w * P, r + 3, (size_t)P * 8);
        if (npairs) {
            memcpy(dxy + (size_t)w * npairs, r + 3 + P, (size_t)npairs * 8);
            memcpy(fst + (size_t)w * npairs, r + 3 + P + npairs, (size_t)npairs * 8);
        }
    }
    return PG_OK;
}

extern "C" int pg_set_freqstats(pg_ctx* ctx, int32_t enable) {
    PG_CHECK(ctx != nullptr, "pg_set_freqstats: null ctx");
    const bool e = enable != 0;
    if (e != ctx->want_freq) {
        ctx->want_freq = e;
        ctx->epoch += 1;              // the slot layout of the site pass changes
        ctx->h_rec.clear();
    }
    return PG_OK;
}

// popFreq columns of the records produced by the most recent pg_popgen on this ctx
extern "C" int pg_popgen_freqstats(pg_ctx* ctx, double* l, double* S, double* theta_pi, double* theta_w, double* taj_d) {
    PG_CHECK(ctx && l && S && theta_pi && theta_w && taj_d, "pg_popgen_freqstats: null argument");
    const int P = ctx->P;
    const int npairs = P * (P - 1) / 2;
    const int RC = 3 + P + 2 * npairs + 1 + 4 * P;
    const int64_t W = ctx->W;
    PG_CHECK(ctx->want_freq, "pg_popgen_freqstats: enable the popFreq counters first (pg_set_freqstats(ctx, 1))");
    PG_CHECK(ctx->h_rec.size() == (size_t)W * RC, "pg_popgen_freqstats: call pg_popgen first (same pops / windows)");
    for (int64_t w = 0; w < W; ++w) {
        const double* f = reinterpret_cast<const double*>(ctx->h_rec.data() + (size_t)w * RC + 3 + P + 2 * npairs);
        l[w] = f[0];
        memcpy(S + (size_t)w * P, f + 1, (size_t)P * 8);
        memcpy(theta_pi + (size_t)w * P, f + 1 + P, (size_t)P * 8);
        memcpy(theta_w + (size_t)w * P, f + 1 + 2 * P, (size_t)P * 8);
        memcpy(taj_d + (size_t)w * P, f + 1 + 3 * P, (size_t)P * 8);
    }
    return PG_OK;
}

// ================================================================================================
// pg_abbababa
// ================================================================================================
// Enqueue site pass + finalize of the ABBA-BABA statistics on the ctx stream; records [W x 8] 8-byte words
// [sites, pos_sum, ABBA, BABA, D, fd, fdM, sitesUsed] are left in the DEVICE buffer d_rec.  No synchronisation.
int pg_abba_enqueue(pg_ctx* ctx, const int* sel, double min_data, void* d_rec) {
    PG_CHECK(ctx->P >= 1, "pg_abbababa: call pg_set_pops first");
    for (int k = 0; k < 4; ++k) {
        PG_CHECK(sel[k] >= 0 && sel[k] < ctx->P, "pg_abbababa: population index %d out of range", sel[k]);
        for (int j = 0; j < k; ++j) PG_CHECK(sel[j] != sel[k], "pg_abbababa: populations must be distinct");
    }
    PG_CUDA(cudaSetDevice(ctx->device));
    pg_timings_reset(ctx);
    const int64_t W = ctx->W;
    if (W == 0) return PG_OK;
    const int Q = 9, RC = 8;
    if (ctx->S == 0) {
        std::vector<unsigned long long> h((size_t)W * RC, 0ull);
        const double qn = NAN;
        unsigned long long nanbits;
        memcpy(&nanbits, &qn, 8);
        for (int64_t w = 0; w < W; ++w)
            for (int k = 2; k < RC; ++k) h[(size_t)w * RC + k] = nanbits;
        PG_CUDA(cudaMemcpyAsync(d_rec, h.data(), h.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        PG_CUDA(cudaStreamSynchronize(ctx->stream));
        return PG_OK;
    }
    K1Cache& c = *cache_of(ctx, 1);
    if (!c.valid || c.epoch != ctx->epoch || memcmp(c.sel, sel, 4 * sizeof(int)) != 0) {
        c.valid = false;
        std::vector<int32_t> local(ctx->H, -1);
        for (int h = 0; h < ctx->H; ++h)
            for (int k = 0; k < 4; ++k)
                if (ctx->hap_pop[h] == sel[k]) local[h] = k;
        PG_TRY(prepare_windowed(ctx, c, local, 4, Q, k1_nw_for(ctx->pitch)));
        for (int k = 0; k < 4; ++k) PG_CHECK(c.pt.popN[k] >= 1, "pg_abbababa: population %d has no haplotypes", sel[k]);
        memcpy(c.sel, sel, 4 * sizeof(int));
        c.epoch = ctx->epoch;
        c.valid = true;
    }
    for (int k = 0; k < 4; ++k) {
        // smallest n with (double)n / N >= minData  (genomics.py:1657-1660, exact in integers)
        int thr = c.pt.popN[k] + 1;
        for (int n = 0; n <= c.pt.popN[k]; ++n)
            if ((double)n * 1.0 / (double)c.pt.popN[k] >= min_data) {
                thr = n;
                break;
            }
        c.L.prm.thr[k] = thr;
    }
    PG_TRY(arm_slots(ctx, c));
    PG_TRY((launch_site_pass<MODE_ABBA, 4>(ctx, c.L, "k1_abba")));
    FinParams fp;
    fill_fin(fp, ctx, c, Q, 3);
    fp.P = 4;
    fp.Ppad = 4;
    fp.rec = (unsigned long long*)d_rec;
    fp.RC = RC;
    const int ti = pg_time_begin(ctx, "k1_finalize");
    k1_finalize<MODE_ABBA><<<(unsigned)std::min<int64_t>(W, 65535), 64, 0, ctx->stream>>>(fp);
    pg_time_end(ctx, ti);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_abbababa(pg_ctx* ctx, int32_t p1, int32_t p2, int32_t p3, int32_t o, double min_data, double* out,
                           double* sites_used, int64_t* n_sites, int64_t* pos_sum) {
    PG_CHECK(ctx && out && sites_used && n_sites && pos_sum, "pg_abbababa: null argument");
    const int sel[4] = {p1, p2, p3, o};
    const int RC = 8;
    const int64_t W = ctx->W;
    PG_CUDA(cudaSetDevice(ctx->device));
    PG_TRY(ctx->out_d.ensure((size_t)std::max<int64_t>(W, 1) * RC * 8 + 64));
    PG_TRY(pg_abba_enqueue(ctx, sel, min_data, ctx->out_d.p));
    if (W == 0) return PG_OK;
    void* hp = nullptr;
    PG_TRY(pg_pinned(ctx, (size_t)W * RC * 8 + 64, &hp));
    const unsigned long long* hrec = (const unsigned long long*)hp;
    PG_CUDA(cudaMemcpyAsync(hp, ctx->out_d.p, (size_t)W * RC * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int64_t w = 0; w < W; ++w) {
        const unsigned long long* r = hrec + (size_t)w * RC;
        n_sites[w] = (int64_t)r[0];
        pos_sum[w] = (int64_t)r[1];
        memcpy(out + (size_t)w * 5, r + 2, 40);
        memcpy(sites_used + w, r + 7, 8);
    }
    return PG_OK;
}

// ================================================================================================
// pg_fourpop
// ================================================================================================
// Enqueue site pass + finalize of genomics.fourPop; records [W x 17] words [sites, pos_sum, 14 statistics, sitesUsed]
// are left in the DEVICE buffer d_rec.  No synchronisation.
int pg_fourpop_enqueue(pg_ctx* ctx, const int* sel, double min_data, int mode, void* d_rec) {
    PG_CHECK(ctx->P >= 1, "pg_fourpop: call pg_set_pops first");
    PG_CHECK(mode >= 0 && mode <= 2, "pg_fourpop: mode must be 0 (default), 1 (polarize) or 2 (fixed)");
    for (int k = 0; k < 4; ++k) {
        PG_CHECK(sel[k] >= 0 && sel[k] < ctx->P, "pg_fourpop: population index %d out of range", sel[k]);
        for (int j = 0; j < k; ++j) PG_CHECK(sel[j] != sel[k], "pg_fourpop: populations must be distinct");
    }
    PG_CUDA(cudaSetDevice(ctx->device));
    pg_timings_reset(ctx);
    const int64_t W = ctx->W;
    if (W == 0) return PG_OK;
    const int Q = 19, RC = 17;
    if (ctx->S == 0) {
        std::vector<unsigned long long> h((size_t)W * RC, 0ull);       // sites, pos_sum, sitesUsed = 0.0
        const double qn = NAN;
        unsigned long long nanbits;
        memcpy(&nanbits, &qn, 8);
        for (int64_t w = 0; w < W; ++w)
            for (int k = 2; k < 16; ++k) h[(size_t)w * RC + k] = nanbits;
        PG_CUDA(cudaMemcpyAsync(d_rec, h.data(), h.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        PG_CUDA(cudaStreamSynchronize(ctx->stream));
        return PG_OK;
    }
    K1Cache& c = *cache_of(ctx, 2);
    if (!c.valid || c.epoch != ctx->epoch || memcmp(c.sel, sel, 4 * sizeof(int)) != 0) {
        c.valid = false;
        std::vector<int32_t> local(ctx->H, -1);
        for (int h = 0; h < ctx->H; ++h)
            for (int k = 0; k < 4; ++k)
                if (ctx->hap_pop[h] == sel[k]) local[h] = k;
        PG_TRY(prepare_windowed(ctx, c, local, 4, Q, k1_nw_for(ctx->pitch)));
        for (int k = 0; k < 4; ++k) PG_CHECK(c.pt.popN[k] >= 1, "pg_fourpop: population %d has no haplotypes", sel[k]);
        memcpy(c.sel, sel, 4 * sizeof(int));
        c.epoch = ctx->epoch;
        c.valid = true;
    }
    for (int k = 0; k < 4; ++k) {
        int thr = c.pt.popN[k] + 1;       // smallest n with (double)n / N >= minData (genomics.py:1597-1600)
        for (int n = 0; n <= c.pt.popN[k]; ++n)
            if ((double)n * 1.0 / (double)c.pt.popN[k] >= min_data) {
                thr = n;
                break;
            }
        c.L.prm.thr[k] = thr;
    }
    c.L.prm.variant = mode;
    PG_TRY(arm_slots(ctx, c));
    if (getenv("PG_K1_FOURPOP_QUEUE")) PG_TRY((launch_site_pass<MODE_FOURPOP_Q, 4>(ctx, c.L, "k1_fourpop")));
    else PG_TRY((launch_site_pass<MODE_FOURPOP, 4>(ctx, c.L, "k1_fourpop")));
    FinParams fp;
    fill_fin(fp, ctx, c, Q, 3);
    fp.P = 4;
    fp.Ppad = 4;
    fp.rec = (unsigned long long*)d_rec;
    fp.RC = RC;
    const int ti = pg_time_begin(ctx, "k1_finalize");
    k1_finalize<MODE_FOURPOP><<<(unsigned)std::min<int64_t>(W, 65535), 64, 0, ctx->stream>>>(fp);
    pg_time_end(ctx, ti);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_fourpop(pg_ctx* ctx, int32_t p1, int32_t p2, int32_t p3, int32_t p4, double min_data, int32_t mode,
                          double* out, double* sites_used, int64_t* n_sites, int64_t* pos_sum) {
    PG_CHECK(ctx && out && sites_used && n_sites && pos_sum, "pg_fourpop: null argument");
    const int sel[4] = {p1, p2, p3, p4};
    const int RC = 17;
    const int64_t W = ctx->W;
    PG_CUDA(cudaSetDevice(ctx->device));
    PG_TRY(ctx->out_d.ensure((size_t)std::max<int64_t>(W, 1) * RC * 8 + 64));
    PG_TRY(pg_fourpop_enqueue(ctx, sel, min_data, mode, ctx->out_d.p));
    if (W == 0) return PG_OK;
    void* hp = nullptr;
    PG_TRY(pg_pinned(ctx, (size_t)W * RC * 8 + 64, &hp));
    const unsigned long long* hrec = (const unsigned long long*)hp;
    PG_CUDA(cudaMemcpyAsync(hp, ctx->out_d.p, (size_t)W * RC * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int64_t w = 0; w < W; ++w) {
        const unsigned long long* r = hrec + (size_t)w * RC;
        n_sites[w] = (int64_t)r[0];
        pos_sum[w] = (int64_t)r[1];
        memcpy(out + (size_t)w * 14, r + 2, 14 * 8);
        memcpy(sites_used + w, r + 16, 8);
    }
    return PG_OK;
}

// ================================================================================================
// pg_site_counts / pg_site_target_freqs
// ================================================================================================
namespace {
// Many small populations (freq.py --indFreqs: one population per individual): one thread per (site, population) gathers
// the population's few haplotype bytes — ONE pass over the rows instead of P/8 site passes.
__global__ void __launch_bounds__(256) k1_counts_gather(const uint8_t* __restrict__ geno, int pitch, int64_t site0, int64_t n,
                                                        int P, const int32_t* __restrict__ pop_off,
                                                        const int32_t* __restrict__ pop_cols, uint16_t* __restrict__ out) {
    const int64_t total = n * P;
    for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
        const int64_t s = idx / P;
        const int X = (int)(idx % P);
        const uint8_t* row = geno + (site0 + s) * pitch;
        unsigned a = 0, c = 0, g = 0, t = 0;
        for (int k = pop_off[X]; k < pop_off[X + 1]; ++k) {
            const unsigned b = row[pop_cols[k]];            // one-hot: A 0x01, C 0x04, G 0x10, T 0x40
            a += b & 1u;
            c += (b >> 2) & 1u;
            g += (b >> 4) & 1u;
            t += (b >> 6) & 1u;
        }
        reinterpret_cast<ushort4*>(out)[idx] = make_ushort4((unsigned short)a, (unsigned short)c, (unsigned short)g, (unsigned short)t);
    }
}

// per-site counts of `cnt` sites starting at `first` -> ctx->misc as uint16 [cnt x P x 4]
int site_counts_slab(pg_ctx* ctx, int64_t first, int64_t cnt) {
    const int P = ctx->P;
    const int64_t stride = (int64_t)P * 4;
    if (P > 16 && !getenv("PG_COUNTS_NO_GATHER")) {
        std::vector<int32_t> off(P + 1, 0), cols;
        cols.reserve((size_t)ctx->H);
        for (int X = 0; X < P; ++X) {
            off[X] = (int32_t)cols.size();
            for (int h = 0; h < ctx->H; ++h)
                if (ctx->hap_pop[h] == X) cols.push_back(h);
        }
        off[P] = (int32_t)cols.size();
        PG_TRY(ctx->tables.ensure((size_t)(P + 1) * 4 + cols.size() * 4 + 256));
        uint8_t* base = (uint8_t*)ctx->tables.p;
        size_t o = 0;
        int32_t *d_off = nullptr, *d_cols = nullptr;
        PG_TRY(push(ctx, base, o, off.data(), off.size(), &d_off));
        PG_TRY(push(ctx, base, o, cols.data(), cols.size(), &d_cols));
        const int ti = pg_time_begin(ctx, "k1_counts");
        k1_counts_gather<<<(unsigned)std::min<int64_t>((cnt * P + 255) / 256, (int64_t)ctx->sm_count * 32), 256, 0, ctx->stream>>>(
            (const uint8_t*)ctx->d_geno, ctx->pitch, first, cnt, P, d_off, d_cols, (uint16_t*)ctx->misc.p);
        pg_time_end(ctx, ti);
        PG_CUDA(cudaGetLastError());
        PG_CUDA(cudaStreamSynchronize(ctx->stream));       // host vectors behind the async copies
        return PG_OK;
    }
    for (int p0 = 0; p0 < P; p0 += PG_MAX_K1_POPS) {
        const int pc = std::min(PG_MAX_K1_POPS, P - p0);
        const int Pp = pad_pops(pc);
        std::vector<int32_t> local(ctx->H, -1);
        for (int h = 0; h < ctx->H; ++h)
            if (ctx->hap_pop[h] >= p0 && ctx->hap_pop[h] < p0 + pc) local[h] = ctx->hap_pop[h] - p0;
        PopTables pt;
        build_tables(local, ctx->H, ctx->pitch / 16, Pp, pt);
        const int n_ent = (int)pt.ent_chunk.size();
        const int table_bytes = n_ent * 20 + 64 + 512;
        PG_CHECK(table_bytes <= 48 * 1024, "population layout needs too many mask entries");
        K1Launch L;
        // long rows, a full group of 4 or 8 populations: one lane per population (counts fit 16 bits in any case)
        bool lp = (Pp == 4 || Pp == 8) && Pp == pc && ctx->pitch >= 1024;
        if (const char* e = getenv("PG_K1_LANEPOP")) lp = atoi(e) != 0 && (Pp == 4 || Pp == 8);
        const int nw = lp ? k1_env_nw12() : k1_nw_for(ctx->pitch);
        L.plan = pg_make_k1_plan(cnt, ctx->H, ctx->sm_count, table_bytes, nw, lp ? Pp : 0);
        PG_CHECK(L.plan.stages >= 2, "rows of %d haplotypes are too long for the site-pass kernel", ctx->H);
        PG_TRY(check_plan(L.plan));
        PG_TRY(ctx->tables.ensure((size_t)n_ent * 20 + 4096));
        uint8_t* base = (uint8_t*)ctx->tables.p;
        size_t o = 0;
        uint32_t* d_mask_words = nullptr;
        int32_t* d_chunk = nullptr;
        PG_TRY(push(ctx, base, o, pt.ent_mask.data(), pt.ent_mask.size(), &d_mask_words));
        PG_TRY(push(ctx, base, o, pt.ent_chunk.data(), pt.ent_chunk.size(), &d_chunk));
        K1Params& p = L.prm;
        memset(&p, 0, sizeof(p));
        p.geno = (const uint8_t*)ctx->d_geno;
        p.pos = ctx->d_pos;
        p.site_begin = first;
        p.site_end = first + cnt;
        p.num_tiles = L.plan.num_tiles;
        p.pitch = L.plan.pitch;
        p.G = L.plan.G;
        p.I = L.plan.I;
        p.T = L.plan.T;
        p.wpt = L.plan.wpt;
        p.nw = nw;
        p.stages = L.plan.stages;
        p.tile_bytes = L.plan.tile_bytes;
        p.ent_chunk = d_chunk;
        p.ent_mask = reinterpret_cast<uint4*>(d_mask_words);
        p.n_ent = n_ent;
        for (int X = 0; X < PG_MAX_K1_POPS; ++X) {
            p.ent_lo[X] = pt.ent_lo[X];
            p.ent_hi[X] = pt.ent_hi[X];
            p.full_lo[X] = pt.full_lo[X];
            p.full_hi[X] = pt.full_hi[X];
            p.popN[X] = pt.popN[X];
        }
        p.counts_out = (uint16_t*)ctx->misc.p + (size_t)p0 * 4;
        p.counts_stride = stride;
        p.counts_pops = pc;
        p.lanepop = lp ? 1 : 0;
        if (Pp == 2) PG_TRY((launch_site_pass<MODE_COUNTS, 2>(ctx, L, "k1_counts")));
        else if (Pp == 4) PG_TRY((launch_site_pass<MODE_COUNTS, 4>(ctx, L, "k1_counts")));
        else PG_TRY((launch_site_pass<MODE_COUNTS, 8>(ctx, L, "k1_counts")));
        // the table buffer (and the host vectors behind the async copies) are reused by the next group
        PG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return PG_OK;
}

// freq.py --target (freq.py:62-92): one thread per site turns the per-population counts into the frequency (or
// count) of the target allele.  target 1 = derived (last population is the outgroup; derivedAllele, genomics.py:636-659),
// 2 = minor (minorAllele, 663-668; an exact tie, random in the reference, takes the lower allele and is flagged).
__global__ void __launch_bounds__(256) k1_target_freqs(const uint16_t* __restrict__ counts, int64_t n, int P, int target,
                                                       double min_data, int as_counts, double* __restrict__ out,
                                                       uint8_t* __restrict__ tie) {
    const int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (s >= n) return;
    const ushort4* c = reinterpret_cast<const ushort4*>(counts) + s * P;
    unsigned in[4] = {0, 0, 0, 0}, outg[4] = {0, 0, 0, 0};
    for (int X = 0; X < P; ++X) {
        const ushort4 v = c[X];
        unsigned* dst = (target == 1 && X == P - 1) ? outg : in;
        dst[0] += v.x;
        dst[1] += v.y;
        dst[2] += v.z;
        dst[3] += v.w;
    }
    int tgt = -1;
    bool tied = false;
    if (target == 1) {
        int n_in = 0, n_out = 0, oa = -1;
        for (int a = 0; a < 4; ++a) {
            n_in += in[a] > 0;
            if (outg[a] > 0) {
                ++n_out;
                oa = a;
            }
        }
        if (n_out == 1 && n_in == 2 && in[oa] > 0)
            for (int a = 0; a < 4; ++a)
                if (in[a] > 0 && a != oa) tgt = a;
    } else {
        int a0 = -1, a1 = -1, na = 0;
        for (int a = 0; a < 4; ++a)
            if (in[a] > 0) {
                if (na == 0) a0 = a; else a1 = a;
                ++na;
            }
        if (na == 2) {
            tied = in[a0] == in[a1];
            tgt = in[a1] < in[a0] ? a1 : a0;
        }
    }
    if (tie) tie[s] = tied ? 1 : 0;
    const double none = as_counts ? 0.0 : __longlong_as_double(0x7ff8000000000000ll);
    for (int X = 0; X < P; ++X) {
        const ushort4 v = c[X];
        const unsigned nk = (unsigned)v.x + v.y + v.z + v.w;
        double r = none;
        if (tgt >= 0 && (double)nk >= min_data) {            // siteNonNan() >= minData compares a COUNT (freq.py:79)
            const unsigned k = tgt == 0 ? v.x : (tgt == 1 ? v.y : (tgt == 2 ? v.z : v.w));
            if (as_counts) r = (double)k;
            else if (nk > 0) r = (double)k / (double)nk;     // nan when the population has no data (genomics.py:597)
        }
        out[s * P + X] = r;
    }
}
}  // namespace

extern "C" int pg_site_counts(pg_ctx* ctx, int64_t site0, int64_t n, uint16_t* counts) {
    PG_CHECK(ctx && counts, "pg_site_counts: null argument");
    PG_CHECK(ctx->P >= 1, "pg_site_counts: call pg_set_pops first");
    PG_CHECK(site0 >= 0 && n >= 0 && site0 + n <= ctx->S, "pg_site_counts: range outside the uploaded sites");
    PG_CUDA(cudaSetDevice(ctx->device));
    pg_timings_reset(ctx);
    if (n == 0) return PG_OK;
    const int64_t stride = (int64_t)ctx->P * 4;
    // bounded device output buffer: process the range in slabs
    const int64_t slab = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)(1ll << 30) / (stride * 2)));
    PG_TRY(ctx->misc.ensure((size_t)slab * stride * 2 + 64));
    for (int64_t s = 0; s < n; s += slab) {
        const int64_t cnt = std::min(slab, n - s);
        PG_TRY(site_counts_slab(ctx, site0 + s, cnt));
        PG_TRY(pg_d2h_staged(ctx, counts + (size_t)s * stride, ctx->misc.p, (size_t)cnt * stride * 2));
    }
    return PG_OK;
}

extern "C" int pg_site_target_freqs(pg_ctx* ctx, int64_t site0, int64_t n, int32_t target, double min_data,
                                    int32_t as_counts, double* out, uint8_t* tie) {
    PG_CHECK(ctx && out, "pg_site_target_freqs: null argument");
    PG_CHECK(ctx->P >= 1, "pg_site_target_freqs: call pg_set_pops first");
    PG_CHECK(target == 1 || target == 2, "pg_site_target_freqs: target must be 1 (derived) or 2 (minor)");
    PG_CHECK(target != 1 || ctx->P >= 2, "pg_site_target_freqs: derived needs an outgroup population (the last one)");
    PG_CHECK(site0 >= 0 && n >= 0 && site0 + n <= ctx->S, "pg_site_target_freqs: range outside the uploaded sites");
    PG_CUDA(cudaSetDevice(ctx->device));
    pg_timings_reset(ctx);
    if (n == 0) return PG_OK;
    const int P = ctx->P;
    const int64_t stride = (int64_t)P * 4;
    const int64_t slab = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)(1ll << 28) / (stride * 2)));
    PG_TRY(ctx->misc.ensure((size_t)slab * stride * 2 + 64));
    PG_TRY(ctx->out_d.ensure((size_t)slab * P * 8 + (size_t)slab + 64));
    double* d_out = (double*)ctx->out_d.p;
    uint8_t* d_tie = (uint8_t*)(d_out + (size_t)slab * P);
    for (int64_t s = 0; s < n; s += slab) {
        const int64_t cnt = std::min(slab, n - s);
        PG_TRY(site_counts_slab(ctx, site0 + s, cnt));
        const int ti = pg_time_begin(ctx, "k1_target_freqs");
        k1_target_freqs<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>((const uint16_t*)ctx->misc.p, cnt, P, target,
                                                                              min_data, as_counts ? 1 : 0, d_out, d_tie);
        pg_time_end(ctx, ti);
        PG_CUDA(cudaGetLastError());
        PG_TRY(pg_d2h_staged(ctx, out + (size_t)s * P, d_out, (size_t)cnt * P * 8));
        if (tie) PG_TRY(pg_d2h_staged(ctx, tie + s, d_tie, (size_t)cnt));
    }
    return PG_OK;
}

// ================================================================================================
// pg_sfs — sfs.py's per-site loop for --inputType genotypes (sfs.py:430-470, getTargetCounts 68-92, SparseFS 94-125)
// ================================================================================================
namespace {
struct SfsParams {
    const uint16_t* counts;     // [n x P x 4] of this slab
    int64_t n, site0;           // sites of the slab, absolute index of its first site
    int P, n_in, outgroup;      // in-group = populations 0 .. n_in-1; outgroup = population index or -1
    int popN[PG_MAX_POPS];      // haplotypes per population: an in-group population must be complete (sfs.py:449)
    int dims[PG_MAX_POPS];      // radix of each population in the histograms (largest possible count + 1)
    int require_complete;       // genotype input only
    const int32_t* targets;     // targetCounts input: [n x P] counts of the target allele, no allele logic at all
    const uint8_t* mask;        // [S] absolute, or nullptr
    int n_groups;
    const int32_t* group_off;   // [n_groups+1] into group_pops
    const int32_t* group_pops;
    const long long* hist_off;  // [n_groups] first cell of each group's dense histogram
    unsigned long long* hist;
    long long* first;           // first site (absolute) that hit the cell
    unsigned long long* n_counted;
};

__global__ void __launch_bounds__(256) k1_sfs(const __grid_constant__ SfsParams sp) {
    const int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (s >= sp.n) return;
    const int64_t site = sp.site0 + s;
    if (sp.mask && !sp.mask[site]) return;
    if (sp.targets) {                                                // sfs.py:472-474: the table already holds the counts
        const int32_t* t = sp.targets + s * sp.P;
        atomicAdd(sp.n_counted, 1ull);
        for (int g = 0; g < sp.n_groups; ++g) {
            long long idx = 0;
            for (int k = sp.group_off[g]; k < sp.group_off[g + 1]; ++k) idx = idx * sp.dims[sp.group_pops[k]] + t[sp.group_pops[k]];
            atomicAdd(sp.hist + sp.hist_off[g] + idx, 1ull);
            atomicMin(sp.first + sp.hist_off[g] + idx, (long long)site);
        }
        return;
    }
    const ushort4* c = reinterpret_cast<const ushort4*>(sp.counts) + s * sp.P;
    unsigned tot[4] = {0, 0, 0, 0};
    for (int X = 0; X < sp.n_in; ++X) {
        const ushort4 v = c[X];
        if (sp.require_complete && (int)v.x + v.y + v.z + v.w != sp.popN[X]) return;   // every in-group haplotype called (449)
        tot[0] += v.x;
        tot[1] += v.y;
        tot[2] += v.z;
        tot[3] += v.w;
    }
    int target;
    if (sp.outgroup >= 0) {
        const ushort4 o = c[sp.outgroup];
        const unsigned oc[4] = {o.x, o.y, o.z, o.w};
        int n_all = 0, n_out = 0;
        for (int a = 0; a < 4; ++a) {
            n_all += (tot[a] > 0 || oc[a] > 0) ? 1 : 0;
            n_out += oc[a] > 0 ? 1 : 0;
        }
        if (n_all < 1 || n_all > 2) return;                          // 79
        // `outgroupMono & nOutAlleles != 1` is (outgroupMono & nOutAlleles) != 1: the count must be odd, i.e. 1 (84)
        if (n_out == 0 || (n_out & 1) != 1) return;
        target = -1;
        for (int a = 3; a >= 0; --a)
            if (tot[a] > 0 && oc[a] == 0) target = a;               // first in-group allele the outgroup lacks (86)
        if (target < 0)
            for (int a = 3; a >= 0; --a)
                if (tot[a] == 0) target = a;                         // invariant: first absent allele (87), count 0
        if (target < 0) return;
    } else {
        int n_all = 0;
        for (int a = 0; a < 4; ++a) n_all += tot[a] > 0 ? 1 : 0;
        if (n_all < 1 || n_all > 2) return;
        // totalBaseCounts.argsort()[-2] (90): second in a stable ascending order = with two alleles the rarer one, the
        // lower allele on an exact tie; with one allele an absent allele (count 0 everywhere)
        int best = -1, second = -1;                                  // positions [-1] and [-2] of the stable argsort
        for (int a = 0; a < 4; ++a) {
            if (best < 0 || tot[a] >= tot[best]) {
                second = best;
                best = a;
            } else if (second < 0 || tot[a] >= tot[second]) second = a;
        }
        target = second;
    }
    atomicAdd(sp.n_counted, 1ull);
    for (int g = 0; g < sp.n_groups; ++g) {
        long long idx = 0;
        for (int k = sp.group_off[g]; k < sp.group_off[g + 1]; ++k) {
            const int X = sp.group_pops[k];
            const ushort4 v = c[X];
            const unsigned t = target == 0 ? v.x : (target == 1 ? v.y : (target == 2 ? v.z : v.w));
            idx = idx * sp.dims[X] + t;
        }
        atomicAdd(sp.hist + sp.hist_off[g] + idx, 1ull);
        atomicMin(sp.first + sp.hist_off[g] + idx, (long long)site);
    }
}
}  // namespace

extern "C" int pg_sfs(pg_ctx* ctx, int32_t n_in, int32_t outgroup, int32_t n_groups, const int32_t* group_off,
                      const int32_t* group_pops, const uint8_t* site_mask, int64_t* hist, int64_t* first, int64_t* n_counted) {
    PG_CHECK(ctx && group_off && group_pops && hist && first, "pg_sfs: null argument");
    PG_CHECK(ctx->P >= 1 && ctx->P <= PG_MAX_POPS, "pg_sfs: call pg_set_pops first (at most %d populations)", PG_MAX_POPS);
    PG_CHECK(n_in >= 1 && n_in <= ctx->P, "pg_sfs: n_in out of range");
    PG_CHECK(outgroup == -1 || (outgroup >= n_in && outgroup < ctx->P), "pg_sfs: the outgroup must be a population after the in-group");
    PG_CHECK(n_groups >= 1, "pg_sfs: no spectra requested");
    PG_CUDA(cudaSetDevice(ctx->device));
    pg_timings_reset(ctx);
    const int P = ctx->P;
    std::vector<int> popN(P, 0);
    for (int h = 0; h < ctx->H; ++h)
        if (ctx->hap_pop[h] >= 0) popN[ctx->hap_pop[h]] += 1;
    std::vector<long long> hist_off(n_groups, 0);
    long long cells = 0;
    for (int g = 0; g < n_groups; ++g) {
        hist_off[g] = cells;
        long long sz = 1;
        PG_CHECK(group_off[g + 1] > group_off[g], "pg_sfs: spectrum %d has no population", g);
        for (int k = group_off[g]; k < group_off[g + 1]; ++k) {
            PG_CHECK(group_pops[k] >= 0 && group_pops[k] < n_in, "pg_sfs: spectrum %d uses a population outside the in-group", g);
            sz *= (long long)popN[group_pops[k]] + 1;
            PG_CHECK(sz <= (1ll << 28), "pg_sfs: spectrum %d is too large for a dense histogram", g);
        }
        cells += sz;
        PG_CHECK(cells <= (1ll << 28), "pg_sfs: the spectra need more than 2^28 cells");
    }
    if (n_counted) *n_counted = 0;
    const int n_gp = group_off[n_groups];
    // device buffers: histograms | first | tables | counter
    PG_TRY(ctx->pairs.ensure((size_t)cells * 16 + 64));
    unsigned long long* d_hist = (unsigned long long*)ctx->pairs.p;
    long long* d_first = (long long*)(d_hist + cells);
    PG_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)cells * 8, ctx->stream));
    PG_CUDA(cudaMemsetAsync(d_first, 0x7f, (size_t)cells * 8, ctx->stream));
    PG_TRY(ctx->misc2.ensure((size_t)(n_groups + 1) * 4 + (size_t)n_gp * 4 + (size_t)n_groups * 8 + 64 + 64));
    uint8_t* tb = (uint8_t*)ctx->misc2.p;
    size_t o = 0;
    int32_t* d_goff = nullptr;
    int32_t* d_gpops = nullptr;
    long long* d_hoff = nullptr;
    PG_TRY(push(ctx, tb, o, group_off, (size_t)n_groups + 1, &d_goff));
    PG_TRY(push(ctx, tb, o, group_pops, (size_t)n_gp, &d_gpops));
    PG_TRY(push(ctx, tb, o, hist_off.data(), (size_t)n_groups, &d_hoff));
    PG_TRY(ctx->out_i.ensure(64));
    unsigned long long* d_cnt = (unsigned long long*)ctx->out_i.p;
    PG_CUDA(cudaMemsetAsync(d_cnt, 0, 8, ctx->stream));
    uint8_t* d_mask = nullptr;
    if (site_mask && ctx->S > 0) {
        PG_TRY(ctx->misc3.ensure((size_t)ctx->S + 64));
        d_mask = (uint8_t*)ctx->misc3.p;
        PG_CUDA(cudaMemcpyAsync(d_mask, site_mask, (size_t)ctx->S, cudaMemcpyHostToDevice, ctx->stream));
    }
    PG_CUDA(cudaStreamSynchronize(ctx->stream));       // host tables behind the async copies
    const int64_t stride = (int64_t)P * 4;
    const int64_t slab = std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(ctx->S, 1), (int64_t)(1ll << 28) / (stride * 2)));
    PG_TRY(ctx->misc.ensure((size_t)slab * stride * 2 + 64));
    for (int64_t s0 = 0; s0 < ctx->S; s0 += slab) {
        const int64_t cnt = std::min(slab, ctx->S - s0);
        PG_TRY(site_counts_slab(ctx, s0, cnt));
        SfsParams sp;
        memset(&sp, 0, sizeof(sp));
        sp.counts = (const uint16_t*)ctx->misc.p;
        sp.n = cnt;
        sp.site0 = s0;
        sp.P = P;
        sp.n_in = n_in;
        sp.outgroup = outgroup;
        for (int X = 0; X < P; ++X) {
            sp.popN[X] = popN[X];
            sp.dims[X] = popN[X] + 1;
        }
        sp.require_complete = 1;
        sp.mask = d_mask;
        sp.n_groups = n_groups;
        sp.group_off = d_goff;
        sp.group_pops = d_gpops;
        sp.hist_off = d_hoff;
        sp.hist = d_hist;
        sp.first = d_first;
        sp.n_counted = d_cnt;
        const int ti = pg_time_begin(ctx, "k1_sfs");
        k1_sfs<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(sp);
        pg_time_end(ctx, ti);
        PG_CUDA(cudaGetLastError());
    }
    unsigned long long h_cnt = 0;
    PG_CUDA(cudaMemcpyAsync(hist, d_hist, (size_t)cells * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(cudaMemcpyAsync(first, d_first, (size_t)cells * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(cudaMemcpyAsync(&h_cnt, d_cnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    for (long long k = 0; k < cells; ++k)
        if (hist[k] == 0) first[k] = -1;
    if (n_counted) *n_counted = (int64_t)h_cnt;
    return PG_OK;
}

// The same spectra from TABLES of counts on the host (sfs.py --inputType baseCounts | targetCounts, sfs.py:456-474): no
// genotype matrix, no completeness test.  kind 0: table = uint16 [n x P x 4] base counts per population (freq.py's
// default output); kind 1: table = int32 [n x P] counts of the target allele.  dims[X] = largest count of population X + 1.
extern "C" int pg_sfs_tables(pg_ctx* ctx, int32_t kind, const void* table, int64_t n, int32_t P, const int32_t* dims,
                             int32_t n_in, int32_t outgroup, int32_t n_groups, const int32_t* group_off,
                             const int32_t* group_pops, const uint8_t* site_mask, int64_t* hist, int64_t* first,
                             int64_t* n_counted) {
    PG_CHECK(ctx && (table || n == 0) && dims && group_off && group_pops && hist && first, "pg_sfs_tables: null argument");
    PG_CHECK(kind == 0 || kind == 1, "pg_sfs_tables: kind must be 0 (base counts) or 1 (target counts)");
    PG_CHECK(P >= 1 && P <= PG_MAX_POPS, "pg_sfs_tables: at most %d populations", PG_MAX_POPS);
    PG_CHECK(n_in >= 1 && n_in <= P && n >= 0, "pg_sfs_tables: bad shape");
    PG_CHECK(outgroup == -1 || (kind == 0 && outgroup >= n_in && outgroup < P), "pg_sfs_tables: bad outgroup");
    PG_CHECK(n_groups >= 1, "pg_sfs_tables: no spectra requested");
    PG_CUDA(cudaSetDevice(ctx->device));
    pg_timings_reset(ctx);
    std::vector<long long> hist_off(n_groups, 0);
    long long cells = 0;
    for (int g = 0; g < n_groups; ++g) {
        hist_off[g] = cells;
        long long sz = 1;
        PG_CHECK(group_off[g + 1] > group_off[g], "pg_sfs_tables: spectrum %d has no population", g);
        for (int k = group_off[g]; k < group_off[g + 1]; ++k) {
            PG_CHECK(group_pops[k] >= 0 && group_pops[k] < n_in, "pg_sfs_tables: spectrum %d uses a population outside the in-group", g);
            PG_CHECK(dims[group_pops[k]] >= 1, "pg_sfs_tables: dims must be >= 1");
            sz *= (long long)dims[group_pops[k]];
            PG_CHECK(sz <= (1ll << 28), "pg_sfs_tables: spectrum %d is too large for a dense histogram", g);
        }
        cells += sz;
        PG_CHECK(cells <= (1ll << 28), "pg_sfs_tables: the spectra need more than 2^28 cells");
    }
    if (n_counted) *n_counted = 0;
    const int n_gp = group_off[n_groups];
    PG_TRY(ctx->pairs.ensure((size_t)cells * 16 + 64));
    unsigned long long* d_hist = (unsigned long long*)ctx->pairs.p;
    long long* d_first = (long long*)(d_hist + cells);
    PG_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)cells * 8, ctx->stream));
    PG_CUDA(cudaMemsetAsync(d_first, 0x7f, (size_t)cells * 8, ctx->stream));
    PG_TRY(ctx->misc2.ensure((size_t)(n_groups + 1) * 4 + (size_t)n_gp * 4 + (size_t)n_groups * 8 + 128));
    uint8_t* tb = (uint8_t*)ctx->misc2.p;
    size_t o = 0;
    int32_t* d_goff = nullptr;
    int32_t* d_gpops = nullptr;
    long long* d_hoff = nullptr;
    PG_TRY(push(ctx, tb, o, group_off, (size_t)n_groups + 1, &d_goff));
    PG_TRY(push(ctx, tb, o, group_pops, (size_t)n_gp, &d_gpops));
    PG_TRY(push(ctx, tb, o, hist_off.data(), (size_t)n_groups, &d_hoff));
    PG_TRY(ctx->out_i.ensure(64));
    unsigned long long* d_cnt = (unsigned long long*)ctx->out_i.p;
    PG_CUDA(cudaMemsetAsync(d_cnt, 0, 8, ctx->stream));
    uint8_t* d_mask = nullptr;
    if (site_mask && n > 0) {
        PG_TRY(ctx->misc3.ensure((size_t)n + 64));
        d_mask = (uint8_t*)ctx->misc3.p;
        PG_CUDA(cudaMemcpyAsync(d_mask, site_mask, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    }
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    const size_t row_bytes = kind == 0 ? (size_t)P * 8 : (size_t)P * 4;
    const int64_t slab = std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(n, 1), (int64_t)(((size_t)256 << 20) / row_bytes)));
    PG_TRY(ctx->misc.ensure((size_t)slab * row_bytes + 64));
    for (int64_t s0 = 0; s0 < n; s0 += slab) {
        const int64_t cnt = std::min(slab, n - s0);
        PG_CUDA(cudaMemcpyAsync(ctx->misc.p, (const uint8_t*)table + (size_t)s0 * row_bytes, (size_t)cnt * row_bytes,
                                cudaMemcpyHostToDevice, ctx->stream));
        SfsParams sp;
        memset(&sp, 0, sizeof(sp));
        sp.counts = kind == 0 ? (const uint16_t*)ctx->misc.p : nullptr;
        sp.targets = kind == 1 ? (const int32_t*)ctx->misc.p : nullptr;
        sp.n = cnt;
        sp.site0 = s0;
        sp.P = P;
        sp.n_in = n_in;
        sp.outgroup = outgroup;
        for (int X = 0; X < P; ++X) sp.dims[X] = dims[X];
        sp.require_complete = 0;
        sp.mask = d_mask;
        sp.n_groups = n_groups;
        sp.group_off = d_goff;
        sp.group_pops = d_gpops;
        sp.hist_off = d_hoff;
        sp.hist = d_hist;
        sp.first = d_first;
        sp.n_counted = d_cnt;
        const int ti = pg_time_begin(ctx, "k1_sfs");
        k1_sfs<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(sp);
        pg_time_end(ctx, ti);
        PG_CUDA(cudaGetLastError());
        PG_CUDA(cudaStreamSynchronize(ctx->stream));           // the staging buffer is reused by the next slab
    }
    unsigned long long h_cnt = 0;
    PG_CUDA(cudaMemcpyAsync(hist, d_hist, (size_t)cells * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(cudaMemcpyAsync(first, d_first, (size_t)cells * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(cudaMemcpyAsync(&h_cnt, d_cnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    for (long long k = 0; k < cells; ++k)
        if (hist[k] == 0) first[k] = -1;
    if (n_counted) *n_counted = (int64_t)h_cnt;
    return PG_OK;
}
