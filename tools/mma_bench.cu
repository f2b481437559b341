// Microbenchmark: cycles per tcgen05.mma (cta_group::1, M=128) for kind::i8 / kind::f16 / kind::f8f6f4, K-major no-swizzle
// operands in shared memory, N = 64..256.  One CTA per SM; one thread issues `iters` MMAs back to back, then commits.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) | (1ull << 46);
}
template <int KIND>
__global__ void __launch_bounds__(128, 1) bench(int N, int iters, int bstride, long long* out) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x;
    for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x01010101u * (i & 1);
    if (tid < 32) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
            asm volatile("fence.mbarrier_init.release.cluster;");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        uint32_t idesc;
        if (KIND == 0) idesc = (2u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);                       // i8: S32, U8 x U8
        else if (KIND == 1) idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);   // f16: F32, BF16
        else idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);                                  // f8f6f4: F32, E4M3
        const uint32_t a = smem_u32(sm), b = a + 4096;
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint64_t da = desc(a), db = desc(b + (i & 1) * bstride);
            if (KIND == 0)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(i));
            else if (KIND == 1)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(i));
            else
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(i));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        uint32_t ok;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        } while (!ok);
        t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
int main() {
    long long* d;
    cudaMalloc(&d, 148 * 8);
    const int iters = 2000;
    const char* names[3] = {"i8", "bf16", "e4m3"};
    for (int kind = 0; kind < 3; ++kind)
        for (int N : {64, 128, 256})
            for (int grid : {1, 148}) {
                void (*k)(int, int, int, long long*) = kind == 0 ? bench<0> : kind == 1 ? bench<1> : bench<2>;
                cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
                k<<<grid, 128, 48 * 1024>>>(N, iters, 8192, d);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[148];
                cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                printf("%s M=128 N=%d K=32B grid=%d: %.1f cycles/MMA (%s)\n", names[kind], N, grid, (double)mx / iters, cudaGetErrorString(e));
            }
    return 0;
}
