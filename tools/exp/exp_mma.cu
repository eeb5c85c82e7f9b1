// Micro-experiment: issue-to-completion cost of tcgen05.mma (SS operands) for kind::tf32 vs kind::f16(bf16) at M=128.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_mma exp_mma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1;} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_kmajor_sw128(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
template <int KIND>   // 0 tf32, 1 bf16
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND>
__global__ void __launch_bounds__(128) k_mma(int n, int iters, int same_operands, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < 192 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tptr;
    if (threadIdx.x == 0) {
        const uint32_t fmt = KIND == 0 ? 2u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            // walk through 4 k-steps x 4 stages of 16 KB like the real pipeline, or hammer one operand
            const int st = same_operands ? 0 : (i >> 2) & 3, k = same_operands ? 0 : i & 3;
            mma<KIND>(tm, desc_kmajor_sw128(a0 + st * 16384 + k * 32), desc_kmajor_sw128(b0 + st * 16384 * 2 + k * 32), idesc, i > 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512) : "memory");
}
int main() {
    long long* d; CK(cudaMalloc(&d, 8));
    const int smem = 193 * 1024 + 1024;
    CK(cudaFuncSetAttribute(k_mma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 4096;
    for (int grid : {1, 148})
        for (int kind = 0; kind < 2; ++kind)
            for (int n : {64, 128, 208, 256})
                for (int same = 0; same < 2; ++same) {
                    if (kind == 0) k_mma<0><<<grid, 128, smem>>>(n, iters, same, d);
                    else k_mma<1><<<grid, 128, smem>>>(n, iters, same, d);
                    CK(cudaDeviceSynchronize());
                    long long c; CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
                    const double cyc = (double)c / iters;
                    const int K = kind == 0 ? 8 : 16;
                    printf("grid %3d  %s  M=128 N=%3d K=%2d  %s : %.1f cycles/MMA  -> %.0f flop/clk/SM\n", grid,
                           kind == 0 ? "tf32" : "bf16", n, K, same ? "same operand " : "4x4 operands ", cyc, 2.0 * 128 * n * K / cyc);
                }
    return 0;
}
