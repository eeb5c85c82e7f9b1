// Micro-experiment: cost of 1.7M random 64-byte row updates under different instruction choices.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_red exp_red.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ void red4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld64(const float* p) {
    float4 v;
    asm volatile("ld.global.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// mode 0: red.v4 ; 1: ld.L2::64B then red ; 2: st.v4 ; 3: ld + add + st (non-atomic RMW) ; 4: prefetch.L2 then red
// 5: scalar red x4
template <int MODE>
__global__ void k_rows(float* __restrict__ g, const long long* __restrict__ idx, long long n_upd, int D) {
    const int lpr = D / 4;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < n_upd * lpr; t += stride) {
        const long long u = t / lpr; const int q = (int)(t % lpr);
        const long long r = idx[u];
        float* p = g + r * D + 4 * q;
        float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
        if (MODE == 0) red4(p, v);
        else if (MODE == 1) { float4 o = ld64(p); if (o.x == 12345.678f) v.x = 0.f; red4(p, v); }
        else if (MODE == 2) *reinterpret_cast<float4*>(p) = v;
        else if (MODE == 3) { float4 o = ld64(p); o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w; *reinterpret_cast<float4*>(p) = o; }
        else if (MODE == 4) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); red4(p, v); }
        else if (MODE == 5) { atomicAdd(p, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w); }
    }
}
// mode 6: TMA bulk reduce, one 64-byte row per lane from shared memory
__global__ void k_bulk(float* __restrict__ g, const long long* __restrict__ idx, long long n_upd, int D) {
    extern __shared__ __align__(128) float sm[];
    float* mine = sm + (size_t)threadIdx.x * D;
    for (int i = 0; i < D; ++i) mine[i] = 1.f + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < n_upd; t += stride) {
        const long long r = idx[t];
        float* p = g + r * D;
        uint32_t s = (uint32_t)__cvta_generic_to_shared(mine);
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(p), "r"(s), "r"(D * 4) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
    const int D = 16, F = 26;
    const long long rows = 26LL * 1000001, n_upd = 65536LL * F;
    float* g; long long* idx[4];
    CK(cudaMalloc(&g, rows * D * 4));
    CK(cudaMemset(g, 0, rows * D * 4));
    std::vector<long long> h(n_upd);
    for (int b = 0; b < 4; ++b) {
        srand(17 + b);
        for (long long i = 0; i < n_upd; ++i) {
            long long f = i % F;   // field-major-ish like the real step: each sample touches one row of each table
            unsigned long long x = ((unsigned long long)rand() << 31) ^ rand();
            h[i] = f * 1000001 + (long long)(x % 1000001);
        }
        CK(cudaMalloc(&idx[b], n_upd * 8));
        CK(cudaMemcpy(idx[b], h.data(), n_upd * 8, cudaMemcpyHostToDevice));
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, int mode, int gran) {
        if (gran) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran));
        float best = 1e9, tot = 0; int n = 0;
        for (int it = 0; it < 10; ++it) {
            const long long* ix = idx[it & 3];
            cudaEventRecord(e0);
            const int blocks = 148 * 8, thr = 256;
            switch (mode) {
                case 0: k_rows<0><<<blocks, thr>>>(g, ix, n_upd, D); break;
                case 1: k_rows<1><<<blocks, thr>>>(g, ix, n_upd, D); break;
                case 2: k_rows<2><<<blocks, thr>>>(g, ix, n_upd, D); break;
                case 3: k_rows<3><<<blocks, thr>>>(g, ix, n_upd, D); break;
                case 4: k_rows<4><<<blocks, thr>>>(g, ix, n_upd, D); break;
                case 5: k_rows<5><<<blocks, thr>>>(g, ix, n_upd, D); break;
                case 6: k_bulk<<<blocks, thr, thr * D * 4>>>(g, ix, n_upd, D); break;
            }
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (it >= 2) { best = ms < best ? ms : best; tot += ms; ++n; }
        }
        size_t gr = 0; cudaDeviceGetLimit(&gr, cudaLimitMaxL2FetchGranularity);
        printf("%-44s gran=%3zu  avg %.1f us  best %.1f us  (%.2f TB/s payload)\n", name, gr, tot / n * 1e3, best * 1e3,
               n_upd * 64.0 / (tot / n * 1e-3) / 1e12);
    };
    for (int gran : {128, 64, 32}) {
        run("red.v4 (4 lanes/row)", 0, gran);
        run("ld.L2::64B + red.v4", 1, 0);
        run("st.v4 (plain store)", 2, 0);
        run("ld + add + st (non-atomic)", 3, 0);
        run("prefetch.L2 + red.v4", 4, 0);
        run("scalar red x4", 5, 0);
        run("cp.reduce.async.bulk 64 B per lane", 6, 0);
    }
    return 0;
}
