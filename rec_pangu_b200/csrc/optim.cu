// Optimizer step for the hot path's parameters (reference: torch.optim.Adam(model.parameters(), lr, betas=(0.9,0.999),
// eps=1e-8) created in rec_pangu/trainer.py:75 and stepped in model_pipeline.py:57).  SURVEY.md §8f rank 1: the
// reference's dense Adam walks all 416 M parameters and two moment buffers every step (>= 11.6 GB of HBM traffic at
// config 2).  Here the embedding tables are updated ROW-SPARSELY: only rows the batch touched are read and written
// ("lazy" Adam, the semantics of torch.optim.SparseAdam: rows without gradient keep their moments and value), fused
// with the re-zeroing of the persistent gradient buffer.  Dense parameters use the plain element-wise kernel.
#include <type_traits>

#include "common.cuh"

namespace rpb {

struct AdamHyper { float lr, b1, b2, eps, bc1, bc2_sqrt; };   // bc1 = 1 - b1^t, bc2_sqrt = sqrt(1 - b2^t)

__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, const AdamHyper& h) {
    m = h.b1 * m + (1.f - h.b1) * g;
    v = h.b2 * v + (1.f - h.b2) * g * g;
    const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
    p -= (h.lr / h.bc1) * (m / denom);
}

__global__ void __launch_bounds__(256)
adam_dense_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  long long n, AdamHyper h) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float pp = p[i], mm = m[i], vv = v[i];
    adam_update(pp, mm, vv, g[i], h);
    p[i] = pp; m[i] = mm; v[i] = vv;
}

struct SparseAdamParams {
    float* w[RPB_MAX_FIELDS];
    float* g[RPB_MAX_FIELDS];
    float* m[RPB_MAX_FIELDS];
    float* v[RPB_MAX_FIELDS];
    int* stamp[RPB_MAX_FIELDS];
    const long long* idx[RPB_MAX_FIELDS];
    long long rows[RPB_MAX_FIELDS];
    int B, F, D, step;
    AdamHyper h;
};

// LPR lanes own one (sample, field) occurrence at a time; lane 0 claims the row for this step with an atomic
// exchange on its stamp, so a row hit by several samples is updated exactly once (its gradient is already summed).
template <int LPR>
__global__ void __launch_bounds__(256)
sparse_adam_kernel(const __grid_constant__ SparseAdamParams p) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b_raw = (int)(gt / LPR), l = (int)(gt % LPR);
    const bool valid = b_raw < p.B;
    const int b = valid ? b_raw : p.B - 1;
    const int DV = p.D / 4;
    for (int f = 0; f < p.F; ++f) {
        if (p.w[f] == nullptr) continue;                        // frozen table (uniform branch)
        long long row = __ldg(p.idx[f] + b);
        if ((unsigned long long)row >= (unsigned long long)p.rows[f]) row = 0;
        int claimed = 0;
        if (l == 0 && valid) claimed = (atomicExch(p.stamp[f] + row, p.step) != p.step) ? 1 : 0;
        claimed = __shfl_sync(0xffffffffu, claimed, (threadIdx.x & 31) / LPR * LPR);
        if (claimed && l < DV) {
            const size_t off = (size_t)row * p.D + l * 4;
            float4 g = *reinterpret_cast<const float4*>(p.g[f] + off);
            float4 w = *reinterpret_cast<const float4*>(p.w[f] + off);
            float4 m = *reinterpret_cast<const float4*>(p.m[f] + off);
            float4 v = *reinterpret_cast<const float4*>(p.v[f] + off);
            adam_update(w.x, m.x, v.x, g.x, p.h); adam_update(w.y, m.y, v.y, g.y, p.h);
            adam_update(w.z, m.z, v.z, g.z, p.h); adam_update(w.w, m.w, v.w, g.w, p.h);
            *reinterpret_cast<float4*>(p.w[f] + off) = w;
            *reinterpret_cast<float4*>(p.m[f] + off) = m;
            *reinterpret_cast<float4*>(p.v[f] + off) = v;
            *reinterpret_cast<float4*>(p.g[f] + off) = make_float4(0.f, 0.f, 0.f, 0.f);   // fused sparse zero_grad
        }
    }
}

}  // namespace rpb

using namespace rpb;

static AdamHyper make_hyper(float lr, float b1, float b2, float eps, int step) {
    AdamHyper h;
    h.lr = lr; h.b1 = b1; h.b2 = b2; h.eps = eps;
    h.bc1 = 1.f - powf(b1, (float)step);
    h.bc2_sqrt = sqrtf(1.f - powf(b2, (float)step));
    return h;
}

RPB_API int rpb_adam_dense(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                           float eps, int step, void* stream) {
    if (p == nullptr || g == nullptr || m == nullptr || v == nullptr || n <= 0 || step < 1) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    adam_dense_kernel<<<ceil_div(n, 256), 256, 0, st>>>(p, g, m, v, n, make_hyper(lr, beta1, beta2, eps, step));
    RPB_LAUNCH_CHECK();
    return 0;
}

RPB_API int rpb_sparse_adam(const RpbSparseAdamDesc* d, void* stream) {
    if (d == nullptr || d->B <= 0 || d->F <= 0 || d->D <= 0 || d->step < 1) return RPB_ERR_BAD_ARG;
    if (d->F > RPB_MAX_FIELDS || d->D % 4 != 0 || d->D > 128) return RPB_ERR_UNSUPPORTED;
    SparseAdamParams p{};
    for (int f = 0; f < d->F; ++f) {
        p.w[f] = d->weights[f]; p.g[f] = d->grads[f]; p.m[f] = d->exp_avg[f]; p.v[f] = d->exp_avg_sq[f];
        p.stamp[f] = d->stamps[f];
        p.idx[f] = reinterpret_cast<const long long*>(d->idx[f]);
        p.rows[f] = d->rows[f];
    }
    p.B = d->B; p.F = d->F; p.D = d->D; p.step = d->step;
    p.h = make_hyper(d->lr, d->beta1, d->beta2, d->eps, d->step);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int dv = d->D / 4;
    auto launch = [&](auto lt) -> int {
        constexpr int LPR = decltype(lt)::value;
        sparse_adam_kernel<LPR><<<ceil_div((long long)p.B * LPR, 256), 256, 0, st>>>(p);
        return (int)cudaGetLastError();
    };
    if (dv <= 1) return launch(std::integral_constant<int, 1>{});
    if (dv <= 2) return launch(std::integral_constant<int, 2>{});
    if (dv <= 4) return launch(std::integral_constant<int, 4>{});
    if (dv <= 8) return launch(std::integral_constant<int, 8>{});
    if (dv <= 16) return launch(std::integral_constant<int, 16>{});
    return launch(std::integral_constant<int, 32>{});
}
