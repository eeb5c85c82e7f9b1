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

__host__ __device__ inline AdamHyper make_hyper(float lr, float b1, float b2, float eps, int step) {
    AdamHyper h;
    h.lr = lr; h.b1 = b1; h.b2 = b2; h.eps = eps;
    h.bc1 = 1.f - powf(b1, (float)step);
    h.bc2_sqrt = sqrtf(1.f - powf(b2, (float)step));
    return h;
}

__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, const AdamHyper& h) {
    m = h.b1 * m + (1.f - h.b1) * g;
    v = h.b2 * v + (1.f - h.b2) * g * g;
    const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
    p -= (h.lr / h.bc1) * (m / denom);
}

// Up to RPB_ADAM_MAX_TENSORS dense parameters in one launch: block b works on tensor t with blk0[t] <= b < blk0[t+1].
// step_dev (device int32, may be null): the step number is read on the device, so a captured CUDA graph of the
// training step keeps counting (bias correction) across replays.
struct AdamMultiParams {
    float* p[RPB_ADAM_MAX_TENSORS]; const float* g[RPB_ADAM_MAX_TENSORS];
    float* m[RPB_ADAM_MAX_TENSORS]; float* v[RPB_ADAM_MAX_TENSORS];
    long long n[RPB_ADAM_MAX_TENSORS]; int blk0[RPB_ADAM_MAX_TENSORS + 1];
    int count; AdamHyper h; const int* step_dev;
};

__global__ void __launch_bounds__(256)
adam_multi_kernel(const __grid_constant__ AdamMultiParams q) {
    int t = 0;
    while (t + 1 < q.count && (int)blockIdx.x >= q.blk0[t + 1]) ++t;
    const long long i = (long long)(blockIdx.x - q.blk0[t]) * blockDim.x + threadIdx.x;
    if (i >= q.n[t]) return;
    const AdamHyper h = q.step_dev != nullptr ? make_hyper(q.h.lr, q.h.b1, q.h.b2, q.h.eps, *q.step_dev) : q.h;
    float pp = q.p[t][i], mm = q.m[t][i], vv = q.v[t][i];
    adam_update(pp, mm, vv, q.g[t][i], h);
    q.p[t][i] = pp; q.m[t][i] = mm; q.v[t][i] = vv;
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
    const int* step_dev;
};

// LPR lanes own one (sample, field) occurrence at a time; lane 0 claims the row for this step with an atomic
// exchange on its stamp, so a row hit by several samples is updated exactly once (its gradient is already summed).
template <int LPR>
__global__ void __launch_bounds__(256)
sparse_adam_kernel(const __grid_constant__ SparseAdamParams p) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int step = p.step_dev != nullptr ? *p.step_dev : p.step;
    const AdamHyper hy = p.step_dev != nullptr ? make_hyper(p.h.lr, p.h.b1, p.h.b2, p.h.eps, step) : p.h;
    const int b_raw = (int)(gt / LPR), l = (int)(gt % LPR);
    const bool valid = b_raw < p.B;
    const int b = valid ? b_raw : p.B - 1;
    const int DV = p.D / 4;
    for (int f = 0; f < p.F; ++f) {
        if (p.w[f] == nullptr) continue;                        // frozen table (uniform branch)
        long long row = __ldg(p.idx[f] + b);
        if ((unsigned long long)row >= (unsigned long long)p.rows[f]) row = 0;
        int claimed = 0;
        if (l == 0 && valid) claimed = (atomicExch(p.stamp[f] + row, step) != step) ? 1 : 0;
        claimed = __shfl_sync(0xffffffffu, claimed, (threadIdx.x & 31) / LPR * LPR);
        if (claimed && l < DV) {
            const size_t off = (size_t)row * p.D + l * 4;
            float4 g = *reinterpret_cast<const float4*>(p.g[f] + off);
            float4 w = *reinterpret_cast<const float4*>(p.w[f] + off);
            float4 m = *reinterpret_cast<const float4*>(p.m[f] + off);
            float4 v = *reinterpret_cast<const float4*>(p.v[f] + off);
            adam_update(w.x, m.x, v.x, g.x, hy); adam_update(w.y, m.y, v.y, g.y, hy);
            adam_update(w.z, m.z, v.z, g.z, hy); adam_update(w.w, m.w, v.w, g.w, hy);
            *reinterpret_cast<float4*>(p.w[f] + off) = w;
            *reinterpret_cast<float4*>(p.m[f] + off) = m;
            *reinterpret_cast<float4*>(p.v[f] + off) = v;
            *reinterpret_cast<float4*>(p.g[f] + off) = make_float4(0.f, 0.f, 0.f, 0.f);   // fused sparse zero_grad
        }
    }
}

// D = 1 tables (the LR_Layer's wide part, models/layers/shallow.py:14-27): one thread per sample walks the fields; the row is
// claimed with the same stamp exchange, so the step stays graph-safe (no torch.unique, step number read on the device).
__global__ void __launch_bounds__(256)
sparse_adam_scalar_kernel(const __grid_constant__ SparseAdamParams p) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const int step = p.step_dev != nullptr ? *p.step_dev : p.step;
    const AdamHyper hy = p.step_dev != nullptr ? make_hyper(p.h.lr, p.h.b1, p.h.b2, p.h.eps, step) : p.h;
    for (int f = 0; f < p.F; ++f) {
        if (p.w[f] == nullptr) continue;
        long long row = __ldg(p.idx[f] + b);
        if ((unsigned long long)row >= (unsigned long long)p.rows[f]) row = 0;
        if (atomicExch(p.stamp[f] + row, step) == step) continue;          // another sample of this batch already updated the row
        float w = p.w[f][row], m = p.m[f][row], v = p.v[f][row];
        adam_update(w, m, v, p.g[f][row], hy);
        p.w[f][row] = w; p.m[f][row] = m; p.v[f][row] = v;
        p.g[f][row] = 0.f;                                                  // fused sparse zero_grad
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Exact-lazy Adam (FusedAdam(exact=True)): results equal the reference's DENSE torch.optim.Adam (trainer.py:75), which moves
// every row at every step — a row without gradient still decays its moments and walks by its momentum — while only rows a
// batch touches are ever read or written.  stamp[row] = the last optimizer step whose update the row has received.  Before
// a forward reads a row (catch-up kernel, launched from a forward pre-hook with the batch ids) and before a state_dict is
// taken (flush kernel over all rows), the missed zero-gradient steps stamp+1 .. cur are replayed:
//     m <- b1 m,  v <- b2 v,  w <- w - lr / (1 - b1^s) * m / (sqrt(v) / sqrt(1 - b2^s) + eps)        for s = stamp+1 .. cur
// The walk terms shrink like 0.9^k, so after CATCHUP_ITERS steps the remaining ones are below fp32 resolution of the sum;
// the moments of the rest of the gap decay in closed form (b^gap).
constexpr int CATCHUP_ITERS = 320;

__device__ __forceinline__ void catchup4(float4& w, float4& m, float4& v, int old, int cur, const AdamHyper& hy) {
    const int gap = cur - old;
    if (gap <= 0) return;
    const bool zero = m.x == 0.f && m.y == 0.f && m.z == 0.f && m.w == 0.f;       // never-touched row: nothing moves
    if (!zero) {
        float pb1 = powf(hy.b1, (float)old), pb2 = powf(hy.b2, (float)old);
        const int n = min(gap, CATCHUP_ITERS);
        for (int k = 0; k < n; ++k) {
            pb1 *= hy.b1; pb2 *= hy.b2;
            const float a = hy.lr / (1.f - pb1), c = rsqrtf(1.f - pb2);
            m.x *= hy.b1; m.y *= hy.b1; m.z *= hy.b1; m.w *= hy.b1;
            v.x *= hy.b2; v.y *= hy.b2; v.z *= hy.b2; v.w *= hy.b2;
            w.x -= a * (m.x / (sqrtf(v.x) * c + hy.eps)); w.y -= a * (m.y / (sqrtf(v.y) * c + hy.eps));
            w.z -= a * (m.z / (sqrtf(v.z) * c + hy.eps)); w.w -= a * (m.w / (sqrtf(v.w) * c + hy.eps));
        }
        if (gap > n) {
            const float d1 = powf(hy.b1, (float)(gap - n)), d2 = powf(hy.b2, (float)(gap - n));
            m.x *= d1; m.y *= d1; m.z *= d1; m.w *= d1;
            v.x *= d2; v.y *= d2; v.z *= d2; v.w *= d2;
        }
    }
}

// rows of the batch about to be read; lane 0 of a row group claims the row (stamp exchange), so a row hit by several samples is
// caught up once.  `cur` = *step_dev = optimizer steps taken so far.
template <int LPR>
__global__ void __launch_bounds__(256)
sparse_adam_catchup_kernel(const __grid_constant__ SparseAdamParams p) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int cur = p.step_dev != nullptr ? *p.step_dev : p.step;
    const int b_raw = (int)(gt / LPR), l = (int)(gt % LPR);
    const bool valid = b_raw < p.B;
    const int b = valid ? b_raw : p.B - 1;
    const int DV = p.D >= 4 ? p.D / 4 : 1;
    for (int f = 0; f < p.F; ++f) {
        if (p.w[f] == nullptr) continue;
        long long row = __ldg(p.idx[f] + b);
        if ((unsigned long long)row >= (unsigned long long)p.rows[f]) row = 0;
        int old = cur;
        if (l == 0 && valid) old = atomicExch(p.stamp[f] + row, cur);
        old = __shfl_sync(0xffffffffu, old, (threadIdx.x & 31) / LPR * LPR);
        if (old < cur && l < DV) {
            if (p.D >= 4) {
                const size_t off = (size_t)row * p.D + l * 4;
                float4 w = *reinterpret_cast<const float4*>(p.w[f] + off);
                float4 m = *reinterpret_cast<const float4*>(p.m[f] + off);
                float4 v = *reinterpret_cast<const float4*>(p.v[f] + off);
                catchup4(w, m, v, old, cur, p.h);
                *reinterpret_cast<float4*>(p.w[f] + off) = w;
                *reinterpret_cast<float4*>(p.m[f] + off) = m;
                *reinterpret_cast<float4*>(p.v[f] + off) = v;
            } else {                                               // D = 1 (LR tables): lane 0 only (DV = 1)
                float4 w = make_float4(p.w[f][row], 0.f, 0.f, 0.f), m = make_float4(p.m[f][row], 0.f, 0.f, 0.f);
                float4 v = make_float4(p.v[f][row], 0.f, 0.f, 0.f);
                catchup4(w, m, v, old, cur, p.h);
                p.w[f][row] = w.x; p.m[f][row] = m.x; p.v[f][row] = v.x;
            }
        }
    }
}

// every row of one table (state_dict / evaluation of the whole table): thread = (row, 4-float piece)
__global__ void __launch_bounds__(256)
sparse_adam_flush_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v, int* __restrict__ stamp, long long rows,
                         int D, AdamHyper hy, const int* __restrict__ step_dev, int step) {
    const int DV = D >= 4 ? D / 4 : 1;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long row = t / DV;
    const int l = (int)(t % DV);
    if (row >= rows) return;
    const int cur = step_dev != nullptr ? *step_dev : step;
    const int old = stamp[row];
    if (old < cur) {
        if (D >= 4) {
            const size_t off = (size_t)row * D + l * 4;
            float4 ww = *reinterpret_cast<const float4*>(w + off), mm = *reinterpret_cast<const float4*>(m + off);
            float4 vv = *reinterpret_cast<const float4*>(v + off);
            catchup4(ww, mm, vv, old, cur, hy);
            *reinterpret_cast<float4*>(w + off) = ww; *reinterpret_cast<float4*>(m + off) = mm; *reinterpret_cast<float4*>(v + off) = vv;
        } else {
            float4 ww = make_float4(w[row], 0.f, 0.f, 0.f), mm = make_float4(m[row], 0.f, 0.f, 0.f), vv = make_float4(v[row], 0.f, 0.f, 0.f);
            catchup4(ww, mm, vv, old, cur, hy);
            w[row] = ww.x; m[row] = mm.x; v[row] = vv.x;
        }
    }
}
// second pass: the stamps (a separate launch: every piece of a row must have read the old stamp first)
__global__ void __launch_bounds__(256)
sparse_adam_flush_stamp_kernel(int* __restrict__ stamp, long long rows, const int* __restrict__ step_dev, int step) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const int cur = step_dev != nullptr ? *step_dev : step;
    if (stamp[row] < cur) stamp[row] = cur;
}

}  // namespace rpb

using namespace rpb;

RPB_API int rpb_adam_dense(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                           float eps, int step, void* stream) {
    if (p == nullptr || g == nullptr || m == nullptr || v == nullptr || n <= 0 || step < 1) return RPB_ERR_BAD_ARG;
    RpbAdamMultiDesc d{};
    float* pp[1] = {p}; const float* gg[1] = {g}; float* mm[1] = {m}; float* vv[1] = {v}; int64_t nn[1] = {n};
    d.count = 1; d.params = pp; d.grads = gg; d.exp_avg = mm; d.exp_avg_sq = vv; d.numel = nn;
    d.lr = lr; d.beta1 = beta1; d.beta2 = beta2; d.eps = eps; d.step = step; d.step_dev = nullptr;
    return rpb_adam_multi(&d, stream);
}

RPB_API int rpb_adam_multi(const RpbAdamMultiDesc* d, void* stream) {
    if (d == nullptr || d->count <= 0 || (d->step_dev == nullptr && d->step < 1)) return RPB_ERR_BAD_ARG;
    if (d->count > RPB_ADAM_MAX_TENSORS) return RPB_ERR_UNSUPPORTED;
    AdamMultiParams q{};
    int blocks = 0;
    for (int t = 0; t < d->count; ++t) {
        if (d->params[t] == nullptr || d->grads[t] == nullptr || d->exp_avg[t] == nullptr || d->exp_avg_sq[t] == nullptr || d->numel[t] <= 0)
            return RPB_ERR_BAD_ARG;
        q.p[t] = d->params[t]; q.g[t] = d->grads[t]; q.m[t] = d->exp_avg[t]; q.v[t] = d->exp_avg_sq[t]; q.n[t] = d->numel[t];
        q.blk0[t] = blocks;
        blocks += ceil_div(d->numel[t], 256);
    }
    q.blk0[d->count] = blocks;
    q.count = d->count;
    q.h = make_hyper(d->lr, d->beta1, d->beta2, d->eps, d->step_dev != nullptr ? 1 : d->step);
    q.step_dev = d->step_dev;
    adam_multi_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(q);
    RPB_LAUNCH_CHECK();
    return 0;
}

RPB_API int rpb_sparse_adam(const RpbSparseAdamDesc* d, void* stream) {
    if (d == nullptr || d->B <= 0 || d->F <= 0 || d->D <= 0 || (d->step_dev == nullptr && d->step < 1)) return RPB_ERR_BAD_ARG;
    if (d->F > RPB_MAX_FIELDS || (d->D != 1 && d->D % 4 != 0) || d->D > 128) return RPB_ERR_UNSUPPORTED;
    SparseAdamParams p{};
    for (int f = 0; f < d->F; ++f) {
        p.w[f] = d->weights[f]; p.g[f] = d->grads[f]; p.m[f] = d->exp_avg[f]; p.v[f] = d->exp_avg_sq[f];
        p.stamp[f] = d->stamps[f];
        p.idx[f] = reinterpret_cast<const long long*>(d->idx[f]);
        p.rows[f] = d->rows[f];
    }
    p.B = d->B; p.F = d->F; p.D = d->D; p.step = d->step;
    p.step_dev = d->step_dev;
    p.h = make_hyper(d->lr, d->beta1, d->beta2, d->eps, d->step_dev != nullptr ? 1 : d->step);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (d->D == 1) {
        sparse_adam_scalar_kernel<<<ceil_div(p.B, 256), 256, 0, st>>>(p);
        return (int)cudaGetLastError();
    }
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

static int fill_sparse_params(const RpbSparseAdamDesc* d, SparseAdamParams& p) {
    if (d == nullptr || d->B <= 0 || d->F <= 0 || d->D <= 0 || (d->step_dev == nullptr && d->step < 0)) return RPB_ERR_BAD_ARG;
    if (d->F > RPB_MAX_FIELDS || (d->D != 1 && d->D % 4 != 0) || d->D > 128) return RPB_ERR_UNSUPPORTED;
    for (int f = 0; f < d->F; ++f) {
        p.w[f] = d->weights[f]; p.g[f] = nullptr; p.m[f] = d->exp_avg[f]; p.v[f] = d->exp_avg_sq[f];
        p.stamp[f] = d->stamps[f];
        p.idx[f] = reinterpret_cast<const long long*>(d->idx[f]);
        p.rows[f] = d->rows[f];
    }
    p.B = d->B; p.F = d->F; p.D = d->D; p.step = d->step;
    p.step_dev = d->step_dev;
    p.h = make_hyper(d->lr, d->beta1, d->beta2, d->eps, 1);
    return 0;
}

RPB_API int rpb_sparse_adam_catchup(const RpbSparseAdamDesc* d, void* stream) {
    SparseAdamParams p{};
    const int rc = fill_sparse_params(d, p);
    if (rc != 0) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int dv = d->D >= 4 ? d->D / 4 : 1;
    auto launch = [&](auto lt) -> int {
        constexpr int LPR = decltype(lt)::value;
        sparse_adam_catchup_kernel<LPR><<<ceil_div((long long)p.B * LPR, 256), 256, 0, st>>>(p);
        return (int)cudaGetLastError();
    };
    if (dv <= 1) return launch(std::integral_constant<int, 1>{});
    if (dv <= 2) return launch(std::integral_constant<int, 2>{});
    if (dv <= 4) return launch(std::integral_constant<int, 4>{});
    if (dv <= 8) return launch(std::integral_constant<int, 8>{});
    if (dv <= 16) return launch(std::integral_constant<int, 16>{});
    return launch(std::integral_constant<int, 32>{});
}

RPB_API int rpb_sparse_adam_flush(float* w, float* m, float* v, int32_t* stamp, int64_t rows, int32_t D, float lr, float beta1,
                                  float beta2, float eps, const int32_t* step_dev, int32_t step, void* stream) {
    if (w == nullptr || m == nullptr || v == nullptr || stamp == nullptr || rows <= 0 || D <= 0) return RPB_ERR_BAD_ARG;
    if ((D != 1 && D % 4 != 0) || D > 128) return RPB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int dv = D >= 4 ? D / 4 : 1;
    const AdamHyper hy = make_hyper(lr, beta1, beta2, eps, 1);
    sparse_adam_flush_kernel<<<ceil_div(rows * dv, 256), 256, 0, st>>>(w, m, v, stamp, rows, D, hy, step_dev, step);
    RPB_LAUNCH_CHECK();
    sparse_adam_flush_stamp_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(stamp, rows, step_dev, step);
    RPB_LAUNCH_CHECK();
    return 0;
}
