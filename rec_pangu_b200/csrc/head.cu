// sigmoid + binary cross entropy head (reference: torch.sigmoid + torch.nn.BCELoss mean reduction,
// e.g. models/ranking/deepfm.py:61-63; multi_task/mmoe.py:127-128 adds 1e-6 to pred first).
// Formulas restate ATen's binary_cross_entropy / _backward: log terms clamped at -100, backward denominator
// clamped at 1e-12.  The mean is reduced deterministically: per-CTA partials, last CTA sums them in order.
#include "common.cuh"

namespace rpb {

constexpr int kHeadMaxBlocks = 1024;

__global__ void __launch_bounds__(256)
sigmoid_bce_fwd_kernel(const float* __restrict__ logit, const float* __restrict__ label, float* __restrict__ pred,
                       float* __restrict__ loss_out, float eps, float scale, int M, unsigned int* counter,
                       float* partials) {
    __shared__ float red[32];
    __shared__ bool is_last;
    float acc = 0.f;
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
        const float z = __ldg(logit + m);
        const float p = 1.f / (1.f + expf(-z));
        pred[m] = p;
        if (label != nullptr) {
            const float y = __ldg(label + m);
            const float pe = p + eps;
            const float l1 = fmaxf(logf(pe), -100.f);
            const float l0 = fmaxf(logf(1.f - pe), -100.f);
            acc += -(y * l1 + (1.f - y) * l0);
        }
    }
    if (loss_out == nullptr) return;
    const float t = block_sum(acc, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = t;
        __threadfence();
        const unsigned int done = atomicAdd(counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        float s = 0.f;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += ((volatile float*)partials)[i];
        s = block_sum(s, red);
        if (threadIdx.x == 0) {
            loss_out[0] = scale * (s / (float)M);
            *counter = 0u;                           // self-reset for the next launch
        }
    }
}

__global__ void __launch_bounds__(256)
sigmoid_bce_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ label,
                       const float* __restrict__ gloss, float eps, float scale, float* __restrict__ dlogit, int M) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float g = (gloss != nullptr ? __ldg(gloss) : 1.f) * scale / (float)M;
    const float p = __ldg(pred + m), y = __ldg(label + m);
    const float pe = p + eps;
    // ATen: grad_input = grad * (input - target) / max((1 - input) * input, 1e-12)
    const float dp = g * (pe - y) / fmaxf((1.f - pe) * pe, 1e-12f);
    dlogit[m] = dp * p * (1.f - p);                  // sigmoid backward
}

// ESSM head (multi_task/essm.py:50-75): click = sigmoid(z1), conversion = sigmoid(z2), pctrcvr = click * conversion,
// loss = mean BCE(pctrcvr, y2) + w * mean BCE(click, y1) — the reference passes the PRODUCT as the "conversion" argument of
// its loss (essm.py:56,69-73) while reporting sigmoid(z2) as task2_pred.  Same ATen clamps as above.
__global__ void __launch_bounds__(256)
essm_head_fwd_kernel(const float* __restrict__ z1, const float* __restrict__ z2, const float* __restrict__ y1,
                     const float* __restrict__ y2, float* __restrict__ click, float* __restrict__ conv,
                     float* __restrict__ loss_out, float w, int M, unsigned int* counter, float* partials) {
    __shared__ float red[32];
    __shared__ bool is_last;
    float acc = 0.f;
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
        const float c = 1.f / (1.f + expf(-__ldg(z1 + m)));
        const float v = 1.f / (1.f + expf(-__ldg(z2 + m)));
        click[m] = c;
        conv[m] = v;
        if (loss_out != nullptr) {
            const float p = c * v, t1 = __ldg(y1 + m), t2 = __ldg(y2 + m);
            acc += -(t2 * fmaxf(logf(p), -100.f) + (1.f - t2) * fmaxf(logf(1.f - p), -100.f));
            acc += -w * (t1 * fmaxf(logf(c), -100.f) + (1.f - t1) * fmaxf(logf(1.f - c), -100.f));
        }
    }
    if (loss_out == nullptr) return;
    const float t = block_sum(acc, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = t;
        __threadfence();
        const unsigned int done = atomicAdd(counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        float s = 0.f;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += ((volatile float*)partials)[i];
        s = block_sum(s, red);
        if (threadIdx.x == 0) {
            loss_out[0] = s / (float)M;
            *counter = 0u;
        }
    }
}

__global__ void __launch_bounds__(256)
essm_head_bwd_kernel(const float* __restrict__ click, const float* __restrict__ conv, const float* __restrict__ y1,
                     const float* __restrict__ y2, const float* __restrict__ gloss, float w, float* __restrict__ dz1,
                     float* __restrict__ dz2, int M) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float g = (gloss != nullptr ? __ldg(gloss) : 1.f) / (float)M;
    const float c = __ldg(click + m), v = __ldg(conv + m), p = c * v;
    const float dp = g * (p - __ldg(y2 + m)) / fmaxf((1.f - p) * p, 1e-12f);
    const float dc = dp * v + w * g * (c - __ldg(y1 + m)) / fmaxf((1.f - c) * c, 1e-12f);
    dz1[m] = dc * c * (1.f - c);
    dz2[m] = dp * c * v * (1.f - v);
}

// counter-based RNG: keep(i) = u(seed, i) >= p, u uniform in [0,1) from splitmix64; recomputed in backward
__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned long long i) {
    unsigned long long z = seed + (i + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// y = x * keep / (1-p)  (in place allowed).  VEC = 4: one float4 per thread (16-byte aligned pointers, n % 4 == 0) — the mask of
// element i depends on (seed, epoch, i) only, so both forms draw the same masks; the scalar form moved 33 MB in 17 us.
template <int VEC>
__global__ void __launch_bounds__(256)
dropout_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float p, float inv_keep,
                   unsigned long long seed, const unsigned long long* __restrict__ epoch) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (i >= n) return;
    if (epoch != nullptr) seed += __ldg(epoch) * 0xD1B54A32D192ED03ull;      // device-side step counter (CUDA-graph replays)
    if constexpr (VEC == 4) {
        float4 v = *reinterpret_cast<const float4*>(x + i);
        v.x = (uniform01(seed, (unsigned long long)i) >= p) ? v.x * inv_keep : 0.f;
        v.y = (uniform01(seed, (unsigned long long)i + 1ull) >= p) ? v.y * inv_keep : 0.f;
        v.z = (uniform01(seed, (unsigned long long)i + 2ull) >= p) ? v.z * inv_keep : 0.f;
        v.w = (uniform01(seed, (unsigned long long)i + 3ull) >= p) ? v.w * inv_keep : 0.f;
        *reinterpret_cast<float4*>(y + i) = v;
    } else {
        y[i] = (uniform01(seed, (unsigned long long)i) >= p) ? x[i] * inv_keep : 0.f;
    }
}

// dx = dy * keep/(1-p) * (relu_out ? relu_out > 0 : 1)
template <int VEC>
__global__ void __launch_bounds__(256)
dropout_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ relu_out, float* __restrict__ dx,
                   long long n, float p, float inv_keep, unsigned long long seed, const unsigned long long* __restrict__ epoch) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (i >= n) return;
    if (epoch != nullptr) seed += __ldg(epoch) * 0xD1B54A32D192ED03ull;
    if constexpr (VEC == 4) {
        float4 v = *reinterpret_cast<const float4*>(dy + i);
        float4 r = make_float4(1.f, 1.f, 1.f, 1.f);
        if (relu_out != nullptr) r = *reinterpret_cast<const float4*>(relu_out + i);
        v.x = (uniform01(seed, (unsigned long long)i) >= p && r.x > 0.f) ? v.x * inv_keep : 0.f;
        v.y = (uniform01(seed, (unsigned long long)i + 1ull) >= p && r.y > 0.f) ? v.y * inv_keep : 0.f;
        v.z = (uniform01(seed, (unsigned long long)i + 2ull) >= p && r.z > 0.f) ? v.z * inv_keep : 0.f;
        v.w = (uniform01(seed, (unsigned long long)i + 3ull) >= p && r.w > 0.f) ? v.w * inv_keep : 0.f;
        *reinterpret_cast<float4*>(dx + i) = v;
    } else {
        float v = (uniform01(seed, (unsigned long long)i) >= p) ? dy[i] * inv_keep : 0.f;
        if (relu_out != nullptr && !(relu_out[i] > 0.f)) v = 0.f;
        dx[i] = v;
    }
}

}  // namespace rpb

using namespace rpb;

RPB_API int rpb_dropout_fwd(const float* x, float* y, int64_t n, float p, uint64_t seed, const uint64_t* epoch, void* stream) {
    if (x == nullptr || y == nullptr || n <= 0 || p < 0.f || p >= 1.f) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned long long* ep = reinterpret_cast<const unsigned long long*>(epoch);
    if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0)
        dropout_fwd_kernel<4><<<ceil_div(n / 4, 256), 256, 0, st>>>(x, y, n, p, 1.f / (1.f - p), seed, ep);
    else
        dropout_fwd_kernel<1><<<ceil_div(n, 256), 256, 0, st>>>(x, y, n, p, 1.f / (1.f - p), seed, ep);
    RPB_LAUNCH_CHECK();
    return 0;
}

RPB_API int rpb_dropout_bwd(const float* dy, const float* relu_out, float* dx, int64_t n, float p, uint64_t seed,
                            const uint64_t* epoch, void* stream) {
    if (dy == nullptr || dx == nullptr || n <= 0 || p < 0.f || p >= 1.f) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned long long* ep = reinterpret_cast<const unsigned long long*>(epoch);
    if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(relu_out)) & 15u) == 0)
        dropout_bwd_kernel<4><<<ceil_div(n / 4, 256), 256, 0, st>>>(dy, relu_out, dx, n, p, 1.f / (1.f - p), seed, ep);
    else
        dropout_bwd_kernel<1><<<ceil_div(n, 256), 256, 0, st>>>(dy, relu_out, dx, n, p, 1.f / (1.f - p), seed, ep);
    RPB_LAUNCH_CHECK();
    return 0;
}

RPB_API int rpb_sigmoid_bce_fwd(const float* logit, const float* label, float* pred, float* loss_out, float eps,
                                float scale, int M, void* work, void* stream) {
    if (logit == nullptr || pred == nullptr || M <= 0) return RPB_ERR_BAD_ARG;
    if (loss_out != nullptr && (label == nullptr || work == nullptr)) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = min(kHeadMaxBlocks, ceil_div(M, 256));
    unsigned int* counter = reinterpret_cast<unsigned int*>(work);
    float* partials = reinterpret_cast<float*>(work) + 2;
    sigmoid_bce_fwd_kernel<<<grid, 256, 0, st>>>(logit, label, pred, loss_out, eps, scale, M, counter, partials);
    RPB_LAUNCH_CHECK();
    return 0;
}

RPB_API int rpb_sigmoid_bce_bwd(const float* pred, const float* label, const float* gloss, float eps, float scale,
                                float* dlogit, int M, void* stream) {
    if (pred == nullptr || label == nullptr || dlogit == nullptr || M <= 0) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    sigmoid_bce_bwd_kernel<<<ceil_div(M, 256), 256, 0, st>>>(pred, label, gloss, eps, scale, dlogit, M);
    RPB_LAUNCH_CHECK();
    return 0;
}

RPB_API int rpb_essm_head_fwd(const float* z1, const float* z2, const float* y1, const float* y2, float* click,
                              float* conv, float* loss_out, float w_ctr, int M, void* work, void* stream) {
    if (z1 == nullptr || z2 == nullptr || click == nullptr || conv == nullptr || M <= 0) return RPB_ERR_BAD_ARG;
    if (loss_out != nullptr && (y1 == nullptr || y2 == nullptr || work == nullptr)) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = min(kHeadMaxBlocks, ceil_div(M, 256));
    essm_head_fwd_kernel<<<grid, 256, 0, st>>>(z1, z2, y1, y2, click, conv, loss_out, w_ctr, M,
                                               reinterpret_cast<unsigned int*>(work), reinterpret_cast<float*>(work) + 2);
    RPB_LAUNCH_CHECK();
    return 0;
}

RPB_API int rpb_essm_head_bwd(const float* click, const float* conv, const float* y1, const float* y2, const float* gloss,
                              float w_ctr, float* dz1, float* dz2, int M, void* stream) {
    if (click == nullptr || conv == nullptr || y1 == nullptr || y2 == nullptr || dz1 == nullptr || dz2 == nullptr || M <= 0)
        return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    essm_head_bwd_kernel<<<ceil_div(M, 256), 256, 0, st>>>(click, conv, y1, y2, gloss, w_ctr, dz1, dz2, M);
    RPB_LAUNCH_CHECK();
    return 0;
}
