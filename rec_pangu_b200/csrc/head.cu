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

// counter-based RNG: keep(i) = u(seed, i) >= p, u uniform in [0,1) from splitmix64; recomputed in backward
__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned long long i) {
    unsigned long long z = seed + (i + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// y = x * keep / (1-p)  (in place allowed)
__global__ void __launch_bounds__(256)
dropout_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float p, float inv_keep,
                   unsigned long long seed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    y[i] = (uniform01(seed, (unsigned long long)i) >= p) ? x[i] * inv_keep : 0.f;
}

// dx = dy * keep/(1-p) * (relu_out ? relu_out > 0 : 1)
__global__ void __launch_bounds__(256)
dropout_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ relu_out, float* __restrict__ dx,
                   long long n, float p, float inv_keep, unsigned long long seed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = (uniform01(seed, (unsigned long long)i) >= p) ? dy[i] * inv_keep : 0.f;
    if (relu_out != nullptr && !(relu_out[i] > 0.f)) v = 0.f;
    dx[i] = v;
}

}  // namespace rpb

using namespace rpb;

RPB_API int rpb_dropout_fwd(const float* x, float* y, int64_t n, float p, uint64_t seed, void* stream) {
    if (x == nullptr || y == nullptr || n <= 0 || p < 0.f || p >= 1.f) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    dropout_fwd_kernel<<<ceil_div(n, 256), 256, 0, st>>>(x, y, n, p, 1.f / (1.f - p), seed);
    RPB_LAUNCH_CHECK();
    return 0;
}

RPB_API int rpb_dropout_bwd(const float* dy, const float* relu_out, float* dx, int64_t n, float p, uint64_t seed,
                            void* stream) {
    if (dy == nullptr || dx == nullptr || n <= 0 || p < 0.f || p >= 1.f) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    dropout_bwd_kernel<<<ceil_div(n, 256), 256, 0, st>>>(dy, relu_out, dx, n, p, 1.f / (1.f - p), seed);
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
