// LayerNorm over the last dimension of a [M, N] fp32 matrix (torch.nn.LayerNorm, eps inside the sqrt, biased variance) for the
// MaskBlock of MaskNet (reference: rec_pangu/models/layers/interaction.py:254-283 — LayerNorm on the block input and output).
// One warp per row: the row lives in registers across the two passes (mean, then centred sum of squares — the numerically
// safe order torch uses), float4 lanes when the row stride allows.  Backward: dx per row from the saved (mean, rstd) with
// two warp reductions; dgamma / dbeta are column sums over the batch, accumulated per CTA in shared memory and added to
// the outputs with one atomic per column and CTA.
#include "common.cuh"

namespace rpb {

constexpr int LN_WARPS = 8;
constexpr int LN_MAX_PER_LANE = 32;          // N <= 1024

__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float eps, float* __restrict__ y, long long ldy, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                     int M, int N) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * LN_WARPS + warp;
    if (m >= M) return;
    const float* xr = x + (size_t)m * ldx;
    float v[LN_MAX_PER_LANE];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
        const int c = lane + 32 * i;
        v[i] = c < N ? xr[c] : 0.f;
        s += v[i];
    }
    const float mean = warp_sum(s) / (float)N;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
        const int c = lane + 32 * i;
        const float d = c < N ? v[i] - mean : 0.f;
        q = fmaf(d, d, q);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)N + eps);
    float* yr = y + (size_t)m * ldy;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
        const int c = lane + 32 * i;
        if (c < N) yr[c] = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    }
    if (lane == 0) { mean_out[m] = mean; rstd_out[m] = rstd; }
}

// dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)),  g = dy * gamma,  xhat = (x - mean) * rstd
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ x, long long ldx,
                     const float* __restrict__ gamma, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                     float* __restrict__ dx, long long lddx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     int M, int N, int rows_per_cta) {
    extern __shared__ float sm[];                       // [2][N] column partials of this CTA
    float* sg = sm;
    float* sb = sm + N;
    for (int c = threadIdx.x; c < 2 * N; c += blockDim.x) sm[c] = 0.f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_end = min(M, (int)(blockIdx.x + 1) * rows_per_cta);
    float accg[LN_MAX_PER_LANE], accb[LN_MAX_PER_LANE];
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) { accg[i] = 0.f; accb[i] = 0.f; }
    for (int m = blockIdx.x * rows_per_cta + warp; m < m_end; m += LN_WARPS) {
        const float mean = __ldg(mean_in + m), rstd = __ldg(rstd_in + m);
        const float* xr = x + (size_t)m * ldx;
        const float* gr = dy + (size_t)m * lddy;
        float xh[LN_MAX_PER_LANE], g[LN_MAX_PER_LANE];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
            const int c = lane + 32 * i;
            if (c < N) {
                const float d = gr[c];
                xh[i] = (xr[c] - mean) * rstd;
                g[i] = d * __ldg(gamma + c);
                accg[i] = fmaf(d, xh[i], accg[i]);
                accb[i] += d;
            } else { xh[i] = 0.f; g[i] = 0.f; }
            s1 += g[i];
            s2 = fmaf(g[i], xh[i], s2);
        }
        s1 = warp_sum(s1) / (float)N;
        s2 = warp_sum(s2) / (float)N;
        float* dr = dx + (size_t)m * lddx;
#pragma unroll
        for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
            const int c = lane + 32 * i;
            if (c < N) dr[c] = rstd * (g[i] - s1 - xh[i] * s2);
        }
    }
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
        const int c = lane + 32 * i;
        if (c < N) { atomicAdd(sg + c, accg[i]); atomicAdd(sb + c, accb[i]); }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < N; c += blockDim.x) {
        red_add_f1(dgamma + c, sg[c]);
        red_add_f1(dbeta + c, sb[c]);
    }
}

}  // namespace rpb

using namespace rpb;

RPB_API int rpb_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, float* y, int64_t ldy,
                              float* mean, float* rstd, int32_t M, int32_t N, void* stream) {
    if (x == nullptr || gamma == nullptr || beta == nullptr || y == nullptr || mean == nullptr || rstd == nullptr || M < 1 || N < 1)
        return RPB_ERR_BAD_ARG;
    if (N > 32 * LN_MAX_PER_LANE || ldx < N || ldy < N) return RPB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    layernorm_fwd_kernel<<<ceil_div(M, LN_WARPS), LN_WARPS * 32, 0, st>>>(x, ldx, gamma, beta, eps, y, ldy, mean, rstd, M, N);
    RPB_LAUNCH_CHECK();
    return 0;
}

RPB_API int rpb_layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* gamma, const float* mean,
                              const float* rstd, float* dx, int64_t lddx, float* dgamma, float* dbeta, int32_t M, int32_t N,
                              void* stream) {
    if (dy == nullptr || x == nullptr || gamma == nullptr || mean == nullptr || rstd == nullptr || dx == nullptr || dgamma == nullptr ||
        dbeta == nullptr || M < 1 || N < 1)
        return RPB_ERR_BAD_ARG;
    if (N > 32 * LN_MAX_PER_LANE || ldx < N || lddy < N || lddx < N) return RPB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int ctas = min(ceil_div(M, LN_WARPS), 148 * 4);
    const int rows_per_cta = ceil_div(M, ctas);
    layernorm_bwd_kernel<<<ceil_div(M, rows_per_cta), LN_WARPS * 32, 2 * N * sizeof(float), st>>>(
        dy, lddy, x, ldx, gamma, mean, rstd, dx, lddx, dgamma, dbeta, M, N, rows_per_cta);
    RPB_LAUNCH_CHECK();
    return 0;
}
