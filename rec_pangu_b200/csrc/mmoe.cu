// MMOE body pieces (reference: models/multi_task/mmoe.py:86-112) and the [K,N]-stored matmul it needs.
//
//  * experts_out = einsum('ij,jkl->ikl', hidden, experts) + experts_bias and the per-task gate logits
//    hidden @ gates[t] + gates_bias[t] are ONE GEMM over the concatenated [hid, Hh*E + T*E] weight
//    (rpb_matmul_kn_fwd: weights stored [K, N] like the reference's Parameters, tcgen05 3xTF32 when it qualifies);
//  * mmoe_combine: gate softmax + gated sum over experts, warp per sample;
//  * BatchNorm1d of the task towers (Linear -> BN -> Dropout, mmoe.py:49-58): column statistics + normalise, and the
//    three-term backward, as coalesced column-parallel kernels.
#include "common.cuh"

namespace rpb {

// from linear_tc.cu / linear_simt.cu
int gemm_tc(const float* A, long long lda, const float* Bsrc, long long ldb, int b_transpose, const TcEpilogue& ep,
            int M, int N, int K, cudaStream_t st);
bool tc_shape_ok(const float* A, long long lda, int M, int N, int K);
int wgrad_tc(const float* dy, long long lddy, const float* x, long long ldx, float* dW, int M, int N, int K, cudaStream_t st);
int linear_dw_simt(const float* dy, long long lddy, const float* x, long long ldx, float* dW, float* db,
                   int M, int N, int K, cudaStream_t st);
int sgemm_kn_simt(const float* x, long long ldx, const float* Wkn, long long ldw, const float* bias, float* y,
                  long long ldy, int M, int N, int K, cudaStream_t st);
int sgemm_nk_simt(const float* x, long long ldx, const float* Wnk, long long ldw, float* y, long long ldy, int M, int N,
                  int K, cudaStream_t st);

// out[k] += sum_m x[m,k]  (and optionally sum of squares)   grid (ceil(K/32), slabs), block (32, 8)
__global__ void __launch_bounds__(256)
colstats_kernel(const float* __restrict__ x, long long ldx, float* __restrict__ sum, float* __restrict__ sumsq,
                int M, int K, int slab) {
    __shared__ float s1[8][33], s2[8][33];
    const int k = blockIdx.x * 32 + threadIdx.x;
    const int mbeg = blockIdx.y * slab, mend = min(M, mbeg + slab);
    float a = 0.f, b = 0.f;
    if (k < K) {
        for (int m = mbeg + threadIdx.y; m < mend; m += 8) {
            const float v = __ldg(x + (size_t)m * ldx + k);
            a += v;
            b = fmaf(v, v, b);
        }
    }
    s1[threadIdx.y][threadIdx.x] = a;
    s2[threadIdx.y][threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.y == 0 && k < K) {
#pragma unroll
        for (int j = 1; j < 8; ++j) { a += s1[j][threadIdx.x]; b += s2[j][threadIdx.x]; }
        red_add_f1(sum + k, a);
        if (sumsq != nullptr) red_add_f1(sumsq + k, b);
    }
}

// y = (x - mean) * invstd * gamma + beta
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y, long long n, int N) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % N);
    y[i] = fmaf((x[i] - __ldg(mean + c)) * __ldg(invstd + c), __ldg(gamma + c), __ldg(beta + c));
}

// column sums of dy and dy*xhat  (xhat = (x-mean)*invstd)   same grid as colstats
__global__ void __launch_bounds__(256)
bn_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                    const float* __restrict__ invstd, float* __restrict__ dbeta, float* __restrict__ dgamma,
                    int M, int N, int slab) {
    __shared__ float s1[8][33], s2[8][33];
    const int k = blockIdx.x * 32 + threadIdx.x;
    const int mbeg = blockIdx.y * slab, mend = min(M, mbeg + slab);
    float a = 0.f, b = 0.f;
    if (k < N) {
        const float mu = __ldg(mean + k), is = __ldg(invstd + k);
        for (int m = mbeg + threadIdx.y; m < mend; m += 8) {
            const float g = __ldg(dy + (size_t)m * N + k);
            a += g;
            b = fmaf(g, (__ldg(x + (size_t)m * N + k) - mu) * is, b);
        }
    }
    s1[threadIdx.y][threadIdx.x] = a;
    s2[threadIdx.y][threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.y == 0 && k < N) {
#pragma unroll
        for (int j = 1; j < 8; ++j) { a += s1[j][threadIdx.x]; b += s2[j][threadIdx.x]; }
        red_add_f1(dbeta + k, a);
        red_add_f1(dgamma + k, b);
    }
}

// training: dx = gamma*invstd*(dy - dbeta/M - xhat*dgamma/M);   eval (use_batch_stats == 0): dx = gamma*invstd*dy
__global__ void __launch_bounds__(256)
bn_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                 const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ dbeta,
                 const float* __restrict__ dgamma, float* __restrict__ dx, long long n, int N, float inv_m,
                 int use_batch_stats) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % N);
    const float is = __ldg(invstd + c), g = __ldg(gamma + c);
    float v = dy[i];
    if (use_batch_stats) {
        const float xhat = (x[i] - __ldg(mean + c)) * is;
        v = v - __ldg(dbeta + c) * inv_m - xhat * __ldg(dgamma + c) * inv_m;
    }
    dx[i] = g * is * v;
}

// eo: [B, ld]: columns k*E + l (k < Hh, l < E) expert outputs, then T*E gate logits.  out: [T, B, Hh]; gate: [B, T*E]
__global__ void __launch_bounds__(256)
mmoe_combine_fwd_kernel(const float* __restrict__ eo, long long ld, int B, int Hh, int E, int T,
                        float* __restrict__ out, float* __restrict__ gate) {
    const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const float* row = eo + (size_t)b * ld;
    for (int t = 0; t < T; ++t) {
        // softmax over E (E <= 32): lane l holds logit l
        const float z = lane < E ? __ldg(row + Hh * E + t * E + lane) : -INFINITY;
        float mx = z;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float ex = lane < E ? expf(z - mx) : 0.f;
        const float sum = warp_sum(ex);
        const float g = ex / sum;
        if (lane < E) gate[(size_t)b * T * E + t * E + lane] = g;
        for (int k0 = 0; k0 < Hh; k0 += 32) {          // warp-uniform trip count: the shuffles below need all lanes
            const int k = k0 + lane;
            float acc = 0.f;
            for (int l = 0; l < E; ++l) {
                const float gl = __shfl_sync(0xffffffffu, g, l);
                if (k < Hh) acc = fmaf(__ldg(row + k * E + l), gl, acc);
            }
            if (k < Hh) out[((size_t)t * B + b) * Hh + k] = acc;
        }
    }
}

// deo[b, k*E+l] = sum_t dout_t[b,k]*gate_t[l];  dlogit_t[l] = gate_t[l]*(dg_t[l] - sum_l' gate_t[l'] dg_t[l']),
// dg_t[l] = sum_k dout_t[b,k]*eo[b,k*E+l]
__global__ void __launch_bounds__(256)
mmoe_combine_bwd_kernel(const float* __restrict__ eo, long long ld, const float* __restrict__ gate,
                        const float* __restrict__ dout, int B, int Hh, int E, int T, float* __restrict__ deo,
                        long long ldd) {
    const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const float* row = eo + (size_t)b * ld;
    float* drow = deo + (size_t)b * ldd;
    // expert part
    for (int k = lane; k < Hh; k += 32) {
        for (int l = 0; l < E; ++l) {
            float acc = 0.f;
            for (int t = 0; t < T; ++t)
                acc = fmaf(__ldg(dout + ((size_t)t * B + b) * Hh + k), __ldg(gate + (size_t)b * T * E + t * E + l), acc);
            drow[k * E + l] = acc;
        }
    }
    // gate part
    for (int t = 0; t < T; ++t) {
        float dg = 0.f;      // lane l accumulates dg_t[l] over k: every lane loops over all k for its own l
        if (lane < E) {
            for (int k = 0; k < Hh; ++k) dg = fmaf(__ldg(dout + ((size_t)t * B + b) * Hh + k), __ldg(row + k * E + lane), dg);
        }
        const float g = lane < E ? __ldg(gate + (size_t)b * T * E + t * E + lane) : 0.f;
        const float dot = warp_sum(g * dg);
        if (lane < E) drow[Hh * E + t * E + lane] = g * (dg - dot);
    }
    for (int j = Hh * E + T * E + lane; j < ldd; j += 32) drow[j] = 0.f;
}

static void stats_grid(int M, int K, dim3& grid, int& slab);

void colsum_launch(const float* x, long long ldx, float* out, int M, int K, cudaStream_t st) {
    dim3 grid; int slab;
    stats_grid(M, K, grid, slab);
    colstats_kernel<<<grid, dim3(32, 8), 0, st>>>(x, ldx, out, nullptr, M, K, slab);
}

static void stats_grid(int M, int K, dim3& grid, int& slab) {
    const int kt = ceil_div(K, 32);
    int slabs = max(1, min(ceil_div(M, 64), (148 * 8) / kt));
    slab = ceil_div(M, slabs);
    slabs = ceil_div(M, slab);
    grid = dim3(kt, slabs);
}

}  // namespace rpb

using namespace rpb;

RPB_API int rpb_matmul_kn_fwd(const float* x, int64_t ldx, const float* Wkn, int64_t ldw, const float* bias, float* y,
                              int64_t ldy, int M, int N, int K, int impl, void* stream) {
    if (x == nullptr || Wkn == nullptr || y == nullptr || M <= 0 || N <= 0 || K <= 0) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool tc_ok = tc_shape_ok(x, ldx, M, N, K);
    if (impl == 2 && !tc_ok) return RPB_ERR_UNSUPPORTED;
    if (impl == 2 || (impl == 0 && tc_ok && M >= 512)) {
        TcEpilogue ep{y, ldy, bias, nullptr, 0, M, N, 0, nullptr};
        return gemm_tc(x, ldx, Wkn, ldw, 1, ep, M, N, K, st);      // B operand [N,K] = Wkn^T
    }
    return sgemm_kn_simt(x, ldx, Wkn, ldw, bias, y, ldy, M, N, K, st);
}

RPB_API int rpb_matmul_kn_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* Wkn, int64_t ldw,
                              float* dx, int64_t lddx, float* dWkn, float* db, int M, int N, int K, int impl,
                              void* stream) {
    if (dy == nullptr || M <= 0 || N <= 0 || K <= 0) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = 0;
    if (dx != nullptr) {
        if (Wkn == nullptr) return RPB_ERR_BAD_ARG;
        // dx[M,K] = dy[M,N] @ Wkn^T: reduction over N, B operand [K rows, N cols] = Wkn as stored
        const bool tc_ok = tc_shape_ok(dy, lddy, M, K, N);
        if (impl == 2 && !tc_ok) return RPB_ERR_UNSUPPORTED;
        if (impl == 2 || (impl == 0 && tc_ok && M >= 512)) {
            TcEpilogue ep{dx, lddx, nullptr, nullptr, 0, M, K, 0, nullptr};
            rc = gemm_tc(dy, lddy, Wkn, ldw, 0, ep, M, K, N, st);
        } else {
            rc = sgemm_nk_simt(dy, lddy, Wkn, ldw, dx, lddx, M, K, N, st);
        }
        if (rc != 0) return rc;
    }
    if (dWkn != nullptr) {
        if (x == nullptr) return RPB_ERR_BAD_ARG;
        // dWkn[k,n] += sum_m x[m,k] dy[m,n]  == wgrad with the roles of x and dy swapped (output [K, N], row stride N)
        rc = RPB_ERR_UNSUPPORTED;
        if (impl == 2 || (impl == 0 && M >= 2048 && g_wgrad_tc)) {
            rc = wgrad_tc(x, ldx, dy, lddy, dWkn, M, K, N, st);                 // tcgen05, sliced over K > 256
            if (rc != 0 && rc != RPB_ERR_UNSUPPORTED) return rc;
        }
        if (rc != 0) rc = linear_dw_simt(x, ldx, dy, lddy, dWkn, nullptr, M, K, N, st);
        if (rc != 0) return rc;
    }
    if (db != nullptr) {
        dim3 grid; int slab;
        stats_grid(M, N, grid, slab);
        colstats_kernel<<<grid, dim3(32, 8), 0, st>>>(dy, lddy, db, nullptr, M, N, slab);
        RPB_LAUNCH_CHECK();
    }
    return 0;
}

RPB_API int rpb_mmoe_combine_fwd(const float* eo, int64_t ld, int B, int Hh, int E, int T, float* out, float* gate,
                                 void* stream) {
    if (eo == nullptr || out == nullptr || gate == nullptr || B <= 0) return RPB_ERR_BAD_ARG;
    if (E < 1 || E > 32 || ld < (int64_t)Hh * E + T * E) return RPB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    mmoe_combine_fwd_kernel<<<ceil_div((long long)B * 32, 256), 256, 0, st>>>(eo, ld, B, Hh, E, T, out, gate);
    RPB_LAUNCH_CHECK();
    return 0;
}

RPB_API int rpb_mmoe_combine_bwd(const float* eo, int64_t ld, const float* gate, const float* dout, int B, int Hh, int E,
                                 int T, float* deo, int64_t ldd, void* stream) {
    if (eo == nullptr || gate == nullptr || dout == nullptr || deo == nullptr || B <= 0) return RPB_ERR_BAD_ARG;
    if (E < 1 || E > 32 || ldd < (int64_t)Hh * E + T * E) return RPB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    mmoe_combine_bwd_kernel<<<ceil_div((long long)B * 32, 256), 256, 0, st>>>(eo, ld, gate, dout, B, Hh, E, T, deo, ldd);
    RPB_LAUNCH_CHECK();
    return 0;
}

RPB_API int rpb_bn_stats(const float* x, int M, int N, float* sum, float* sumsq, void* stream) {
    if (x == nullptr || sum == nullptr || sumsq == nullptr || M <= 0 || N <= 0) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    dim3 grid; int slab;
    stats_grid(M, N, grid, slab);
    colstats_kernel<<<grid, dim3(32, 8), 0, st>>>(x, N, sum, sumsq, M, N, slab);
    RPB_LAUNCH_CHECK();
    return 0;
}

RPB_API int rpb_bn_apply(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                         float* y, int M, int N, void* stream) {
    if (x == nullptr || y == nullptr || mean == nullptr || invstd == nullptr || M <= 0 || N <= 0) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    bn_apply_kernel<<<ceil_div((long long)M * N, 256), 256, 0, st>>>(x, mean, invstd, gamma, beta, y, (long long)M * N, N);
    RPB_LAUNCH_CHECK();
    return 0;
}

RPB_API int rpb_bn_bwd_stats(const float* dy, const float* x, const float* mean, const float* invstd, float* dgamma,
                             float* dbeta, int M, int N, void* stream) {
    if (dy == nullptr || x == nullptr || mean == nullptr || invstd == nullptr || dgamma == nullptr || dbeta == nullptr || M <= 0)
        return RPB_ERR_BAD_ARG;
    dim3 grid; int slab;
    stats_grid(M, N, grid, slab);
    bn_bwd_stats_kernel<<<grid, dim3(32, 8), 0, reinterpret_cast<cudaStream_t>(stream)>>>(dy, x, mean, invstd, dbeta, dgamma, M, N, slab);
    RPB_LAUNCH_CHECK();
    return 0;
}

RPB_API int rpb_bn_bwd_dx(const float* dy, const float* x, const float* mean, const float* invstd, const float* gamma,
                          const float* dgamma_sum, const float* dbeta_sum, float* dx, int M, int N, float inv_count,
                          int use_batch_stats, void* stream) {
    if (dy == nullptr || x == nullptr || dx == nullptr || gamma == nullptr || invstd == nullptr || M <= 0) return RPB_ERR_BAD_ARG;
    if (use_batch_stats && (dgamma_sum == nullptr || dbeta_sum == nullptr || mean == nullptr)) return RPB_ERR_BAD_ARG;
    bn_bwd_dx_kernel<<<ceil_div((long long)M * N, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        dy, x, mean, invstd, gamma, dbeta_sum, dgamma_sum, dx, (long long)M * N, N, inv_count, use_batch_stats);
    RPB_LAUNCH_CHECK();
    return 0;
}

RPB_API int rpb_bn_bwd(const float* dy, const float* x, const float* mean, const float* invstd, const float* gamma,
                       float* dx, float* dgamma, float* dbeta, int M, int N, int use_batch_stats, void* stream) {
    if (dy == nullptr || x == nullptr || dx == nullptr || dgamma == nullptr || dbeta == nullptr || M <= 0) return RPB_ERR_BAD_ARG;
    int rc = rpb_bn_bwd_stats(dy, x, mean, invstd, dgamma, dbeta, M, N, stream);
    if (rc == 0) rc = rpb_bn_bwd_dx(dy, x, mean, invstd, gamma, dgamma, dbeta, dx, M, N, 1.f / (float)M, use_batch_stats, stream);
    return rc;
}
