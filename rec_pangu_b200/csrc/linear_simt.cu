// fp32 CUDA-core dense layers (exact fp32 FMA): the always-available implementation of the MLP tower's
// nn.Linear (+bias, +ReLU) and its backward (reference: models/layers/deep.py:62-70, autograd of aten::addmm).
// The tcgen05 3xTF32 path in linear_tc.cu is preferred when the shape qualifies; this file also serves as
// its on-device cross-check.
#include "common.cuh"

namespace rpb {

constexpr int BM = 128, BN = 64, BK = 16;   // CTA tile; 256 threads, each 8 (m) x 4 (n)

// C[m,n] = epi( sum_k A[m*lda+k] * Bm(k,n) ),  Bm(k,n) = B_NK ? Bp[n*ldb+k] : Bp[k*ldb+n]
template <bool B_NK>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ Bp, long long ldb,
             const float* __restrict__ bias, const float* __restrict__ mask, long long ldmask,
             float* __restrict__ C, long long ldc, int M, int N, int K, int relu) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int ty = tid / 16, tx = tid % 16;          // ty: 16 row groups of 8, tx: 16 col groups of 4
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
        // A tile: 128 x 16, k contiguous in memory: thread -> (row = tid/16 + 16*i, k = tid%16)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = tid / 16 + 16 * i, kk = tid % 16;
            const int m = m0 + r, k = k0 + kk;
            As[kk][r] = (m < M && k < K) ? __ldg(A + (size_t)m * lda + k) : 0.f;
        }
        if (B_NK) {   // W[n, k], k contiguous: thread -> (n = tid/16 + 16*i, k = tid%16)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = tid / 16 + 16 * i, kk = tid % 16;
                const int n = n0 + c, k = k0 + kk;
                Bs[kk][c] = (n < N && k < K) ? __ldg(Bp + (size_t)n * ldb + k) : 0.f;
            }
        } else {      // W[k, n], n contiguous: thread -> (k = tid/64 + 4*i, n = tid%64)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int kk = tid / 64 + 4 * i, c = tid % 64;
                const int n = n0 + c, k = k0 + kk;
                Bs[kk][c] = (n < N && k < K) ? __ldg(Bp + (size_t)k * ldb + n) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[8], b[4];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (bias != nullptr) v += __ldg(bias + n);
            if (relu) v = fmaxf(v, 0.f);
            if (mask != nullptr) v = (__ldg(mask + (size_t)m * ldmask + n) > 0.f) ? v : 0.f;
            C[(size_t)m * ldc + n] = v;
        }
    }
}

// dW[n,k] += sum_{m in slab} dy[m,n] * x[m,k];  db[n] += sum_m dy[m,n] (by the CTAs with blockIdx.x == 0)
// grid: (ceil(K/64), ceil(N/64), slabs)
__global__ void __launch_bounds__(256)
wgrad_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ x, long long ldx,
             float* __restrict__ dW, float* __restrict__ db, int M, int N, int K, int slab) {
    __shared__ float Ys[16][64 + 4];
    __shared__ float Xs[16][64 + 4];
    const int tid = threadIdx.x;
    const int k0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const int mbeg = blockIdx.z * slab, mend = min(M, mbeg + slab);
    const int ty = tid / 16, tx = tid % 16;       // outputs: n = n0 + ty*4 + i, k = k0 + tx*4 + j
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float colsum = 0.f;                            // thread tid < 64 sums column n0+tid
    for (int m0 = mbeg; m0 < mend; m0 += 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = tid / 64 + 4 * i, c = tid % 64;
            const int m = m0 + r;
            Ys[r][c] = (m < mend && n0 + c < N) ? __ldg(dy + (size_t)m * lddy + n0 + c) : 0.f;
            Xs[r][c] = (m < mend && k0 + c < K) ? __ldg(x + (size_t)m * ldx + k0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(&Ys[r][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Xs[r][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (db != nullptr && blockIdx.x == 0 && tid < 64) {
#pragma unroll
            for (int r = 0; r < 16; ++r) colsum += Ys[r][tid];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < K) red_add_f1(dW + (size_t)n * K + k, acc[i][j]);
        }
    }
    if (db != nullptr && blockIdx.x == 0 && tid < 64 && n0 + tid < N) red_add_f1(db + n0 + tid, colsum);
}

// out[m] = x[m,:K].w + bias + addends ; one warp per row
__global__ void __launch_bounds__(256)
rowdot_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                  const float* __restrict__ bias, const float* __restrict__ a0, const float* __restrict__ a1,
                  const float* __restrict__ a2, float* __restrict__ out, int M, int K) {
    const int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (warp >= M) return;
    const float* row = x + (size_t)warp * ldx;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(__ldg(row + k), __ldg(w + k), acc);
    acc = warp_sum(acc);
    if (lane == 0) {
        if (bias != nullptr) acc += __ldg(bias);
        if (a0 != nullptr) acc += __ldg(a0 + warp);
        if (a1 != nullptr) acc += __ldg(a1 + warp);
        if (a2 != nullptr) acc += __ldg(a2 + warp);
        out[warp] = acc;
    }
}

// dx[m,k] = dout[m] * w[k] (* mask);  elementwise, thread per (m, k)
__global__ void __launch_bounds__(256)
rowdot_dx_kernel(const float* __restrict__ dout, const float* __restrict__ w, const float* __restrict__ mask,
                 long long ldmask, float* __restrict__ dx, long long lddx, int M, int K) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int m = (int)(t / K), k = (int)(t % K);
    if (m >= M) return;
    float v = __ldg(dout + m) * __ldg(w + k);
    if (mask != nullptr) v = (__ldg(mask + (size_t)m * ldmask + k) > 0.f) ? v : 0.f;
    dx[(size_t)m * lddx + k] = v;
}

// dw[k] += sum_m dout[m] x[m,k]; db += sum_m dout[m].  Warp per row (coalesced 128-byte row segments), each lane keeps
// K/32 partial sums in registers over the rows of its warp, then one block reduction through shared memory.
template <int NJ>
__global__ void __launch_bounds__(256)
rowdot_dw_kernel(const float* __restrict__ dout, const float* __restrict__ x, long long ldx,
                 float* __restrict__ dw, float* __restrict__ db, int M, int K) {
    __shared__ float part[8][32 * NJ + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[j] = 0.f;
    float dsum = 0.f;
    for (long long m = (long long)blockIdx.x * 8 + warp; m < M; m += (long long)gridDim.x * 8) {
        const float d = __ldg(dout + m);
        dsum += d;
        const float* row = x + (size_t)m * ldx;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int k = lane + 32 * j;
            if (k < K) acc[j] = fmaf(d, __ldg(row + k), acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) part[warp][lane + 32 * j] = acc[j];
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += 256) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += part[w][k];
        red_add_f1(dw + k, s);
    }
    if (db != nullptr) {
        // every lane of a warp accumulated the same dsum; take lane 0 of each warp
        __syncthreads();
        if (lane == 0) part[warp][0] = dsum;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += part[w][0];
            red_add_f1(db, s);
        }
    }
}

int linear_fwd_simt(const float* x, long long ldx, const float* W, const float* bias, float* y, long long ldy,
                    int M, int N, int K, int act, cudaStream_t st) {
    dim3 grid(ceil_div(M, BM), ceil_div(N, BN));
    sgemm_kernel<true><<<grid, 256, 0, st>>>(x, ldx, W, K, bias, nullptr, 0, y, ldy, M, N, K, act == 1);
    RPB_LAUNCH_CHECK();
    return 0;
}

int linear_dx_simt(const float* dy, long long lddy, const float* W, const float* mask, long long ldmask,
                   float* dx, long long lddx, int M, int N, int K, cudaStream_t st) {
    // dx[M,K] = dy[M,N] @ W[N,K]: reduction over N, output width K
    dim3 grid(ceil_div(M, BM), ceil_div(K, BN));
    sgemm_kernel<false><<<grid, 256, 0, st>>>(dy, lddy, W, K, nullptr, mask, ldmask, dx, lddx, M, K, N, 0);
    RPB_LAUNCH_CHECK();
    return 0;
}

// y[M,N] = x[M,K] @ Wkn[K,N] + bias   (weights stored [K,N], row stride ldw)
int sgemm_kn_simt(const float* x, long long ldx, const float* Wkn, long long ldw, const float* bias, float* y,
                  long long ldy, int M, int N, int K, cudaStream_t st) {
    dim3 grid(ceil_div(M, BM), ceil_div(N, BN));
    sgemm_kernel<false><<<grid, 256, 0, st>>>(x, ldx, Wkn, ldw, bias, nullptr, 0, y, ldy, M, N, K, 0);
    RPB_LAUNCH_CHECK();
    return 0;
}

// y[M,N] = x[M,K] @ Wnk[N,K]^T   (row stride ldw)
int sgemm_nk_simt(const float* x, long long ldx, const float* Wnk, long long ldw, float* y, long long ldy, int M, int N,
                  int K, cudaStream_t st) {
    dim3 grid(ceil_div(M, BM), ceil_div(N, BN));
    sgemm_kernel<true><<<grid, 256, 0, st>>>(x, ldx, Wnk, ldw, nullptr, nullptr, 0, y, ldy, M, N, K, 0);
    RPB_LAUNCH_CHECK();
    return 0;
}

int linear_dw_simt(const float* dy, long long lddy, const float* x, long long ldx, float* dW, float* db,
                   int M, int N, int K, cudaStream_t st) {
    const int tiles = ceil_div(K, 64) * ceil_div(N, 64);
    int slabs = max(1, min(ceil_div(M, 256), ceil_div(148 * 4, tiles)));
    int slab = ceil_div(M, slabs);
    slab = ((slab + 15) / 16) * 16;
    slabs = ceil_div(M, slab);
    dim3 grid(ceil_div(K, 64), ceil_div(N, 64), slabs);
    wgrad_kernel<<<grid, 256, 0, st>>>(dy, lddy, x, ldx, dW, db, M, N, K, slab);
    RPB_LAUNCH_CHECK();
    return 0;
}

}  // namespace rpb

using namespace rpb;

RPB_API int rpb_rowdot_fwd(const float* x, int64_t ldx, const float* w, const float* bias, const float* add0,
                           const float* add1, const float* add2, float* out, int M, int K, void* stream) {
    if (x == nullptr || w == nullptr || out == nullptr || M <= 0 || K <= 0) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    rowdot_fwd_kernel<<<ceil_div((long long)M * 32, 256), 256, 0, st>>>(x, ldx, w, bias, add0, add1, add2, out, M, K);
    RPB_LAUNCH_CHECK();
    return 0;
}

RPB_API int rpb_rowdot_bwd(const float* dout, const float* x, int64_t ldx, const float* w, const float* mask,
                           int64_t ldmask, float* dx, int64_t lddx, float* dw, float* db, int M, int K,
                           void* stream) {
    if (dout == nullptr || M <= 0 || K <= 0) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dx != nullptr) {
        if (w == nullptr) return RPB_ERR_BAD_ARG;
        rowdot_dx_kernel<<<ceil_div((long long)M * K, 256), 256, 0, st>>>(dout, w, mask, ldmask, dx, lddx, M, K);
        RPB_LAUNCH_CHECK();
    }
    if (dw != nullptr) {
        if (x == nullptr) return RPB_ERR_BAD_ARG;
        const int grid = min(ceil_div(M, 8), 148 * 4);
        const int nj = ceil_div(K, 32);
        if (nj <= 2) rowdot_dw_kernel<2><<<grid, 256, 0, st>>>(dout, x, ldx, dw, db, M, K);
        else if (nj <= 8) rowdot_dw_kernel<8><<<grid, 256, 0, st>>>(dout, x, ldx, dw, db, M, K);
        else if (nj <= 32) rowdot_dw_kernel<32><<<grid, 256, 0, st>>>(dout, x, ldx, dw, db, M, K);
        else return RPB_ERR_UNSUPPORTED;
        RPB_LAUNCH_CHECK();
    }
    return 0;
}
