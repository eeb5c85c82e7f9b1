// DCN cross network, all layers fused (reference: models/layers/interaction.py:119-141):
//   x_{l+1} = x_l + (w_l . x_l) x_0 + b_l.
// One warp per sample: x_0 and x_l live in registers (K/32 values per lane), the dot product is a warp
// shuffle reduction, so the L layers cost one read of x_0 and one write of x_L.  Backward uses
//   x_l = (1 + sum_{j<l} s_j) x_0 + sum_{j<l} b_j   (s_l = w_l . x_l saved by forward, [B,L])
// so the per-sample pass only produces dx_0 and the 2L scalars c_l = ds_l (1+S_l), ds_l; the parameter
// gradients are then batch reductions: dw_l = X0^T c_l + (sum_b ds_l) Bp_l,  db_l = colsum(g_L) + sum_{j>l}(sum_b ds_j) w_j.
#include <type_traits>

#include "common.cuh"

namespace rpb {

constexpr int CN_MAXL = 8;
struct CrossPtrs { const float* w[CN_MAXL]; const float* b[CN_MAXL]; float* dw[CN_MAXL]; float* db[CN_MAXL]; };

int linear_dw_simt(const float* dy, long long lddy, const float* x, long long ldx, float* dW, float* db,
                   int M, int N, int K, cudaStream_t st);

template <int NJ>
__global__ void __launch_bounds__(256)
crossnet_fwd_kernel(const float* __restrict__ x0, long long ldx, int K, int L, CrossPtrs p, float* __restrict__ out,
                    long long ldo, float* __restrict__ S, int B) {
    const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    float a[NJ], x[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int k = lane + 32 * j;
        a[j] = (k < K) ? __ldg(x0 + (size_t)b * ldx + k) : 0.f;
        x[j] = a[j];
    }
    for (int l = 0; l < L; ++l) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int k = lane + 32 * j;
            if (k < K) s = fmaf(__ldg(p.w[l] + k), x[j], s);
        }
        s = warp_sum(s);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int k = lane + 32 * j;
            if (k < K) x[j] = x[j] + (fmaf(s, a[j], __ldg(p.b[l] + k)));
        }
        if (lane == 0 && S != nullptr) S[(size_t)b * L + l] = s;
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int k = lane + 32 * j;
        if (k < ldo) out[(size_t)b * ldo + k] = (k < K) ? x[j] : 0.f;
    }
}

template <int NJ>
__global__ void __launch_bounds__(256)
crossnet_bwd_kernel(const float* __restrict__ x0, long long ldx, int K, int L, CrossPtrs p, const float* __restrict__ S,
                    const float* __restrict__ dout, long long lddo, float* __restrict__ dx0, long long lddx,
                    float* __restrict__ C, int B) {
    const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    float a[NJ], g[NJ], acc[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int k = lane + 32 * j;
        a[j] = (k < K) ? __ldg(x0 + (size_t)b * ldx + k) : 0.f;
        g[j] = (k < K) ? __ldg(dout + (size_t)b * lddo + k) : 0.f;
        acc[j] = 0.f;
    }
    float Spre[CN_MAXL];                 // S_l = sum_{j<l} s_j
    float run = 0.f;
#pragma unroll
    for (int l = 0; l < CN_MAXL; ++l) {
        Spre[l] = run;
        if (l < L) run += __ldg(S + (size_t)b * L + l);
    }
#pragma unroll
    for (int l = CN_MAXL - 1; l >= 0; --l) {
        if (l < L) {
            const float sl = __ldg(S + (size_t)b * L + l);
            float ds = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) ds = fmaf(g[j], a[j], ds);
            ds = warp_sum(ds);
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int k = lane + 32 * j;
                acc[j] = fmaf(sl, g[j], acc[j]);
                if (k < K) g[j] = fmaf(ds, __ldg(p.w[l] + k), g[j]);
            }
            if (lane == 0) {
                C[(size_t)b * 2 * L + l] = ds * (1.f + Spre[l]);
                C[(size_t)b * 2 * L + L + l] = ds;
            }
        }
    }
    if (dx0 != nullptr) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int k = lane + 32 * j;
            if (k < lddx) dx0[(size_t)b * lddx + k] = (k < K) ? g[j] + acc[j] : 0.f;
        }
    }
}

// out[k] += sum_m x[m,k]   grid (ceil(K/256), slabs)
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, long long ldx, float* __restrict__ out, int M, int K, int slab) {
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= K) return;
    const int mbeg = blockIdx.y * slab, mend = min(M, mbeg + slab);
    float acc = 0.f;
    for (int m = mbeg; m < mend; ++m) acc += __ldg(x + (size_t)m * ldx + k);
    red_add_f1(out + k, acc);
}

// dw_l[k] += T[l][k] + dsum[l] * Bp_l[k];   db_l[k] += csum[k] + sum_{j>l} dsum[j] * w_j[k]
// T = [2L, K] (rows l<L used), tb = colsum of C [2L] (entries L..2L-1 = sum_b ds_l), csum = colsum(g_L) [K]
__global__ void __launch_bounds__(256)
crossnet_finalize_kernel(CrossPtrs p, const float* __restrict__ T, const float* __restrict__ tb,
                         const float* __restrict__ csum, int K, int L) {
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= K) return;
    float bp = 0.f;
    for (int l = 0; l < L; ++l) {
        if (p.dw[l] != nullptr) p.dw[l][k] += T[(size_t)l * K + k] + tb[L + l] * bp;
        bp += __ldg(p.b[l] + k);
    }
    float tail = 0.f;
    for (int l = L - 1; l >= 0; --l) {
        if (p.db[l] != nullptr) p.db[l][k] += csum[k] + tail;
        tail = fmaf(tb[L + l], __ldg(p.w[l] + k), tail);
    }
}

template <typename Fn>
static int cn_dispatch(int K, Fn&& fn) {
    const int nj = (K + 31) / 32;
    if (nj <= 4) return fn(std::integral_constant<int, 4>{});
    if (nj <= 8) return fn(std::integral_constant<int, 8>{});
    if (nj <= 16) return fn(std::integral_constant<int, 16>{});
    if (nj <= 32) return fn(std::integral_constant<int, 32>{});
    return RPB_ERR_UNSUPPORTED;
}

}  // namespace rpb

using namespace rpb;

RPB_API int rpb_crossnet_fwd(const float* x0, int64_t ldx, int K, int L, const float* const* w, const float* const* bias,
                             float* out, int64_t ldo, float* S, int B, void* stream) {
    if (x0 == nullptr || out == nullptr || w == nullptr || bias == nullptr || B <= 0 || K <= 0) return RPB_ERR_BAD_ARG;
    if (L < 1 || L > CN_MAXL || ldo > 32 * 32) return RPB_ERR_UNSUPPORTED;
    CrossPtrs p{};
    for (int l = 0; l < L; ++l) { p.w[l] = w[l]; p.b[l] = bias[l]; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return cn_dispatch((int)max((int64_t)K, ldo), [&](auto t) -> int {
        constexpr int NJ = decltype(t)::value;
        crossnet_fwd_kernel<NJ><<<ceil_div((long long)B * 32, 256), 256, 0, st>>>(x0, ldx, K, L, p, out, ldo, S, B);
        return (int)cudaGetLastError();
    });
}

RPB_API int rpb_crossnet_bwd(const float* x0, int64_t ldx, int K, int L, const float* const* w, const float* const* bias,
                             const float* S, const float* dout, int64_t lddo, float* dx0, int64_t lddx,
                             float* const* dw, float* const* db, int B, void* stream) {
    if (x0 == nullptr || S == nullptr || dout == nullptr || w == nullptr || bias == nullptr || B <= 0) return RPB_ERR_BAD_ARG;
    if (L < 1 || L > CN_MAXL || lddx > 32 * 32) return RPB_ERR_UNSUPPORTED;
    CrossPtrs p{};
    for (int l = 0; l < L; ++l) {
        p.w[l] = w[l]; p.b[l] = bias[l];
        p.dw[l] = dw ? dw[l] : nullptr; p.db[l] = db ? db[l] : nullptr;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // workspace: C [B, 2L] | T [2L, K] | tb [2L] | csum [K]
    const size_t nC = (size_t)B * 2 * L, nT = (size_t)2 * L * K;
    int werr = 0;
    float* ws = static_cast<float*>(workspace(3, (nC + nT + 2 * L + K + 8) * sizeof(float), &werr));
    if (ws == nullptr) return werr;
    float* C = ws; float* T = ws + nC; float* tb = T + nT; float* csum = tb + 2 * L;
    cudaMemsetAsync(T, 0, (nT + 2 * L + K) * sizeof(float), st);
    int rc = cn_dispatch((int)max((int64_t)K, dx0 ? lddx : 0), [&](auto t) -> int {
        constexpr int NJ = decltype(t)::value;
        crossnet_bwd_kernel<NJ><<<ceil_div((long long)B * 32, 256), 256, 0, st>>>(x0, ldx, K, L, p, S, dout, lddo, dx0, lddx, C, B);
        return (int)cudaGetLastError();
    });
    if (rc == 0 && dw != nullptr) {
        rc = linear_dw_simt(C, 2 * L, x0, ldx, T, tb, B, 2 * L, K, st);
        if (rc == 0) {
            const int kt = ceil_div(K, 256);
            int slabs = max(1, min(ceil_div(B, 64), (148 * 8) / kt));
            const int slab = ceil_div(B, slabs);
            slabs = ceil_div(B, slab);
            colsum_kernel<<<dim3(kt, slabs), 256, 0, st>>>(dout, lddo, csum, B, K, slab);
            crossnet_finalize_kernel<<<kt, 256, 0, st>>>(p, T, tb, csum, K, L);
            rc = (int)cudaGetLastError();
        }
    }
    return rc;
}
