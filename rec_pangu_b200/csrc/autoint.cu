// AutoInt interacting layer core (reference: models/layers/attention.py:12-32, 63-95).
//
// The projections Q|K|V|R = X . [W_q;W_k;W_v;W_res]^T run on the dense-layer GEMM (linear_tc.cu); this kernel
// is the per-sample attention with the reference's raw `.view(B*H, -1, d)` head regroup (attention.py:73-75,
// SURVEY.md App. A-1): the F x (H*d) projection is read as H pseudo-heads of F consecutive d-wide tokens,
// token t = p*F + r lives at field f0 = t / H, columns (t % H)*d .. +d.  No scaling, softmax over the F tokens of
// the pseudo-head, residual add, ReLU.
//
// One warp per sample, lane r = token r of the current pseudo-head (F <= 32): the query row, the F scores and the
// softmax live in the lane's registers; key/value rows are shared-memory broadcasts.  Backward recomputes the
// softmax and does the two transposed reductions (dK, dV) through a shared-memory copy of the score matrices.
#include <type_traits>

#include "common.cuh"

namespace rpb {

constexpr int AI_MAXF = 32;
constexpr int AI_WARPS = 4;

// qkvr: [B*F, ldq] with Q at col 0, K at HD, V at 2HD, R at 3HD (R absent when res != nullptr: then res is [B*F, ldres])
// VEC (rpb_set_option("autoint_vec", 1), the default; bit-identical to the scalar form, 12 % faster at config 4): a lane's DH outputs are contiguous and 16-byte
// aligned, so they move as DH/4 float4 requests instead of DH scalar ones — the scalar form issues 4x the memory
// instructions and writes every 32-byte sector of dQ|dK|dV|dR eight times over (ncu: 493 GB/s at 7.6 % of peak in backward).
template <int DH, bool VEC>   // attention_dim d
__global__ void __launch_bounds__(AI_WARPS * 32)
autoint_attn_fwd_kernel(const float* __restrict__ qkvr, long long ldq, const float* __restrict__ res, long long ldres,
                        float* __restrict__ out, int B, int F, int H, int ncols) {
    extern __shared__ __align__(16) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HD = H * DH;
    float* s = sm + (size_t)warp * F * ncols;              // this warp's copy of the sample's [F, ncols] block
    for (long long b = (long long)blockIdx.x * AI_WARPS + warp; b < B; b += (long long)gridDim.x * AI_WARPS) {
        __syncwarp();
        const float* src = qkvr + (size_t)b * F * ldq;
        for (int i = lane; i < F * (ncols / 4); i += 32) {
            const int f = i / (ncols / 4), c4 = i % (ncols / 4);
            reinterpret_cast<float4*>(s + f * ncols)[c4] = __ldg(reinterpret_cast<const float4*>(src + (size_t)f * ldq) + c4);
        }
        __syncwarp();
        for (int p = 0; p < H; ++p) {
            const int t = p * F + lane;                      // my token (valid when lane < F)
            const bool on = lane < F;
            const int f0 = on ? t / H : 0, c0 = on ? (t % H) * DH : 0;
            float q[DH];
#pragma unroll
            for (int e = 0; e < DH; ++e) q[e] = s[f0 * ncols + c0 + e];
            float sc[AI_MAXF];
            float mx = -INFINITY;
#pragma unroll
            for (int r2 = 0; r2 < AI_MAXF; ++r2) {
                sc[r2] = -INFINITY;
                if (r2 < F) {
                    const int t2 = p * F + r2;
                    const float* kr = s + (t2 / H) * ncols + HD + (t2 % H) * DH;
                    float a = 0.f;
#pragma unroll
                    for (int e = 0; e < DH; ++e) a = fmaf(q[e], kr[e], a);
                    sc[r2] = a;
                    mx = fmaxf(mx, a);
                }
            }
            float sum = 0.f;
#pragma unroll
            for (int r2 = 0; r2 < AI_MAXF; ++r2) if (r2 < F) { sc[r2] = expf(sc[r2] - mx); sum += sc[r2]; }
            const float inv = 1.f / sum;
            float o[DH];
#pragma unroll
            for (int e = 0; e < DH; ++e) o[e] = 0.f;
#pragma unroll
            for (int r2 = 0; r2 < AI_MAXF; ++r2) {
                if (r2 < F) {
                    const int t2 = p * F + r2;
                    const float* vr = s + (t2 / H) * ncols + 2 * HD + (t2 % H) * DH;
                    const float a = sc[r2] * inv;
#pragma unroll
                    for (int e = 0; e < DH; ++e) o[e] = fmaf(a, vr[e], o[e]);
                }
            }
            if (on) {
                float* dst = out + ((size_t)b * F + f0) * HD + c0;
                if constexpr (VEC) {
#pragma unroll
                    for (int e = 0; e < DH; e += 4) {
                        const float4 r = (res != nullptr) ? ldg_f4(res + ((size_t)b * F + f0) * ldres + c0 + e)
                                                          : *reinterpret_cast<const float4*>(s + f0 * ncols + 3 * HD + c0 + e);
                        stg_f4(dst + e, make_float4(fmaxf(o[e] + r.x, 0.f), fmaxf(o[e + 1] + r.y, 0.f),
                                                    fmaxf(o[e + 2] + r.z, 0.f), fmaxf(o[e + 3] + r.w, 0.f)));
                    }
                } else {
#pragma unroll
                for (int e = 0; e < DH; ++e) {
                    const float r = (res != nullptr) ? __ldg(res + ((size_t)b * F + f0) * ldres + c0 + e)
                                                     : s[f0 * ncols + 3 * HD + c0 + e];
                    dst[e] = fmaxf(o[e] + r, 0.f);
                }
                }
            }
        }
    }
}

// dqkvr: [B*F, ldq] receives dQ|dK|dV|dR (dR = masked dout; written even when the residual is X itself so that the
// caller can add it to dX).
template <int DH, bool VEC>
__global__ void __launch_bounds__(AI_WARPS * 32)
autoint_attn_bwd_kernel(const float* __restrict__ qkvr, long long ldq, const float* __restrict__ out,
                        const float* __restrict__ dout, float* __restrict__ dqkvr, long long lddq, int B, int F, int H,
                        int ncols) {
    extern __shared__ __align__(16) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HD = H * DH;
    const int FP = F + 1;
    const int per_warp = F * ncols + 2 * F * FP + F * DH;
    float* s = sm + (size_t)warp * per_warp;
    float* A = s + F * ncols;            // [F][FP] softmax
    float* DS = A + F * FP;              // [F][FP] dscore
    float* DO = DS + F * FP;             // [F][DH] masked dout of the pseudo-head's tokens
    for (long long b = (long long)blockIdx.x * AI_WARPS + warp; b < B; b += (long long)gridDim.x * AI_WARPS) {
        __syncwarp();
        const float* src = qkvr + (size_t)b * F * ldq;
        for (int i = lane; i < F * (ncols / 4); i += 32) {
            const int f = i / (ncols / 4), c4 = i % (ncols / 4);
            reinterpret_cast<float4*>(s + f * ncols)[c4] = __ldg(reinterpret_cast<const float4*>(src + (size_t)f * ldq) + c4);
        }
        __syncwarp();
        for (int p = 0; p < H; ++p) {
            const int t = p * F + lane;
            const bool on = lane < F;
            const int f0 = on ? t / H : 0, c0 = on ? (t % H) * DH : 0;
            float q[DH], g[DH];
            if constexpr (VEC) {
#pragma unroll
                for (int e = 0; e < DH; ++e) { q[e] = s[f0 * ncols + c0 + e]; g[e] = 0.f; }
                if (on) {
                    const size_t oi = ((size_t)b * F + f0) * HD + c0;
#pragma unroll
                    for (int e = 0; e < DH; e += 4) {
                        const float4 ov = ldg_f4(out + oi + e), dv4 = ldg_f4(dout + oi + e);
                        g[e] = ov.x > 0.f ? dv4.x : 0.f; g[e + 1] = ov.y > 0.f ? dv4.y : 0.f;       // ReLU backward
                        g[e + 2] = ov.z > 0.f ? dv4.z : 0.f; g[e + 3] = ov.w > 0.f ? dv4.w : 0.f;
                        const float4 gv = make_float4(g[e], g[e + 1], g[e + 2], g[e + 3]);
                        stg_f4(dqkvr + ((size_t)b * F + f0) * lddq + 3 * HD + c0 + e, gv);            // residual branch grad
                        *reinterpret_cast<float4*>(DO + lane * DH + e) = gv;
                    }
                }
            } else {
#pragma unroll
            for (int e = 0; e < DH; ++e) {
                q[e] = s[f0 * ncols + c0 + e];
                g[e] = 0.f;
                if (on) {
                    const size_t oi = ((size_t)b * F + f0) * HD + c0 + e;
                    g[e] = (__ldg(out + oi) > 0.f) ? __ldg(dout + oi) : 0.f;        // ReLU backward
                    dqkvr[((size_t)b * F + f0) * lddq + 3 * HD + c0 + e] = g[e];      // residual branch grad
                    DO[lane * DH + e] = g[e];
                }
            }
            }
            float sc[AI_MAXF];
            float mx = -INFINITY;
#pragma unroll
            for (int r2 = 0; r2 < AI_MAXF; ++r2) {
                sc[r2] = -INFINITY;
                if (r2 < F) {
                    const int t2 = p * F + r2;
                    const float* kr = s + (t2 / H) * ncols + HD + (t2 % H) * DH;
                    float a = 0.f;
#pragma unroll
                    for (int e = 0; e < DH; ++e) a = fmaf(q[e], kr[e], a);
                    sc[r2] = a;
                    mx = fmaxf(mx, a);
                }
            }
            float sum = 0.f;
#pragma unroll
            for (int r2 = 0; r2 < AI_MAXF; ++r2) if (r2 < F) { sc[r2] = expf(sc[r2] - mx); sum += sc[r2]; }
            const float inv = 1.f / sum;
            // da[r2] = g . v_{t2};  dot = sum_r2 a*da
            float da[AI_MAXF];
            float dot = 0.f;
#pragma unroll
            for (int r2 = 0; r2 < AI_MAXF; ++r2) {
                da[r2] = 0.f;
                if (r2 < F) {
                    const int t2 = p * F + r2;
                    const float* vr = s + (t2 / H) * ncols + 2 * HD + (t2 % H) * DH;
                    float x = 0.f;
#pragma unroll
                    for (int e = 0; e < DH; ++e) x = fmaf(g[e], vr[e], x);
                    sc[r2] *= inv;                                   // sc now holds the softmax a[r2]
                    da[r2] = x;
                    dot = fmaf(sc[r2], x, dot);
                }
            }
            float dq[DH];
#pragma unroll
            for (int e = 0; e < DH; ++e) dq[e] = 0.f;
#pragma unroll
            for (int r2 = 0; r2 < AI_MAXF; ++r2) {
                if (r2 < F) {
                    const float ds = sc[r2] * (da[r2] - dot);
                    const int t2 = p * F + r2;
                    const float* kr = s + (t2 / H) * ncols + HD + (t2 % H) * DH;
#pragma unroll
                    for (int e = 0; e < DH; ++e) dq[e] = fmaf(ds, kr[e], dq[e]);
                    if (on) { A[lane * FP + r2] = sc[r2]; DS[lane * FP + r2] = ds; }
                }
            }
            if (on) {
                if constexpr (VEC) {
#pragma unroll
                    for (int e = 0; e < DH; e += 4)
                        stg_f4(dqkvr + ((size_t)b * F + f0) * lddq + c0 + e, make_float4(dq[e], dq[e + 1], dq[e + 2], dq[e + 3]));
                } else {
#pragma unroll
                for (int e = 0; e < DH; ++e) dqkvr[((size_t)b * F + f0) * lddq + c0 + e] = dq[e];
                }
            }
            __syncwarp();
            // transposed reductions: lane = key/value token r2
            if (on) {
                float dk[DH], dv[DH];
#pragma unroll
                for (int e = 0; e < DH; ++e) { dk[e] = 0.f; dv[e] = 0.f; }
                for (int r = 0; r < F; ++r) {
                    const float dsv = DS[r * FP + lane], av = A[r * FP + lane];
                    const int tr = p * F + r;
                    const float* qr = s + (tr / H) * ncols + (tr % H) * DH;
#pragma unroll
                    for (int e = 0; e < DH; ++e) {
                        dk[e] = fmaf(dsv, qr[e], dk[e]);
                        dv[e] = fmaf(av, DO[r * DH + e], dv[e]);
                    }
                }
                if constexpr (VEC) {
#pragma unroll
                    for (int e = 0; e < DH; e += 4) {
                        stg_f4(dqkvr + ((size_t)b * F + f0) * lddq + HD + c0 + e, make_float4(dk[e], dk[e + 1], dk[e + 2], dk[e + 3]));
                        stg_f4(dqkvr + ((size_t)b * F + f0) * lddq + 2 * HD + c0 + e, make_float4(dv[e], dv[e + 1], dv[e + 2], dv[e + 3]));
                    }
                } else {
#pragma unroll
                for (int e = 0; e < DH; ++e) {
                    dqkvr[((size_t)b * F + f0) * lddq + HD + c0 + e] = dk[e];
                    dqkvr[((size_t)b * F + f0) * lddq + 2 * HD + c0 + e] = dv[e];
                }
                }
            }
            __syncwarp();
        }
    }
}

template <typename Fn>
static int ai_dispatch(int d, Fn&& fn) {
    switch (d) {
        case 4: return fn(std::integral_constant<int, 4>{});
        case 8: return fn(std::integral_constant<int, 8>{});
        case 16: return fn(std::integral_constant<int, 16>{});
        default: return RPB_ERR_UNSUPPORTED;
    }
}

}  // namespace rpb

using namespace rpb;

RPB_API int rpb_autoint_attn_fwd(const float* qkvr, int64_t ldq, const float* res, int64_t ldres, float* out, int B,
                                 int F, int H, int d, void* stream) {
    if (qkvr == nullptr || out == nullptr || B <= 0 || F <= 0 || H <= 0) return RPB_ERR_BAD_ARG;
    if (F > AI_MAXF || (ldq % 4) != 0 || (reinterpret_cast<uintptr_t>(qkvr) & 15u)) return RPB_ERR_UNSUPPORTED;
    const int HD = H * d;
    const int ncols = (res != nullptr ? 3 : 4) * HD;
    if (ncols % 4 != 0 || ncols > ldq) return RPB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return ai_dispatch(d, [&](auto dt) -> int {
        constexpr int DH = decltype(dt)::value;
        const size_t smem = (size_t)AI_WARPS * F * ncols * sizeof(float);
        if (smem > 200 * 1024) return RPB_ERR_UNSUPPORTED;
        const int grid = min(ceil_div(B, AI_WARPS), 148 * 8);
        const bool vec = g_autoint_vec && (HD % 4) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0 &&
                         (res == nullptr || ((ldres % 4) == 0 && (reinterpret_cast<uintptr_t>(res) & 15u) == 0));
        if (vec) {
            cudaError_t ev = cudaFuncSetAttribute(autoint_attn_fwd_kernel<DH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (ev != cudaSuccess) return (int)ev;
            autoint_attn_fwd_kernel<DH, true><<<grid, AI_WARPS * 32, smem, st>>>(qkvr, ldq, res, ldres, out, B, F, H, ncols);
            return (int)cudaGetLastError();
        }
        cudaError_t e = cudaFuncSetAttribute(autoint_attn_fwd_kernel<DH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        autoint_attn_fwd_kernel<DH, false><<<grid, AI_WARPS * 32, smem, st>>>(qkvr, ldq, res, ldres, out, B, F, H, ncols);
        return (int)cudaGetLastError();
    });
}

RPB_API int rpb_autoint_attn_bwd(const float* qkvr, int64_t ldq, int has_res_proj, const float* out, const float* dout,
                                 float* dqkvr, int64_t lddq, int B, int F, int H, int d, void* stream) {
    if (qkvr == nullptr || out == nullptr || dout == nullptr || dqkvr == nullptr || B <= 0) return RPB_ERR_BAD_ARG;
    if (F > AI_MAXF || (ldq % 4) != 0 || (reinterpret_cast<uintptr_t>(qkvr) & 15u)) return RPB_ERR_UNSUPPORTED;
    const int HD = H * d;
    if (4 * HD > lddq) return RPB_ERR_BAD_ARG;           // dqkvr always has the 4 blocks dQ|dK|dV|dR
    const int ncols = (has_res_proj ? 4 : 3) * HD;
    if (ncols > ldq) return RPB_ERR_BAD_ARG;
    if (ncols % 4 != 0) return RPB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return ai_dispatch(d, [&](auto dt) -> int {
        constexpr int DH = decltype(dt)::value;
        const size_t smem = (size_t)AI_WARPS * (F * ncols + 2 * F * (F + 1) + F * DH) * sizeof(float);
        if (smem > 200 * 1024) return RPB_ERR_UNSUPPORTED;
        const int grid = min(ceil_div(B, AI_WARPS), 148 * 6);
        const bool vec = g_autoint_vec && (HD % 4) == 0 && (lddq % 4) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0 &&
                         (reinterpret_cast<uintptr_t>(dout) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dqkvr) & 15u) == 0 &&
                         ((F * DH) % 4) == 0;
        if (vec) {
            cudaError_t ev = cudaFuncSetAttribute(autoint_attn_bwd_kernel<DH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (ev != cudaSuccess) return (int)ev;
            autoint_attn_bwd_kernel<DH, true><<<grid, AI_WARPS * 32, smem, st>>>(qkvr, ldq, out, dout, dqkvr, lddq, B, F, H, ncols);
            return (int)cudaGetLastError();
        }
        cudaError_t e = cudaFuncSetAttribute(autoint_attn_bwd_kernel<DH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        autoint_attn_bwd_kernel<DH, false><<<grid, AI_WARPS * 32, smem, st>>>(qkvr, ldq, out, dout, dqkvr, lddq, B, F, H, ncols);
        return (int)cudaGetLastError();
    });
}
