// xDeepFM Compressed Interaction Network, fused (reference: models/layers/interaction.py:144-171).
//
// Reference per layer k:  Z[b,h*M+m,d] = X0[b,h,d]*Xk[b,m,d]  (einsum, 2.8 GB at config 3),
//                         X_{k+1} = Conv1d_{k=1}(Z) -> [B,U_k,D],  p_k = sum_d X_{k+1}.
// Here the outer product never leaves registers: D lanes own one sample (lane = embedding dim d); a lane keeps
// Xk[:,d] in registers, reads X0[h,d] from shared memory and streams the transposed 1x1-conv weights
// Wt[h*M+m][o] from shared memory as broadcast 16-byte loads: per (h,m) one FMUL + U FFMA per lane.
// All layers' weights (96 KB at config 3) stay resident in shared memory of a persistent CTA.
//
// Backward = (A) per-sample kernel: recompute forward, back-propagate G_k = dL/dX_{k+1}, produce dE and spill
// G_k / X_k; (B) weight-gradient kernel: dW_k[o,h*M+m] = sum_{b,d} G_k[b,o,d] X0[b,h,d] Xk[b,m,d] as a
// register-tiled reduction over (b,d) rows with the outer product again formed on the fly.
#include <type_traits>

#include "common.cuh"

namespace rpb {

constexpr int CIN_MAXL = 8;
constexpr int CIN_MAXM = 32;     // max fields / max units of a layer feeding the next

struct CinMeta {
    int L, F;
    int U[CIN_MAXL];       // units of layer k
    int M[CIN_MAXL];       // inputs of layer k (F for k=0 else U[k-1])
    int ustr[CIN_MAXL];    // U rounded up to 4 (row stride of Wt)
    int w_off[CIN_MAXL];   // float offset of layer k in the packed Wt buffer
    int p_off[CIN_MAXL];   // offset of layer k in pooled / bias (prefix sum of U)
    int w_total, u_total;
};

struct CinPtrs {            // passed by value: no host->device pointer-table copies (CUDA-graph friendly)
    const float* W[CIN_MAXL];
    const float* bias[CIN_MAXL];
    float* dW[CIN_MAXL];
    float* db[CIN_MAXL];
};

// Wt_k[(h*M+m)*ustr + o] = W_k[o*(F*M) + h*M + m];  bcat[p_off_k + o] = bias_k[o]
__global__ void cin_pack_kernel(CinPtrs ptrs, CinMeta meta, float* __restrict__ Wt, float* __restrict__ bcat) {
    const int k = blockIdx.y;
    const int FM = meta.F * meta.M[k], ustr = meta.ustr[k], U = meta.U[k];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < FM * ustr; i += gridDim.x * blockDim.x) {
        const int hm = i / ustr, o = i % ustr;
        Wt[meta.w_off[k] + i] = (o < U) ? __ldg(ptrs.W[k] + (size_t)o * FM + hm) : 0.f;
    }
    if (blockIdx.x == 0)
        for (int o = threadIdx.x; o < U; o += blockDim.x)
            bcat[meta.p_off[k] + o] = ptrs.bias[k] != nullptr ? __ldg(ptrs.bias[k] + o) : 0.f;
}

template <int MAXU>
__device__ __forceinline__ void cin_layer_fwd(const float* __restrict__ xs_col, int dstride, const float* __restrict__ Wt,
                                              int F, int M, int U, int ustr, const float (&xk)[CIN_MAXM],
                                              float (&acc)[MAXU]) {
#pragma unroll
    for (int o = 0; o < MAXU; ++o) acc[o] = 0.f;
    for (int h = 0; h < F; ++h) {
        const float a = xs_col[h * dstride];
        const float* wrow = Wt + (size_t)h * M * ustr;
#pragma unroll
        for (int m = 0; m < CIN_MAXM; ++m) {
            if (m < M) {
                const float z = a * xk[m];
                const float4* w4 = reinterpret_cast<const float4*>(wrow + m * ustr);
#pragma unroll
                for (int o4 = 0; o4 < MAXU / 4; ++o4) {
                    if (o4 * 4 < U) {
                        const float4 w = w4[o4];
                        acc[o4 * 4 + 0] = fmaf(w.x, z, acc[o4 * 4 + 0]);
                        acc[o4 * 4 + 1] = fmaf(w.y, z, acc[o4 * 4 + 1]);
                        acc[o4 * 4 + 2] = fmaf(w.z, z, acc[o4 * 4 + 2]);
                        acc[o4 * 4 + 3] = fmaf(w.w, z, acc[o4 * 4 + 3]);
                    }
                }
            }
        }
    }
}

// smem: [Wt (w_total) | bias (u_total) | xs: SPC * F * D]
template <int D, int MAXU>
__global__ void __launch_bounds__(256)
cin_fwd_kernel(const float* __restrict__ e, long long lde, int B, CinMeta meta, const float* __restrict__ Wt_g,
               const float* __restrict__ bias_g, float* __restrict__ pooled, int ldp) {
    extern __shared__ __align__(16) float sm[];
    constexpr int SPC = 256 / D;
    float* Wt = sm;
    float* bias = sm + meta.w_total;
    float* xs = bias + ((meta.u_total + 3) & ~3);
    for (int i = threadIdx.x; i < meta.w_total / 4; i += 256)
        reinterpret_cast<float4*>(Wt)[i] = __ldg(reinterpret_cast<const float4*>(Wt_g) + i);
    for (int i = threadIdx.x; i < meta.u_total; i += 256) bias[i] = __ldg(bias_g + i);
    __syncthreads();
    const int slot = threadIdx.x / D, d = threadIdx.x % D;
    const int F = meta.F;
    float* xcol = xs + (size_t)slot * F * D + d;           // xs[slot][h][d], stride D over h
    for (long long grp = blockIdx.x; grp * SPC < B; grp += gridDim.x) {
        const long long b_raw = grp * SPC + slot;
        const bool valid = b_raw < B;
        const long long b = valid ? b_raw : B - 1;
        const float* erow = e + (size_t)b * lde + d;
        float xk[CIN_MAXM];
#pragma unroll
        for (int m = 0; m < CIN_MAXM; ++m) xk[m] = 0.f;
        for (int h = 0; h < F; ++h) xcol[h * D] = __ldg(erow + h * D);
#pragma unroll
        for (int m = 0; m < CIN_MAXM; ++m) if (m < F) xk[m] = xcol[m * D];
        for (int k = 0; k < meta.L; ++k) {
            float acc[MAXU];
            cin_layer_fwd<MAXU>(xcol, D, Wt + meta.w_off[k], F, meta.M[k], meta.U[k], meta.ustr[k], xk, acc);
#pragma unroll
            for (int o = 0; o < MAXU; ++o) {
                if (o < meta.U[k]) {
                    const float v = acc[o] + bias[meta.p_off[k] + o];
                    xk[o] = v;
                    const float s = group_sum<D>(v);
                    if (d == 0 && valid) pooled[(size_t)b * ldp + meta.p_off[k] + o] = s;
                }
            }
        }
    }
}

// Backward kernel A.  smem: [Wt | bias | xs: SPC*F*D | dx0: SPC*F*D | xk store: SPC*u_total*D]
// Spills for kernel B:  Gout[b][p_off_k + o][d]  (all layers),  Xout[b][p_off_k + o][d] = X_{k+1} (k < L-1 used).
template <int D, int MAXU>
__global__ void __launch_bounds__(256, 1)
cin_bwd_sample_kernel(const float* __restrict__ e, long long lde, int B, CinMeta meta, const float* __restrict__ Wt_g,
                      const float* __restrict__ bias_g, const float* __restrict__ dpooled, int lddp,
                      float* __restrict__ de, long long ldde, int accumulate, float* __restrict__ Gout,
                      float* __restrict__ Xout) {
    extern __shared__ __align__(16) float sm[];
    constexpr int SPC = 256 / D;
    float* Wt = sm;
    float* bias = sm + meta.w_total;
    float* xs = bias + ((meta.u_total + 3) & ~3);
    float* dx0 = xs + (size_t)SPC * meta.F * D;
    float* xst = dx0 + (size_t)SPC * meta.F * D;
    for (int i = threadIdx.x; i < meta.w_total / 4; i += 256)
        reinterpret_cast<float4*>(Wt)[i] = __ldg(reinterpret_cast<const float4*>(Wt_g) + i);
    for (int i = threadIdx.x; i < meta.u_total; i += 256) bias[i] = __ldg(bias_g + i);
    __syncthreads();
    const int slot = threadIdx.x / D, d = threadIdx.x % D;
    const int F = meta.F, L = meta.L;
    float* xcol = xs + (size_t)slot * F * D + d;
    float* dcol = dx0 + (size_t)slot * F * D + d;
    float* kcol = xst + (size_t)slot * meta.u_total * D + d;     // kcol[(p_off_k + o) * D] = X_{k+1}[o][d]
    for (long long grp = blockIdx.x; grp * SPC < B; grp += gridDim.x) {
        const long long b_raw = grp * SPC + slot;
        const bool valid = b_raw < B;
        const long long b = valid ? b_raw : B - 1;
        const float* erow = e + (size_t)b * lde + d;
        float xk[CIN_MAXM];
#pragma unroll
        for (int m = 0; m < CIN_MAXM; ++m) xk[m] = 0.f;
        for (int h = 0; h < F; ++h) { xcol[h * D] = __ldg(erow + h * D); dcol[h * D] = 0.f; }
#pragma unroll
        for (int m = 0; m < CIN_MAXM; ++m) if (m < F) xk[m] = xcol[m * D];
        // ---- recompute forward, keep every layer output in shared memory
        for (int k = 0; k < L; ++k) {
            float acc[MAXU];
            cin_layer_fwd<MAXU>(xcol, D, Wt + meta.w_off[k], F, meta.M[k], meta.U[k], meta.ustr[k], xk, acc);
#pragma unroll
            for (int o = 0; o < MAXU; ++o) {
                if (o < meta.U[k]) {
                    const float v = acc[o] + bias[meta.p_off[k] + o];
                    xk[o] = v;
                    kcol[(meta.p_off[k] + o) * D] = v;
                    if (valid && k < L - 1) Xout[((size_t)b * meta.u_total + meta.p_off[k] + o) * D + d] = v;
                }
            }
        }
        // ---- backward through the layers
        float gx[MAXU];
#pragma unroll
        for (int o = 0; o < MAXU; ++o) gx[o] = 0.f;
        for (int k = L - 1; k >= 0; --k) {
            const int M = meta.M[k], U = meta.U[k], ustr = meta.ustr[k];
            float G[MAXU];
#pragma unroll
            for (int o = 0; o < MAXU; ++o) {
                G[o] = 0.f;
                if (o < U) {
                    G[o] = __ldg(dpooled + (size_t)b * lddp + meta.p_off[k] + o) + gx[o];
                    if (valid) Gout[((size_t)b * meta.u_total + meta.p_off[k] + o) * D + d] = G[o];
                }
            }
            float dxk[CIN_MAXM];
#pragma unroll
            for (int m = 0; m < CIN_MAXM; ++m) {
                dxk[m] = 0.f;
                xk[m] = 0.f;
                if (m < M) xk[m] = (k == 0) ? xcol[m * D] : kcol[(meta.p_off[k - 1] + m) * D];
            }
            const float* Wk = Wt + meta.w_off[k];
            for (int h = 0; h < F; ++h) {
                const float a = xcol[h * D];
                float s0 = 0.f;
                const float* wrow = Wk + (size_t)h * M * ustr;
#pragma unroll
                for (int m = 0; m < CIN_MAXM; ++m) {
                    if (m < M) {
                        const float4* w4 = reinterpret_cast<const float4*>(wrow + m * ustr);
                        float T = 0.f;
#pragma unroll
                        for (int o4 = 0; o4 < MAXU / 4; ++o4) {
                            if (o4 * 4 < U) {
                                const float4 w = w4[o4];
                                T = fmaf(w.x, G[o4 * 4 + 0], T);
                                T = fmaf(w.y, G[o4 * 4 + 1], T);
                                T = fmaf(w.z, G[o4 * 4 + 2], T);
                                T = fmaf(w.w, G[o4 * 4 + 3], T);
                            }
                        }
                        dxk[m] = fmaf(T, a, dxk[m]);
                        s0 = fmaf(T, xk[m], s0);
                    }
                }
                dcol[h * D] += s0;
            }
            if (k > 0) {
#pragma unroll
                for (int o = 0; o < MAXU; ++o) gx[o] = (o < M) ? dxk[o] : 0.f;
            } else {
#pragma unroll
                for (int m = 0; m < CIN_MAXM; ++m) if (m < F) dcol[m * D] += dxk[m];
            }
        }
        if (valid) {
            float* drow = de + (size_t)b * ldde + d;
            for (int h = 0; h < F; ++h) drow[h * D] = accumulate ? drow[h * D] + dcol[h * D] : dcol[h * D];
        }
    }
}

// Backward kernel B: dW_k[o, hm] += sum_{(b,d) in slab} G_k[b,o,d] * X0[b,h,d] * Xk[b,m,d];  db_k[o] += sum G_k.
// grid (slabs, L).  Tile = ST samples; smem tiles are [s][d][*] so lanes (consecutive hm -> consecutive m) are
// conflict-free.  Thread t owns hm in {t, t+256, ...} (NJ of them) x all U outputs.
template <int D, int MAXU, int NJ>
__global__ void __launch_bounds__(256)
cin_wgrad_kernel(const float* __restrict__ e, long long lde, int B, CinMeta meta, const float* __restrict__ Gin,
                 const float* __restrict__ Xin, CinPtrs ptrs, int slab) {
    constexpr int ST = 8;
    extern __shared__ __align__(16) float sm[];
    const int k = blockIdx.y;
    const int F = meta.F, M = meta.M[k], U = meta.U[k], FM = F * M;
    float* x0s = sm;                                  // [ST][D][F]
    float* xks = x0s + ST * D * F;                    // [ST][D][M]
    float* gs = xks + ST * D * CIN_MAXM;              // [ST][D][MAXU]
    const int t = threadIdx.x;
    int hh[NJ], mm[NJ];
    bool on[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int hm = t + j * 256;
        on[j] = hm < FM;
        hh[j] = on[j] ? hm / M : 0;
        mm[j] = on[j] ? hm % M : 0;
    }
    float acc[NJ][MAXU];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int o = 0; o < MAXU; ++o) acc[j][o] = 0.f;
    float bsum = 0.f;
    const long long bbeg = (long long)blockIdx.x * slab, bend = min((long long)B, bbeg + slab);
    for (long long b0 = bbeg; b0 < bend; b0 += ST) {
        const int ns = (int)min((long long)ST, bend - b0);
        __syncthreads();
        for (int i = t; i < ns * F * D; i += 256) {           // e[b][h][d] -> x0s[s][d][h]
            const int s = i / (F * D), r = i % (F * D), h = r / D, d = r % D;
            x0s[(s * D + d) * F + h] = __ldg(e + (size_t)(b0 + s) * lde + r);
        }
        for (int i = t; i < ns * M * D; i += 256) {           // Xk[b][m][d] -> xks[s][d][m]
            const int s = i / (M * D), r = i % (M * D), m = r / D, d = r % D;
            float v;
            if (k == 0) v = __ldg(e + (size_t)(b0 + s) * lde + r);
            else v = __ldg(Xin + ((size_t)(b0 + s) * meta.u_total + meta.p_off[k - 1]) * D + r);
            xks[(s * D + d) * CIN_MAXM + m] = v;
        }
        for (int i = t; i < ns * U * D; i += 256) {           // G[b][o][d] -> gs[s][d][o]
            const int s = i / (U * D), r = i % (U * D), o = r / D, d = r % D;
            gs[(s * D + d) * MAXU + o] = __ldg(Gin + ((size_t)(b0 + s) * meta.u_total + meta.p_off[k]) * D + r);
        }
        __syncthreads();
        for (int r = 0; r < ns * D; ++r) {
            const float4* g4 = reinterpret_cast<const float4*>(gs + r * MAXU);
            float g[MAXU];
#pragma unroll
            for (int o4 = 0; o4 < MAXU / 4; ++o4) {
                const float4 v = (o4 * 4 < U) ? g4[o4] : make_float4(0.f, 0.f, 0.f, 0.f);
                g[o4 * 4] = v.x; g[o4 * 4 + 1] = v.y; g[o4 * 4 + 2] = v.z; g[o4 * 4 + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if (on[j]) {
                    const float z = x0s[r * F + hh[j]] * xks[r * CIN_MAXM + mm[j]];
#pragma unroll
                    for (int o = 0; o < MAXU; ++o) acc[j][o] = fmaf(g[o], z, acc[j][o]);
                }
            }
        }
        if (t < U) for (int r = 0; r < ns * D; ++r) bsum += gs[r * MAXU + t];
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        if (on[j]) {
            const int hm = t + j * 256;
#pragma unroll
            for (int o = 0; o < MAXU; ++o) if (o < U) red_add_f1(ptrs.dW[k] + (size_t)o * FM + hm, acc[j][o]);
        }
    }
    if (t < U && ptrs.db[k] != nullptr) red_add_f1(ptrs.db[k] + t, bsum);
}

struct CinHost {
    CinMeta meta;
    bool ok;
};

static CinHost cin_meta(int F, int L, const int* units) {
    CinHost h{};
    h.ok = false;
    if (L < 1 || L > CIN_MAXL || F < 1 || F > CIN_MAXM) return h;
    CinMeta& m = h.meta;
    m.L = L; m.F = F;
    int woff = 0, poff = 0;
    for (int k = 0; k < L; ++k) {
        if (units[k] < 1 || units[k] > 32) return h;
        m.U[k] = units[k];
        m.M[k] = (k == 0) ? F : units[k - 1];
        m.ustr[k] = (units[k] + 3) & ~3;
        m.w_off[k] = woff;
        m.p_off[k] = poff;
        woff += F * m.M[k] * m.ustr[k];
        poff += units[k];
    }
    m.w_total = woff;        // multiple of 4 because ustr is
    m.u_total = poff;
    h.ok = true;
    return h;
}

template <typename Fn>
static int cin_dispatch(int D, int maxu, Fn&& fn) {
    if (maxu <= 16) {
        if (D == 8) return fn(std::integral_constant<int, 8>{}, std::integral_constant<int, 16>{});
        if (D == 16) return fn(std::integral_constant<int, 16>{}, std::integral_constant<int, 16>{});
        if (D == 32) return fn(std::integral_constant<int, 32>{}, std::integral_constant<int, 16>{});
    } else {
        if (D == 8) return fn(std::integral_constant<int, 8>{}, std::integral_constant<int, 32>{});
        if (D == 16) return fn(std::integral_constant<int, 16>{}, std::integral_constant<int, 32>{});
        if (D == 32) return fn(std::integral_constant<int, 32>{}, std::integral_constant<int, 32>{});
    }
    return RPB_ERR_UNSUPPORTED;
}

static int pack_weights(const CinMeta& meta, const CinPtrs& ptrs, float** Wt_out, float** bias_out, cudaStream_t st) {
    const size_t bytes = ((size_t)meta.w_total + meta.u_total + 4) * sizeof(float);
    int werr = 0;
    float* ws = static_cast<float*>(workspace(1, bytes, &werr));
    if (ws == nullptr) return werr;
    cin_pack_kernel<<<dim3(8, meta.L), 256, 0, st>>>(ptrs, meta, ws, ws + meta.w_total);
    *Wt_out = ws;
    *bias_out = ws + meta.w_total;
    return (int)cudaGetLastError();
}

}  // namespace rpb

using namespace rpb;

namespace rpb {
bool cin_tc_shape_ok(int F, int D, int L, const int* units);                                                   // cin_tc.cu
int cin_layer_fwd_tc(int F, int M, int U, int D, const float* W, const float* bias, const float* x0, long long ld0, const float* xk,
                     long long ldk, float* xout, long long ldo, float* pooled, long long ldp, int B, cudaStream_t st);
int cin_layer_bwd_tc_c(int F, int M, const float* W, const float* x0, long long ld0, const float* xk, long long ldk, const float* dpooled,
                       long long lddp, const float* gx, long long ldgx, float* gout, long long ldgo, float* db, float* dxk, long long lddxk,
                       float* de, long long ldde, int de_accumulate, int B, cudaStream_t st);
int cin_layer_wgrad_tc(int F, int M, const float* x0, long long ld0, const float* xk, long long ldk, const float* g, long long ldg,
                       float* dW, int B, cudaStream_t st);
}

static int cin_fwd_impl(const float* e, int64_t lde, int B, int F, int D, int L, const int32_t* units,
                        const float* const* W, const float* const* bias, float* pooled, int64_t ldp, float* xsave, int64_t ldx, void* stream) {
    if (e == nullptr || pooled == nullptr || W == nullptr || bias == nullptr || B <= 0) return RPB_ERR_BAD_ARG;
    CinHost h = cin_meta(F, L, units);
    if (!h.ok) return RPB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (cin_tc_shape_ok(F, D, L, units)) {
        // tensor-core path (cin_tc.cu): one launch per layer; X_{k+1} [B, U, D] ping-pongs through workspace slot 3, or goes to
        // the caller's xsave ([B, ldx], layer k at column p_off[k] * D: the layout the backward kernels read) when training
        int werr = 0;
        const size_t per = (size_t)B * units[0] * D;
        float* xs = (L > 1 && xsave == nullptr) ? static_cast<float*>(workspace(3, 2 * per * sizeof(float), &werr)) : nullptr;
        if (L > 1 && xsave == nullptr && xs == nullptr) return werr;
        const float* xk = e; long long ldk = lde;
        for (int k = 0; k < L; ++k) {
            float* xout = nullptr; long long ldo = (long long)units[k] * D;
            if (k + 1 < L) {
                if (xsave != nullptr) { xout = xsave + (size_t)h.meta.p_off[k] * D; ldo = ldx; }
                else xout = xs + (size_t)(k & 1) * per;
            }
            const int rc = cin_layer_fwd_tc(F, h.meta.M[k], units[k], D, W[k], bias[k], e, lde, xk, ldk, xout, ldo,
                                            pooled + h.meta.p_off[k], ldp, B, st);
            if (rc != 0) return rc;
            xk = xout; ldk = ldo;
        }
        return 0;
    }
    if (xsave != nullptr) return RPB_ERR_UNSUPPORTED;                      // only the tensor-core path keeps X_k
    int maxu = 0;
    for (int k = 0; k < L; ++k) maxu = max(maxu, units[k]);
    float *Wt = nullptr, *bcat = nullptr;
    CinPtrs ptrs{};
    for (int k = 0; k < L; ++k) { ptrs.W[k] = W[k]; ptrs.bias[k] = bias[k]; }
    int rc = pack_weights(h.meta, ptrs, &Wt, &bcat, st);
    if (rc != 0) return rc;
    rc = cin_dispatch(D, maxu, [&](auto dt, auto ut) -> int {
        constexpr int DD = decltype(dt)::value, MU = decltype(ut)::value;
        constexpr int SPC = 256 / DD;
        const size_t smem = ((size_t)h.meta.w_total + ((h.meta.u_total + 3) & ~3) + (size_t)SPC * F * DD) * sizeof(float);
        if (smem > 220 * 1024) return RPB_ERR_UNSUPPORTED;
        cudaError_t ee = cudaFuncSetAttribute(cin_fwd_kernel<DD, MU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ee != cudaSuccess) return (int)ee;
        const int groups = ceil_div(B, SPC);
        const int per_sm = max(1, (int)((220 * 1024) / smem));
        const int grid = min(groups, 148 * min(per_sm, 2));
        cin_fwd_kernel<DD, MU><<<grid, 256, smem, st>>>(e, lde, B, h.meta, Wt, bcat, pooled, (int)ldp);
        return (int)cudaGetLastError();
    });
    return rc;
}

RPB_API int rpb_cin_fwd(const float* e, int64_t lde, int B, int F, int D, int L, const int32_t* units,
                        const float* const* W, const float* const* bias, float* pooled, int64_t ldp, void* stream) {
    return cin_fwd_impl(e, lde, B, F, D, L, units, W, bias, pooled, ldp, nullptr, 0, stream);
}

RPB_API int rpb_cin_fwd_save(const float* e, int64_t lde, int B, int F, int D, int L, const int32_t* units,
                             const float* const* W, const float* const* bias, float* pooled, int64_t ldp,
                             float* xsave, int64_t ldx, void* stream) {
    if (xsave == nullptr || L < 2) return RPB_ERR_BAD_ARG;
    int64_t need = 0;
    for (int k = 0; k + 1 < L; ++k) need += (int64_t)units[k] * D;
    if (ldx < need || (ldx & 3) || (reinterpret_cast<uintptr_t>(xsave) & 15u)) return RPB_ERR_BAD_ARG;
    return cin_fwd_impl(e, lde, B, F, D, L, units, W, bias, pooled, ldp, xsave, ldx, stream);
}

static int cin_bwd_impl(const float* e, int64_t lde, int B, int F, int D, int L, const int32_t* units,
                        const float* const* W, const float* const* bias, const float* dpooled, int64_t lddp,
                        float* de, int64_t ldde, int accumulate, float* const* dW, float* const* db,
                        const float* xsaved, int64_t ldx, void* stream) {
    if (e == nullptr || dpooled == nullptr || de == nullptr || W == nullptr || dW == nullptr || B <= 0) return RPB_ERR_BAD_ARG;
    CinHost h = cin_meta(F, L, units);
    if (!h.ok) return RPB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int maxu = 0, maxfm = 0;
    for (int k = 0; k < L; ++k) { maxu = max(maxu, units[k]); maxfm = max(maxfm, F * h.meta.M[k]); }
    float *Wt = nullptr, *bcat = nullptr;
    CinPtrs ptrs{};
    for (int k = 0; k < L; ++k) {
        ptrs.W[k] = W[k]; ptrs.bias[k] = bias ? bias[k] : nullptr;
        ptrs.dW[k] = dW[k]; ptrs.db[k] = db ? db[k] : nullptr;
    }
    int rc = pack_weights(h.meta, ptrs, &Wt, &bcat, st);
    if (rc != 0) return rc;
    const size_t per = (size_t)B * h.meta.u_total * D;
    int werr = 0;
    float* spill = static_cast<float*>(workspace(2, (2 * per) * sizeof(float), &werr));
    if (spill == nullptr) return werr;
    float* Gout = spill;
    float* Xout = spill + per;
    if (cin_tc_shape_ok(F, D, L, units)) {
        // tensor-core path (cin_tc.cu): recompute X_1 .. X_{L-1} (one launch per layer), then per layer, top down, the dZ GEMM
        // with the gradient contraction in its epilogue; G_k and X_k are spilled in the layout the weight-gradient kernel reads
        const long long ldu = (long long)h.meta.u_total * D;
        float* dxb = static_cast<float*>(workspace(3, 2 * (size_t)B * units[0] * D * sizeof(float), &werr));
        if (dxb == nullptr) return werr;
        const size_t dper = (size_t)B * units[0] * D;
        const float* Xs = Xout; long long ldxs = ldu;                      // X_1 .. X_{L-1}: recomputed here, or kept by the forward
        if (xsaved != nullptr) { Xs = xsaved; ldxs = ldx; }
        else {
            for (int k = 0; k + 1 < L; ++k) {
                rc = cin_layer_fwd_tc(F, h.meta.M[k], units[k], D, W[k], bias ? bias[k] : nullptr, e, lde,
                                      k == 0 ? e : Xout + (size_t)h.meta.p_off[k - 1] * D, k == 0 ? lde : ldu,
                                      Xout + (size_t)h.meta.p_off[k] * D, ldu, nullptr, 0, B, st);
                if (rc != 0) return rc;
            }
        }
        for (int k = L - 1; k >= 0; --k) {
            rc = cin_layer_bwd_tc_c(F, h.meta.M[k], W[k], e, lde, k == 0 ? e : Xs + (size_t)h.meta.p_off[k - 1] * D, k == 0 ? lde : ldxs,
                                    dpooled + h.meta.p_off[k], lddp, k == L - 1 ? nullptr : dxb + (size_t)((k + 1) & 1) * dper,
                                    (long long)units[k] * D, Gout + (size_t)h.meta.p_off[k] * D, ldu, db ? db[k] : nullptr,
                                    k > 0 ? dxb + (size_t)(k & 1) * dper : nullptr, (long long)h.meta.M[k] * D, de, ldde,
                                    k == L - 1 ? accumulate : 1, B, st);
            if (rc != 0) return rc;
        }
        // weight gradients: dW_k = P^T . X_k over the rows (b, d), P = G_k x X_0 formed in registers (cin_wgrad_tc_kernel)
        for (int k = 0; k < L; ++k) {
            rc = cin_layer_wgrad_tc(F, h.meta.M[k], e, lde, k == 0 ? e : Xs + (size_t)h.meta.p_off[k - 1] * D, k == 0 ? lde : ldxs,
                                    Gout + (size_t)h.meta.p_off[k] * D, ldu, dW[k], B, st);
            if (rc != 0) return rc;
        }
        return 0;
    }
    if (xsaved != nullptr) return RPB_ERR_UNSUPPORTED;
    rc = cin_dispatch(D, maxu, [&](auto dt, auto ut) -> int {
        constexpr int DD = decltype(dt)::value, MU = decltype(ut)::value;
        constexpr int SPC = 256 / DD;
        {
        const size_t smem = ((size_t)h.meta.w_total + ((h.meta.u_total + 3) & ~3) + (size_t)SPC * F * DD * 2 +
                             (size_t)SPC * h.meta.u_total * DD) * sizeof(float);
        if (smem > 220 * 1024) return RPB_ERR_UNSUPPORTED;
        cudaError_t ee = cudaFuncSetAttribute(cin_bwd_sample_kernel<DD, MU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ee != cudaSuccess) return (int)ee;
        const int grid = min(ceil_div(B, SPC), 148);
        cin_bwd_sample_kernel<DD, MU><<<grid, 256, smem, st>>>(e, lde, B, h.meta, Wt, bcat, dpooled, (int)lddp, de, ldde,
                                                               accumulate, Gout, Xout);
        int r2 = (int)cudaGetLastError();
        if (r2 != 0) return r2;
        }
        // weight gradients
        const int slabs = max(1, min(ceil_div(B, 8), (148 * 2) / L));
        int slab = ceil_div(B, slabs);
        slab = ((slab + 7) / 8) * 8;
        const int nslabs = ceil_div(B, slab);
        const size_t smem2 = (size_t)8 * DD * (F + CIN_MAXM + MU) * sizeof(float);
        auto launch = [&](auto njt) -> int {
            constexpr int NJ = decltype(njt)::value;
            cudaError_t e3 = cudaFuncSetAttribute(cin_wgrad_kernel<DD, MU, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
            if (e3 != cudaSuccess) return (int)e3;
            cin_wgrad_kernel<DD, MU, NJ><<<dim3(nslabs, L), 256, smem2, st>>>(e, lde, B, h.meta, Gout, Xout, ptrs, slab);
            return (int)cudaGetLastError();
        };
        if (maxfm <= 256) return launch(std::integral_constant<int, 1>{});
        if (maxfm <= 512) return launch(std::integral_constant<int, 2>{});
        if (maxfm <= 768) return launch(std::integral_constant<int, 3>{});
        return launch(std::integral_constant<int, 4>{});
    });
    return rc;
}

RPB_API int rpb_cin_bwd(const float* e, int64_t lde, int B, int F, int D, int L, const int32_t* units,
                        const float* const* W, const float* const* bias, const float* dpooled, int64_t lddp,
                        float* de, int64_t ldde, int accumulate, float* const* dW, float* const* db, void* stream) {
    return cin_bwd_impl(e, lde, B, F, D, L, units, W, bias, dpooled, lddp, de, ldde, accumulate, dW, db, nullptr, 0, stream);
}

RPB_API int rpb_cin_bwd_saved(const float* e, int64_t lde, int B, int F, int D, int L, const int32_t* units,
                              const float* const* W, const float* const* bias, const float* dpooled, int64_t lddp,
                              float* de, int64_t ldde, int accumulate, float* const* dW, float* const* db,
                              const float* xsaved, int64_t ldx, void* stream) {
    if (xsaved == nullptr) return RPB_ERR_BAD_ARG;
    return cin_bwd_impl(e, lde, B, F, D, L, units, W, bias, dpooled, lddp, de, ldde, accumulate, dW, db, xsaved, ldx, stream);
}
