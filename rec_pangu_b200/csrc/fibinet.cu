// FiBiNet interaction: SENET re-weighting + bilinear 'field_interaction' on raw and re-weighted embeddings, fused
// (reference: models/layers/interaction.py:55-81 BilinearInteractionLayer, :238-251 SENET_Layer, ranking/fibinet.py:59-66).
//
// Reference: ~1000 tiny launches (325 x Linear(16,16) + mul, twice) and two cat copies.  Here one kernel per direction:
// D lanes own a sample (lane r = output component r).  Since V = E * A[f] (a per-field scalar),
//   bilinear(V)[p] = (W_p (A_i v_i)) o (A_j v_j) = A_i A_j * bilinear(E)[p],
// so the second bilinear pass costs one multiply per output.  The kernel writes the MLP input row directly:
//   comb[b] = [ bilinear(E) (P*D) | bilinear(V) (P*D) | dense (Nd) | 0-pad ],  P = F(F-1)/2, pairs in combinations order.
// Backward: (A) per-sample kernel — dE, SENET backward (dW1/dW2 accumulated per CTA), spills du_p = dL/d(W_p v_i);
//           (B) dW_p = sum_b du_p (x) v_i as a register-tiled batch reduction, 4x4 outputs per thread.
#include <type_traits>

#include "common.cuh"

namespace rpb {

constexpr int FB_MAXF = 32;

__device__ __forceinline__ void pair_from_index(int p, int F, int& i, int& j) {   // unused on the hot path (loops carry i,j)
    i = 0;
    int rem = p;
    while (rem >= F - 1 - i) { rem -= F - 1 - i; ++i; }
    j = i + 1 + rem;
}

// smem per CTA: Es [SPC][F][D] | As [SPC][F] | Zs [SPC][F] | A1s [SPC][R]
template <int D>
__global__ void __launch_bounds__(256)
fibinet_fwd_kernel(const float* __restrict__ x, long long ldx, int B, int F, int Nd, const float* __restrict__ W1, int R,
                   const float* __restrict__ W2, const float* __restrict__ Wb, float* __restrict__ comb, long long ldc,
                   float* __restrict__ Aout) {
    constexpr int SPC = 256 / D;
    extern __shared__ __align__(16) float sm[];
    float* Es = sm;
    float* As = Es + SPC * F * D;
    float* Zs = As + SPC * F;
    float* A1s = Zs + SPC * F;
    const int s = threadIdx.x / D, r = threadIdx.x % D;
    const int P = F * (F - 1) / 2;
    for (long long grp = blockIdx.x; grp * SPC < B; grp += gridDim.x) {
        const long long b_raw = grp * SPC + s;
        const bool valid = b_raw < B;
        const long long b = valid ? b_raw : B - 1;
        const float* xr = x + (size_t)b * ldx;
        float* Er = Es + (size_t)s * F * D;
        __syncwarp();
        for (int f = 0; f < F; ++f) Er[f * D + r] = __ldg(xr + f * D + r);
        __syncwarp();
        for (int f = r; f < F; f += D) {
            float z = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) z += Er[f * D + d];
            Zs[s * F + f] = z / (float)D;                       // torch.mean over the embedding dim
        }
        __syncwarp();
        for (int j = r; j < R; j += D) {
            float a = 0.f;
            for (int f = 0; f < F; ++f) a = fmaf(__ldg(W1 + j * F + f), Zs[s * F + f], a);
            A1s[s * R + j] = fmaxf(a, 0.f);
        }
        __syncwarp();
        for (int f = r; f < F; f += D) {
            float a = 0.f;
            for (int j = 0; j < R; ++j) a = fmaf(__ldg(W2 + f * R + j), A1s[s * R + j], a);
            a = fmaxf(a, 0.f);
            As[s * F + f] = a;
            if (valid && Aout != nullptr) Aout[(size_t)b * F + f] = a;
        }
        __syncwarp();
        float* cr = comb + (size_t)b * ldc;
        int p = 0;
        for (int i = 0; i < F - 1; ++i) {
            float vi[D];
#pragma unroll
            for (int c4 = 0; c4 < D / 4; ++c4) {
                const float4 t = *reinterpret_cast<const float4*>(Er + i * D + c4 * 4);
                vi[c4 * 4] = t.x; vi[c4 * 4 + 1] = t.y; vi[c4 * 4 + 2] = t.z; vi[c4 * 4 + 3] = t.w;
            }
            const float ai = As[s * F + i];
            for (int j = i + 1; j < F; ++j, ++p) {
                const float4* w4 = reinterpret_cast<const float4*>(Wb + ((size_t)p * D + r) * D);
                float u = 0.f;
#pragma unroll
                for (int c4 = 0; c4 < D / 4; ++c4) {
                    const float4 w = __ldg(w4 + c4);
                    u = fmaf(w.x, vi[c4 * 4], u); u = fmaf(w.y, vi[c4 * 4 + 1], u);
                    u = fmaf(w.z, vi[c4 * 4 + 2], u); u = fmaf(w.w, vi[c4 * 4 + 3], u);
                }
                const float oe = u * Er[j * D + r];
                if (valid) {
                    cr[(size_t)p * D + r] = oe;
                    cr[(size_t)(P + p) * D + r] = oe * ai * As[s * F + j];
                }
            }
        }
        if (valid) {
            const int base = 2 * P * D;
            for (int j = r; j < Nd; j += D) cr[base + j] = __ldg(xr + F * D + j);
            for (long long j = base + Nd + r; j < ldc; j += D) cr[j] = 0.f;
        }
    }
}

// Wt[p][c][r] = Wb[p][r][c]
__global__ void fibinet_transpose_w_kernel(const float* __restrict__ Wb, float* __restrict__ Wt, int P, int D) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)P * D * D) return;
    const int p = (int)(i / (D * D)), rem = (int)(i % (D * D)), c = rem / D, r = rem % D;
    Wt[i] = __ldg(Wb + ((size_t)p * D + r) * D + c);
}

// smem: Es | dEs [SPC][F][D] | As, Zs, dAs [SPC][F] | A1s [SPC][R] | dus [SPC][D] | accW1 [R*F] | accW2 [F*R]
template <int D>
__global__ void __launch_bounds__(256)
fibinet_bwd_sample_kernel(const float* __restrict__ x, long long ldx, int B, int F, const float* __restrict__ W1, int R,
                          const float* __restrict__ W2, const float* __restrict__ Wb, const float* __restrict__ Wt,
                          const float* __restrict__ Ain, const float* __restrict__ dcomb, long long lddc,
                          float* __restrict__ dx, long long lddx, float* __restrict__ DU, float* __restrict__ dW1,
                          float* __restrict__ dW2) {
    constexpr int SPC = 256 / D;
    extern __shared__ __align__(16) float sm[];
    float* Es = sm;
    float* dEs = Es + SPC * F * D;
    float* As = dEs + SPC * F * D;
    float* Zs = As + SPC * F;
    float* dAs = Zs + SPC * F;
    float* A1s = dAs + SPC * F;
    float* dus = A1s + SPC * R;
    float* accW1 = dus + SPC * D;
    float* accW2 = accW1 + R * F;
    const int s = threadIdx.x / D, r = threadIdx.x % D;
    const int P = F * (F - 1) / 2;
    for (int i = threadIdx.x; i < 2 * R * F; i += 256) accW1[i] = 0.f;
    __syncthreads();
    for (long long grp = blockIdx.x; grp * SPC < B; grp += gridDim.x) {
        const long long b_raw = grp * SPC + s;
        const bool valid = b_raw < B;
        const long long b = valid ? b_raw : B - 1;
        const float* xr = x + (size_t)b * ldx;
        float* Er = Es + (size_t)s * F * D;
        float* dEr = dEs + (size_t)s * F * D;
        __syncwarp();
        for (int f = 0; f < F; ++f) { Er[f * D + r] = __ldg(xr + f * D + r); dEr[f * D + r] = 0.f; }
        __syncwarp();
        for (int f = r; f < F; f += D) {
            float z = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) z += Er[f * D + d];
            Zs[s * F + f] = z / (float)D;
            As[s * F + f] = __ldg(Ain + (size_t)b * F + f);
            dAs[s * F + f] = 0.f;
        }
        __syncwarp();
        for (int j = r; j < R; j += D) {
            float a = 0.f;
            for (int f = 0; f < F; ++f) a = fmaf(__ldg(W1 + j * F + f), Zs[s * F + f], a);
            A1s[s * R + j] = fmaxf(a, 0.f);
        }
        __syncwarp();
        const float* gr = dcomb + (size_t)b * lddc;
        int p = 0;
        for (int i = 0; i < F - 1; ++i) {
            float vi[D];
#pragma unroll
            for (int c4 = 0; c4 < D / 4; ++c4) {
                const float4 t = *reinterpret_cast<const float4*>(Er + i * D + c4 * 4);
                vi[c4 * 4] = t.x; vi[c4 * 4 + 1] = t.y; vi[c4 * 4 + 2] = t.z; vi[c4 * 4 + 3] = t.w;
            }
            const float ai = As[s * F + i];
            float dei = 0.f;                                        // accumulates dE[i][c = r] over j
            for (int j = i + 1; j < F; ++j, ++p) {
                const float4* w4 = reinterpret_cast<const float4*>(Wb + ((size_t)p * D + r) * D);
                float u = 0.f;
#pragma unroll
                for (int c4 = 0; c4 < D / 4; ++c4) {
                    const float4 w = __ldg(w4 + c4);
                    u = fmaf(w.x, vi[c4 * 4], u); u = fmaf(w.y, vi[c4 * 4 + 1], u);
                    u = fmaf(w.z, vi[c4 * 4 + 2], u); u = fmaf(w.w, vi[c4 * 4 + 3], u);
                }
                const float ej = Er[j * D + r], aj = As[s * F + j];
                const float gE = __ldg(gr + (size_t)p * D + r), gV = __ldg(gr + (size_t)(P + p) * D + r);
                const float G = fmaf(ai * aj, gV, gE);              // dL/d(bilinear(E)[p][r])
                const float t = group_sum<D>(gV * u * ej);          // sum_r gV * out_E
                if (r == 0) { dAs[s * F + i] += aj * t; dAs[s * F + j] += ai * t; }
                dEr[j * D + r] = fmaf(G, u, dEr[j * D + r]);
                const float du = G * ej;
                if (valid) DU[((size_t)b * P + p) * D + r] = du;
                __syncwarp();
                dus[s * D + r] = du;
                __syncwarp();
                // dE[i][c] += sum_r' W[r'][c] du[r']  (c = this lane), transposed weights give a contiguous row
                const float4* t4 = reinterpret_cast<const float4*>(Wt + ((size_t)p * D + r) * D);
                float acc = 0.f;
#pragma unroll
                for (int q4 = 0; q4 < D / 4; ++q4) {
                    const float4 w = __ldg(t4 + q4);
                    const float4 dv = *reinterpret_cast<const float4*>(dus + s * D + q4 * 4);
                    acc = fmaf(w.x, dv.x, acc); acc = fmaf(w.y, dv.y, acc);
                    acc = fmaf(w.z, dv.z, acc); acc = fmaf(w.w, dv.w, acc);
                }
                dei += acc;
            }
            dEr[i * D + r] += dei;
        }
        __syncwarp();
        // ---- SENET backward (per sample), parameter grads accumulated per CTA in shared memory
        for (int f = r; f < F; f += D) {
            const float dap = (As[s * F + f] > 0.f) ? dAs[s * F + f] : 0.f;
            dAs[s * F + f] = dap;                                    // now holds dL/d(pre-activation of A)
            if (valid) for (int j = 0; j < R; ++j) atomicAdd(&accW2[f * R + j], dap * A1s[s * R + j]);
        }
        __syncwarp();
        float* dA1 = dus + s * D;                                    // reuse: dA1pre[j], j < R (R <= D is checked on host)
        for (int j = r; j < R; j += D) {
            float a = 0.f;
            for (int f = 0; f < F; ++f) a = fmaf(__ldg(W2 + f * R + j), dAs[s * F + f], a);
            dA1[j] = (A1s[s * R + j] > 0.f) ? a : 0.f;
        }
        __syncwarp();
        for (int f = r; f < F; f += D) {
            float dz = 0.f;
            for (int j = 0; j < R; ++j) {
                dz = fmaf(__ldg(W1 + j * F + f), dA1[j], dz);
                if (valid) atomicAdd(&accW1[j * F + f], dA1[j] * Zs[s * F + f]);
            }
            Zs[s * F + f] = dz / (float)D;                           // now holds dL/dE[f][d] contribution of the mean
        }
        __syncwarp();
        if (valid) {
            float* dr = dx + (size_t)b * lddx;
            for (int f = 0; f < F; ++f) dr[f * D + r] = dEr[f * D + r] + Zs[s * F + f];
            for (long long j = F * D + r; j < lddx; j += D) dr[j] = 0.f;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < R * F; i += 256) {
        if (dW1 != nullptr) red_add_f1(dW1 + i, accW1[i]);
        if (dW2 != nullptr) red_add_f1(dW2 + i, accW2[i]);
    }
}

// dWb[p][r][c] += sum_{b in slab} DU[b][p][r] * E[b][i(p)][c].  grid (ceil(P / PPC), slabs), 256 threads:
// thread -> (pair q of the CTA's PPC pairs, 4x4 block (rb, cb)).
template <int D>
__global__ void __launch_bounds__(256)
fibinet_wgrad_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ DU, int B, int F,
                     float* __restrict__ dWb, int slab) {
    constexpr int NB = D / 4, TP = NB * NB, PPC = 256 / TP, ST = 8;
    extern __shared__ __align__(16) float sm[];
    float* Es = sm;                        // [ST][F*D]
    float* Us = Es + ST * F * D;           // [ST][PPC][D]
    const int P = F * (F - 1) / 2;
    const int q = threadIdx.x / TP, blk = threadIdx.x % TP, rb = blk / NB, cb = blk % NB;
    const int p0 = blockIdx.x * PPC, p = p0 + q;
    int pi = 0, pj = 0;
    if (p < P) pair_from_index(p, F, pi, pj);
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
    const long long bbeg = (long long)blockIdx.y * slab, bend = min((long long)B, bbeg + slab);
    const int npairs = min(PPC, P - p0);
    for (long long b0 = bbeg; b0 < bend; b0 += ST) {
        const int ns = (int)min((long long)ST, bend - b0);
        __syncthreads();
        for (int i = threadIdx.x; i < ns * F * D; i += 256) {
            const int s = i / (F * D), rem = i % (F * D);
            Es[s * F * D + rem] = __ldg(x + (size_t)(b0 + s) * ldx + rem);
        }
        for (int i = threadIdx.x; i < ns * npairs * D; i += 256) {
            const int s = i / (npairs * D), rem = i % (npairs * D);
            Us[(s * PPC) * D + rem] = __ldg(DU + ((size_t)(b0 + s) * P + p0) * D + rem);
        }
        __syncthreads();
        if (p < P) {
            for (int s = 0; s < ns; ++s) {
                const float4 du = *reinterpret_cast<const float4*>(Us + (s * PPC + q) * D + rb * 4);
                const float4 e = *reinterpret_cast<const float4*>(Es + s * F * D + pi * D + cb * 4);
                const float dv[4] = {du.x, du.y, du.z, du.w}, ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(dv[a], ev[c], acc[a][c]);
            }
        }
    }
    if (p < P) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                red_add_f1(dWb + ((size_t)p * D + rb * 4 + a) * D + cb * 4 + c, acc[a][c]);
    }
}

template <typename Fn>
static int fb_dispatch(int D, Fn&& fn) {
    if (D == 8) return fn(std::integral_constant<int, 8>{});
    if (D == 16) return fn(std::integral_constant<int, 16>{});
    if (D == 32) return fn(std::integral_constant<int, 32>{});
    return RPB_ERR_UNSUPPORTED;
}

}  // namespace rpb

using namespace rpb;

RPB_API int rpb_fibinet_fwd(const float* x, int64_t ldx, int B, int F, int D, int Nd, const float* W1, int R,
                            const float* W2, const float* Wb, float* comb, int64_t ldc, float* A, void* stream) {
    if (x == nullptr || W1 == nullptr || W2 == nullptr || Wb == nullptr || comb == nullptr || B <= 0) return RPB_ERR_BAD_ARG;
    if (F < 2 || F > FB_MAXF || R < 1 || R > D || ldc < (int64_t)F * (F - 1) * D + Nd) return RPB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(Wb) & 15u) != 0) return RPB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return fb_dispatch(D, [&](auto dt) -> int {
        constexpr int DD = decltype(dt)::value;
        constexpr int SPC = 256 / DD;
        const size_t smem = ((size_t)SPC * F * DD + 2 * SPC * F + SPC * R) * sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(fibinet_fwd_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        const int grid = min(ceil_div(B, SPC), 148 * 4);
        fibinet_fwd_kernel<DD><<<grid, 256, smem, st>>>(x, ldx, B, F, Nd, W1, R, W2, Wb, comb, ldc, A);
        return (int)cudaGetLastError();
    });
}

RPB_API int rpb_fibinet_bwd(const float* x, int64_t ldx, int B, int F, int D, const float* W1, int R, const float* W2,
                            const float* Wb, const float* A, const float* dcomb, int64_t lddc, float* dx, int64_t lddx,
                            float* dW1, float* dW2, float* dWb, void* stream) {
    if (x == nullptr || A == nullptr || dcomb == nullptr || dx == nullptr || Wb == nullptr || B <= 0) return RPB_ERR_BAD_ARG;
    if (F < 2 || F > FB_MAXF || R < 1 || R > D) return RPB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int P = F * (F - 1) / 2;
    int werr = 0;
    float* Wt = static_cast<float*>(workspace(4, (size_t)P * D * D * sizeof(float), &werr));
    if (Wt == nullptr) return werr;
    float* DU = static_cast<float*>(workspace(5, (size_t)B * P * D * sizeof(float), &werr));
    if (DU == nullptr) return werr;
    fibinet_transpose_w_kernel<<<ceil_div((long long)P * D * D, 256), 256, 0, st>>>(Wb, Wt, P, D);
    return fb_dispatch(D, [&](auto dt) -> int {
        constexpr int DD = decltype(dt)::value;
        constexpr int SPC = 256 / DD;
        const size_t smem = ((size_t)2 * SPC * F * DD + 3 * SPC * F + SPC * R + SPC * DD + 2 * R * F) * sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(fibinet_bwd_sample_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        const int grid = min(ceil_div(B, SPC), 148 * 2);
        fibinet_bwd_sample_kernel<DD><<<grid, 256, smem, st>>>(x, ldx, B, F, W1, R, W2, Wb, Wt, A, dcomb, lddc, dx, lddx, DU,
                                                               dW1, dW2);
        int rc = (int)cudaGetLastError();
        if (rc != 0 || dWb == nullptr) return rc;
        constexpr int PPC = 256 / ((DD / 4) * (DD / 4));
        const int ptiles = ceil_div(P, PPC);
        int slabs = max(1, min(ceil_div(B, 64), (148 * 4) / ptiles + 1));
        int slab = ceil_div(B, slabs);
        slab = ((slab + 7) / 8) * 8;
        slabs = ceil_div(B, slab);
        const size_t smem2 = ((size_t)8 * F * DD + 8 * PPC * DD) * sizeof(float);
        e = cudaFuncSetAttribute(fibinet_wgrad_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) return (int)e;
        fibinet_wgrad_kernel<DD><<<dim3(ptiles, slabs), 256, smem2, st>>>(x, ldx, DU, B, F, dWb, slab);
        return (int)cudaGetLastError();
    });
}
