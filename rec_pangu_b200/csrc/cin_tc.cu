// xDeepFM Compressed Interaction Network on tcgen05 (reference: models/layers/interaction.py:144-171).
//
//     X_{k+1}[b,u,d] = sum_{h,m} W_k[u, h*M + m] * X0[b,h,d] * Xk[b,m,d] + bias[u]        pooled[b,u] = sum_d X_{k+1}[b,u,d]
// rows = (sample, embedding dim) pairs (1 M rows at config 3).  The fp32 CUDA-core kernels of cin.cu form the outer product
// Z = X0 x Xk in registers and pay 4 LDS.128 of weights per 17 FP instructions (3.6 ms forward, 14 ms backward at config 3,
// ~1 % of the HBM roofline, 0 % tensor pipe: profiles/r01_kernels.md).  Here every part is a tensor-core GEMM whose row operand
// is written by the warps that own the rows (thread = row (b,d) = tensor-memory lane, TS-mode MMA, 3xTF32 for the 1e-4 parity
// bound) and whose other contraction is done by the epilogue straight out of the accumulator, every index a compile-time
// constant:
//   forward   cin_fwd2_tc_kernel   T[(b,d),(h,u)] = sum_m Xk[m] W_k[u,h,m]     (K = M),  epilogue  X_{k+1}[u] = sum_h X0[h] T[(h,u)]
//   backward  cin_bwd_tc_kernel    dZ[(b,d),(h,m)] = sum_u G[u] W_k[u,h,m]     (K = 16), epilogue  dXk[m] += dZ X0[h], dX0[h] += dZ Xk[m]
//   weights   cin_wgrad_tc_kernel  dW_k[(u,h),m]  = sum_(b,d) (G[u] X0[h]) Xk[m]   (K = rows; P = G x X0 formed element-wise)
// A first forward that fed the outer product itself to the tensor core (676 splits per row) was bound by its split warps
// (562 us per layer against 250 us for the form above); the steps from 18.6 ms to 2.54 ms per xDeepFM step are listed in
// profiles/r02_cin_ncu.md.  All kernels are persistent (one CTA per SM) and walk tiles of 128 rows = 8 samples.
#include "tc_ptx.cuh"

namespace rpb {

constexpr int CT_D = 16;                 // embedding dim these kernels are built for (one 64-byte row per field)
constexpr int CT_U = 16;                 // units per layer
// registers are re-dealt between the warpgroups of the 384-thread kernels (168 per thread at launch)
#define RPB_REG_DEC(n) asm volatile("setmaxnreg.dec.sync.aligned.u32 " #n ";")
#define RPB_REG_INC(n) asm volatile("setmaxnreg.inc.sync.aligned.u32 " #n ";")

// Pre-split K-major weight operands with PERMUTED rows, so that consecutive accumulator columns feed different register
// accumulators in the epilogues (a chain of dependent FMAs per output otherwise: profiles/r02_cin_ncu.md).
//   mode 0 (forward):  row h*16 + u            = W[u, h*M + 0..M-1]                      (K = M columns, zero padded to 32)
//   mode 1 (backward): row s*F + h, m=(h+s)%M  = W[0..15, h*M + m]  (W^T, K = 16 columns, zero padded to 32)
__global__ void __launch_bounds__(256)
cin_pack_operand_kernel(const float* __restrict__ W, int F, int M, int mode, int rows_valid, int Rp, float* __restrict__ hi, float* __restrict__ lo) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Rp * TC_BLOCK_K) return;
    const int r = t / TC_BLOCK_K, c = t - r * TC_BLOCK_K;
    float v = 0.f;
    if (r < rows_valid) {
        if (mode == 0) {
            const int h = r / CT_U, u = r - h * CT_U;
            if (c < M) v = __ldg(W + (size_t)u * F * M + h * M + c);
        } else {
            const int sft = r / F, h = r - sft * F, m = (h + sft) % M;
            if (c < CT_U) v = __ldg(W + (size_t)c * F * M + h * M + m);
        }
    }
    const float hv = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[t] = hv;
    lo[t] = v - hv;
}

static int cin_prepare_operand(const float* W, int F, int M, int mode, int Rp, int box_rows, int slot, CUtensorMap* tm_hi, CUtensorMap* tm_lo,
                               cudaStream_t st) {
    int werr = 0;
    float* ws = static_cast<float*>(workspace(slot, (size_t)2 * Rp * TC_BLOCK_K * sizeof(float), &werr));
    if (ws == nullptr) return werr;
    float* hi = ws;
    float* lo = ws + (size_t)Rp * TC_BLOCK_K;
    const int rows_valid = mode == 0 ? CT_U * F : F * M;
    cin_pack_operand_kernel<<<ceil_div(Rp * TC_BLOCK_K, 256), 256, 0, st>>>(W, F, M, mode, rows_valid, Rp, hi, lo);
    int rc = tc_make_map2d(tm_hi, hi, Rp, TC_BLOCK_K, TC_BLOCK_K, TC_BLOCK_K, box_rows, 128);
    if (rc == 0) rc = tc_make_map2d(tm_lo, lo, Rp, TC_BLOCK_K, TC_BLOCK_K, TC_BLOCK_K, box_rows, 128);
    return rc;
}

struct CinTcParams {
    const float* x0; long long ld0;      // X0[b,h,d] = x0[b * ld0 + h * 16 + d]            (feature row of the gather)
    const float* xk; long long ldk;      // Xk[b,m,d] = xk[b * ldk + m * 16 + d]            (layer 0: = x0)
    const float* bias;                   // [16] or null
    float* xout; long long ldo;          // X_{k+1}[b,u,d] = xout[b * ldo + u * 16 + d], or null
    float* pooled; long long ldp;        // pooled[b * ldp + u] = sum_d X_{k+1}[b,u,d], or null
    int B, m_tiles;
};

// ---------------------------------------------------------------------------------------------------------------------
// Forward: contract over m on the tensor core, over h in the epilogue.
//     T[(b,d), (h,u)] = sum_m Xk[b,m,d] * W_k[u, h*M + m]           (rows x M x 16*F GEMM; A = Xk rows, K = M)
//     X_{k+1}[b,u,d]  = sum_h X0[b,h,d] * T[(b,d), (h,u)] + bias[u]  (epilogue, straight out of the accumulator)
// The A operand is M values per row (not the F*M products of the outer product), W_k is its own K-major operand
// ([(h,u) rows, M contiguous columns]: no transposition), and the epilogue is one FMA per accumulator column.
constexpr int C2_THREADS = 12 * 32;      // warpgroups: [0 weights, 1 MMA, 2-3 idle] [4-7 operand warps] [8-11 epilogue warps]

template <int F, int NT, int NTI, int C0>
struct CinTCols {
    // columns C0 .. C0+15 of N-tile NTI (global j = NTI*NT + C0 + i = h*16 + u, cin_pack_operand_kernel mode 0): out[u] += T * x0[h]
    static __device__ __forceinline__ void run(uint32_t acc_addr, const float (&x0)[F], float (&out)[CT_U]) {
        if constexpr (C0 < NT && NTI * NT + C0 < CT_U * F) {
            uint32_t a[16];
            tmem_ld16(acc_addr + (uint32_t)C0, a);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int j = NTI * NT + C0 + i;
                if (j < CT_U * F) out[j % CT_U] = fmaf(__uint_as_float(a[i]), x0[j / CT_U], out[j % CT_U]);
            }
            CinTCols<F, NT, NTI, C0 + 16>::run(acc_addr, x0, out);
        }
    }
};

template <int F, int NT, int NTILES, int NTI>
struct CinTTiles {
    static __device__ __forceinline__ void run(uint32_t acc0, uint64_t* acc_full, uint64_t* acc_empty, uint32_t& n_acc, int lane,
                                               const float (&x0)[F], float (&out)[CT_U]) {
        if constexpr (NTI < NTILES) {
            const uint32_t buf = n_acc & 1u;
            mbar_wait(&acc_full[buf], (n_acc >> 1) & 1u);
            tc_fence_after();
            CinTCols<F, NT, NTI, 0>::run(acc0 + buf * NT, x0, out);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            ++n_acc;
            CinTTiles<F, NT, NTILES, NTI + 1>::run(acc0, acc_full, acc_empty, n_acc, lane, x0, out);
        }
    }
};

template <int F, int M, int NT, int NTILES>
__global__ void __launch_bounds__(C2_THREADS, 1)
cin_fwd2_tc_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo, const __grid_constant__ CinTcParams p) {
    constexpr int KP = M <= 16 ? 16 : 32;                                 // K padded to whole k-steps (zero columns)
    constexpr int KS = KP / TC_UMMA_K;
    constexpr int NP = NT * NTILES;                                       // padded 16*F
    constexpr uint32_t ACC0 = 0, A_COL = 2 * NT;                          // two NT-column accumulators, then 2 x (hi KP | lo KP)
    static_assert(A_COL + 4 * KP <= 512 && NP >= CT_U * F && NT % 16 == 0 && NT <= 256 && NTILES <= 4, "tile plan");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* w_hi = smem;                                                 // [NP rows (h,u)][128 B]: W hi (M valid columns, rest zero)
    uint8_t* w_lo = w_hi + (size_t)NP * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_lo + (size_t)NP * 128);
    uint64_t* w_full = bars;                       // [1]
    uint64_t* a_ready = w_full + 1;                // [2] Xk operand of a tile written (4 arrivals)
    uint64_t* a_empty = a_ready + 2;               // [2] its MMAs are done
    uint64_t* acc_full = a_empty + 2;              // [2]
    uint64_t* acc_empty = acc_full + 2;            // [2] 4 arrivals
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int my_tiles = ((int)blockIdx.x < p.m_tiles) ? (p.m_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (threadIdx.x == 0) {
        mbar_init(w_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&a_ready[s], 4); mbar_init(&a_empty[s], 1); mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // roles by WARPGROUP: setmaxnreg is executed by all four warps of a warpgroup at the same instruction
    if (warp < 4) {
        RPB_REG_DEC(56);
    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(w_full, (uint32_t)(2 * NP * 128));
            for (int nt = 0; nt < NTILES; ++nt) {
                tma_load_2d(w_hi + (size_t)nt * NT * 128, &tmWhi, w_full, 0, nt * NT);
                tma_load_2d(w_lo + (size_t)nt * NT * 128, &tmWlo, w_full, 0, nt * NT);
            }
        }
    } else if (warp == 1) {
        // MMA issue: warp-uniform control flow, one elected lane issues; descriptors advance by additions
        const uint32_t idesc = make_idesc_tf32(TC_BLOCK_M, NT);
        const uint64_t dh0 = make_kmajor_sw128_desc(smem_u32(w_hi)), dl0 = make_kmajor_sw128_desc(smem_u32(w_lo));
        mbar_wait(w_full, 0u);
        uint32_t n_acc = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const uint32_t ab = (uint32_t)t & 1u;
            mbar_wait(&a_ready[ab], ((uint32_t)t >> 1) & 1u);
            tc_fence_after();
            const uint32_t ta_hi = tmem_base + A_COL + ab * (2 * KP), ta_lo = ta_hi + KP;
#pragma unroll
            for (int nt = 0; nt < NTILES; ++nt, ++n_acc) {
                const uint32_t buf = n_acc & 1u;
                mbar_wait(&acc_empty[buf], ((n_acc >> 1) & 1u) ^ 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d_tmem = tmem_base + ACC0 + buf * NT;
#pragma unroll
                    for (int k = 0; k < KS; ++k) {
                        const uint64_t off = (uint64_t)((nt * NT * 128 + k * TC_UMMA_K * 4) >> 4);
                        umma_tf32_ts(d_tmem, ta_lo + k * TC_UMMA_K, dh0 + off, idesc, k > 0 ? 1u : 0u);
                        umma_tf32_ts(d_tmem, ta_hi + k * TC_UMMA_K, dl0 + off, idesc, 1u);
                        umma_tf32_ts(d_tmem, ta_hi + k * TC_UMMA_K, dh0 + off, idesc, 1u);
                    }
                    umma_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(&a_empty[ab]);
            __syncwarp();
        }
    }
    } else if (warp < 8) {
        RPB_REG_DEC(96);
        // ---------------- operand warps: thread = row (b, d): Xk[b,:,d] -> (hi, lo) -> tensor memory
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        for (int t = 0; t < my_tiles; ++t) {
            const long long b = (long long)((int)blockIdx.x + t * (int)gridDim.x) * 8 + (row >> 4);
            const int d = row & 15;
            float xv[KP];
#pragma unroll
            for (int m = 0; m < KP; ++m) xv[m] = (m < M && b < p.B) ? __ldg(p.xk + (size_t)b * p.ldk + m * CT_D + d) : 0.f;
            const uint32_t ab = (uint32_t)t & 1u;
            mbar_wait(&a_empty[ab], (((uint32_t)t >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t ta = tmem_base + A_COL + ab * (2 * KP) + lane_addr;
#pragma unroll
            for (int c = 0; c < KP; c += 16) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    hi[i] = __float_as_uint(xv[c + i]) & 0xFFFFE000u;
                    lo[i] = __float_as_uint(xv[c + i] - __uint_as_float(hi[i]));
                }
                tmem_st16(ta + c, hi);
                tmem_st16(ta + KP + c, lo);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_ready[ab]);
        }
    } else {
        RPB_REG_INC(240);
        // ---------------- epilogue warps: thread = row (b, d); X0[b,:,d] (prefetched one tile ahead) and the 16 outputs in registers
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const int d = row & 15;
        float bv[CT_U];
#pragma unroll
        for (int u = 0; u < CT_U; ++u) bv[u] = p.bias != nullptr ? __ldg(p.bias + u) : 0.f;
        uint32_t n_acc = 0;
        float x0n[F];
        {
            const long long b = (long long)(int)blockIdx.x * 8 + (row >> 4);
#pragma unroll
            for (int h = 0; h < F; ++h) x0n[h] = (my_tiles > 0 && b < p.B) ? __ldg(p.x0 + (size_t)b * p.ld0 + h * CT_D + d) : 0.f;
        }
        for (int t = 0; t < my_tiles; ++t) {
            const long long b = (long long)((int)blockIdx.x + t * (int)gridDim.x) * 8 + (row >> 4);
            const bool valid = b < p.B;
            float x0[F], out[CT_U];
#pragma unroll
            for (int h = 0; h < F; ++h) x0[h] = x0n[h];
            {
                const long long bn = (long long)((int)blockIdx.x + (t + 1) * (int)gridDim.x) * 8 + (row >> 4);
                const bool okn = t + 1 < my_tiles && bn < p.B;
#pragma unroll
                for (int h = 0; h < F; ++h) x0n[h] = okn ? __ldg(p.x0 + (size_t)bn * p.ld0 + h * CT_D + d) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < CT_U; ++u) out[u] = bv[u];
            CinTTiles<F, NT, NTILES, 0>::run(tmem_base + ACC0 + lane_addr, acc_full, acc_empty, n_acc, lane, x0, out);
            if (valid && p.xout != nullptr) {
                float* xo = p.xout + (size_t)b * p.ldo + d;
#pragma unroll
                for (int u = 0; u < CT_U; ++u) xo[u * CT_D] = out[u];
            }
            if (p.pooled != nullptr) {
                // sums over the 16 lanes (d) of a sample for all 16 units at once: recursive halving, 8 + 4 + 2 + 1 shuffles instead
                // of 16 four-step butterflies; lane d ends up with the total of unit u = d -> one 64-byte store per sample
                float r8[8], r4[4], r2[2];
                {
                    const bool up = (d & 8) != 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float recv = __shfl_xor_sync(0xffffffffu, up ? out[i] : out[i + 8], 8);
                        r8[i] = (up ? out[i + 8] : out[i]) + recv;
                    }
                }
                {
                    const bool up = (d & 4) != 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float recv = __shfl_xor_sync(0xffffffffu, up ? r8[i] : r8[i + 4], 4);
                        r4[i] = (up ? r8[i + 4] : r8[i]) + recv;
                    }
                }
                {
                    const bool up = (d & 2) != 0;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float recv = __shfl_xor_sync(0xffffffffu, up ? r4[i] : r4[i + 2], 2);
                        r2[i] = (up ? r4[i + 2] : r4[i]) + recv;
                    }
                }
                const bool up = (d & 1) != 0;
                const float recv = __shfl_xor_sync(0xffffffffu, up ? r2[0] : r2[1], 1);
                const float tot = (up ? r2[1] : r2[0]) + recv;
                if (valid) p.pooled[(size_t)b * p.ldp + d] = tot;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int F, int M, int NT, int NTILES>
static int cin_fwd2_tc_launch(const float* W, const CinTcParams& p, cudaStream_t st) {
    CUtensorMap tmWhi, tmWlo;
    // W_k [16, F*M] -> rows (h, u), M columns -> hi / lo [NT*NTILES rows, 32 columns] (zero padded), [NT x 32] boxes
    int rc = cin_prepare_operand(W, F, M, 0, NT * NTILES, NT, 9, &tmWhi, &tmWlo, st);
    if (rc != 0) return rc;
    const size_t smem = (size_t)2 * NT * NTILES * 128 + 9 * 8 + 32 + 1024;
    cudaError_t e = cudaFuncSetAttribute(cin_fwd2_tc_kernel<F, M, NT, NTILES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cin_fwd2_tc_kernel<F, M, NT, NTILES><<<min(p.m_tiles, 148), C2_THREADS, smem, st>>>(tmWhi, tmWlo, p);
    return (int)cudaGetLastError();
}

// One CIN layer forward on tensor cores.  Returns RPB_ERR_UNSUPPORTED for shapes outside the instantiated (F, M) pairs.
int cin_layer_fwd_tc(int F, int M, int U, int D, const float* W, const float* bias, const float* x0, long long ld0, const float* xk,
                     long long ldk, float* xout, long long ldo, float* pooled, long long ldp, int B, cudaStream_t st) {
    if (D != CT_D || U != CT_U || !g_gemm_v2) return RPB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(W) & 15u) != 0) return RPB_ERR_UNSUPPORTED;
    CinTcParams p{};
    p.x0 = x0; p.ld0 = ld0; p.xk = xk; p.ldk = ldk; p.bias = bias; p.xout = xout; p.ldo = ldo; p.pooled = pooled; p.ldp = ldp;
    p.B = B; p.m_tiles = ceil_div(B, 8);
    if (F == 26 && M == 26) return cin_fwd2_tc_launch<26, 26, 144, 3>(W, p, st);
    if (F == 26 && M == 16) return cin_fwd2_tc_launch<26, 16, 208, 2>(W, p, st);
    return RPB_ERR_UNSUPPORTED;
}

// =====================================================================================================================
// Backward of one layer, part A (per-sample gradients):  dZ[(b,d), j] = sum_u G[(b,d), u] * W_k[u, j]  (rows x 16 x F*M GEMM),
// consumed straight out of the accumulator:  dXk[b,m,d] += dZ[., h*M+m] * X0[b,h,d],   dX0[b,h,d] += dZ[., h*M+m] * Xk[b,m,d].
// G = dL/dX_{k+1} = dpooled (broadcast over d) + the dXk the layer above produced.  Thread = row (b, d) on both sides of the
// tensor core: "G" warps form G, spill it for the weight-gradient kernel, split it into the TS-mode A operand (K = 16);
// epilogue warps hold X0[b,:,d], Xk[b,:,d] and the two gradient vectors in registers and read dZ 16 columns at a time; every
// (h, m) is a compile-time constant.  W_k^T hi / lo ([F*M rows, 16 -> 32 zero-padded columns], SWIZZLE_128B) resident in smem;
// 3 MMAs per k-step (hi.hi + lo.hi + hi.lo) into NT-column accumulators, double buffered.
constexpr int CB_THREADS = 12 * 32;      // warpgroups: [0 weights, 1 MMA, 2-3 idle] [4-7 G warps] [8-11 epilogue warps]; registers re-dealt by setmaxnreg

struct CinBwdParams {
    const float* x0; long long ld0;
    const float* xk; long long ldk;
    const float* dpooled; long long lddp;        // + p_off_k already applied
    const float* gx; long long ldgx;             // dXk of the layer above [B, 16, 16] or null (top layer)
    float* gout; long long ldgo;                 // G spill for the weight gradient: gout[b * ldgo + u * 16 + d]
    float* db;                                   // [16] accumulated (atomics), or null
    float* dxk; long long lddxk;                 // k > 0: dXk[b,m,d] written (-> gx of the layer below); k == 0: null
    float* de; long long ldde; int de_accumulate;    // dX0 (+ dXk when k == 0) added into / written to de[b,h,d]
    int B, m_tiles;
};

template <int F, int M, int NT, int NTI, int C0>
struct CinDzCols {
    // columns C0 .. C0+15 of N-tile NTI; global column j = NTI*NT + C0 + i = s*F + h stands for (h, m = (h + s) % M)
    // (cin_pack_operand_kernel mode 1: neighbouring columns differ in h AND m, so no two consecutive FMAs share an accumulator):
    // dxk[m] += dz * x0[h], dx0[h] += dz * xk[m]
    static __device__ __forceinline__ void run(uint32_t acc_addr, const float (&x0)[F], const float (&xk)[M], float (&dx0)[F], float (&dxk)[M]) {
        if constexpr (C0 < NT && NTI * NT + C0 < F * M) {
            uint32_t a[16];
            tmem_ld16(acc_addr + (uint32_t)C0, a);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int j = NTI * NT + C0 + i;
                if (j < F * M) {
                    const int h = j % F, m = (h + j / F) % M;
                    const float dz = __uint_as_float(a[i]);
                    dxk[m] = fmaf(dz, x0[h], dxk[m]);
                    dx0[h] = fmaf(dz, xk[m], dx0[h]);
                }
            }
            CinDzCols<F, M, NT, NTI, C0 + 16>::run(acc_addr, x0, xk, dx0, dxk);
        }
    }
};

template <int F, int M, int NT, int NTILES>
__global__ void __launch_bounds__(CB_THREADS, 1)
cin_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo, const __grid_constant__ CinBwdParams p) {
    constexpr int NP = NT * NTILES;                                       // padded F*M
    constexpr uint32_t ACC0 = 0, A_COL = 2 * NT;                          // two NT-column accumulators, then 2 x 32 operand columns
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* w_hi = smem;                                                 // [NP rows][128 B]: W^T hi (16 valid columns, 16 zero)
    uint8_t* w_lo = w_hi + (size_t)NP * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_lo + (size_t)NP * 128);
    uint64_t* w_full = bars;                       // [1]
    uint64_t* a_ready = w_full + 1;                // [2] G operand of a tile written (4 arrivals)
    uint64_t* a_empty = a_ready + 2;               // [2] its MMAs are done
    uint64_t* acc_full = a_empty + 2;              // [2]
    uint64_t* acc_empty = acc_full + 2;            // [2] 4 arrivals
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* db_part = reinterpret_cast<float*>(tmem_ptr + 4);              // [4 warps][16]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int my_tiles = ((int)blockIdx.x < p.m_tiles) ? (p.m_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (threadIdx.x == 0) {
        mbar_init(w_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&a_ready[s], 4); mbar_init(&a_empty[s], 1); mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // roles by WARPGROUP: setmaxnreg is executed by all four warps of a warpgroup at the same instruction
    if (warp < 4) {
        RPB_REG_DEC(56);
    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(w_full, (uint32_t)(2 * NP * 128));
            for (int nt = 0; nt < NTILES; ++nt) {
                tma_load_2d(w_hi + (size_t)nt * NT * 128, &tmWhi, w_full, 0, nt * NT);
                tma_load_2d(w_lo + (size_t)nt * NT * 128, &tmWlo, w_full, 0, nt * NT);
            }
        }
    } else if (warp == 1) {
        // MMA issue: warp-uniform control flow, one elected lane issues; descriptors advance by additions
        const uint32_t idesc = make_idesc_tf32(TC_BLOCK_M, NT);
        const uint64_t dh0 = make_kmajor_sw128_desc(smem_u32(w_hi)), dl0 = make_kmajor_sw128_desc(smem_u32(w_lo));
        mbar_wait(w_full, 0u);
        uint32_t n_acc = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const uint32_t ab = (uint32_t)t & 1u;
            mbar_wait(&a_ready[ab], ((uint32_t)t >> 1) & 1u);
            tc_fence_after();
            const uint32_t ta_hi = tmem_base + A_COL + ab * 32u, ta_lo = ta_hi + 16u;
#pragma unroll
            for (int nt = 0; nt < NTILES; ++nt, ++n_acc) {
                const uint32_t buf = n_acc & 1u;
                mbar_wait(&acc_empty[buf], ((n_acc >> 1) & 1u) ^ 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d_tmem = tmem_base + ACC0 + buf * NT;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {                          // K = 16 = 2 k-steps of 8
                        const uint64_t off = (uint64_t)((nt * NT * 128 + k * TC_UMMA_K * 4) >> 4);
                        umma_tf32_ts(d_tmem, ta_lo + k * TC_UMMA_K, dh0 + off, idesc, k > 0 ? 1u : 0u);
                        umma_tf32_ts(d_tmem, ta_hi + k * TC_UMMA_K, dl0 + off, idesc, 1u);
                        umma_tf32_ts(d_tmem, ta_hi + k * TC_UMMA_K, dh0 + off, idesc, 1u);
                    }
                    umma_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(&a_empty[ab]);
            __syncwarp();
        }
    }
    } else if (warp < 8) {
        RPB_REG_DEC(96);
        // ---------------- G warps: thread = row (b, d): G[u] = dpooled[b, u] + gx[b, u, d]; spill, bias gradient, A operand
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        float dbs[CT_U];
#pragma unroll
        for (int u = 0; u < CT_U; ++u) dbs[u] = 0.f;
        for (int t = 0; t < my_tiles; ++t) {
            const long long b = (long long)((int)blockIdx.x + t * (int)gridDim.x) * 8 + (row >> 4);
            const int d = row & 15;
            {
                // these warps run one to two tiles ahead of the epilogue: pull the rows it will read (X0, Xk, the dE rows it
                // accumulates into) of the NEXT tile into L2, one 64-byte row per lane and step, so that its loads at the top and
                // bottom of a tile are L2 hits instead of DRAM round trips
                const long long bn = (long long)((int)blockIdx.x + (t + 1) * (int)gridDim.x) * 8 + (row >> 4);
                if (t + 1 < my_tiles && bn < p.B) {
                    for (int h = d; h < F; h += CT_D) {
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(p.x0 + (size_t)bn * p.ld0 + h * CT_D));
                        if (p.de_accumulate) asm volatile("prefetch.global.L2 [%0];" :: "l"(p.de + (size_t)bn * p.ldde + h * CT_D));
                    }
                    if (p.xk != p.x0) {
                        for (int m = d; m < M; m += CT_D) asm volatile("prefetch.global.L2 [%0];" :: "l"(p.xk + (size_t)bn * p.ldk + m * CT_D));
                    }
                }
            }
            float g[CT_U];
            if (b < p.B) {
#pragma unroll
                for (int u = 0; u < CT_U; ++u) {
                    float v = __ldg(p.dpooled + (size_t)b * p.lddp + u);
                    if (p.gx != nullptr) v += __ldg(p.gx + (size_t)b * p.ldgx + u * CT_D + d);
                    g[u] = v;
                    dbs[u] += v;
                    if (p.gout != nullptr) p.gout[(size_t)b * p.ldgo + u * CT_D + d] = v;
                }
            } else {
#pragma unroll
                for (int u = 0; u < CT_U; ++u) g[u] = 0.f;
            }
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int u = 0; u < CT_U; ++u) {
                hi[u] = __float_as_uint(g[u]) & 0xFFFFE000u;
                lo[u] = __float_as_uint(g[u] - __uint_as_float(hi[u]));
            }
            const uint32_t ab = (uint32_t)t & 1u;
            mbar_wait(&a_empty[ab], (((uint32_t)t >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t ta = tmem_base + A_COL + ab * 32u + lane_addr;
            tmem_st16(ta, hi);
            tmem_st16(ta + 16u, lo);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_ready[ab]);
        }
        if (p.db != nullptr) {
#pragma unroll
            for (int u = 0; u < CT_U; ++u) {
                const float s = warp_sum(dbs[u]);
                if (lane == 0) db_part[(warp - 4) * CT_U + u] = s;
            }
        }
    } else {
        RPB_REG_INC(240);
        // ---------------- epilogue warps: thread = row (b, d); X0[b,:,d], Xk[b,:,d], dX0, dXk in registers
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        uint32_t n_acc = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const long long b = (long long)((int)blockIdx.x + t * (int)gridDim.x) * 8 + (row >> 4);
            const int d = row & 15;
            const bool valid = b < p.B;
            float x0[F], xk[M], dx0[F], dxk[M];
#pragma unroll
            for (int h = 0; h < F; ++h) { x0[h] = valid ? __ldg(p.x0 + (size_t)b * p.ld0 + h * CT_D + d) : 0.f; dx0[h] = 0.f; }
#pragma unroll
            for (int m = 0; m < M; ++m) { xk[m] = valid ? __ldg(p.xk + (size_t)b * p.ldk + m * CT_D + d) : 0.f; dxk[m] = 0.f; }
            // N-tiles unrolled by hand (NTILES <= 4): every column index must be a compile-time constant
            auto one_tile = [&](auto nti_tag) {
                constexpr int NTI = decltype(nti_tag)::value;
                if constexpr (NTI < NTILES) {
                    const uint32_t buf = n_acc & 1u;
                    mbar_wait(&acc_full[buf], (n_acc >> 1) & 1u);
                    tc_fence_after();
                    CinDzCols<F, M, NT, NTI, 0>::run(tmem_base + ACC0 + buf * NT + lane_addr, x0, xk, dx0, dxk);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                    ++n_acc;
                }
            };
            one_tile(std::integral_constant<int, 0>{});
            one_tile(std::integral_constant<int, 1>{});
            one_tile(std::integral_constant<int, 2>{});
            one_tile(std::integral_constant<int, 3>{});
            if (valid) {
                if (p.dxk != nullptr) {
#pragma unroll
                    for (int m = 0; m < M; ++m) p.dxk[(size_t)b * p.lddxk + m * CT_D + d] = dxk[m];
                }
                // old values first, all loads in flight together (x0 / xk are dead here: their registers are free)
                float* dst = p.de + (size_t)b * p.ldde + d;
                float old[F];
#pragma unroll
                for (int h = 0; h < F; ++h) old[h] = p.de_accumulate ? __ldcg(dst + h * CT_D) : 0.f;
#pragma unroll
                for (int h = 0; h < F; ++h) {
                    float v = dx0[h] + old[h];
                    if constexpr (M == F) { if (p.dxk == nullptr) v += dxk[h]; }          // layer 0: Xk is X0
                    dst[h * CT_D] = v;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
    if (p.db != nullptr && threadIdx.x < CT_U) {
        const float s = db_part[threadIdx.x] + db_part[CT_U + threadIdx.x] + db_part[2 * CT_U + threadIdx.x] + db_part[3 * CT_U + threadIdx.x];
        red_add_f1(p.db + threadIdx.x, s);
    }
}

template <int F, int M, int NT, int NTILES>
static int cin_bwd_tc_launch(const float* W, const CinBwdParams& p, cudaStream_t st) {
    static_assert(NTILES <= 4 && NT % 16 == 0 && NT * NTILES >= F * M && 2 * NT + 64 <= 512, "tile plan");
    CUtensorMap tmWhi, tmWlo;
    // W_k [16, F*M] -> W_k^T with permuted rows (s, h) [NT*NTILES rows, 32 columns] hi / lo (columns 16.. and rows F*M.. are zero)
    int rc = cin_prepare_operand(W, F, M, 1, NT * NTILES, NT, 8, &tmWhi, &tmWlo, st);
    if (rc != 0) return rc;
    const size_t smem = (size_t)2 * NT * NTILES * 128 + 9 * 8 + 16 + 4 * CT_U * 4 + 1024;
    cudaError_t e = cudaFuncSetAttribute(cin_bwd_tc_kernel<F, M, NT, NTILES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cin_bwd_tc_kernel<F, M, NT, NTILES><<<min(p.m_tiles, 148), CB_THREADS, smem, st>>>(tmWhi, tmWlo, p);
    return (int)cudaGetLastError();
}

int cin_layer_bwd_tc(int F, int M, const float* W, const CinBwdParams& p, cudaStream_t st) {
    if (F == 26 && M == 26) return cin_bwd_tc_launch<26, 26, 176, 4>(W, p, st);
    if (F == 26 && M == 16) return cin_bwd_tc_launch<26, 16, 208, 2>(W, p, st);
    return RPB_ERR_UNSUPPORTED;
}

// =====================================================================================================================
// Backward of one layer, part B (weight gradient):  dW_k[u, h*M + m] = sum_{(b,d)} G[(b,d),u] * X0[(b,d),h] * Xk[(b,d),m].
// Written as a GEMM over the rows r = (b, d):   D[(u,h), m] = sum_r P[(u,h), r] * Xk[r, m],   P[(u,h), r] = G[r,u] * X0[r,h].
//   * two loader warps stage, per k-block of 32 rows (2 samples x 16 d) and CW_AHEAD k-blocks ahead (cp.async): the Xk rows as the
//     K-major B tile (row m = Xk[b0,m,:] | Xk[b1,m,:], SWIZZLE_128B positions), the G rows [2][16 u][16 d] and the X0 rows
//     [2][F][16 d] (chunks XOR-swizzled by (h >> 1) & 3 so that a quarter warp of consecutive h reads conflict-free).  The raw
//     fp32 Xk tile is the "hi" operand (the tensor core ignores the low 13 mantissa bits); its loader writes lo = x - hi below
//     it: stacked [raw ; lo] operand of 2*NB rows (rows m >= M are zero), two TS-mode MMAs (a_lo, a_hi) per k-step.
//   * A = P^T in tensor memory (TS mode): lane = (u, h) pair (416 pairs = 4 M-tiles of 128), columns = the 32 rows of the
//     k-block.  An operand thread reads its two 64-byte rows G[b,u,:] and X0[b,h,:] from the stage (LDS.128), multiplies them
//     element-wise, splits and stores: no outer product, no transposition, no exposed DRAM latency.
//   * the four accumulators (2*NB columns each) live in tensor memory for the CTA's whole slab of rows; the epilogue adds the two
//     halves and reduces into dW with fp32 `red`.
constexpr int CW_THREADS = 11 * 32;      // 0 = Xk loader, 1 = MMA, 2-9 = operand warps (w, w+4: sample 0 / 1 of the k-block; 2-5 also epilogue), 10 = G / X0 loader
constexpr int CW_BST = 6;                // stage ring
constexpr int CW_AHEAD = 4;              // k-blocks whose cp.async groups are in flight (< CW_BST)
constexpr int CW_OPN = 4;                // A operand ring (64 columns each)

struct CinWgParams {
    const float* x0; long long ld0;
    const float* xk; long long ldk;
    const float* g; long long ldg;       // G spill: g[b * ldg + u * 16 + d]
    float* dW;                           // [16, F*M] accumulated
    int B, n_kb;                         // k-blocks = ceil(B / 2)
};

template <int F, int M>
struct CwLayout {
    static constexpr int NB = M <= 16 ? 16 : 32;                          // rows of one half of the stacked B tile
    static constexpr int B_BYTES = 2 * NB * 128;
    static constexpr int G_OFF = B_BYTES;                                 // [2][16][64 B]
    static constexpr int X_OFF = G_OFF + 2 * CT_U * 64;                   // [2][F][64 B], chunk-swizzled
    static constexpr int STAGE = (X_OFF + 2 * F * 64 + 1023) / 1024 * 1024;
};

template <int F, int M>
__global__ void __launch_bounds__(CW_THREADS, 1)
cin_wgrad_tc_kernel(const __grid_constant__ CinWgParams p) {
    using L = CwLayout<F, M>;
    constexpr int NB = L::NB;
    constexpr int NPAIR = CT_U * F;                                      // (u, h) pairs
    constexpr int MT = (NPAIR + 127) / 128;                              // M-tiles
    static_assert(MT <= 4 && M <= 32, "accumulator plan");
    constexpr uint32_t ACCW = 2 * NB;
    constexpr uint32_t A_COL = 4 * ACCW;                                 // after the accumulators
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* st_base = smem;                                             // CW_BST stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(st_base + CW_BST * L::STAGE);
    uint64_t* b_full = bars;                       // [CW_BST] Xk tile (+ lo half) written          -> MMA thread
    uint64_t* b_empty = b_full + CW_BST;           // [CW_BST] its MMAs are done (commit)          -> Xk loader
    uint64_t* g_full = b_empty + CW_BST;           // [CW_BST] G / X0 rows written                  -> operand warps
    uint64_t* g_empty = g_full + CW_BST;           // [CW_BST] 8 operand warps have read them       -> G / X0 loader
    uint64_t* a_ready = g_empty + CW_BST;          // [CW_OPN] 8 arrivals
    uint64_t* a_empty = a_ready + CW_OPN;          // [CW_OPN]
    uint64_t* acc_done = a_empty + CW_OPN;         // [1]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // slab of k-blocks of this CTA
    const int per = (p.n_kb + (int)gridDim.x - 1) / (int)gridDim.x;
    const int kb0 = (int)blockIdx.x * per, kb1 = min(p.n_kb, kb0 + per);
    const int nkb = max(0, kb1 - kb0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < CW_BST; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); mbar_init(&g_full[s], 1); mbar_init(&g_empty[s], 8); }
        for (int s = 0; s < CW_OPN; ++s) { mbar_init(&a_ready[s], 8); mbar_init(&a_empty[s], 1); }
        mbar_init(acc_done, 1);
        fence_barrier_init();
    }
    // rows m >= M of every stacked tile stay zero for the whole kernel
    for (int i = threadIdx.x; i < CW_BST * L::STAGE / 16; i += blockDim.x) reinterpret_cast<float4*>(st_base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ---------------- Xk loader: rows of a k-block by cp.async (16-byte pieces, SWIZZLE_128B positions), CW_AHEAD k-blocks in
        // flight (one cp.async group each), then the lo half of the tile once the group has landed.  Per-lane piece offsets are
        // fixed for the whole kernel (the loaders were instruction-bound with the index arithmetic in the loop).
        constexpr int NE = (M * 8 + 31) / 32;
        uint32_t src_off[NE], dst_off[NE], hf[NE];
#pragma unroll
        for (int j = 0; j < NE; ++j) {
            const int e = lane + 32 * j, m = e >> 3, c = e & 7;           // (row m, chunk c): c < 4 sample 0, c >= 4 sample 1
            hf[j] = e < M * 8 ? (uint32_t)(c >> 2) : 0x40000000u;         // out-of-range pieces: "sample" beyond any batch
            src_off[j] = m * CT_D + (c & 3) * 4;
            dst_off[j] = m * 128 + ((c ^ (m & 7)) << 4);
        }
        auto issue = [&](int j2) {
            const int s = j2 % CW_BST;
            mbar_wait(&b_empty[s], ((j2 / CW_BST) & 1u) ^ 1u);
            const uint32_t tile = smem_u32(st_base + (size_t)s * L::STAGE);
            const long long bA = (long long)(kb0 + j2) * 2;
#pragma unroll
            for (int j = 0; j < NE; ++j) {
                const long long b = bA + hf[j];
                const bool ok = b < p.B;
                if (hf[j] < 2u) {
                    const float* src = p.xk + (size_t)(ok ? b : 0) * p.ldk + src_off[j];
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(tile + dst_off[j]), "l"(src), "r"(ok ? 16 : 0) : "memory");
                }
            }
        };
        for (int j = 0; j < CW_AHEAD; ++j) {
            if (j < nkb) issue(j);
            asm volatile("cp.async.commit_group;" ::: "memory");          // (possibly empty) group j
        }
        for (int i = 0; i < nkb; ++i) {
            const int s = i % CW_BST;
            uint8_t* tile = st_base + (size_t)s * L::STAGE;
            asm volatile("cp.async.wait_group %0;" :: "n"(CW_AHEAD - 1) : "memory");      // group i has landed
            __syncwarp();
#pragma unroll
            for (int j = 0; j < NE; ++j) {
                if (hf[j] < 2u) {
                    const float4 v = *reinterpret_cast<const float4*>(tile + dst_off[j]);
                    float4 l;
                    l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                    l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                    *reinterpret_cast<float4*>(tile + NB * 128 + dst_off[j]) = l;       // lower half of the stacked tile (same swizzle phase)
                }
            }
            fence_proxy_async();                                          // generic-proxy writes -> the MMA's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&b_full[s]);
            if (i + CW_AHEAD < nkb) issue(i + CW_AHEAD);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    } else if (warp == 10) {
        // ---------------- G / X0 loader: [2][16 u][64 B] and [2][F][64 B] (chunks XOR-swizzled by (h >> 1) & 3) per k-block
        constexpr int NG = (2 * CT_U * 4) / 32, NX = (2 * F * 4 + 31) / 32;
        uint32_t src_off[NG + NX], dst_off[NG + NX], hf[NG + NX];
#pragma unroll
        for (int j = 0; j < NG; ++j) {
            const int e = lane + 32 * j, half = e >> 6, u = (e >> 2) & 15, c = e & 3;
            hf[j] = (uint32_t)half;
            src_off[j] = u * CT_D + c * 4;
            dst_off[j] = L::G_OFF + (half * CT_U + u) * 64 + (c << 4);
        }
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            const int e = lane + 32 * j, half = e / (F * 4), r = e - half * (F * 4), h = r >> 2, c = r & 3;
            hf[NG + j] = e < 2 * F * 4 ? (uint32_t)half : 0x40000000u;
            src_off[NG + j] = h * CT_D + c * 4;
            dst_off[NG + j] = L::X_OFF + (half * F + h) * 64 + ((c ^ ((h >> 1) & 3)) << 4);
        }
        auto issue = [&](int j2) {
            const int s = j2 % CW_BST;
            mbar_wait(&g_empty[s], ((j2 / CW_BST) & 1u) ^ 1u);
            const uint32_t tile = smem_u32(st_base + (size_t)s * L::STAGE);
            const long long bA = (long long)(kb0 + j2) * 2;
#pragma unroll
            for (int j = 0; j < NG + NX; ++j) {
                const long long b = bA + hf[j];
                const bool ok = b < p.B;
                if (hf[j] < 2u) {
                    const float* src = (j < NG ? p.g + (size_t)(ok ? b : 0) * p.ldg : p.x0 + (size_t)(ok ? b : 0) * p.ld0) + src_off[j];
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(tile + dst_off[j]), "l"(src), "r"(ok ? 16 : 0) : "memory");
                }
            }
        };
        for (int j = 0; j < CW_AHEAD; ++j) {
            if (j < nkb) issue(j);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int i = 0; i < nkb; ++i) {
            asm volatile("cp.async.wait_group %0;" :: "n"(CW_AHEAD - 1) : "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&g_full[i % CW_BST]);
            if (i + CW_AHEAD < nkb) issue(i + CW_AHEAD);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    } else if (warp == 1) {
        // ---------------- MMA issue: warp-uniform control flow, one elected lane issues (no per-instruction waterfall: with 32 small
        // MMAs per k-block the issue loop was the bound, profiles/r02_cin_ncu.md); descriptors advance by additions
        const uint32_t idesc = make_idesc_tf32(TC_BLOCK_M, 2 * NB);
        const uint64_t desc0 = make_kmajor_sw128_desc(smem_u32(st_base));
        uint32_t ga = 0;
        for (int i = 0; i < nkb; ++i) {
            const int s = i % CW_BST;
            mbar_wait(&b_full[s], (i / CW_BST) & 1u);
            tc_fence_after();
            const uint64_t dbs = desc0 + (uint64_t)(((uint32_t)s * (uint32_t)L::STAGE) >> 4);
            const uint32_t acc_on = i > 0 ? 1u : 0u;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt, ++ga) {
                const uint32_t o = ga % CW_OPN;
                mbar_wait(&a_ready[o], (ga / CW_OPN) & 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)mt * ACCW;
                    const uint32_t ta_hi = tmem_base + A_COL + o * 64u, ta_lo = ta_hi + 32u;
#pragma unroll
                    for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
                        const uint64_t db = dbs + (uint64_t)(k * TC_UMMA_K * 4 >> 4);
                        umma_tf32_ts(d_tmem, ta_lo + k * TC_UMMA_K, db, idesc, k > 0 ? 1u : acc_on);
                        umma_tf32_ts(d_tmem, ta_hi + k * TC_UMMA_K, db, idesc, 1u);
                    }
                    umma_commit(&a_empty[o]);
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(&b_empty[s]);
            __syncwarp();
        }
        if (elect_one()) umma_commit(acc_done);
        __syncwarp();
    } else {
        // ---------------- operand warps: lane = (u, h) pair of the M-tile, `half` = which sample of the k-block
        const int q = warp & 3, half = (warp - 2) >> 2;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        uint32_t g_off[MT], x_off[MT], x_sw[MT];
        bool okp[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int pr = mt * 128 + q * 32 + lane;
            okp[mt] = pr < NPAIR;
            const int u = okp[mt] ? pr / F : 0, h = okp[mt] ? pr - u * F : 0;
            g_off[mt] = L::G_OFF + (half * CT_U + u) * 64;
            x_off[mt] = L::X_OFF + (half * F + h) * 64;
            x_sw[mt] = (h >> 1) & 3;
        }
        static_assert(MT == 4, "the operand loop below is unrolled over four M-tiles with two register sets");
        // P columns of M-tile mt for this thread's (u, h) pair and sample: G[b,u,:] * X0[b,h,:], split into (hi, lo)
        auto compute = [&](const uint8_t* tile, int mt, uint32_t (&hi)[16], uint32_t (&lo)[16]) {
            if (okp[mt]) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 a = *reinterpret_cast<const float4*>(tile + g_off[mt] + (c << 4));
                    const float4 x = *reinterpret_cast<const float4*>(tile + x_off[mt] + ((c ^ x_sw[mt]) << 4));
                    const float z[4] = {a.x * x.x, a.y * x.y, a.z * x.z, a.w * x.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        hi[4 * c + e] = __float_as_uint(z[e]) & 0xFFFFE000u;
                        lo[4 * c + e] = __float_as_uint(z[e] - __uint_as_float(hi[4 * c + e]));
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) { hi[e] = 0u; lo[e] = 0u; }
            }
        };
        // software pipeline: the tcgen05.st of operand g is in flight while the products of operand g+1 are formed
        auto put = [&](uint32_t g, const uint32_t (&hi)[16], const uint32_t (&lo)[16]) {
            const uint32_t o = g % CW_OPN;
            mbar_wait(&a_empty[o], ((g / CW_OPN) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t ta = tmem_base + A_COL + o * 64u + lane_addr + (uint32_t)half * 16u;
            tmem_st16(ta, hi);
            tmem_st16(ta + 32u, lo);
        };
        auto publish = [&](uint32_t g) {
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_ready[g % CW_OPN]);
        };
        uint32_t hA[16], lA[16], hB[16], lB[16];
        uint32_t ga = 0;
        if (nkb > 0) {
            mbar_wait(&g_full[0], 0u);
            compute(st_base, 0, hA, lA);
        }
        for (int i = 0; i < nkb; ++i) {
            const int s = i % CW_BST;
            const uint8_t* tile = st_base + (size_t)s * L::STAGE;
            put(ga, hA, lA); compute(tile, 1, hB, lB); publish(ga); ++ga;
            put(ga, hB, lB); compute(tile, 2, hA, lA); publish(ga); ++ga;
            put(ga, hA, lA); compute(tile, 3, hB, lB); publish(ga); ++ga;
            if (lane == 0) mbar_arrive(&g_empty[s]);                      // every lane has read the stage's G / X0 rows (syncwarp in publish)
            put(ga, hB, lB);
            if (i + 1 < nkb) {
                const int s2 = (i + 1) % CW_BST;
                mbar_wait(&g_full[s2], ((i + 1) / CW_BST) & 1u);
                compute(st_base + (size_t)s2 * L::STAGE, 0, hA, lA);
            }
            publish(ga); ++ga;
        }
        // ---------------- epilogue (first four operand warps): acc[mt] = a.b_raw | a.b_lo  ->  dW[u, h*M + m] += sum
        if (half == 0 && nkb > 0) {
            mbar_wait(acc_done, 0u);
            tc_fence_after();
            for (int mt = 0; mt < MT; ++mt) {
                const int pr = mt * 128 + q * 32 + lane;
                const int u = pr / F, h = pr - u * F;
                float* dst = p.dW + (size_t)u * (F * M) + h * M;
#pragma unroll
                for (int c = 0; c < NB; c += 16) {
                    uint32_t a[16], l[16];
                    tmem_ld16(tmem_base + (uint32_t)mt * ACCW + lane_addr + (uint32_t)c, a);
                    tmem_ld16(tmem_base + (uint32_t)mt * ACCW + lane_addr + (uint32_t)(NB + c), l);
                    if (pr < NPAIR) {
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            if (c + e < M) red_add_f1(dst + c + e, __uint_as_float(a[e]) + __uint_as_float(l[e]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int F, int M>
static int cin_wgrad_tc_launch(const CinWgParams& p, cudaStream_t st) {
    const size_t smem = (size_t)CW_BST * CwLayout<F, M>::STAGE + (4 * CW_BST + 2 * CW_OPN + 1) * 8 + 16 + 1024;
    cudaError_t e = cudaFuncSetAttribute(cin_wgrad_tc_kernel<F, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cin_wgrad_tc_kernel<F, M><<<min(p.n_kb, 148), CW_THREADS, smem, st>>>(p);
    return (int)cudaGetLastError();
}

int cin_layer_wgrad_tc(int F, int M, const float* x0, long long ld0, const float* xk, long long ldk, const float* g, long long ldg,
                       float* dW, int B, cudaStream_t st) {
    if ((ld0 & 3) || (ldk & 3) || (ldg & 3) || (reinterpret_cast<uintptr_t>(x0) & 15u) || (reinterpret_cast<uintptr_t>(xk) & 15u) ||
        (reinterpret_cast<uintptr_t>(g) & 15u))
        return RPB_ERR_UNSUPPORTED;
    CinWgParams p{};
    p.x0 = x0; p.ld0 = ld0; p.xk = xk; p.ldk = ldk; p.g = g; p.ldg = ldg; p.dW = dW; p.B = B; p.n_kb = ceil_div(B, 2);
    if (F == 26 && M == 26) return cin_wgrad_tc_launch<26, 26>(p, st);
    if (F == 26 && M == 16) return cin_wgrad_tc_launch<26, 16>(p, st);
    return RPB_ERR_UNSUPPORTED;
}

int cin_layer_bwd_tc_c(int F, int M, const float* W, const float* x0, long long ld0, const float* xk, long long ldk, const float* dpooled,
                       long long lddp, const float* gx, long long ldgx, float* gout, long long ldgo, float* db, float* dxk, long long lddxk,
                       float* de, long long ldde, int de_accumulate, int B, cudaStream_t st) {
    CinBwdParams p{};
    p.x0 = x0; p.ld0 = ld0; p.xk = xk; p.ldk = ldk; p.dpooled = dpooled; p.lddp = lddp; p.gx = gx; p.ldgx = ldgx;
    p.gout = gout; p.ldgo = ldgo; p.db = db; p.dxk = dxk; p.lddxk = lddxk; p.de = de; p.ldde = ldde; p.de_accumulate = de_accumulate;
    p.B = B; p.m_tiles = ceil_div(B, 8);
    return cin_layer_bwd_tc(F, M, W, p, st);
}

// (F, M, U, D) combinations the tensor-core kernels are instantiated for: the Criteo shape of BASELINE.json config 3
bool cin_tc_shape_ok(int F, int D, int L, const int* units) {
    if (!g_cin_tc || !g_gemm_v2 || F != 26 || D != CT_D || L < 1) return false;
    for (int k = 0; k < L; ++k) if (units[k] != CT_U) return false;
    return true;
}

}  // namespace rpb
