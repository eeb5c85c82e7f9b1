// Backward of the MLP tower tail with the dz chain on tcgen05 (opt-in: rpb_set_option("tower_bwd_tc", 1); written after the
// last GPU call of round 1 — compiled, NOT YET RUN ON HARDWARE; the default is tower_tail_bwd_kernel in tower.cu).
//
// Same contract as rpb_tower_tail_bwd (reference: autograd of models/layers/deep.py:62-84 + BCELoss/sigmoid backward):
//   dz_top[m, n] = dlogit[m] * w_out[n] * (h_top[m, n] > 0);   dz_j = (dz_{j+1} . W_j) * (hin[j] > 0),  j = n_tail-1 .. 0
//   db[j] += colsum(dz[j]);  dw_out += sum_m dlogit[m] * h_top[m, :];  db_out += sum_m dlogit[m].
// The CUDA-core kernel spends 2 * 64 * 64 FMAs per sample and layer on the fp32 pipe (48 us at config 2, floor 15 us).  Here
// one CTA owns a 128-sample tile: thread = (sample row, column half) as in the tcgen05 forward tail (deepfm_fused.cu,
// TCTAIL) — it forms its 32 columns of dz in registers, stores them for the weight-gradient kernels, splits them into
// (hi, lo) and writes them into TENSOR MEMORY as the A operand; one thread issues 8 k-steps x 2 TS-mode MMAs against the
// resident stacked [W_j^T hi ; W_j^T lo] operand (K-major for the contraction over the layer's outputs) and the 3xTF32
// product comes back through tcgen05.ld for the ReLU mask of the layer below.  No shared-memory traffic for activations;
// what remains is HBM: 3 x 16.8 MB of masks read + 3 x 16.8 MB of dz written at config 2 (~16 us at the measured copy rate).
// Two CTAs per SM (256 tensor-memory columns each) overlap one tile's load / MMA latency with the other's epilogue.
// Column sums (bias gradients, dw_out): a 4-stage exchange butterfly over the 16 values a lane holds per chunk leaves lane l
// with the sum of column (l & 15) over the warp's 32 rows: 16 shuffles per chunk instead of 80.
#include "tc_ptx.cuh"
#include "tower_tile.cuh"

namespace rpb {

constexpr int TB_THREADS = 320;                         // warp 0: weight loader, warp 1: MMA issuer, warps 2-9: compute
constexpr int TB_KB_BYTES = 2 * TW_H * TC_BLOCK_K * 4;  // one k-block of a stacked operand: 128 rows x 128 B = 16 KiB
constexpr int TB_LAYER_BYTES = 2 * TB_KB_BYTES;         // K = 64 = two k-blocks
constexpr uint32_t TB_ACC = 0, TB_A = 128;              // tensor-memory columns: accumulator [0,128), operand hi [128,192) | lo [192,256)

// v[i] (i < 16) = this lane's value of column i.  Returns, in every lane l, the sum over the 32 lanes of column (l & 15).
__device__ __forceinline__ float colsum16(float (&v)[16], int lane) {
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float recv = __shfl_xor_sync(0xffffffffu, send, off);
            v[i] = (up ? v[i + off] : v[i]) + recv;
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

__global__ void __launch_bounds__(TB_THREADS, 2)
tower_tail_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmThi, const __grid_constant__ CUtensorMap tmTlo,
                         const __grid_constant__ TowerBwdParams p, int tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* t_base = smem;                                            // n_tail x 32 KiB resident operands (SWIZZLE_128B)
    uint64_t* bars = reinterpret_cast<uint64_t*>(t_base + (size_t)p.n_tail * TB_LAYER_BYTES);
    uint64_t* w_full = bars;                                           // weights landed
    uint64_t* a_ready = bars + 1;                                      // 256 arrivals: operand of the next MMA round written
    uint64_t* mma_done = bars + 2;                                     // that round's accumulator is complete
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = p.n_tail;
    const int my_tiles = ((int)blockIdx.x < tiles) ? (tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (threadIdx.x == 0) {
        mbar_init(w_full, 1);
        mbar_init(a_ready, 256);
        mbar_init(mma_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(w_full, (uint32_t)(L * TB_LAYER_BYTES));
            for (int l = 0; l < L; ++l)
                for (int kb = 0; kb < 2; ++kb) {
                    uint8_t* st = t_base + (size_t)(l * 2 + kb) * TB_KB_BYTES;
                    tma_load_2d(st, &tmThi, w_full, kb * TC_BLOCK_K, l * TW_H);
                    tma_load_2d(st + TB_KB_BYTES / 2, &tmTlo, w_full, kb * TC_BLOCK_K, l * TW_H);
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TC_BLOCK_M, 2 * TW_H);
            mbar_wait(w_full, 0u);
            uint32_t n = 0;
            for (int t = 0; t < my_tiles; ++t)
                for (int j = L - 1; j >= 0; --j, ++n) {
                    mbar_wait(a_ready, n & 1u);
                    tc_fence_after();
                    const uint32_t tb = smem_u32(t_base + (size_t)j * TB_LAYER_BYTES);
#pragma unroll
                    for (int k = 0; k < TW_H / TC_UMMA_K; ++k) {
                        const uint64_t db = make_kmajor_sw128_desc(tb + (uint32_t)(k >> 2) * TB_KB_BYTES + (uint32_t)(k & 3) * TC_UMMA_K * 4);
                        umma_tf32_ts(tmem_base + TB_ACC, tmem_base + TB_A + 64u + k * TC_UMMA_K, db, idesc, k > 0 ? 1u : 0u);
                        umma_tf32_ts(tmem_base + TB_ACC, tmem_base + TB_A + k * TC_UMMA_K, db, idesc, 1u);
                    }
                    umma_commit(mma_done);
                }
        }
    } else {
        // ---------------- compute warps: thread = (row, half); columns {half*16 .. +15} and {32 + half*16 .. +15}
        const int quarter = warp & 3;                      // tensor-memory lane quarter this warp may access
        const int half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const float gscale = (p.gloss != nullptr ? __ldg(p.gloss) : 1.f) * p.scale / (float)p.M;
        const float* htop = p.hin[L];
        float cs[TW_MAX_TAIL + 1][2], csw[2];              // column sums: lane l <-> column chunk_base + (l & 15)
#pragma unroll
        for (int j = 0; j <= TW_MAX_TAIL; ++j) { cs[j][0] = 0.f; cs[j][1] = 0.f; }
        csw[0] = 0.f; csw[1] = 0.f;
        float dbo = 0.f;
        uint32_t n = 0;                                    // mma_done phases consumed

        auto split_store = [&](const float (&d)[16], int c0) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                hi[i] = __float_as_uint(d[i]) & 0xFFFFE000u;
                lo[i] = __float_as_uint(d[i] - __uint_as_float(hi[i]));
            }
            tmem_st16(tmem_base + TB_A + lane_addr + (uint32_t)c0, hi);
            tmem_st16(tmem_base + TB_A + 64u + lane_addr + (uint32_t)c0, lo);
        };

        for (int t = 0; t < my_tiles; ++t) {
            const int m = ((int)blockIdx.x + t * (int)gridDim.x) * TC_BLOCK_M + row;
            const bool ok = m < p.M;
            float dl = 0.f;
            if (ok) {
                if (p.dlogit_in != nullptr) {
                    dl = __ldg(p.dlogit_in + m);
                } else {   // ATen binary_cross_entropy_backward (denominator clamped at 1e-12) x sigmoid backward
                    const float q = __ldg(p.pred + m), y = __ldg(p.label + m);
                    const float pe = q + p.eps;
                    dl = gscale * (pe - y) / fmaxf((1.f - pe) * pe, 1e-12f) * q * (1.f - q);
                }
                if (half == 0) {
                    if (p.dlogit_out != nullptr) p.dlogit_out[m] = dl;
                    dbo += dl;
                }
            }
            // ---- output layer: dz_top = dlogit * w_out * (h_top > 0); dw_out partial = dlogit * h_top
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int c0 = half * 16 + 32 * u;
                float d[16], w[16];
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok) hv = ldg_f4_stream(htop + (size_t)m * TW_H + c0 + 4 * i4);
                    const float4 wo = ldg_f4(p.w_out + c0 + 4 * i4);
                    d[4 * i4 + 0] = hv.x > 0.f ? dl * wo.x : 0.f; d[4 * i4 + 1] = hv.y > 0.f ? dl * wo.y : 0.f;
                    d[4 * i4 + 2] = hv.z > 0.f ? dl * wo.z : 0.f; d[4 * i4 + 3] = hv.w > 0.f ? dl * wo.w : 0.f;
                    w[4 * i4 + 0] = dl * hv.x; w[4 * i4 + 1] = dl * hv.y; w[4 * i4 + 2] = dl * hv.z; w[4 * i4 + 3] = dl * hv.w;
                    if (ok) stg_f4(p.dz[L] + (size_t)m * TW_H + c0 + 4 * i4,
                                   make_float4(d[4 * i4], d[4 * i4 + 1], d[4 * i4 + 2], d[4 * i4 + 3]));
                }
                split_store(d, c0);
                cs[L][u] += colsum16(d, lane);
                csw[u] += colsum16(w, lane);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(a_ready);
            // ---- hidden tail layers, top down: dz_j = (dz_{j+1} . W_j) * (hin[j] > 0)
            for (int j = L - 1; j >= 0; --j) {
                const float* hj = p.hin[j];
                const long long ld = j == 0 ? p.ldh1 : (long long)TW_H;
                float4 hv[2][4];                           // ReLU-mask pieces, requested before the accumulator is awaited
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        hv[u][i4] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ok) hv[u][i4] = ldg_f4_stream(hj + (size_t)m * ld + half * 16 + 32 * u + 4 * i4);
                    }
                mbar_wait(mma_done, n & 1u);
                ++n;
                tc_fence_after();
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int c0 = half * 16 + 32 * u;
                    uint32_t a0[16], a1[16];
                    tmem_ld16(tmem_base + TB_ACC + lane_addr + (uint32_t)c0, a0);
                    tmem_ld16(tmem_base + TB_ACC + lane_addr + (uint32_t)(TW_H + c0), a1);
                    float d[16];
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        const float4 h4 = hv[u][i4];
                        const float hm[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int i = 4 * i4 + c;
                            d[i] = hm[c] > 0.f ? __uint_as_float(a0[i]) + __uint_as_float(a1[i]) : 0.f;
                        }
                        if (ok) stg_f4(p.dz[j] + (size_t)m * TW_H + c0 + 4 * i4,
                                       make_float4(d[4 * i4], d[4 * i4 + 1], d[4 * i4 + 2], d[4 * i4 + 3]));
                    }
                    if (j > 0) split_store(d, c0);
                    cs[j][u] += colsum16(d, lane);
                }
                tc_fence_before();                         // accumulator reads are done before the next arrive releases it
                if (j > 0) {
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(a_ready);
                }
            }
        }
        // ---- reductions out: lanes 0-15 of every warp own one column per chunk; db_out from the half-0 warps
        if (lane < 16) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int col = half * 16 + 32 * u + lane;
                for (int j = 0; j <= L; ++j)
                    if (p.db[j] != nullptr) red_add_f1(p.db[j] + col, cs[j][u]);
                if (p.dw_out != nullptr) red_add_f1(p.dw_out + col, csw[u]);
            }
        }
        if (half == 0) {
            dbo = warp_sum(dbo);
            if (lane == 0 && p.db_out != nullptr) red_add_f1(p.db_out, dbo);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

int tower_tail_bwd_tc(const TowerBwdParams& p, cudaStream_t st) {
    if (p.n_tail < 1 || p.n_tail > 3 || p.M < 512) return RPB_ERR_UNSUPPORTED;
    CUtensorMap tmThi, tmTlo;
    int rc = tc_prepare_tail_weights(p.W, p.n_tail, &tmThi, &tmTlo, st, /*transpose=*/1, /*slot=*/7);
    if (rc != 0) return rc;
    const size_t smem = (size_t)p.n_tail * TB_LAYER_BYTES + 4 * 8 + 16 + 1024;
    cudaError_t e = cudaFuncSetAttribute(tower_tail_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int tiles = ceil_div(p.M, TC_BLOCK_M);
    tower_tail_bwd_tc_kernel<<<min(tiles, 2 * 148), TB_THREADS, smem, st>>>(tmThi, tmTlo, p, tiles);
    return (int)cudaGetLastError();
}

}  // namespace rpb
