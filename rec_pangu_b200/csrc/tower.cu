// MLP tower "tail": every 64-wide hidden layer after the first one, the Linear(64 -> 1) output, the logit sum, the
// sigmoid and the BCE loss of the ranking models in ONE kernel per direction (reference: models/layers/deep.py:62-84
// for the Linear/ReLU chain, ranking/deepfm.py:57-63 for `y_pred = fm + dnn; sigmoid; BCELoss`).
//
// Why not tcgen05 here: a 64x64 layer over a 65536-sample batch is 0.5 GFLOP against 33 MB of activations, and run
// as separate GEMM launches each such layer cost 17-25 us of pipeline fill / epilogue / weight-split overhead
// (profiles/r01_bench_launches.csv: 4 GEMMs + 3 column-sum + 4 split + 5 head kernels = 160 us of a 546 us step).
// The chain below keeps a 64-row tile of activations in shared memory, multiplies it by the resident 16 KiB weight
// matrix with register-tiled fp32 FMAs (8 rows x 4 columns per thread, float4 operands: 12 LDS.128 per 128 FMA), and
// touches HBM exactly once per activation: read h1, write h2..hL (saved for backward) and the logit.  It is exact
// fp32 (no TF32 split), so it is also the more accurate path.
//
// Backward is the mirror image: dlogit from (pred, label) -> masked outer product with w_out -> per layer
// dz_{l-1} = (dz_l . W_l) * (h_{l-1} > 0), with the bias gradients (column sums), dw_out and db_out accumulated on
// the way; the dz tiles are written for the weight-gradient kernels and for the layer-1 dx + scatter GEMM.
#include "tower_tile.cuh"

namespace rpb {

constexpr int TW_ROWS = 64;              // samples per tile of the standalone kernels
constexpr int TW_THREADS = 128;          // 8 row groups x 16 column groups

__global__ void __launch_bounds__(TW_THREADS, 4)
tower_tail_fwd_kernel(const TowerFwdParams p) {
    extern __shared__ __align__(16) float tw_smem[];
    float* As = tw_smem;                              // [TW_ROWS][TW_LDA]
    float* Bs = tw_smem + TW_ROWS * TW_LDA;           // [n_tail][k][n] = W_l[n][k]
    __shared__ float red[32];
    __shared__ bool is_last;
    const int tid = threadIdx.x, tx = tid & 15;

    tower_load_weights_t<TW_THREADS>(p, Bs, tid);
    const float4 wo = ldg_f4(p.w_out + tx * 4);
    const float bo = p.b_out != nullptr ? __ldg(p.b_out) : 0.f;
    float loss_acc = 0.f;
    const int tiles = (p.M + TW_ROWS - 1) / TW_ROWS;

    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int m0 = tile * TW_ROWS;
        __syncthreads();                              // previous tile fully consumed (and Bs visible on the first pass)
#pragma unroll
        for (int i = 0; i < TW_ROWS * 16 / TW_THREADS; ++i) {
            const int e = tid + i * TW_THREADS, r = e >> 4, c4 = e & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + r < p.M) v = ldg_f4_stream(p.h1 + (size_t)(m0 + r) * p.ldh1 + c4 * 4);
            *reinterpret_cast<float4*>(As + r * TW_LDA + c4 * 4) = v;
        }
        __syncthreads();
        tower_tail_tile_fwd<TW_THREADS>(p, As, Bs, m0, tid, wo, bo, loss_acc, [] { __syncthreads(); });
    }
    if (p.loss == nullptr) return;
    // deterministic mean: per-CTA partial, the last CTA to finish adds them in index order (same as head.cu)
    const float t = block_sum(loss_acc, red);
    if (tid == 0) {
        p.partials[blockIdx.x] = t;
        __threadfence();
        const unsigned int done = atomicAdd(p.counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        float s = 0.f;
        for (int i = tid; i < (int)gridDim.x; i += TW_THREADS) s += ((volatile float*)p.partials)[i];
        s = block_sum(s, red);
        if (tid == 0) {
            p.loss[0] = p.scale * (s / (float)p.M);
            *p.counter = 0u;
        }
    }
}

// Adds this thread's 4 column partials (columns tx*4..tx*4+3) to the warp's private row: the two 16-lane halves of a
// warp own the same columns, so one shuffle folds them and lanes 0-15 do a plain read-modify-write.
__device__ __forceinline__ void colsum_add(float4* row, int lane, float s0, float s1, float s2, float s3) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, 16); s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    s2 += __shfl_xor_sync(0xffffffffu, s2, 16); s3 += __shfl_xor_sync(0xffffffffu, s3, 16);
    if (lane < 16) {
        float4 v = row[lane];
        v.x += s0; v.y += s1; v.z += s2; v.w += s3;
        row[lane] = v;
    }
}

// 16-byte asynchronous global -> shared copy (zero-fill when `valid` is false)
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Latency plan (the first version spent most of its time on exposed global loads — ncu: long-scoreboard stalls
// dominated): the top activation tile of the NEXT tile is copied into Hs with cp.async while the current tile's layers
// run, its (pred, label) / dlogit values are loaded one tile ahead into registers, and the ReLU-mask tile of each
// hidden layer is loaded into registers BEFORE that layer's FMA block.  3 CTAs per SM (register budget 170).
__global__ void __launch_bounds__(TW_THREADS, 3)
tower_tail_bwd_kernel(const TowerBwdParams p) {
    extern __shared__ __align__(16) float tw_smem[];
    float* As = tw_smem;                              // [TW_ROWS][TW_LDA]: dz of the layer above
    float* Hs = As + TW_ROWS * TW_LDA;                // [TW_ROWS][TW_LDA]: top activation tile (prefetched)
    float* Bs = Hs + TW_ROWS * TW_LDA;                // [n_tail][n][k] = W_l[n][k]
    // column sums (bias gradients, dw_out): one private row per warp, so no atomics: [j] = db[j], [TW_MAX_TAIL + 1] = dw_out
    __shared__ float4 cs[TW_MAX_TAIL + 2][TW_THREADS / 32][TW_H / 4];
    __shared__ float dbo_s;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, warp = tid >> 5, lane = tid & 31;
    const int tiles = (p.M + TW_ROWS - 1) / TW_ROWS;
    const float* htop = p.hin[p.n_tail];
    const long long ldtop = p.n_tail == 0 ? p.ldh1 : (long long)TW_H;

    auto prefetch_top = [&](int tile) {               // Hs <- rows of the top activation of `tile`
        const int m0 = tile * TW_ROWS;
#pragma unroll
        for (int i = 0; i < TW_ROWS * 16 / TW_THREADS; ++i) {
            const int e = tid + i * TW_THREADS, r = e >> 4, c4 = e & 15;
            const bool ok = m0 + r < p.M;
            cp_async16(Hs + r * TW_LDA + c4 * 4, htop + (size_t)(ok ? m0 + r : 0) * ldtop + c4 * 4, ok);
        }
    };
    // (q, y) = (pred, label) or (dlogit_in, unused) of this thread's 8 rows, one tile ahead
    float qn[8], yn[8];
    auto load_dl_inputs = [&](int tile) {
        const int m0 = tile * TW_ROWS;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + ty + 8 * i;
            qn[i] = 0.f; yn[i] = 0.f;
            if (m < p.M) {
                if (p.dlogit_in != nullptr) qn[i] = __ldg(p.dlogit_in + m);
                else { qn[i] = __ldg(p.pred + m); yn[i] = __ldg(p.label + m); }
            }
        }
    };

    prefetch_top(blockIdx.x);
    load_dl_inputs(blockIdx.x);
    for (int l = 0; l < p.n_tail; ++l)
        for (int i = tid; i < TW_H * TW_H / 4; i += TW_THREADS)
            reinterpret_cast<float4*>(Bs + l * TW_H * TW_H)[i] = __ldg(reinterpret_cast<const float4*>(p.W[l]) + i);
    for (int i = tid; i < (TW_MAX_TAIL + 2) * (TW_THREADS / 32) * (TW_H / 4); i += TW_THREADS)
        (&cs[0][0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) dbo_s = 0.f;
    const float4 wo = ldg_f4(p.w_out + tx * 4);
    const float gscale = (p.gloss != nullptr ? __ldg(p.gloss) : 1.f) * p.scale / (float)p.M;
    float dbo = 0.f;

    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int m0 = tile * TW_ROWS;
        float dl[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (p.dlogit_in != nullptr) {
                dl[i] = qn[i];
            } else {
                // ATen binary_cross_entropy_backward (denominator clamped at 1e-12) x sigmoid backward
                const float pe = qn[i] + p.eps;
                dl[i] = gscale * (pe - yn[i]) / fmaxf((1.f - pe) * pe, 1e-12f) * qn[i] * (1.f - qn[i]);
            }
        }
        cp_async_wait_all();
        __syncthreads();                              // Hs landed; the previous tile is done with As
        // ---- output layer: dz_top[r, n] = dlogit[r] * w_out[n] * (h_top[r, n] > 0)
        {
            const int j = p.n_tail;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = ty + 8 * i, m = m0 + r;
                float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < p.M) {
                    if (tx == 0) {
                        if (p.dlogit_out != nullptr) p.dlogit_out[m] = dl[i];
                        dbo += dl[i];
                    }
                    const float4 hv = *reinterpret_cast<const float4*>(Hs + r * TW_LDA + tx * 4);
                    w0 = fmaf(dl[i], hv.x, w0); w1 = fmaf(dl[i], hv.y, w1); w2 = fmaf(dl[i], hv.z, w2); w3 = fmaf(dl[i], hv.w, w3);
                    d.x = hv.x > 0.f ? dl[i] * wo.x : 0.f; d.y = hv.y > 0.f ? dl[i] * wo.y : 0.f;
                    d.z = hv.z > 0.f ? dl[i] * wo.z : 0.f; d.w = hv.w > 0.f ? dl[i] * wo.w : 0.f;
                    s0 += d.x; s1 += d.y; s2 += d.z; s3 += d.w;
                    stg_f4(p.dz[j] + (size_t)m * TW_H + tx * 4, d);
                }
                *reinterpret_cast<float4*>(As + r * TW_LDA + tx * 4) = d;
            }
            colsum_add(cs[j][warp], lane, s0, s1, s2, s3);
            colsum_add(cs[TW_MAX_TAIL + 1][warp], lane, w0, w1, w2, w3);
        }
        __syncthreads();                              // As complete, Hs fully consumed
        if (tile + (int)gridDim.x < tiles) {          // next tile's inputs fly while this tile's layers run
            prefetch_top(tile + gridDim.x);
            load_dl_inputs(tile + gridDim.x);
        }
        // ---- hidden tail layers, top down: dz_j = (dz_{j+1} . W_j) * (hin[j] > 0)
        for (int j = p.n_tail - 1; j >= 0; --j) {
            const float* hj = p.hin[j];
            const long long ld = j == 0 ? p.ldh1 : (long long)TW_H;
            float4 hv[8];                             // ReLU-mask tile, requested before the FMA block that hides it
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = m0 + ty + 8 * i;
                hv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < p.M) hv[i] = ldg_f4_stream(hj + (size_t)m * ld + tx * 4);
            }
            float acc[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
            tile_fma<TW_THREADS / 16>(As, Bs + j * TW_H * TW_H, ty, tx, acc);
            __syncthreads();                          // dz_{j+1} fully consumed
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = ty + 8 * i, m = m0 + r;
                float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < p.M) {
                    d.x = hv[i].x > 0.f ? acc[i][0] : 0.f; d.y = hv[i].y > 0.f ? acc[i][1] : 0.f;
                    d.z = hv[i].z > 0.f ? acc[i][2] : 0.f; d.w = hv[i].w > 0.f ? acc[i][3] : 0.f;
                    s0 += d.x; s1 += d.y; s2 += d.z; s3 += d.w;
                    stg_f4(p.dz[j] + (size_t)m * TW_H + tx * 4, d);
                }
                if (j > 0) *reinterpret_cast<float4*>(As + r * TW_LDA + tx * 4) = d;
            }
            colsum_add(cs[j][warp], lane, s0, s1, s2, s3);
            __syncthreads();
        }
    }
    cp_async_wait_all();
    if (tx == 0 && dbo != 0.f) atomicAdd(&dbo_s, dbo);
    __syncthreads();
    if (tid < TW_H) {
        auto total = [&](int j) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < TW_THREADS / 32; ++w) t += reinterpret_cast<const float*>(cs[j][w])[tid];
            return t;
        };
        for (int j = 0; j <= p.n_tail; ++j)
            if (p.db[j] != nullptr) red_add_f1(p.db[j] + tid, total(j));
        if (p.dw_out != nullptr) red_add_f1(p.dw_out + tid, total(TW_MAX_TAIL + 1));
    }
    if (tid == 0 && p.db_out != nullptr) red_add_f1(p.db_out, dbo_s);
}

static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        n = min(n, TW_MAX_GRID / 4);
    }
    return n;
}

}  // namespace rpb

using namespace rpb;

namespace rpb {
int tower_fwd_params(const RpbTowerFwdDesc* d, TowerFwdParams* out) {
    if (d == nullptr || d->M <= 0 || d->h1 == nullptr || d->w_out == nullptr || d->logit == nullptr) return RPB_ERR_BAD_ARG;
    if (d->H != TW_H || d->n_tail < 0 || d->n_tail > TW_MAX_TAIL || (d->ldh1 % 4) != 0 || !aligned16(d->h1) || !aligned16(d->w_out))
        return RPB_ERR_UNSUPPORTED;
    if (d->loss != nullptr && (d->label == nullptr || d->pred == nullptr || d->work == nullptr)) return RPB_ERR_BAD_ARG;
    TowerFwdParams p{};
    p.h1 = d->h1; p.ldh1 = d->ldh1;
    for (int l = 0; l < d->n_tail; ++l) {
        if (d->W == nullptr || d->b == nullptr || d->h == nullptr || d->W[l] == nullptr || d->b[l] == nullptr || d->h[l] == nullptr)
            return RPB_ERR_BAD_ARG;
        if (!aligned16(d->W[l]) || !aligned16(d->b[l]) || !aligned16(d->h[l])) return RPB_ERR_UNSUPPORTED;
        p.W[l] = d->W[l]; p.b[l] = d->b[l]; p.h[l] = d->h[l];
    }
    p.w_out = d->w_out; p.b_out = d->b_out; p.addend = d->addend;
    p.logit = d->logit; p.label = d->label; p.pred = d->pred; p.loss = d->loss;
    p.eps = d->eps; p.scale = d->scale;
    p.counter = reinterpret_cast<unsigned int*>(d->work);
    p.partials = reinterpret_cast<float*>(d->work) + 2;
    p.M = d->M; p.n_tail = d->n_tail; p.enabled = 1;
    *out = p;
    return 0;
}
}  // namespace rpb

RPB_API int rpb_tower_tail_fwd(const RpbTowerFwdDesc* d, void* stream) {
    TowerFwdParams p{};
    const int rc = tower_fwd_params(d, &p);
    if (rc != 0) return rc;
    const size_t smem = (size_t)(TW_ROWS * TW_LDA + d->n_tail * TW_H * TW_H) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(tower_tail_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)      // 4 CTAs x ~50 KiB per SM only fit with the L1/shared split at its shared-memory maximum
        e = cudaFuncSetAttribute(tower_tail_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    const int tiles = ceil_div(d->M, TW_ROWS);
    // persistent: 4 CTAs per SM walk the tile list, so the transposed weight load is paid once per CTA
    tower_tail_fwd_kernel<<<min(tiles, 4 * sm_count()), TW_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    RPB_LAUNCH_CHECK();
    return 0;
}

RPB_API int rpb_tower_tail_bwd(const RpbTowerBwdDesc* d, void* stream) {
    if (d == nullptr || d->M <= 0 || d->hin == nullptr || d->dz == nullptr || d->w_out == nullptr) return RPB_ERR_BAD_ARG;
    if (d->H != TW_H || d->n_tail < 0 || d->n_tail > TW_MAX_TAIL || (d->ldh1 % 4) != 0 || !aligned16(d->w_out))
        return RPB_ERR_UNSUPPORTED;
    if (d->dlogit_in == nullptr && (d->pred == nullptr || d->label == nullptr)) return RPB_ERR_BAD_ARG;
    TowerBwdParams p{};
    p.ldh1 = d->ldh1;
    for (int j = 0; j <= d->n_tail; ++j) {
        if (d->hin[j] == nullptr || d->dz[j] == nullptr) return RPB_ERR_BAD_ARG;
        if (!aligned16(d->hin[j]) || !aligned16(d->dz[j])) return RPB_ERR_UNSUPPORTED;
        p.hin[j] = d->hin[j]; p.dz[j] = d->dz[j];
        p.db[j] = d->db != nullptr ? d->db[j] : nullptr;
    }
    for (int l = 0; l < d->n_tail; ++l) {
        if (d->W == nullptr || d->W[l] == nullptr) return RPB_ERR_BAD_ARG;
        if (!aligned16(d->W[l])) return RPB_ERR_UNSUPPORTED;
        p.W[l] = d->W[l];
    }
    p.w_out = d->w_out; p.dw_out = d->dw_out; p.db_out = d->db_out;
    p.pred = d->pred; p.label = d->label; p.gloss = d->gloss; p.eps = d->eps; p.scale = d->scale;
    p.dlogit_in = d->dlogit_in; p.dlogit_out = d->dlogit_out;
    p.M = d->M; p.n_tail = d->n_tail;
    if (g_tower_bwd_tc) {          // opt-in: the dz chain on tcgen05 (tower_tc.cu); declines shapes it is not built for
        const int rc = tower_tail_bwd_tc(p, reinterpret_cast<cudaStream_t>(stream));
        if (rc != RPB_ERR_UNSUPPORTED) return rc;
    }
    const size_t smem = (size_t)(2 * TW_ROWS * TW_LDA + d->n_tail * TW_H * TW_H) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(tower_tail_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)      // 4 CTAs x ~50 KiB per SM only fit with the L1/shared split at its shared-memory maximum
        e = cudaFuncSetAttribute(tower_tail_bwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    const int tiles = ceil_div(d->M, TW_ROWS);
    tower_tail_bwd_kernel<<<min(tiles, 3 * sm_count()), TW_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    RPB_LAUNCH_CHECK();
    return 0;
}
