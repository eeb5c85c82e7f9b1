// Shared device code of the MLP tower tail (tower.cu: standalone kernels; linear_tc.cu: the same tile routine run by the
// epilogue warps of the layer-1 tcgen05 GEMM, so that h1 never leaves the SM between layer 1 and the rest of the tower).
#pragma once
#include "common.cuh"

namespace rpb {

constexpr int TW_H = 64;                 // hidden width these routines are built for
constexpr int TW_LDA = TW_H + 4;         // padded activation row: conflict-free float4 rows
constexpr int TW_MAX_TAIL = RPB_TOWER_MAX_TAIL;
constexpr int TW_MAX_GRID = 2048;        // loss partials live in the caller's work buffer (2048 floats)

struct TowerFwdParams {
    const float* h1; long long ldh1;
    const float* W[TW_MAX_TAIL]; const float* b[TW_MAX_TAIL];
    float* h[TW_MAX_TAIL];
    const float* w_out; const float* b_out; const float* addend;
    float* logit; const float* label; float* pred; float* loss;
    float eps, scale; unsigned int* counter; float* partials;
    int M, n_tail, enabled;
};

struct TowerBwdParams {
    const float* hin[TW_MAX_TAIL + 1]; long long ldh1;       // hin[0] = h1 (row stride ldh1), hin[j>0] row stride 64
    const float* W[TW_MAX_TAIL]; const float* w_out;
    float* dz[TW_MAX_TAIL + 1]; float* db[TW_MAX_TAIL + 1];
    float* dw_out; float* db_out;
    const float* pred; const float* label; const float* gloss; float eps, scale;
    const float* dlogit_in; float* dlogit_out;
    int M, n_tail;
};
int tower_tail_bwd_tc(const TowerBwdParams& p, cudaStream_t st);       // tower_tc.cu (opt-in tcgen05 backward tail)

// NT threads cooperate on a tile of NT/2 rows: thread (ty = t >> 4, tx = t & 15) owns rows ty + (NT/16)*i, i < 8, and
// columns tx*4 .. tx*4+3.
// acc[i][c] += sum_k As[ty + RG i][k] * Bs[k][tx * 4 + c]     (As row stride TW_LDA, Bs row stride TW_H)
template <int RG>
__device__ __forceinline__ void tile_fma(const float* __restrict__ As, const float* __restrict__ Bs, int ty, int tx,
                                         float (&acc)[8][4]) {
#pragma unroll 4
    for (int k0 = 0; k0 < TW_H; k0 += 4) {
        float4 a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(As + (ty + RG * i) * TW_LDA + k0);
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(Bs + (k0 + j) * TW_H + tx * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[i][0] = fmaf(av[j], b[j].x, acc[i][0]);
                acc[i][1] = fmaf(av[j], b[j].y, acc[i][1]);
                acc[i][2] = fmaf(av[j], b[j].z, acc[i][2]);
                acc[i][3] = fmaf(av[j], b[j].w, acc[i][3]);
            }
        }
    }
}

// Transposed weight load Bs[l][k][n] = W_l[n][k] by NT threads: consecutive lanes take consecutive output rows n, so the
// four scalar stores of a float4 (k..k+3 of row n) land in consecutive banks (the naive "coalesced read, strided store"
// is a 32-way bank conflict).
template <int NT>
__device__ __forceinline__ void tower_load_weights_t(const TowerFwdParams& p, float* Bs, int tid) {
    for (int l = 0; l < p.n_tail; ++l)
        for (int i = tid; i < TW_H * TW_H / 4; i += NT) {
            const int n = i & 63, k4 = i >> 6;
            const float4 v = ldg_f4(p.W[l] + n * TW_H + k4 * 4);
            float* dst = Bs + l * TW_H * TW_H + (k4 * 4) * TW_H + n;
            dst[0] = v.x; dst[TW_H] = v.y; dst[2 * TW_H] = v.z; dst[3 * TW_H] = v.w;
        }
}

// One tile of the forward tail.  On entry As holds the post-ReLU layer-1 tile (rows m0 .. m0+NT/2-1, zero beyond M) and
// the caller has synchronised the NT threads; `sync()` is a barrier over exactly those NT threads.  Runs the n_tail
// hidden layers (h[l] -> HBM, activations stay in As), then logit / pred / BCE term of each row.  `addend_local`, when
// given, replaces p.addend: NT/2 floats indexed by the row inside the tile (shared memory).
template <int NT, class Sync>
__device__ __forceinline__ void tower_tail_tile_fwd(const TowerFwdParams& p, float* As, const float* Bs, int m0, int tid,
                                                    const float4 wo, const float bo, float& loss_acc, Sync sync,
                                                    const float* addend_local = nullptr) {
    constexpr int RG = NT / 16;
    const int tx = tid & 15, ty = tid >> 4;
    for (int l = 0; l < p.n_tail; ++l) {
        float acc[8][4];
        const float4 bv = ldg_f4(p.b[l] + tx * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i][0] = bv.x; acc[i][1] = bv.y; acc[i][2] = bv.z; acc[i][3] = bv.w; }
        tile_fma<RG>(As, Bs + l * TW_H * TW_H, ty, tx, acc);
        sync();                                       // every thread has finished reading the layer input
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 v = make_float4(fmaxf(acc[i][0], 0.f), fmaxf(acc[i][1], 0.f), fmaxf(acc[i][2], 0.f),
                                         fmaxf(acc[i][3], 0.f));
            const int r = ty + RG * i;
            if (m0 + r < p.M) stg_f4(p.h[l] + (size_t)(m0 + r) * TW_H + tx * 4, v);
            *reinterpret_cast<float4*>(As + r * TW_LDA + tx * 4) = v;
        }
        sync();
    }
    // head: logit = h_last . w_out + b_out (+ addend); 16 lanes share a row
    float part[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(As + (ty + RG * i) * TW_LDA + tx * 4);
        part[i] = fmaf(v.x, wo.x, fmaf(v.y, wo.y, fmaf(v.z, wo.z, v.w * wo.w)));
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) part[i] += __shfl_xor_sync(0xffffffffu, part[i], o);
    }
    // the butterfly left every lane of a 16-lane group with all 8 row sums: lane tx < 8 finishes row ty + RG*tx (one
    // exp/log sequence per warp instead of eight serial ones on two active lanes)
    float mine = part[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mine = (tx == i) ? part[i] : mine;
    const int m = m0 + ty + RG * tx;
    if (tx < 8 && m < p.M) {
        float z = mine + bo;
        if (addend_local != nullptr) z += addend_local[ty + RG * tx];      // per-tile addend in shared memory (fused kernel)
        else if (p.addend != nullptr) z += __ldg(p.addend + m);
        p.logit[m] = z;
        if (p.pred != nullptr) {
            const float q = 1.f / (1.f + expf(-z));
            p.pred[m] = q;
            if (p.label != nullptr) {
                const float y = __ldg(p.label + m);
                const float pe = q + p.eps;
                const float l1 = fmaxf(logf(pe), -100.f);
                const float l0 = fmaxf(logf(1.f - pe), -100.f);
                loss_acc += -(y * l1 + (1.f - y) * l0);
            }
        }
    }
}

// Fills TowerFwdParams from the C-ABI descriptor (validation included); h1 / ldh1 are taken as given.
int tower_fwd_params(const RpbTowerFwdDesc* d, TowerFwdParams* out);

}  // namespace rpb
