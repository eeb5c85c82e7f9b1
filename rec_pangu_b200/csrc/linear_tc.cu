// tcgen05 / TMEM / TMA dense layer for sm_100a with fp32-grade accuracy ("3xTF32").
//
// C[M,N] = epi( A[M,K] . B[N,K]^T ),  A,B fp32 K-major in HBM.  The reference computes this contraction in
// full fp32 (aten::addmm, models/layers/deep.py:62-70); a single TF32 pass loses ~1e-3 on the logits
// (SURVEY.md §7 hard-part 2), so every operand is split x = hi + lo with hi = x & 0xFFFFE000 (exact TF32)
// and three tensor-core products are accumulated in TMEM:  hi*hi + lo*hi + hi*lo  (error ~2^-22 relative).
//
// CTA = one 128-row M tile x BLOCK_N (<=256) columns, warp-specialised:
//   warp 0   : TMA producer — A tile [128 x 32 fp32] and pre-split B_hi/B_lo tiles [BLOCK_N x 32], SWIZZLE_128B
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (kind::tf32, M=128, N=BLOCK_N, K=8)
//   warps 2-5: split A in shared memory (hi in place, lo to a twin buffer), then the epilogue:
//              tcgen05.ld 32x32b -> bias / ReLU / ReLU-mask -> global
// Pipeline barriers per stage: full (TMA landed) -> ready (split done, fenced to the async proxy) ->
// empty (tcgen05.commit after the stage's 12 MMAs).
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tower_tile.cuh"

namespace rpb {

template <int kStages>
__global__ void __launch_bounds__(TC_THREADS)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                   const __grid_constant__ CUtensorMap tmBlo, const TcEpilogue ep, int block_n, int num_k_blocks,
                   uint32_t tmem_cols) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by SWIZZLE_128B; the dynamic smem base is only guaranteed 16 B
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the pointer in the shared address space (LDS/STS, not generic LD/ST)
    const int b_bytes = block_n * TC_BLOCK_K * 4;
    const int stage_bytes = 2 * TC_A_BYTES + 2 * b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* ready_bar = bars + kStages;
    uint64_t* empty_bar = bars + 2 * kStages;
    uint64_t* tmem_full_bar = bars + 3 * kStages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TC_BLOCK_M;
    const int n0 = blockIdx.y * block_n;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&ready_bar[s], 128); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tx_bytes = (uint32_t)(TC_A_BYTES + 2 * b_bytes);
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (uint32_t)((kb / kStages) & 1);
                mbar_wait(&empty_bar[s], ph ^ 1u);
                uint8_t* st = smem + (size_t)s * stage_bytes;
                mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
                tma_load_2d(st, &tmA, &full_bar[s], kb * TC_BLOCK_K, m0);
                tma_load_2d(st + 2 * TC_A_BYTES, &tmBhi, &full_bar[s], kb * TC_BLOCK_K, n0);
                tma_load_2d(st + 2 * TC_A_BYTES + b_bytes, &tmBlo, &full_bar[s], kb * TC_BLOCK_K, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TC_BLOCK_M, block_n);
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (uint32_t)((kb / kStages) & 1);
                mbar_wait(&ready_bar[s], ph);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t a_lo = a_hi + TC_A_BYTES;
                const uint32_t b_hi = a_hi + 2 * TC_A_BYTES;
                const uint32_t b_lo = b_hi + b_bytes;
#pragma unroll
                for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
                    const uint32_t koff = k * TC_UMMA_K * 4;      // 32 B per K step inside the 128 B swizzle row
                    const uint64_t da_hi = make_kmajor_sw128_desc(a_hi + koff);
                    const uint64_t da_lo = make_kmajor_sw128_desc(a_lo + koff);
                    const uint64_t db_hi = make_kmajor_sw128_desc(b_hi + koff);
                    const uint64_t db_lo = make_kmajor_sw128_desc(b_lo + koff);
                    umma_tf32(tmem_base, da_lo, db_hi, idesc, (kb > 0 || k > 0) ? 1u : 0u);   // small terms first
                    umma_tf32(tmem_base, da_hi, db_lo, idesc, 1u);
                    umma_tf32(tmem_base, da_hi, db_hi, idesc, 1u);
                }
                umma_commit(&empty_bar[s]);          // frees the stage once these MMAs have read smem
            }
            umma_commit(tmem_full_bar);              // accumulator complete
        }
    } else {
        // ---------------- split warps: A -> (hi, lo)
        const int t = threadIdx.x - 64;              // 0..127
        for (int kb = 0; kb < num_k_blocks; ++kb) {
            const int s = kb % kStages;
            const uint32_t ph = (uint32_t)((kb / kStages) & 1);
            mbar_wait(&full_bar[s], ph);
            float4* a = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes);
            float4* lo = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + TC_A_BYTES);
#pragma unroll
            for (int i = 0; i < TC_A_BYTES / 16 / 128; ++i) {
                const int j = t + i * 128;
                const float4 v = a[j];
                float4 h, l;
                h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
                h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
                h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
                h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
                a[j] = h;
                lo[j] = l;
            }
            fence_proxy_async();                     // generic-proxy writes -> visible to the tensor core (async proxy)
            mbar_arrive(&ready_bar[s]);
        }
        // ---------------- epilogue: TMEM -> registers -> smem staging (thread = row) -> coalesced global (warp = row)
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int quarter = warp & 3;                // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;
        // all MMAs have completed (tcgen05.commit), so the operand pipeline buffers can be reused as staging
        float* stage = reinterpret_cast<float*>(smem);
        const int pitch = block_n + 4;               // +4 floats: conflict-free 16-byte row writes
        for (int c0 = 0; c0 < block_n; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
            float4* dst = reinterpret_cast<float4*>(stage + (size_t)row * pitch + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                     __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        }
        __syncwarp();
        const int vpr = block_n / 4;                 // float4 per row
        const bool c_vec = ((ep.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.C) & 15u) == 0);
        const bool m_vec = ep.mask != nullptr && ((ep.ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.mask) & 15u) == 0);
        for (int i = lane; i < 32 * vpr; i += 32) {
            const int rr = i / vpr, c4 = i % vpr;
            const int m = m0 + quarter * 32 + rr;
            const int n = n0 + c4 * 4;
            if (m >= ep.M || n >= ep.N) continue;
            const float4 acc = *reinterpret_cast<const float4*>(stage + (size_t)(quarter * 32 + rr) * pitch + c4 * 4);
            float v[4] = {acc.x, acc.y, acc.z, acc.w};
            const bool full = n + 4 <= ep.N;
            if (ep.bias != nullptr) {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (n + j < ep.N) v[j] += __ldg(ep.bias + n + j);
            }
            if (ep.relu) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (ep.mask != nullptr) {
                const float* mrow = ep.mask + (size_t)m * ep.ldmask + n;
                if (m_vec && full) {
                    const float4 mk = __ldg(reinterpret_cast<const float4*>(mrow));
                    v[0] = mk.x > 0.f ? v[0] : 0.f; v[1] = mk.y > 0.f ? v[1] : 0.f;
                    v[2] = mk.z > 0.f ? v[2] : 0.f; v[3] = mk.w > 0.f ? v[3] : 0.f;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (n + j < ep.N) v[j] = (__ldg(mrow + j) > 0.f) ? v[j] : 0.f;
                }
            }
            float* crow = ep.C + (size_t)m * ep.ldc + n;
            if (c_vec && full) stg_f4(crow, make_float4(v[0], v[1], v[2], v[3]));
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (n + j < ep.N) crow[j] = v[j];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// ---------------------------------------------------------------------------------- persistent GEMM (v2)
// Same math as gemm_tf32x3_kernel, restructured so that nothing on the tensor-core critical path waits for HBM:
//   * persistent CTAs (one per SM) walk the tile list; the k-block counter runs across tiles so the rings never drain;
//   * a deep RAW ring (kRaw stages: TMA destination for the fp32 A tile and the pre-split B hi/lo chunks) decouples the
//     ~1-2 us load latency from the 2-stage OPERAND ring (A hi / A lo produced by the split warps);
//   * the accumulator is double-buffered in TMEM (2 x block_n columns) and drained by four dedicated epilogue warps,
//     so tile i's epilogue (bias / ReLU / mask / global stores) overlaps tile i+1's main loop.
// Warps: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2-5 = split, 6-13 = epilogue: two warps per TMEM lane quarter
// (quarter = warp & 3), which take alternate 16-column chunks of the accumulator.
//
// Operand placement (a_stages): tcgen05.mma reads BOTH operands of an SS-mode instruction from shared memory — 6 KiB per
// 128x64x8 tf32 MMA, and three MMAs per K step for the 3xTF32 products.  Together with the TMA writes and the split
// warps' own traffic that is ~150 KiB of shared-memory traffic per 32-wide k-block, and ncu shows the narrow layers
// (N = 64) bound by exactly that, at half the tensor-pipe rate (tools/exp/exp_mma.cu: 65 cycles per 128x64x8 MMA alone,
// ~128 in the SS pipeline).  With a_stages > 0 the split warps therefore write A hi / A lo straight into TENSOR MEMORY
// (tcgen05.st, lane = tile row, column = k) and the MMAs take A from TMEM (TS mode): shared memory then only carries
// the raw A tile once and the B chunks.  TMEM columns: [acc 0 | acc 1 | a_stages x (A hi 32 | A lo 32)].

// 4x4 transpose of float4 "elements" across each group of 4 lanes.  In: r[4i..4i+3] = element i of this lane's row
// (lane & 3 = row within the group).  Out: r[4j..4j+3] = element (lane & 3) of row j.  Two butterfly stages, 16 SHFL.
__device__ __forceinline__ void transpose4x4(uint32_t (&r)[16], int q) {
    const bool b0 = (q & 1) != 0, b1 = (q & 2) != 0;
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t send = b0 ? r[4 * i + c] : r[4 * (i + 1) + c];
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
            if (b0) r[4 * i + c] = recv; else r[4 * (i + 1) + c] = recv;
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t send = b1 ? r[4 * i + c] : r[4 * (i + 2) + c];
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 2);
            if (b1) r[4 * i + c] = recv; else r[4 * (i + 2) + c] = recv;
        }
    }
}

// Diagnostics (rpb_debug_tc_trace): per-role stall cycles of CTA 0 of the last v2 launch.
__device__ int g_tc_trace_on = 0;
__device__ unsigned long long g_tc_trace[16];
#define TC_TRACE_T() (trace ? clock64() : 0ll)

constexpr int V2_THREADS = 448;
constexpr int V2_EPI_WARPS = 8;
constexpr int V2_OP_STAGES = 2;       // shared-memory operand ring (SS mode)
constexpr int V2_MAX_OP = 4;          // tensor-memory operand ring (TS mode): up to 4 stages of 64 columns

template <int kRaw, bool kShardedScatter>
__global__ void __launch_bounds__(V2_THREADS, 1)
gemm_tf32x3_v2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                      const __grid_constant__ CUtensorMap tmBlo, const TcEpilogue ep, int block_n, int num_k_blocks,
                      int m_tiles, int n_tiles, uint32_t tmem_cols, int b_resident, int a_stages, int stack_n,
                      const __grid_constant__ TcScatter sc, const __grid_constant__ TowerFwdParams tw) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the pointer in the shared address space (LDS/STS, not generic LD/ST)
    const int b_bytes = block_n * TC_BLOCK_K * 4;
    // b_resident (n_tiles == 1 and the whole pre-split weight fits): B hi/lo of every k-block is loaded ONCE per CTA
    // and stays in shared memory for all of its tiles; the raw ring then carries only A and is released by the split
    // warps themselves.  Otherwise each raw stage is [A raw | B hi | B lo] and is released by the MMA commit.
    const int raw_bytes = b_resident ? TC_A_BYTES : TC_A_BYTES + 2 * b_bytes;
    uint8_t* raw_base = smem;
    uint8_t* op_base = smem + (size_t)kRaw * raw_bytes;                // [A hi | A lo] x V2_OP_STAGES
    const int n_op = a_stages > 0 ? a_stages : V2_OP_STAGES;          // operand ring depth (TMEM or shared memory)
    // stack_n (narrow layers, TS mode): B hi and B lo are adjacent in shared memory, so ONE descriptor of N = 2*block_n rows
    // covers both and one MMA yields a.b_hi in columns [0, block_n) and a.b_lo in [block_n, 2*block_n) of the accumulator.
    // A 128 x N x 8 tf32 MMA costs max(64, N/2) cycles (tools/exp/exp_mma.cu), i.e. N = 128 is as cheap as N = 64: the
    // k-step needs two MMAs (a_hi, a_lo) instead of three, and the epilogue adds the two accumulator halves.
    const uint32_t acc_stride = (uint32_t)(stack_n ? 2 * block_n : block_n);
    const uint32_t a_col = (2u * acc_stride + 31u) & ~31u;             // TS mode: first TMEM column of the A ring
    uint8_t* bres_base = op_base + (a_stages > 0 ? 0 : (size_t)V2_OP_STAGES * 2 * TC_A_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(bres_base + (b_resident ? (size_t)num_k_blocks * 2 * b_bytes : 0));
    uint64_t* full_raw = bars;                     // [kRaw]  TMA landed
    uint64_t* empty_raw = bars + kRaw;             // [kRaw]  MMAs that read B of this stage are done
    uint64_t* ready_op = bars + 2 * kRaw;          // [<=4]   split done
    uint64_t* empty_op = ready_op + V2_MAX_OP;     // [<=4]   MMAs that read A hi/lo of this stage are done
    uint64_t* tmem_full = empty_op + V2_MAX_OP;    // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint64_t* b_full = tmem_empty + 2;             // [1]     resident B landed
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_full + 1);
    // fused tower tail (tw.enabled): [As: 128 x TW_LDA | Bs: n_tail x 64 x 64 | loss partials: 256] floats after the barriers
    // (`bars` is 1024-byte aligned: every region before it is a multiple of 1 KiB; 2*kRaw + 13 barrier slots + the TMEM address)
    float* tw_As = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + (((2 * kRaw + 2 * V2_MAX_OP + 5) * 8 + 4 + 15) & ~15));
    float* tw_Bs = tw_As + TC_BLOCK_M * TW_LDA;
    float* tw_loss = tw_Bs + tw.n_tail * TW_H * TW_H;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = m_tiles * n_tiles;
    const bool trace = g_tc_trace_on != 0 && blockIdx.x == 0;
    const long long t_start = TC_TRACE_T();

    if (threadIdx.x == 0) {
        for (int s = 0; s < kRaw; ++s) { mbar_init(&full_raw[s], 1); mbar_init(&empty_raw[s], b_resident ? 128 : 1); }
        mbar_init(b_full, 1);
        for (int s = 0; s < V2_MAX_OP; ++s) { mbar_init(&ready_op[s], 128); mbar_init(&empty_op[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], V2_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tx_bytes = (uint32_t)raw_bytes;
            uint32_t g = 0;                                              // global k-block counter
            long long w_raw = 0;
            if (b_resident && blockIdx.x < num_tiles) {
                mbar_arrive_expect_tx(b_full, (uint32_t)(num_k_blocks * 2 * b_bytes));
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    tma_load_2d(bres_base + (size_t)kb * 2 * b_bytes, &tmBhi, b_full, kb * TC_BLOCK_K, 0);
                    tma_load_2d(bres_base + (size_t)kb * 2 * b_bytes + b_bytes, &tmBlo, b_full, kb * TC_BLOCK_K, 0);
                }
            }
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int te = sc.enabled == 2 ? num_tiles - 1 - tile : tile;
                const int m0 = (te / n_tiles) * TC_BLOCK_M, n0 = (te % n_tiles) * block_n;
                for (int kb = 0; kb < num_k_blocks; ++kb, ++g) {
                    const int s = g % kRaw;
                    const long long c0 = TC_TRACE_T();
                    mbar_wait(&empty_raw[s], ((g / kRaw) & 1u) ^ 1u);
                    w_raw += TC_TRACE_T() - c0;
                    uint8_t* st = raw_base + (size_t)s * raw_bytes;
                    mbar_arrive_expect_tx(&full_raw[s], tx_bytes);
                    tma_load_2d(st, &tmA, &full_raw[s], kb * TC_BLOCK_K, m0);
                    if (!b_resident) {
                        tma_load_2d(st + TC_A_BYTES, &tmBhi, &full_raw[s], kb * TC_BLOCK_K, n0);
                        tma_load_2d(st + TC_A_BYTES + b_bytes, &tmBlo, &full_raw[s], kb * TC_BLOCK_K, n0);
                    }
                }
            }
            if (trace) { g_tc_trace[1] = (unsigned long long)w_raw; g_tc_trace[10] = g; }
        }
    } else if (warp == 1) {
        // warp-uniform control flow, one elected lane issues (an `if (lane == 0)` region makes ptxas wrap every tcgen05
        // instruction in a per-thread waterfall loop)
        {
            const uint32_t idesc = make_idesc_tf32(TC_BLOCK_M, block_n);
            uint32_t g = 0, t = 0;
            long long w_ready = 0, w_acc = 0, w_issue = 0;
            if (b_resident && blockIdx.x < num_tiles) mbar_wait(b_full, 0);
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
                const uint32_t acc = t & 1u;
                const long long c0 = TC_TRACE_T();
                mbar_wait(&tmem_empty[acc], ((t >> 1) & 1u) ^ 1u);       // epilogue drained this accumulator
                w_acc += TC_TRACE_T() - c0;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * acc_stride;
                for (int kb = 0; kb < num_k_blocks; ++kb, ++g) {
                    const int s = g % kRaw, o = g % n_op;
                    const long long c1 = TC_TRACE_T();
                    mbar_wait(&ready_op[o], (g / n_op) & 1u);
                    const long long c2 = TC_TRACE_T();
                    w_ready += c2 - c1;
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(op_base + (size_t)o * 2 * TC_A_BYTES);
                    const uint32_t a_lo = a_hi + TC_A_BYTES;
                    const uint32_t b_hi = b_resident ? smem_u32(bres_base + (size_t)kb * 2 * b_bytes)
                                                     : smem_u32(raw_base + (size_t)s * raw_bytes + TC_A_BYTES);
                    const uint32_t b_lo = b_hi + b_bytes;
                    if (elect_one()) {
                    if (stack_n) {
                        const uint32_t ta_hi = tmem_base + a_col + (uint32_t)o * 64u, ta_lo = ta_hi + 32u;
                        const uint32_t idesc2 = make_idesc_tf32(TC_BLOCK_M, 2 * block_n);
#pragma unroll
                        for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
                            const uint64_t db = make_kmajor_sw128_desc(b_hi + k * TC_UMMA_K * 4);       // [B hi ; B lo]
                            umma_tf32_ts(d_tmem, ta_lo + k * TC_UMMA_K, db, idesc2, (kb > 0 || k > 0) ? 1u : 0u);
                            umma_tf32_ts(d_tmem, ta_hi + k * TC_UMMA_K, db, idesc2, 1u);
                        }
                    } else if (a_stages > 0) {
                        const uint32_t ta_hi = tmem_base + a_col + (uint32_t)o * 64u, ta_lo = ta_hi + 32u;
#pragma unroll
                        for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
                            const uint32_t koff = k * TC_UMMA_K * 4;
                            const uint64_t db_hi = make_kmajor_sw128_desc(b_hi + koff), db_lo = make_kmajor_sw128_desc(b_lo + koff);
                            umma_tf32_ts(d_tmem, ta_lo + k * TC_UMMA_K, db_hi, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                            umma_tf32_ts(d_tmem, ta_hi + k * TC_UMMA_K, db_lo, idesc, 1u);
                            umma_tf32_ts(d_tmem, ta_hi + k * TC_UMMA_K, db_hi, idesc, 1u);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
                            const uint32_t koff = k * TC_UMMA_K * 4;
                            const uint64_t da_hi = make_kmajor_sw128_desc(a_hi + koff), da_lo = make_kmajor_sw128_desc(a_lo + koff);
                            const uint64_t db_hi = make_kmajor_sw128_desc(b_hi + koff), db_lo = make_kmajor_sw128_desc(b_lo + koff);
                            umma_tf32(d_tmem, da_lo, db_hi, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                            umma_tf32(d_tmem, da_hi, db_lo, idesc, 1u);
                            umma_tf32(d_tmem, da_hi, db_hi, idesc, 1u);
                        }
                    }
                    umma_commit(&empty_op[o]);
                    if (!b_resident) umma_commit(&empty_raw[s]);
                    }
                    __syncwarp();
                    w_issue += TC_TRACE_T() - c2;
                }
                if (elect_one()) umma_commit(&tmem_full[acc]);
                __syncwarp();
            }
            if (trace && lane == 0) { g_tc_trace[2] = (unsigned long long)w_ready; g_tc_trace[3] = (unsigned long long)w_acc; g_tc_trace[4] = (unsigned long long)w_issue; }
        }
    } else if (warp < 6) {
        // ---------------- split warps: raw A -> (hi, lo) operand stage
        const int tid = threadIdx.x - 64;
        uint32_t g = 0;
        long long w_full = 0, w_eop = 0, w_work = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < num_k_blocks; ++kb, ++g) {
                const int s = g % kRaw, o = g % n_op;
                const long long c0 = TC_TRACE_T();
                mbar_wait(&full_raw[s], (g / kRaw) & 1u);
                const long long c1 = TC_TRACE_T();
                mbar_wait(&empty_op[o], ((g / n_op) & 1u) ^ 1u);
                const long long c2 = TC_TRACE_T();
                w_full += c1 - c0; w_eop += c2 - c1;
                const float4* src = reinterpret_cast<const float4*>(raw_base + (size_t)s * raw_bytes);
                if (a_stages > 0) {
                    // TS mode: this thread owns tile row (warp & 3) * 32 + lane — the TMEM lane its warp may write.
                    // The raw tile is SWIZZLE_128B: 16-byte unit j of row r sits at unit j ^ (r & 7).
                    tc_fence_after();
                    const int r = (warp & 3) * 32 + lane;
                    uint32_t h[32], l[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 v = src[r * 8 + (j ^ (r & 7))];
                        h[4 * j + 0] = __float_as_uint(v.x) & 0xFFFFE000u; l[4 * j + 0] = __float_as_uint(v.x - __uint_as_float(h[4 * j + 0]));
                        h[4 * j + 1] = __float_as_uint(v.y) & 0xFFFFE000u; l[4 * j + 1] = __float_as_uint(v.y - __uint_as_float(h[4 * j + 1]));
                        h[4 * j + 2] = __float_as_uint(v.z) & 0xFFFFE000u; l[4 * j + 2] = __float_as_uint(v.z - __uint_as_float(h[4 * j + 2]));
                        h[4 * j + 3] = __float_as_uint(v.w) & 0xFFFFE000u; l[4 * j + 3] = __float_as_uint(v.w - __uint_as_float(h[4 * j + 3]));
                    }
                    const uint32_t ta = tmem_base + a_col + (uint32_t)o * 64u + ((uint32_t)((warp & 3) * 32) << 16);
                    tmem_st32(ta, h);
                    tmem_st32(ta + 32u, l);
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(&ready_op[o]);
                    if (b_resident) mbar_arrive(&empty_raw[s]);
                    w_work += TC_TRACE_T() - c2;
                    continue;
                }
                float4* hi = reinterpret_cast<float4*>(op_base + (size_t)o * 2 * TC_A_BYTES);
                float4* lo = reinterpret_cast<float4*>(op_base + (size_t)o * 2 * TC_A_BYTES + TC_A_BYTES);
#pragma unroll
                for (int i = 0; i < TC_A_BYTES / 16 / 128; ++i) {
                    const int j = tid + i * 128;
                    const float4 v = src[j];
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
                    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
                    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
                    hi[j] = h;
                    lo[j] = l;
                }
                fence_proxy_async();
                mbar_arrive(&ready_op[o]);
                if (b_resident) mbar_arrive(&empty_raw[s]);              // the raw A tile has been consumed
                w_work += TC_TRACE_T() - c2;
            }
        }
        if (trace && tid == 0) { g_tc_trace[5] = (unsigned long long)w_full; g_tc_trace[6] = (unsigned long long)w_eop; g_tc_trace[7] = (unsigned long long)w_work; }
    } else {
        // ---------------- epilogue warps: TMEM -> registers -> global (thread = output row), overlapped with the next tile
        const int quarter = warp & 3;
        const int half = (warp - 6) >> 2;              // which of the quarter's two warps: chunks half, half+2, ...
        const int row = quarter * 32 + lane;
        const bool c_vec = ((ep.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.C) & 15u) == 0);
        const bool m_vec = ep.mask != nullptr && ((ep.ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.mask) & 15u) == 0);
        // ---- fused tower tail: the 8 epilogue warps are one 256-thread team (named barrier 1) that owns tw_As / tw_Bs
        const int et = threadIdx.x - 6 * 32;
        auto epi_sync = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
        float4 tw_wo = make_float4(0.f, 0.f, 0.f, 0.f);
        float tw_bo = 0.f, tw_loss_acc = 0.f;
        if (tw.enabled) {
            tower_load_weights_t<V2_EPI_WARPS * 32>(tw, tw_Bs, et);
            tw_wo = ldg_f4(tw.w_out + (et & 15) * 4);
            tw_bo = tw.b_out != nullptr ? __ldg(tw.b_out) : 0.f;
        }
        uint32_t t = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
            const int te = sc.enabled == 2 ? num_tiles - 1 - tile : tile;
            const int m0 = (te / n_tiles) * TC_BLOCK_M, n0 = (te % n_tiles) * block_n;
            const uint32_t acc = t & 1u;
            const int m = m0 + row;
            if (tw.enabled) {
                // layer-1 epilogue (stacked accumulator: a.b_hi + a.b_lo, bias, ReLU) -> h1 to HBM and into tw_As, then
                // the accumulator is handed back and the tail layers / head run on CUDA cores while the tensor pipe is
                // already busy with the next tiles (host guarantees block_n == 64, stack_n, bias, relu)
                const long long q0 = TC_TRACE_T();
                mbar_wait(&tmem_full[acc], (t >> 1) & 1u);
                tc_fence_after();
                const long long q1 = TC_TRACE_T();
                epi_sync();                            // previous tile's activations fully consumed (weights visible)
                for (int c0 = half * 16; c0 < TW_H; c0 += 32) {
                    uint32_t r[16], r2[16];
                    tmem_ld16(tmem_base + acc * acc_stride + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
                    tmem_ld16(tmem_base + acc * acc_stride + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(TW_H + c0), r2);
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        v[j] = fmaxf(__uint_as_float(r2[j]) + __uint_as_float(r[j]) + __ldg(ep.bias + c0 + j), 0.f);
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 q4 = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        if (m < ep.M) stg_f4(ep.C + (size_t)m * ep.ldc + c0 + j, q4);
                        *reinterpret_cast<float4*>(tw_As + row * TW_LDA + c0 + j) = q4;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                epi_sync();
                const long long q2 = TC_TRACE_T();
                tower_tail_tile_fwd<V2_EPI_WARPS * 32>(tw, tw_As, tw_Bs, m0, et, tw_wo, tw_bo, tw_loss_acc, epi_sync);
                if (trace && warp == 6 && lane == 0) {
                    g_tc_trace[8] += (unsigned long long)(q1 - q0);           // waiting for the accumulator
                    g_tc_trace[9] += (unsigned long long)(q2 - q1);           // layer-1 epilogue (TMEM -> HBM + smem)
                    g_tc_trace[11] += (unsigned long long)(TC_TRACE_T() - q2); // tail layers + head
                }
                continue;
            }
            if (sc.enabled) {
                // ---- fused scatter-add of the table gradients (see TcScatter).
                // TMEM hands each lane 16 consecutive columns of ITS row; a 4x4 float4 transpose inside every group of
                // 4 lanes (transpose4x4) turns that into "4 lanes cover 64 contiguous bytes of one row" for 4 rows, so
                // x loads and table-gradient reductions are whole 64-byte requests instead of 16-byte fragments.
                // Everything that does not depend on the accumulator (row ids, forward rows e) is fetched BEFORE the
                // accumulator is awaited, two 16-column chunks at a time, to overlap the random-access latency with
                // the main loop of this tile.
                const int FD = sc.F * sc.D;
                const int q = lane & 3;
                const int mb = m0 + quarter * 32 + (lane >> 2) * 4;     // first of this lane group's 4 rows
                float cf[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) cf[j] = (sc.dfm != nullptr && mb + j < ep.M) ? __ldg(sc.dfm + mb + j) : 0.f;
                bool waited = false;
                for (int c0 = half * 16; c0 < block_n; c0 += 64) {
                    long long ix[2][4];
                    float4 e[2][4];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int n4 = n0 + c0 + 32 * u + 4 * q;
                        const bool colok = c0 + 32 * u < block_n && n4 < FD;
                        const int f = colok ? n4 / sc.D : 0;
                        const bool live = colok && sc.grads[f] != nullptr;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            ix[u][j] = -1;
                            e[u][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (live && mb + j < ep.M) {
                                long long v = __ldg(sc.idx[f] + mb + j);
                                if ((unsigned long long)v >= (unsigned long long)sc.rows[f]) v = 0;
                                ix[u][j] = v;
                                if (sc.dfm != nullptr) e[u][j] = ldg_f4_stream(sc.x + (size_t)(mb + j) * sc.ldx + n4);
                            }
                        }
                    }
                    if (!waited) { mbar_wait(&tmem_full[acc], (t >> 1) & 1u); tc_fence_after(); waited = true; }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (c0 + 32 * u < block_n) {            // warp-uniform
                            uint32_t r[16];
                            tmem_ld16(tmem_base + acc * acc_stride + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c0 + 32 * u), r);
                            transpose4x4(r, q);                 // r[4j..4j+3] = columns 4q..4q+3 of row mb + j
                            const int n4 = n0 + c0 + 32 * u + 4 * q;
                            const int f = n4 < FD ? n4 / sc.D : 0, d = n4 - f * sc.D;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (ix[u][j] >= 0) {
                                    float4 gv = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                            __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                                    if (sc.dfm != nullptr) {
                                        const float4 sv = ldg_f4(sc.fm_s + (size_t)(mb + j) * sc.D + d);
                                        gv.x = fmaf(cf[j], sv.x - e[u][j].x, gv.x); gv.y = fmaf(cf[j], sv.y - e[u][j].y, gv.y);
                                        gv.z = fmaf(cf[j], sv.z - e[u][j].z, gv.z); gv.w = fmaf(cf[j], sv.w - e[u][j].w, gv.w);
                                    }
                                    float* grow;
                                    if constexpr (kShardedScatter) {   // row-sharded gradients: owner = id mod G, local row = id div G
                                        const unsigned iu = (unsigned)ix[u][j], gg = (unsigned)sc.G;
                                        float* base = reinterpret_cast<float*>(__ldg(reinterpret_cast<const unsigned long long*>(sc.grad_shard_tab) + (size_t)f * gg + (iu % gg)));
                                        grow = base + (size_t)(iu / gg) * sc.D + d;
                                    } else {
                                        grow = sc.grads[f] + (size_t)ix[u][j] * sc.D + d;
                                    }
                                    red_add_f4(grow, gv);
                                }
                            }
                        }
                    }
                }
                if (!waited) { mbar_wait(&tmem_full[acc], (t >> 1) & 1u); tc_fence_after(); }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                continue;
            }
            const long long e0 = TC_TRACE_T();
            mbar_wait(&tmem_full[acc], (t >> 1) & 1u);
            const long long e1 = TC_TRACE_T();
            tc_fence_after();
            for (int c0 = half * 16; c0 < block_n; c0 += 32) {
                uint32_t r[16];
                tmem_ld16(tmem_base + acc * acc_stride + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
                uint32_t r2[16];
                if (stack_n) tmem_ld16(tmem_base + acc * acc_stride + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(block_n + c0), r2);
                if (m < ep.M && n0 + c0 < ep.N) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = stack_n ? __uint_as_float(r2[j]) + __uint_as_float(r[j]) : __uint_as_float(r[j]);
                    const int n = n0 + c0;
                    const bool full = n + 16 <= ep.N;
                    if (ep.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (n + j < ep.N) v[j] += __ldg(ep.bias + n + j);
                    }
                    if (ep.relu) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                    }
                    if (ep.mask != nullptr) {
                        const float* mrow = ep.mask + (size_t)m * ep.ldmask + n;
                        if (m_vec && full) {
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4) {
                                const float4 mk = __ldg(reinterpret_cast<const float4*>(mrow) + j4);
                                v[4 * j4] = mk.x > 0.f ? v[4 * j4] : 0.f; v[4 * j4 + 1] = mk.y > 0.f ? v[4 * j4 + 1] : 0.f;
                                v[4 * j4 + 2] = mk.z > 0.f ? v[4 * j4 + 2] : 0.f; v[4 * j4 + 3] = mk.w > 0.f ? v[4 * j4 + 3] : 0.f;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) if (n + j < ep.N) v[j] = (__ldg(mrow + j) > 0.f) ? v[j] : 0.f;
                        }
                    }
                    float* crow = ep.C + (size_t)m * ep.ldc + n;
                    if (c_vec && full) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) stg_f4(crow + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (n + j < ep.N) crow[j] = v[j];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (trace && warp == 6 && lane == 0) {
                g_tc_trace[8] += (unsigned long long)(e1 - e0);
                g_tc_trace[9] += (unsigned long long)(TC_TRACE_T() - e1);
            }
        }
        if (tw.enabled && tw.loss != nullptr) tw_loss[et] = tw_loss_acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
    if (tw.enabled && tw.loss != nullptr && warp == 0) {
        // deterministic mean BCE: fixed-order sum of the 256 epilogue partials -> per-CTA partial -> the last CTA to
        // finish adds the per-CTA partials in index order (same scheme as head.cu / tower.cu)
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < V2_EPI_WARPS; ++i) s += tw_loss[lane + 32 * i];
        s = warp_sum(s);
        unsigned int last = 0;
        if (lane == 0) {
            tw.partials[blockIdx.x] = s;
            __threadfence();
            last = (atomicAdd(tw.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            __threadfence();
            float tot = 0.f;
            for (int i = lane; i < (int)gridDim.x; i += 32) tot += ((volatile float*)tw.partials)[i];
            tot = warp_sum(tot);
            if (lane == 0) {
                tw.loss[0] = tw.scale * (tot / (float)tw.M);
                *tw.counter = 0u;
            }
        }
    }
    if (trace && threadIdx.x == 0) g_tc_trace[0] = (unsigned long long)(clock64() - t_start);
}

// ---------------------------------------------------------------------------------- weight gradient on tcgen05
// dW[n,k] += sum_m dy[m,n] * x[m,k]: the reduction runs over the SAMPLES, so both operands are MN-major (their
// contiguous dimension is the output dimension): A = x^T tile [128 k-columns x 32 samples], B = dy^T [block_n x 32
// samples].  For 32-bit MN-major operands the only UMMA shared-memory layout is "128B swizzle with 32B atoms"
// (LayoutType 1; TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows of 128 B, 32-byte chunks XOR-ed with (row mod 4), so a
// K atom is 4 samples (512 B).  TMA brings [32 samples x 32 floats] boxes: one box = one 32-float MN chunk (chunks
// LBO = 4 KiB apart) holding eight 4-sample K atoms (SBO = 512 B apart); one MMA (K = 8) consumes two atoms.
// Each CTA reduces one slab of samples for one 128-column tile of dW^T in TMEM and adds it to dW with fp32 reductions.
constexpr int WG_ROWS = 32;                                  // samples per stage (4 stages in flight hide the TMA latency)
constexpr int WG_BOX_BYTES = WG_ROWS * 128;                  // one [32 samples x 32 fp32] box = 4 KiB

__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);            // start address
    d |= (uint64_t)(WG_BOX_BYTES >> 4) << 16;               // LBO: next 32-float chunk along M/N
    d |= (uint64_t)(512 >> 4) << 32;                        // SBO: next 4-sample K atom
    d |= (uint64_t)1 << 46;                                 // version = 1
    d |= (uint64_t)1 << 61;                                 // SWIZZLE_128B_BASE32B
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int m, int n) {
    return make_idesc_tf32(m, n) | (1u << 15) | (1u << 16);  // A and B are MN-major
}

template <int kStages>
__global__ void __launch_bounds__(TC_THREADS)
wgrad_tf32x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy,
                    float* __restrict__ dW, int N, int K, int M, int block_n, int slab, uint32_t tmem_cols, int stack_n, int raw_hi) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the pointer in the shared address space (LDS/STS, not generic LD/ST)
    const int a_bytes = 4 * WG_BOX_BYTES;                    // 128 k-columns = 4 chunks
    const int b_chunks = block_n / 32;
    const int b_bytes = b_chunks * WG_BOX_BYTES;
    const int stage_bytes = 2 * a_bytes + 2 * b_bytes;       // [A hi | A lo | B hi | B lo]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* ready_bar = bars + kStages;
    uint64_t* empty_bar = bars + 2 * kStages;
    uint64_t* tmem_full_bar = bars + 3 * kStages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * 128;                         // dW column tile
    const int mbeg = blockIdx.y * slab;
    const int mend = min(M, mbeg + slab);
    const int num_kb = (mend - mbeg + WG_ROWS - 1) / WG_ROWS;
    const bool trace = g_tc_trace_on != 0 && blockIdx.x == 0 && blockIdx.y == 0;
    const long long t_start = TC_TRACE_T();

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&ready_bar[s], 128); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tx_bytes = (uint32_t)(a_bytes + b_bytes);
            long long w_raw = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (uint32_t)((kb / kStages) & 1);
                const long long c0 = TC_TRACE_T();
                mbar_wait(&empty_bar[s], ph ^ 1u);
                w_raw += TC_TRACE_T() - c0;
                if (trace && kb == num_kb - 1) { g_tc_trace[1] = (unsigned long long)w_raw; g_tc_trace[10] = (unsigned long long)num_kb; }
                uint8_t* st = smem + (size_t)s * stage_bytes;
                const int m0 = mbeg + kb * WG_ROWS;
                mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
                for (int c = 0; c < 4; ++c) tma_load_2d(st + c * WG_BOX_BYTES, &tmX, &full_bar[s], k0 + 32 * c, m0);
                for (int c = 0; c < b_chunks; ++c)
                    tma_load_2d(st + 2 * a_bytes + c * WG_BOX_BYTES, &tmDy, &full_bar[s], 32 * c, m0);
            }
        }
    } else if (warp == 1) {
        // Warp-uniform control flow, one elected lane issues: inside an `if (lane == 0)` region ptxas wraps every tcgen05
        // instruction in a per-thread waterfall loop (~14 instructions per MMA), which made this thread co-critical.
        // stack_n: [B hi ; B lo] (adjacent 32-float chunks, LBO apart) as ONE operand of 2*block_n columns — two MMAs per
        // 8-sample group instead of three; accumulator columns [0, block_n) = a.b_hi, [block_n, 2*block_n) = a.b_lo.
        const uint32_t idesc = make_idesc_tf32_mn(128, stack_n ? 2 * block_n : block_n);
        const uint64_t d0 = make_mnmajor_sw128_desc(smem_u32(smem));
        long long w_ready = 0, w_issue = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kStages;
            const uint32_t ph = (uint32_t)((kb / kStages) & 1);
            const long long c1 = TC_TRACE_T();
            mbar_wait(&ready_bar[s], ph);
            const long long c2 = TC_TRACE_T();
            w_ready += c2 - c1;
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_hi = (uint32_t)(s * stage_bytes);             // byte offsets from the stage ring's base
                const uint32_t a_lo = a_hi + a_bytes;
                const uint32_t b_hi = a_hi + 2 * a_bytes;
                const uint32_t b_lo = b_hi + b_bytes;
#pragma unroll
                for (int g = 0; g < WG_ROWS / TC_UMMA_K; ++g) {
                    const uint32_t koff = g * 1024;                         // one 8-sample K group
                    const uint64_t da_hi = d0 + ((a_hi + koff) >> 4), da_lo = d0 + ((a_lo + koff) >> 4);
                    const uint64_t db_hi = d0 + ((b_hi + koff) >> 4), db_lo = d0 + ((b_lo + koff) >> 4);
                    if (stack_n) {
                        umma_tf32(tmem_base, da_lo, db_hi, idesc, (kb > 0 || g > 0) ? 1u : 0u);
                        umma_tf32(tmem_base, da_hi, db_hi, idesc, 1u);
                        continue;
                    }
                    umma_tf32(tmem_base, da_lo, db_hi, idesc, (kb > 0 || g > 0) ? 1u : 0u);
                    umma_tf32(tmem_base, da_hi, db_lo, idesc, 1u);
                    umma_tf32(tmem_base, da_hi, db_hi, idesc, 1u);
                }
                umma_commit(&empty_bar[s]);
            }
            __syncwarp();
            w_issue += TC_TRACE_T() - c2;
        }
        if (elect_one()) umma_commit(tmem_full_bar);
        __syncwarp();
        if (trace && lane == 0) { g_tc_trace[2] = (unsigned long long)w_ready; g_tc_trace[4] = (unsigned long long)w_issue; }
    } else {
        const int t = threadIdx.x - 64;
        const int b_vec = b_bytes / 16;
        long long w_full = 0, w_work = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kStages;
            const uint32_t ph = (uint32_t)((kb / kStages) & 1);
            const long long c0 = TC_TRACE_T();
            mbar_wait(&full_bar[s], ph);
            const long long c1 = TC_TRACE_T();
            w_full += c1 - c0;
            float4* a = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes);
            float4* alo = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + a_bytes);
            float4* b = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + 2 * a_bytes);
            float4* blo = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + 2 * a_bytes + b_bytes);
            // all loads of a batch are issued before the first dependent store (the straightforward one-vector-per-iteration
            // loop ran at ~125 cycles per vector and made the split warps the bottleneck of the whole kernel)
            {
                float4 v[8];                                     // A: 4 chunks x 4 KiB = 1024 vectors = 8 per thread
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = a[t + 128 * i];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(v[i].x) & 0xFFFFE000u); l.x = v[i].x - h.x;
                    h.y = __uint_as_float(__float_as_uint(v[i].y) & 0xFFFFE000u); l.y = v[i].y - h.y;
                    h.z = __uint_as_float(__float_as_uint(v[i].z) & 0xFFFFE000u); l.z = v[i].z - h.z;
                    h.w = __uint_as_float(__float_as_uint(v[i].w) & 0xFFFFE000u); l.w = v[i].w - h.w;
                    if (!raw_hi) a[t + 128 * i] = h;
                    alo[t + 128 * i] = l;
                }
            }
            for (int j0 = 0; j0 < b_vec; j0 += 512) {           // B: b_chunks x 256 vectors, 4 per thread per pass
                float4 v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) if (j0 + t + 128 * i < b_vec) v[i] = b[j0 + t + 128 * i];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (j0 + t + 128 * i < b_vec) {
                        float4 h, l;
                        h.x = __uint_as_float(__float_as_uint(v[i].x) & 0xFFFFE000u); l.x = v[i].x - h.x;
                        h.y = __uint_as_float(__float_as_uint(v[i].y) & 0xFFFFE000u); l.y = v[i].y - h.y;
                        h.z = __uint_as_float(__float_as_uint(v[i].z) & 0xFFFFE000u); l.z = v[i].z - h.z;
                        h.w = __uint_as_float(__float_as_uint(v[i].w) & 0xFFFFE000u); l.w = v[i].w - h.w;
                        if (!raw_hi) b[j0 + t + 128 * i] = h;
                        blo[j0 + t + 128 * i] = l;
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(&ready_bar[s]);
            w_work += TC_TRACE_T() - c1;
        }
        // epilogue: TMEM lane = dW column k0+row, TMEM column = output row n
        const long long e0 = TC_TRACE_T();
        mbar_wait(tmem_full_bar, 0);
        const long long e1 = TC_TRACE_T();
        if (trace && t == 0) { g_tc_trace[5] = (unsigned long long)w_full; g_tc_trace[7] = (unsigned long long)w_work; g_tc_trace[8] = (unsigned long long)(e1 - e0); }
        tc_fence_after();
        const int quarter = warp & 3;
        const int k = k0 + quarter * 32 + lane;
        for (int c0 = 0; c0 < block_n; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
            uint32_t r2[16];
            if (stack_n) tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(block_n + c0), r2);
            if (k < K) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = c0 + j;
                    const float v = stack_n ? __uint_as_float(r2[j]) + __uint_as_float(r[j]) : __uint_as_float(r[j]);
                    if (n < N) red_add_f1(dW + (size_t)n * K + k, v);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
    if (trace && threadIdx.x == 0) g_tc_trace[0] = (unsigned long long)(clock64() - t_start);
}

// W[N,K] (row stride ldw) -> hi/lo [Np, Kp] zero padded;  transpose: out[k, n] = W[n, k] (out is [Kout=K rows.., ])
__global__ void __launch_bounds__(256)
split_pack_kernel(const float* __restrict__ W, long long ldw, int rows_in, int cols_in, int transpose,
                  float* __restrict__ hi, float* __restrict__ lo, int Rp, int Cp) {
    // output is [Rp, Cp]; logical out(r, c) = transpose ? W[c, r] : W[r, c]
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)Rp * Cp) return;
    const int r = (int)(t / Cp), c = (int)(t % Cp);
    const int out_rows = transpose ? cols_in : rows_in, out_cols = transpose ? rows_in : cols_in;
    float v = 0.f;
    if (r < out_rows && c < out_cols) v = transpose ? __ldg(W + (size_t)c * ldw + r) : __ldg(W + (size_t)r * ldw + c);
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[t] = h;
    lo[t] = v - h;
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 tensor [rows, cols] with row stride ld (floats); box = [box_rows, 32 cols], 128 B swizzle, zero OOB fill
static int make_map(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_rows,
                    CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return RPB_ERR_NO_DRIVER;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)TC_BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : RPB_ERR_BAD_ARG;
}

int tc_make_map2d(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_cols, int box_rows,
                  int swizzle_bytes) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return RPB_ERR_NO_DRIVER;
    const CUtensorMapSwizzle swz = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B :
                                   swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : RPB_ERR_BAD_ARG;
}

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

bool tc_shape_ok(const float* A, long long lda, int M, int N, int K) {
    return M >= 1 && N >= 1 && K >= 1 && (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15u) == 0);
}

// C[M,N] = epi(A[M,K] . op(W)^T) where op(W) = W [N,K] (transpose_w = 0) or W^T with W stored [K,N]... see callers.
// Bsrc is given as a [b_rows_in, b_cols_in] row-major matrix (stride ldb); the B operand is [N, K] = Bsrc or Bsrc^T.
int gemm_tc(const float* A, long long lda, const float* Bsrc, long long ldb, int b_transpose, const TcEpilogue& ep,
            int M, int N, int K, cudaStream_t st) {
    if (!tc_shape_ok(A, lda, M, N, K)) return RPB_ERR_UNSUPPORTED;
    // N tiling: one tile if N <= 256, else tiles of <= 256 columns (multiple of 16)
    const int n_tiles = ceil_div(N, 256);
    const int block_n = round_up(ceil_div(N, n_tiles), 16);
    const int Np = block_n * n_tiles;
    const int Kp = round_up(K, 4);
    const int nkb = ceil_div(K, TC_BLOCK_K);
    int werr = 0;
    float* ws = static_cast<float*>(workspace(0, (size_t)2 * Np * Kp * sizeof(float), &werr));
    if (ws == nullptr) return werr;
    float* hi = ws;
    float* lo = ws + (size_t)Np * Kp;
    const int rows_in = b_transpose ? K : N, cols_in = b_transpose ? N : K;
    split_pack_kernel<<<ceil_div((long long)Np * Kp, 256), 256, 0, st>>>(Bsrc, ldb, rows_in, cols_in, b_transpose, hi, lo, Np, Kp);
    CUtensorMap tmA, tmBhi, tmBlo;
    int rc = make_map(&tmA, A, M, K, lda, TC_BLOCK_M);
    if (rc == 0) rc = make_map(&tmBhi, hi, Np, Kp, Kp, block_n);
    if (rc == 0) rc = make_map(&tmBlo, lo, Np, Kp, Kp, block_n);
    if (rc == 0 && g_gemm_v2) {
        const int b_bytes = block_n * TC_BLOCK_K * 4;
        // TS mode (A hi/lo in tensor memory) whenever two accumulators leave room for >= 2 operand stages of 64 columns;
        // the fused scatter epilogue keeps the SS pipeline (its 2 x 208 accumulator columns fill TMEM).
        const int stack_n = (g_gemm_a_tmem && g_gemm_stack_n && ep.sc == nullptr && n_tiles == 1 && block_n <= 64) ? 1 : 0;
        const int acc_cols = round_up((stack_n ? 4 : 2) * block_n, 32);
        int a_stages = 0;
        if (g_gemm_a_tmem && ep.sc == nullptr && acc_cols + 2 * 64 <= 512) a_stages = min(V2_MAX_OP, (512 - acc_cols) / 64);
        const int op_bytes = a_stages > 0 ? 0 : V2_OP_STAGES * 2 * TC_A_BYTES;
        const int bres_bytes = nkb * 2 * b_bytes;
        const int raw_bytes_nores = TC_A_BYTES + 2 * b_bytes;
        // fused tower tail: needs the stacked TS-mode accumulator of a single 64-wide tile with bias + ReLU
        const bool tail = ep.tail != nullptr;
        if (tail && !(stack_n && a_stages > 0 && n_tiles == 1 && block_n == TW_H && N == TW_H && ep.relu && ep.bias != nullptr &&
                      ep.C != nullptr && ep.mask == nullptr && (ep.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.C) & 15u) == 0))
            return RPB_ERR_UNSUPPORTED;
        const int tail_bytes = tail ? (TC_BLOCK_M * TW_LDA + ep.tail->n_tail * TW_H * TW_H + 256) * 4 + 64 : 0;
        const int b_resident = (n_tiles == 1 && bres_bytes <= 72 * 1024 &&
                                (!tail || 226 * 1024 - bres_bytes - 1280 - tail_bytes >= 3 * TC_A_BYTES)) ? 1 : 0;
        const int raw_bytes = b_resident ? TC_A_BYTES : raw_bytes_nores;
        uint32_t tmem_cols = 32;
        while ((int)tmem_cols < (a_stages > 0 ? acc_cols + 64 * a_stages : 2 * block_n)) tmem_cols <<= 1;
        const int m_tiles = ceil_div(M, TC_BLOCK_M);
        const int budget = 226 * 1024 - op_bytes - (b_resident ? bres_bytes : 0) - 1024 - 256 - tail_bytes;
        const int max_raw = budget / raw_bytes;
        if (tail && (tmem_cols > 512 || max_raw < 2)) return RPB_ERR_UNSUPPORTED;
        if (tmem_cols <= 512 && max_raw >= 2) {
            const int grid = min(m_tiles * n_tiles, 148);
            auto launch = [&](auto raw_tag) -> int {
                constexpr int R = decltype(raw_tag)::value;
                const size_t smem = (size_t)R * raw_bytes + op_bytes + (b_resident ? bres_bytes : 0) +
                                    (2 * R + 2 * V2_MAX_OP + 4 + 2) * 8 + 1024 + tail_bytes;
                static const TcScatter no_scatter{};
                static const TowerFwdParams no_tail{};
                if (ep.sc != nullptr && ep.sc->G > 1) {          // scatter epilogue into row-sharded gradient buffers
                    cudaError_t es = cudaFuncSetAttribute(gemm_tf32x3_v2_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    if (es != cudaSuccess) return (int)es;
                    gemm_tf32x3_v2_kernel<R, true><<<grid, V2_THREADS, smem, st>>>(tmA, tmBhi, tmBlo, ep, block_n, nkb, m_tiles, n_tiles,
                                                                                  tmem_cols, b_resident, a_stages, stack_n, *ep.sc, no_tail);
                    return (int)cudaGetLastError();
                }
                cudaError_t ee = cudaFuncSetAttribute(gemm_tf32x3_v2_kernel<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (ee != cudaSuccess) return (int)ee;
                if (g_l2_persist && ep.sc != nullptr && ep.sc->x != nullptr)      // opt-in: the FM term re-reads x from the L2 set-aside
                    return (int)launch_windowed(gemm_tf32x3_v2_kernel<R, false>, dim3(grid), dim3(V2_THREADS), smem, st, ep.sc->x,
                                                (size_t)M * (size_t)ep.sc->ldx * sizeof(float), tmA, tmBhi, tmBlo, ep, block_n, nkb,
                                                m_tiles, n_tiles, tmem_cols, b_resident, a_stages, stack_n, *ep.sc, no_tail);
                gemm_tf32x3_v2_kernel<R, false><<<grid, V2_THREADS, smem, st>>>(tmA, tmBhi, tmBlo, ep, block_n, nkb, m_tiles, n_tiles,
                                                                               tmem_cols, b_resident, a_stages, stack_n,
                                                                               ep.sc != nullptr ? *ep.sc : no_scatter,
                                                                               tail ? *ep.tail : no_tail);
                return (int)cudaGetLastError();
            };
            if (max_raw >= 6) return launch(std::integral_constant<int, 6>{});
            if (max_raw >= 5) return launch(std::integral_constant<int, 5>{});
            if (max_raw >= 4) return launch(std::integral_constant<int, 4>{});
            if (max_raw >= 3) return launch(std::integral_constant<int, 3>{});
            return launch(std::integral_constant<int, 2>{});
        }
    }
    if (rc == 0 && (ep.sc != nullptr || ep.tail != nullptr)) return RPB_ERR_UNSUPPORTED;   // fused epilogues exist in the v2 kernel only
    if (rc == 0) {
        const int b_bytes = block_n * TC_BLOCK_K * 4;
        const int stage_bytes = 2 * TC_A_BYTES + 2 * b_bytes;
        uint32_t tmem_cols = 32;
        while ((int)tmem_cols < block_n) tmem_cols <<= 1;
        dim3 grid(ceil_div(M, TC_BLOCK_M), n_tiles);
        auto launch = [&](auto stages_tag) -> int {
            constexpr int S = decltype(stages_tag)::value;
            const size_t smem = (size_t)S * stage_bytes + (3 * S + 2) * 8 + 1024;
            cudaError_t ee = cudaFuncSetAttribute(gemm_tf32x3_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (ee != cudaSuccess) return (int)ee;
            gemm_tf32x3_kernel<S><<<grid, TC_THREADS, smem, st>>>(tmA, tmBhi, tmBlo, ep, block_n, nkb, tmem_cols);
            return (int)cudaGetLastError();
        };
        // Stage count: two co-resident CTAs per SM (so one CTA's epilogue overlaps the other's main loop) whenever
        // 2 stages fit in half of the ~220 KiB budget; otherwise as many stages as fit for a single CTA.
        const int max_stages = (200 * 1024) / stage_bytes;
        if (2 * stage_bytes + 2048 <= 110 * 1024) rc = launch(std::integral_constant<int, 2>{});
        else if (max_stages >= 4 && nkb >= 4) rc = launch(std::integral_constant<int, 4>{});
        else if (max_stages >= 3 && nkb >= 3) rc = launch(std::integral_constant<int, 3>{});
        else if (max_stages >= 2) rc = launch(std::integral_constant<int, 2>{});
        else rc = RPB_ERR_UNSUPPORTED;
    }
    return rc;
}


// Pre-split weight operand for kernels outside this file (deepfm_fused.cu): W [N, K] (row stride ldw) -> hi / lo [N, Kp]
// in workspace slot 0 (the same scratch gemm_tc uses; one stream => ordered) + their TMA maps with [N x 32] boxes.
int tc_prepare_weight(const float* W, long long ldw, int N, int K, CUtensorMap* tm_hi, CUtensorMap* tm_lo, cudaStream_t st) {
    const int Kp = round_up(K, 4);
    int werr = 0;
    float* ws = static_cast<float*>(workspace(0, (size_t)2 * N * Kp * sizeof(float), &werr));
    if (ws == nullptr) return werr;
    float* hi = ws;
    float* lo = ws + (size_t)N * Kp;
    split_pack_kernel<<<ceil_div((long long)N * Kp, 256), 256, 0, st>>>(W, ldw, N, K, 0, hi, lo, N, Kp);
    int rc = make_map(tm_hi, hi, N, Kp, Kp, N);
    if (rc == 0) rc = make_map(tm_lo, lo, N, Kp, Kp, N);
    return rc;
}

// General pre-split K-major operand for kernels outside this file (cin_tc.cu): out(r, c) = transpose ? src[c, r] : src[r, c],
// zero padded to [Rp, Cp], hi / lo in workspace `slot`, TMA maps with [box_rows x 32] boxes (SWIZZLE_128B).
int tc_prepare_operand(const float* src, long long ld, int rows_in, int cols_in, int transpose, int Rp, int Cp, int box_rows,
                       int slot, CUtensorMap* tm_hi, CUtensorMap* tm_lo, cudaStream_t st) {
    int werr = 0;
    float* ws = static_cast<float*>(workspace(slot, (size_t)2 * Rp * Cp * sizeof(float), &werr));
    if (ws == nullptr) return werr;
    float* hi = ws;
    float* lo = ws + (size_t)Rp * Cp;
    split_pack_kernel<<<ceil_div((long long)Rp * Cp, 256), 256, 0, st>>>(src, ld, rows_in, cols_in, transpose, hi, lo, Rp, Cp);
    int rc = make_map(tm_hi, hi, Rp, Cp, Cp, box_rows);
    if (rc == 0) rc = make_map(tm_lo, lo, Rp, Cp, Cp, box_rows);
    return rc;
}

// Pre-split operands of the square 64 x 64 tower-tail layers for the tcgen05 tail of deepfm_fused.cu: hi / lo
// [n_tail * 64, 64] (layer l = rows l*64 .. l*64+63) in workspace `slot` (6: forward, 7: backward with transpose = 1, i.e.
// W_l^T, the K-major operand of dz . W_l) + their TMA maps with [64 x 32] boxes.
struct TailSplitArgs { const float* W[RPB_TOWER_MAX_TAIL]; };
__global__ void __launch_bounds__(256)
split_pack_tail_kernel(const TailSplitArgs a, int n_tail, int transpose, float* __restrict__ hi, float* __restrict__ lo) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tail * 64 * 64) return;
    const int i = t & 4095;                              // out[l][r = i >> 6][c = i & 63] = transpose ? W_l[c][r] : W_l[r][c]
    const float v = __ldg(a.W[t >> 12] + (transpose ? ((i & 63) << 6) + (i >> 6) : i));
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[t] = h;
    lo[t] = v - h;
}

int tc_prepare_tail_weights(const float* const* W, int n_tail, CUtensorMap* tm_hi, CUtensorMap* tm_lo, cudaStream_t st,
                            int transpose, int slot) {
    if (n_tail < 1 || n_tail > RPB_TOWER_MAX_TAIL) return RPB_ERR_UNSUPPORTED;
    int werr = 0;
    float* ws = static_cast<float*>(workspace(slot, (size_t)2 * n_tail * 64 * 64 * sizeof(float), &werr));
    if (ws == nullptr) return werr;
    float* hi = ws;
    float* lo = ws + (size_t)n_tail * 64 * 64;
    TailSplitArgs a{};
    for (int l = 0; l < n_tail; ++l) {
        if (W[l] == nullptr) return RPB_ERR_BAD_ARG;
        a.W[l] = W[l];
    }
    split_pack_tail_kernel<<<ceil_div(n_tail * 64 * 64, 256), 256, 0, st>>>(a, n_tail, transpose, hi, lo);
    int rc = make_map(tm_hi, hi, (long long)n_tail * 64, 64, 64, 64);
    if (rc == 0) rc = make_map(tm_lo, lo, (long long)n_tail * 64, 64, 64, 64);
    return rc;
}

// dW[N,K] += dy[M,N]^T @ x[M,K] on tensor cores (see wgrad_tf32x3_kernel).  Returns RPB_ERR_UNSUPPORTED when the
// operands do not satisfy the TMA constraints (16-byte aligned rows) or N > 256.
int wgrad_tc(const float* dy, long long lddy, const float* x, long long ldx, float* dW, int M, int N, int K, cudaStream_t st) {
    if ((lddy % 4) != 0 || (ldx % 4) != 0 || (reinterpret_cast<uintptr_t>(dy) & 15u) || (reinterpret_cast<uintptr_t>(x) & 15u))
        return RPB_ERR_UNSUPPORTED;
    if (N > 256) {                                   // wide layers (MMOE experts): one launch per 256-column slice of dy
        for (int n0 = 0; n0 < N; n0 += 256) {
            const int rc = wgrad_tc(dy + n0, lddy, x, ldx, dW + (size_t)n0 * K, M, min(256, N - n0), K, st);
            if (rc != 0) return rc;
        }
        return 0;
    }
    const int block_n = round_up(N, 32);
    const int stack_n = (g_gemm_stack_n && block_n <= 64) ? 1 : 0;
    CUtensorMap tmX, tmDy;
    int rc = make_map(&tmX, x, M, K, ldx, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc == 0) rc = make_map(&tmDy, dy, M, N, lddy, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc != 0) return rc;
    const int ktiles = ceil_div(K, 128);
    int slabs = max(1, min(ceil_div(M, WG_ROWS), (148 * 2) / ktiles));
    int slab = round_up(ceil_div(M, slabs), WG_ROWS);
    slabs = ceil_div(M, slab);
    const int stage_bytes = 2 * 4 * WG_BOX_BYTES + 2 * (block_n / 32) * WG_BOX_BYTES;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < (stack_n ? 2 * block_n : block_n)) tmem_cols <<= 1;
    auto launch = [&](auto stages_tag) -> int {
        constexpr int S = decltype(stages_tag)::value;
        const size_t smem = (size_t)S * stage_bytes + (3 * S + 2) * 8 + 1024;
        cudaError_t ee = cudaFuncSetAttribute(wgrad_tf32x3_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ee != cudaSuccess) return (int)ee;
        if (g_l2_persist)                          // opt-in: x (the wide operand) is re-read by the scatter epilogue that follows
            return (int)launch_windowed(wgrad_tf32x3_kernel<S>, dim3(ktiles, slabs), dim3(TC_THREADS), smem, st, x,
                                        (size_t)M * (size_t)ldx * sizeof(float), tmX, tmDy, dW, N, K, M, block_n, slab, tmem_cols,
                                        stack_n, g_tf32_raw_hi);
        wgrad_tf32x3_kernel<S><<<dim3(ktiles, slabs), TC_THREADS, smem, st>>>(tmX, tmDy, dW, N, K, M, block_n, slab, tmem_cols, stack_n, g_tf32_raw_hi);
        return (int)cudaGetLastError();
    };
    const int max_stages = min((200 * 1024) / stage_bytes, g_wgrad_stages);
    if (max_stages >= 4) return launch(std::integral_constant<int, 4>{});
    if (max_stages >= 3) return launch(std::integral_constant<int, 3>{});
    if (max_stages >= 2) return launch(std::integral_constant<int, 2>{});
    return RPB_ERR_UNSUPPORTED;
}

void colsum_launch(const float* x, long long ldx, float* out, int M, int K, cudaStream_t st);   // mmoe.cu

// SIMT implementations (linear_simt.cu)
int linear_fwd_simt(const float* x, long long ldx, const float* W, const float* bias, float* y, long long ldy,
                    int M, int N, int K, int act, cudaStream_t st);
int linear_dx_simt(const float* dy, long long lddy, const float* W, const float* mask, long long ldmask,
                   float* dx, long long lddx, int M, int N, int K, cudaStream_t st);
int linear_dw_simt(const float* dy, long long lddy, const float* x, long long ldx, float* dW, float* db,
                   int M, int N, int K, cudaStream_t st);

}  // namespace rpb

using namespace rpb;

RPB_API int rpb_linear_fwd(const float* x, int64_t ldx, const float* W, const float* bias, float* y, int64_t ldy,
                           int M, int N, int K, int act, int impl, void* stream) {
    if (x == nullptr || W == nullptr || y == nullptr || M <= 0 || N <= 0 || K <= 0) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool tc_ok = tc_shape_ok(x, ldx, M, N, K);
    if (impl == 2 && !tc_ok) return RPB_ERR_UNSUPPORTED;
    if (impl == 2 || (impl == 0 && tc_ok && M >= 512)) {
        TcEpilogue ep{y, ldy, bias, nullptr, 0, M, N, act == 1, nullptr};
        return gemm_tc(x, ldx, W, K, 0, ep, M, N, K, st);
    }
    return linear_fwd_simt(x, ldx, W, bias, y, ldy, M, N, K, act, st);
}

RPB_API int rpb_linear_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* W,
                           const float* mask, int64_t ldmask, float* dx, int64_t lddx, float* dW, float* db,
                           int M, int N, int K, int impl, void* stream) {
    if (dy == nullptr || M <= 0 || N <= 0 || K <= 0) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dx != nullptr) {
        if (W == nullptr) return RPB_ERR_BAD_ARG;
        // dx[M,K] = dy[M,N] @ W[N,K]: the GEMM's reduction dim is N, its output width is K, B operand = W^T [K, N]
        const bool tc_ok = tc_shape_ok(dy, lddy, M, K, N);
        if (impl == 2 && !tc_ok) return RPB_ERR_UNSUPPORTED;
        int rc;
        if (impl == 2 || (impl == 0 && tc_ok && M >= 512)) {
            TcEpilogue ep{dx, lddx, nullptr, mask, ldmask, M, K, 0, nullptr};
            rc = gemm_tc(dy, lddy, W, K, 1, ep, M, K, N, st);
        } else {
            rc = linear_dx_simt(dy, lddy, W, mask, ldmask, dx, lddx, M, N, K, st);
        }
        if (rc != 0) return rc;
    }
    if (dW != nullptr) {
        if (x == nullptr) return RPB_ERR_BAD_ARG;
        int rc = RPB_ERR_UNSUPPORTED;
        if (impl == 2 || (impl == 0 && M >= 2048 && g_wgrad_tc)) {
            rc = wgrad_tc(dy, lddy, x, ldx, dW, M, N, K, st);
            if (rc == 0 && db != nullptr) colsum_launch(dy, lddy, db, M, N, st);
            if (rc != 0 && rc != RPB_ERR_UNSUPPORTED) return rc;
        }
        if (rc != 0) rc = linear_dw_simt(dy, lddy, x, ldx, dW, db, M, N, K, st);
        if (rc != 0) return rc;
    }
    return 0;
}

// Layer-1 GEMM fused with the rest of the tower (see TcEpilogue::tail): one launch from the feature row to the loss.
RPB_API int rpb_linear_tower_fwd(const float* x, int64_t ldx, const float* W1, const float* b1, int K,
                                 const RpbTowerFwdDesc* d, void* stream) {
    if (x == nullptr || W1 == nullptr || b1 == nullptr || d == nullptr || K <= 0) return RPB_ERR_BAD_ARG;
    TowerFwdParams p{};
    const int prc = tower_fwd_params(d, &p);
    if (prc != 0) return prc;
    if (d->n_tail < 1 || d->M < 512 || !g_gemm_v2 || !tc_shape_ok(x, ldx, d->M, TW_H, K)) return RPB_ERR_UNSUPPORTED;
    TcEpilogue ep{const_cast<float*>(d->h1), d->ldh1, b1, nullptr, 0, d->M, TW_H, 1, nullptr, &p};
    return gemm_tc(x, ldx, W1, K, 0, ep, d->M, TW_H, K, reinterpret_cast<cudaStream_t>(stream));
}

// dx GEMM of the first MLP layer fused with the embedding-gradient scatter (TcScatter): dx is never written.
// Diagnostics: enable/disable the per-role stall counters of the persistent GEMM and read them back (16 x u64, cycles of
// CTA 0 of the LAST launch: [0] kernel, [1] producer waits for a free raw stage, [2] MMA issuer waits for operands,
// [3] MMA issuer waits for a drained accumulator, [4] MMA issue, [5] split waits for TMA, [6] split waits for a free
// operand stage, [7] split work, [8] epilogue waits for the accumulator, [9] epilogue work, [10] k-blocks).
RPB_API int rpb_debug_tc_trace(uint64_t* out16, int enable) {
    int on = enable != 0;
    cudaError_t e = cudaSuccess;
    if (out16 != nullptr) {
        e = cudaMemcpyFromSymbol(out16, g_tc_trace, sizeof(unsigned long long) * 16);
        if (e != cudaSuccess) return (int)e;
    }
    static const unsigned long long zeros[16] = {};
    e = cudaMemcpyToSymbol(g_tc_trace, zeros, sizeof(zeros));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_tc_trace_on, &on, sizeof(int));
    return (int)e;
}

RPB_API int rpb_linear_dx_scatter(const float* dy, int64_t lddy, const float* W, int M, int N, int K,
                                  const RpbScatterDesc* d, void* stream) {
    if (dy == nullptr || W == nullptr || d == nullptr || M <= 0 || N <= 0 || K <= 0) return RPB_ERR_BAD_ARG;
    if (d->B != M || d->F > RPB_MAX_FIELDS || d->F * d->D > K || (d->D % 4) != 0) return RPB_ERR_UNSUPPORTED;
    const bool sharded = d->G > 1;
    if (sharded && d->grad_shard_tab == nullptr) return RPB_ERR_BAD_ARG;
    if (d->dfm != nullptr && (d->x == nullptr || d->fm_s == nullptr || (d->ldx % 4) != 0)) return RPB_ERR_BAD_ARG;
    if (!g_gemm_v2 || !tc_shape_ok(dy, lddy, M, K, N)) return RPB_ERR_UNSUPPORTED;
    TcScatter sc{};
    for (int f = 0; f < d->F; ++f) {
        sc.grads[f] = d->grads ? d->grads[f] : nullptr;
        sc.idx[f] = reinterpret_cast<const long long*>(d->idx[f]);
        sc.rows[f] = d->rows[f];
        if (sc.grads[f] != nullptr && (reinterpret_cast<uintptr_t>(sc.grads[f]) & 15u)) return RPB_ERR_UNSUPPORTED;
    }
    sc.x = d->x; sc.ldx = d->ldx; sc.dfm = d->dfm; sc.fm_s = d->fm_s; sc.F = d->F; sc.D = d->D;
    sc.G = sharded ? d->G : 1;
    sc.grad_shard_tab = sharded ? d->grad_shard_tab : nullptr;
    // enabled == 2: walk the sample tiles from the END — the layer-1 weight gradient that runs just before this kernel
    // streamed x front to back, so the tail of x is what is still in L2 when the scatter starts re-reading it
    sc.enabled = g_scatter_reverse ? 2 : 1;
    TcEpilogue ep{nullptr, 0, nullptr, nullptr, 0, M, K, 0, &sc};
    // only the embedding columns matter: output width = F*D (the dense-feature columns of dx have no consumer)
    ep.N = d->F * d->D;
    return gemm_tc(dy, lddy, W, K, 1, ep, M, d->F * d->D, N, reinterpret_cast<cudaStream_t>(stream));
}
