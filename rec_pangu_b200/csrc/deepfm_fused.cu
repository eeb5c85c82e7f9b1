// DeepFM forward in ONE kernel: embedding gather -> FM second order -> layer-1 GEMM (tcgen05, 3xTF32) -> tower tail ->
// logit = fm + dnn -> sigmoid -> BCE (reference: ranking/deepfm.py:41-67 = EmbeddingLayer.forward embedding.py:49-63,
// get_linear_input utils.py:122-137, FM_Layer interaction.py:225-235, MLP deep.py:62-84, BCELoss).
//
// The separate kernels spend 77 us gathering 1.7 M table rows into the feature row x, then 79 us re-reading x for the
// layer-1 GEMM (profiles/r01_deepfm_step_ncu_full.md): the gather waits on DRAM row activations with the SMs idle, the
// GEMM is bound by tensor / shared-memory work with DRAM idle.  Here the two overlap: the four "split" warps of the
// persistent GEMM (thread = sample row) fetch their own rows straight from the tables with cp.async into a private ring
// of shared-memory stages — FG_LA k-blocks (= 2 fields x 64 B per sample each) in flight per thread, ~80 KB per SM —
// and then do what they did before: split fp32 -> (hi, lo) into tensor memory for the TS-mode MMAs.  Because a thread
// only ever reads the bytes it requested, the A operand needs no mbarriers at all (cp.async groups), and the FM sums
// fall out of registers the thread already holds.  x is written from those registers only when backward needs it.
//
// Warps: 0 = TMA producer for the pre-split weight k-blocks, 1 = MMA issuer (+TMEM alloc), 2-5 = gather + split,
// 6-13 = epilogue + tower tail (tower_tile.cuh) as in gemm_tf32x3_v2_kernel.
#include "tc_ptx.cuh"
#include "tower_tile.cuh"

namespace rpb {

constexpr int FG_THREADS = 448;
constexpr int FG_EPI_WARPS = 8;
constexpr int FG_LB = 3;                  // weight ring depth (16 KiB stages: [B hi ; B lo] of one k-block)
constexpr int FG_OP = 4;                  // tensor-memory operand ring (64 columns per stage: A hi 32 | A lo 32)
constexpr int FG_ROW = TC_BLOCK_K + 4;    // floats per staged row: 144 B, conflict-free float4 reads with thread = row
constexpr int FG_N = 64;                  // layer-1 width
constexpr int FG_FM_BUF = 4;              // per-tile FM values handed from the split warps to the tail (see kernel)
constexpr int FG_B_BYTES = 2 * FG_N * TC_BLOCK_K * 4;
// TCTAIL variant (rpb_set_option("fused_tc_tail", 1); NOT YET RUN ON HARDWARE): the 64x64 tail layers run on tcgen05 too.
// The epilogue warps turn the layer-1 accumulator into h1 = relu(acc + b1), store it, split it into (hi, lo) and write
// it back into TENSOR MEMORY as the A operand of the next layer (TS-mode MMA, the same trick the gather warps use for
// layer 1); the MMA issuer runs 8 k-steps x 2 MMAs against the resident stacked [W hi ; W lo] operand of that layer into
// the accumulator buffer the tile just vacated, and so on down the tower.  The activations never touch shared memory,
// which takes the tail's LDS.128 stream off the MIO pipe the gather warps are bound by (profiles/r01_experiments.md).
// Tensor memory: 2 x 128 accumulator columns | FT_OP x 64 layer-1 operand ring | 128 columns tail operand (hi 64 | lo 64).
constexpr int FT_OP = 2;
constexpr int FT_TAIL_B_BYTES = 2 * FG_B_BYTES;     // one tail layer: 2 k-blocks of [W hi ; W lo] (128 rows x 128 B each)

struct FusedFwdParams {
    const float* tables[RPB_MAX_FIELDS];
    const long long* idx[RPB_MAX_FIELDS];
    long long rows[RPB_MAX_FIELDS];
    const float* dense[RPB_MAX_DENSE];
    float* x; long long ldx;              // optional materialised feature row (backward needs it)
    float* fm; float* fm_s;               // optional outputs: FM term [M], sum_f e [M, 16]
    long long* err;
    float* h1; long long ldh1; const float* bias1;
    int M, F, Nd, nkb, nkb_emb;
    // row-sharded tables (RpbGatherDesc.G / shard_tab): owner = id mod G, local row = id div G; entry f*G+g of the DEVICE
    // array = rank g's shard of table f, local or mapped over NVLink (the row request is the same cp.async either way)
    int G;
    const float* const* shard_tab;
    int l2_prefetch;                      // FS kernel, unsharded: tiles by which the L2 prefetch warp runs ahead of the fetch warps (0 = off)
};

// table rows: 16-byte pieces of a 64-byte row; the L2 fetch is capped at 64 B so that a row does not drag the other half
// of its 128-byte line in (without the cap ncu showed 231 MB read for 126 MB of rows + ids)
__device__ __forceinline__ void fg_cp16(float* dst, const float* src, bool valid) {
    asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16, %2;" :: "r"(smem_u32(dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void fg_cp8(void* dst, const void* src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(smem_u32(dst)), "l"(src), "r"(valid ? 8 : 0) : "memory");
}
__device__ __forceinline__ void fg_cp4(float* dst, const float* src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(smem_u32(dst)), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void fg_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void fg_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ void fg_bad_index(long long* err, int f, int b, long long ix) {
    if (err != nullptr) {
        unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(err), 0ull, 1ull);
        if (old == 0ull) { err[1] = f; err[2] = b; err[3] = ix; __threadfence_system(); }
    }
}

// Diagnostics (rpb_debug_fused_trace): cycles of CTA 0 of the last launch — [0] kernel, [1] split: cp.async wait,
// [2] split: wait for a free TMEM operand slot, [3] split: work, [4] MMA: wait weights, [5] MMA: wait operands,
// [6] MMA: wait accumulator, [7] MMA: issue, [8] epilogue: wait accumulator, [9] epilogue: layer-1 part + wait FM,
// [10] epilogue: tail, [11] weight producer: wait free stage; 8-warp kernel only: gather phases [12] proxy fence + warp sync +
// TMA store, [13] operand loads + FM sums + hi/lo split, [14] tcgen05.st + wait, [15] next row requests (issue()).
__device__ int g_fg_trace_on = 0;
__device__ unsigned long long g_fg_trace[16];
// per-CTA record of the last traced launch (8-warp kernel): [cta][0] = SM id, [1] = start, [2] = end (globaltimer, ns), [3] = tiles
__device__ unsigned long long g_fg_cta[256 * 4];
__device__ __forceinline__ unsigned long long fg_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned int fg_smid() { unsigned int r; asm volatile("mov.u32 %0, %smid;" : "=r"(r)); return r; }
#define FG_T() (trace ? clock64() : 0ll)

template <int LA, bool SHARDED, bool TCTAIL>
__global__ void __launch_bounds__(FG_THREADS, 1)
deepfm_fwd_fused_kernel(const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                        const __grid_constant__ CUtensorMap tmThi, const __grid_constant__ CUtensorMap tmTlo,
                        const __grid_constant__ FusedFwdParams p, const __grid_constant__ TowerFwdParams tw, int m_tiles) {
    constexpr int OPN = TCTAIL ? FT_OP : FG_OP;                               // depth of the tensor-memory operand ring
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* b_base = smem;                                                   // FG_LB x 16 KiB, 1 KiB aligned (SWIZZLE_128B)
    uint8_t* t_base = b_base + FG_LB * FG_B_BYTES;                            // TCTAIL: n_tail x 32 KiB resident tail operands
    float* a_base = reinterpret_cast<float*>(t_base + (TCTAIL ? tw.n_tail * FT_TAIL_B_BYTES : 0));   // LA x 128 rows x FG_ROW floats
    long long* id_base = reinterpret_cast<long long*>(a_base + LA * TC_BLOCK_M * FG_ROW);   // LA x 128 x 2 ids
    uint64_t* bars = reinterpret_cast<uint64_t*>(id_base + LA * TC_BLOCK_M * 2);
    uint64_t* full_b = bars;                       // [FG_LB]  weight k-block landed
    uint64_t* empty_b = full_b + FG_LB;            // [FG_LB]  MMAs that read it are done
    uint64_t* ready_op = empty_b + FG_LB;          // [FG_OP]  A hi/lo of a k-block are in tensor memory
    uint64_t* empty_op = ready_op + FG_OP;         // [FG_OP]
    uint64_t* tmem_full = empty_op + FG_OP;        // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint64_t* fm_ready = tmem_empty + 2;           // [FG_FM_BUF]  FM values of a tile written
    uint64_t* tail_bars = fm_ready + FG_FM_BUF;    // [4] TCTAIL: [0] tail weights landed, [1] tail operand written, [2] tail MMAs done
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tail_bars + 4);
    float* fm_tile = reinterpret_cast<float*>(tmem_ptr + 4);                  // [FG_FM_BUF][128]
    float* tw_As = fm_tile + FG_FM_BUF * TC_BLOCK_M;                          // 16-byte aligned: every block above is
    float* tw_Bs = tw_As + (TCTAIL ? TC_BLOCK_M : TC_BLOCK_M * TW_LDA);       // TCTAIL: tw_As = 128 head partials only
    float* tw_loss = tw_Bs + (TCTAIL ? 0 : tw.n_tail * TW_H * TW_H);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.nkb;
    const int my_tiles = ((int)blockIdx.x < m_tiles) ? (m_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const uint32_t G = (uint32_t)my_tiles * (uint32_t)nkb;                    // k-blocks this CTA walks
    constexpr uint32_t ACC_STRIDE = 2 * FG_N;                                 // stacked accumulator: [a.b_hi | a.b_lo]
    constexpr uint32_t A_COL = 2 * ACC_STRIDE;                                // first TMEM column of the operand ring
    constexpr uint32_t TAIL_A = A_COL + FT_OP * 64u;                          // TCTAIL: tail operand, hi [0,64) | lo [64,128)

    if (threadIdx.x == 0) {
        for (int s = 0; s < FG_LB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < FG_OP; ++s) { mbar_init(&ready_op[s], 128); mbar_init(&empty_op[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], FG_EPI_WARPS); }
        for (int s = 0; s < FG_FM_BUF; ++s) mbar_init(&fm_ready[s], 128);
        mbar_init(&tail_bars[0], 1); mbar_init(&tail_bars[1], FG_EPI_WARPS * 32); mbar_init(&tail_bars[2], 1); mbar_init(&tail_bars[3], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const bool trace = g_fg_trace_on != 0 && blockIdx.x == 0;
    const long long t_start = FG_T();

    if (warp == 0) {
        // ---------------- weight producer: [B hi ; B lo] of every k-block through a FG_LB-deep ring
        if (lane == 0) {
            if constexpr (TCTAIL) {
                // resident tail operands: layer l, k-block kb -> [W_l hi ; W_l lo] columns kb*32 .. kb*32+31 (128 rows x 128 B)
                mbar_arrive_expect_tx(&tail_bars[0], (uint32_t)(tw.n_tail * FT_TAIL_B_BYTES));
                for (int l = 0; l < tw.n_tail; ++l)
                    for (int kb = 0; kb < 2; ++kb) {
                        uint8_t* st = t_base + (size_t)(l * 2 + kb) * FG_B_BYTES;
                        tma_load_2d(st, &tmThi, &tail_bars[0], kb * TC_BLOCK_K, l * TW_H);
                        tma_load_2d(st + FG_B_BYTES / 2, &tmTlo, &tail_bars[0], kb * TC_BLOCK_K, l * TW_H);
                    }
            }
            long long w_b = 0;
            for (uint32_t g = 0; g < G; ++g) {
                const int s = g % FG_LB, kb = g % nkb;
                const long long c0 = FG_T();
                mbar_wait(&empty_b[s], ((g / FG_LB) & 1u) ^ 1u);
                w_b += FG_T() - c0;
                if (trace && g + 1 == G) g_fg_trace[11] = (unsigned long long)w_b;
                uint8_t* st = b_base + (size_t)s * FG_B_BYTES;
                mbar_arrive_expect_tx(&full_b[s], (uint32_t)FG_B_BYTES);
                tma_load_2d(st, &tmBhi, &full_b[s], kb * TC_BLOCK_K, 0);
                tma_load_2d(st + FG_B_BYTES / 2, &tmBlo, &full_b[s], kb * TC_BLOCK_K, 0);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: per k-step two TS-mode MMAs (a_lo, a_hi) against the stacked 128-row weight operand
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TC_BLOCK_M, 2 * FG_N);
            uint32_t g = 0;
            long long w_fb = 0, w_op = 0, w_acc = 0, w_is = 0;
            // TCTAIL: tail steps are issued in (tile, layer) order as soon as the epilogue warps have written the operand
            // (tail_bars[1]); between k-blocks of the running tile without blocking, and blocking before an accumulator
            // buffer is re-used (tile t + 2 needs tile t's tail finished) and after the last tile.
            int tl_t = 0, tl_l = 0; uint32_t tl_n = 0;
            auto tail_step = [&](bool block) -> bool {
                if (!block && !mbar_test(&tail_bars[1], tl_n & 1u)) return false;
                mbar_wait(&tail_bars[1], tl_n & 1u);
                if (tl_n == 0) mbar_wait(&tail_bars[0], 0u);               // resident tail operands have landed
                tc_fence_after();
                const uint32_t d_t = tmem_base + ((uint32_t)tl_t & 1u) * ACC_STRIDE;
                const uint32_t tb = smem_u32(t_base + (size_t)tl_l * FT_TAIL_B_BYTES);
#pragma unroll
                for (int k = 0; k < TW_H / TC_UMMA_K; ++k) {
                    const uint64_t db = make_kmajor_sw128_desc(tb + (uint32_t)(k >> 2) * FG_B_BYTES + (uint32_t)(k & 3) * TC_UMMA_K * 4);
                    umma_tf32_ts(d_t, tmem_base + TAIL_A + 64u + k * TC_UMMA_K, db, idesc, k > 0 ? 1u : 0u);
                    umma_tf32_ts(d_t, tmem_base + TAIL_A + k * TC_UMMA_K, db, idesc, 1u);
                }
                umma_commit(&tail_bars[2]);
                ++tl_n;
                if (++tl_l == tw.n_tail) { tl_l = 0; ++tl_t; }
                return true;
            };
            for (int t = 0; t < my_tiles; ++t) {
                const uint32_t acc = (uint32_t)t & 1u;
                const long long c0 = FG_T();
                if constexpr (TCTAIL) { while (tl_t + 2 <= t) tail_step(true); }
                mbar_wait(&tmem_empty[acc], (((uint32_t)t >> 1) & 1u) ^ 1u);
                w_acc += FG_T() - c0;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % FG_LB, o = g % OPN;
                    if constexpr (TCTAIL) { if (tl_t < t) tail_step(false); }
                    const long long c1 = FG_T();
                    mbar_wait(&full_b[s], (g / FG_LB) & 1u);
                    const long long c2 = FG_T();
                    mbar_wait(&ready_op[o], (g / OPN) & 1u);
                    const long long c3 = FG_T();
                    w_fb += c2 - c1; w_op += c3 - c2;
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(b_base + (size_t)s * FG_B_BYTES);
                    const uint32_t ta_hi = tmem_base + A_COL + (uint32_t)o * 64u, ta_lo = ta_hi + 32u;
#pragma unroll
                    for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
                        const uint64_t db = make_kmajor_sw128_desc(b_addr + k * TC_UMMA_K * 4);
                        umma_tf32_ts(d_tmem, ta_lo + k * TC_UMMA_K, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_tf32_ts(d_tmem, ta_hi + k * TC_UMMA_K, db, idesc, 1u);
                    }
                    umma_commit(&empty_op[o]);
                    umma_commit(&empty_b[s]);
                    w_is += FG_T() - c3;
                }
                umma_commit(&tmem_full[acc]);
            }
            if constexpr (TCTAIL) { while (tl_t < my_tiles) tail_step(true); }
            if (trace) { g_fg_trace[4] = (unsigned long long)w_fb; g_fg_trace[5] = (unsigned long long)w_op; g_fg_trace[6] = (unsigned long long)w_acc; g_fg_trace[7] = (unsigned long long)w_is; }
        }
    } else if (warp < 6) {
        // ---------------- gather + split warps: thread = sample row of the tile (= its TMEM lane)
        const int r = (warp & 3) * 32 + lane;
        const int K_emb_cols = p.F * 16;
        // issue(): request the table rows (or dense columns) of the next un-requested k-block `gi` into A stage gi % LA, and
        // the ids of k-block gi + LA - 1 into the id FIFO; one cp.async group per call (empty past the end, so the group
        // counts stay uniform).  Row requests are made COOPERATIVELY by the warp: 4 consecutive lanes fetch the four 16-byte
        // pieces of one 64-byte row, so a warp instruction is 8 whole-row requests instead of 32 quarter-row ones (the
        // thread = row mapping cost ~2 k cycles of LSU time per k-block: tools/exp/trace_fused.py).  A warp therefore reads
        // ids and writes stage rows of its OWN 32 rows only, and __syncwarp() is the only synchronisation the A path needs.
        const int wrow0 = (warp & 3) * 32;               // first tile row of this warp
        uint32_t gi = 0; int i_kb = 0, i_t = 0;          // next k-block to request: global index, k-block, local tile
        uint32_t gj = (uint32_t)(LA - 1); int j_kb = (LA - 1) % nkb, j_t = (LA - 1) / nkb;     // next ids to request
        auto issue = [&]() {
            __syncwarp();
            if (gi < G) {
                const int mt = ((int)blockIdx.x + i_t * (int)gridDim.x) * TC_BLOCK_M;
                const int slot = (int)(gi % LA);
                if (i_kb < p.nkb_emb) {
                    const int piece = lane & 3;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = wrow0 + (i & 3) * 8 + (lane >> 2), fsel = i >> 2, f = 2 * i_kb + fsel;
                        const bool ok = mt + row < p.M;
                        long long id = id_base[(slot * TC_BLOCK_M + row) * 2 + fsel];
                        if (!ok) id = 0;
                        else if ((unsigned long long)id >= (unsigned long long)p.rows[f]) { if (piece == 0) fg_bad_index(p.err, f, mt + row, id); id = 0; }
                        const float* src;
                        if constexpr (SHARDED) {
                            const unsigned iu = (unsigned)id, gg = (unsigned)p.G;
                            const float* base = reinterpret_cast<const float*>(__ldg(reinterpret_cast<const unsigned long long*>(p.shard_tab) + (size_t)f * gg + (iu % gg)));
                            src = base + (size_t)(iu / gg) * 16 + piece * 4;
                        } else {
                            src = p.tables[f] + (size_t)id * 16 + piece * 4;
                        }
                        fg_cp16(a_base + (slot * TC_BLOCK_M + row) * FG_ROW + fsel * 16 + piece * 4, src, ok);
                    }
                } else {
                    const int m = mt + r;
                    const bool ok = m < p.M;
                    float* dst = a_base + (slot * TC_BLOCK_M + r) * FG_ROW;
                    const int c0 = i_kb * TC_BLOCK_K - K_emb_cols;           // first dense column of this k-block
#pragma unroll 4
                    for (int j = 0; j < TC_BLOCK_K; ++j) {
                        const int c = c0 + j;
                        if (c < p.Nd) fg_cp4(dst + j, p.dense[c] + (ok ? m : 0), ok);
                        else dst[j] = 0.f;
                    }
                }
            }
            if (gj < G && j_kb < p.nkb_emb) {
                const int m = ((int)blockIdx.x + j_t * (int)gridDim.x) * TC_BLOCK_M + r;
                long long* ids = id_base + ((int)(gj % LA) * TC_BLOCK_M + r) * 2;
                const bool ok = m < p.M;
                fg_cp8(ids, p.idx[2 * j_kb] + (ok ? m : 0), ok);
                fg_cp8(ids + 1, p.idx[2 * j_kb + 1] + (ok ? m : 0), ok);
            }
            fg_commit();
            ++gi; if (++i_kb == nkb) { i_kb = 0; ++i_t; }
            ++gj; if (++j_kb == nkb) { j_kb = 0; ++j_t; }
        };
        // prologue: ids of the first LA-1 k-blocks with plain loads, then LA-1 groups in flight
        {
            int kb = 0, tl = 0;
            for (uint32_t q = 0; q + 1 < (uint32_t)LA && q < G; ++q) {
                const int m = ((int)blockIdx.x + tl * (int)gridDim.x) * TC_BLOCK_M + r;
                long long* ids = id_base + ((int)(q % LA) * TC_BLOCK_M + r) * 2;
                const bool ok = kb < p.nkb_emb && m < p.M;
                ids[0] = ok ? __ldg(p.idx[2 * kb] + m) : 0;
                ids[1] = ok ? __ldg(p.idx[2 * kb + 1] + m) : 0;
                if (++kb == nkb) { kb = 0; ++tl; }
            }
        }
        for (int q = 0; q + 1 < LA; ++q) issue();

        float fs[16];                                   // sum_f e of this sample, and the sum of squares
        float fq = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) fs[j] = 0.f;
        uint32_t g = 0;
        long long w_cp = 0, w_eo = 0, w_wk = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int m = ((int)blockIdx.x + t * (int)gridDim.x) * TC_BLOCK_M + r;
            for (int kb = 0; kb < nkb; ++kb, ++g) {
                const long long c0 = FG_T();
                fg_wait<LA - 2>();                      // group g has landed: A(g) and the ids of k-block g + LA - 1
                const long long c1 = FG_T();
                issue();                                // k-block g + LA - 1 (starts with a __syncwarp: rows of group g visible warp-wide)
                const int o = g % OPN;
                const long long c2 = FG_T();
                mbar_wait(&empty_op[o], ((g / OPN) & 1u) ^ 1u);
                const long long c3 = FG_T();
                w_cp += c1 - c0; w_eo += c3 - c2; w_wk += c2 - c1;
                tc_fence_after();
                const float4* src = reinterpret_cast<const float4*>(a_base + ((g % LA) * TC_BLOCK_M + r) * FG_ROW);
                float4 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = src[j];
                if (p.x != nullptr) {                   // materialise the feature row (training): 8 lanes write one row's 128 B
                    const int mt = m - r, piece = lane & 7;
                    const int nv = min(8, ((int)p.ldx - kb * TC_BLOCK_K) >> 2);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = wrow0 + i * 4 + (lane >> 3);
                        if (piece < nv && mt + row < p.M)
                            stg_f4(p.x + (size_t)(mt + row) * p.ldx + kb * TC_BLOCK_K + piece * 4,
                                   *reinterpret_cast<const float4*>(a_base + ((g % LA) * TC_BLOCK_M + row) * FG_ROW + piece * 4));
                    }
                }
                if (kb < p.nkb_emb) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 a = v[j], b = v[4 + j];
                        fs[4 * j + 0] += a.x + b.x; fs[4 * j + 1] += a.y + b.y; fs[4 * j + 2] += a.z + b.z; fs[4 * j + 3] += a.w + b.w;
                        fq = fmaf(a.x, a.x, fq); fq = fmaf(a.y, a.y, fq); fq = fmaf(a.z, a.z, fq); fq = fmaf(a.w, a.w, fq);
                        fq = fmaf(b.x, b.x, fq); fq = fmaf(b.y, b.y, fq); fq = fmaf(b.z, b.z, fq); fq = fmaf(b.w, b.w, fq);
                    }
                }
                uint32_t h[32], l[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 q = v[j];
                    h[4 * j + 0] = __float_as_uint(q.x) & 0xFFFFE000u; l[4 * j + 0] = __float_as_uint(q.x - __uint_as_float(h[4 * j + 0]));
                    h[4 * j + 1] = __float_as_uint(q.y) & 0xFFFFE000u; l[4 * j + 1] = __float_as_uint(q.y - __uint_as_float(h[4 * j + 1]));
                    h[4 * j + 2] = __float_as_uint(q.z) & 0xFFFFE000u; l[4 * j + 2] = __float_as_uint(q.z - __uint_as_float(h[4 * j + 2]));
                    h[4 * j + 3] = __float_as_uint(q.w) & 0xFFFFE000u; l[4 * j + 3] = __float_as_uint(q.w - __uint_as_float(h[4 * j + 3]));
                }
                const uint32_t ta = tmem_base + A_COL + (uint32_t)o * 64u + ((uint32_t)((warp & 3) * 32) << 16);
                tmem_st32(ta, h);
                tmem_st32(ta + 32u, l);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&ready_op[o]);
                w_wk += FG_T() - c3;
            }
            // FM second order of this sample: 0.5 * (sum_d s_d^2 - sum_{f,d} e^2); handed to the tail through shared memory
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) ss = fmaf(fs[j], fs[j], ss);
            const float fmv = 0.5f * (ss - fq);
            fm_tile[(t % FG_FM_BUF) * TC_BLOCK_M + r] = fmv;
            if (m < p.M) {
                if (p.fm != nullptr) p.fm[m] = fmv;
                if (p.fm_s != nullptr) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        stg_f4(p.fm_s + (size_t)m * 16 + 4 * j, make_float4(fs[4 * j], fs[4 * j + 1], fs[4 * j + 2], fs[4 * j + 3]));
                }
            }
            mbar_arrive(&fm_ready[t % FG_FM_BUF]);
            fq = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) fs[j] = 0.f;
        }
        fg_wait<0>();
        if (trace && r == 0) { g_fg_trace[1] = (unsigned long long)w_cp; g_fg_trace[2] = (unsigned long long)w_eo; g_fg_trace[3] = (unsigned long long)w_wk; }
    } else {
        // ---------------- epilogue warps: layer-1 epilogue -> h1 (HBM + shared memory) -> tower tail (tower_tile.cuh)
        const int quarter = warp & 3;
        const int half = (warp - 6) >> 2;
        const int row = quarter * 32 + lane;
        const int et = threadIdx.x - 6 * 32;
        auto epi_sync = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
        float loss_acc = 0.f;
        if constexpr (TCTAIL) {
            const float tw_bo = tw.b_out != nullptr ? __ldg(tw.b_out) : 0.f;
            // Round r = 0 .. n_tail of a tile: read the accumulator of layer r (r = 0: layer 1; both stacked halves), add the
            // bias, ReLU, store the activation row piece for backward, and either hand it back to the tensor core as the
            // next layer's operand (hi | lo in tensor memory) or, after the last layer, fold it into the output row-dot.
            // Thread = (row, half): the two warps of a lane quarter own columns {half*16 .. +15} and {32 + half*16 .. +15}.
            const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
            const int L = tw.n_tail;
            uint32_t n_out = 0;                                   // tail_bars[2] phases consumed
            for (int t = 0; t < my_tiles; ++t) {
                const int m0 = ((int)blockIdx.x + t * (int)gridDim.x) * TC_BLOCK_M;
                const uint32_t acc = (uint32_t)t & 1u;
                const int m = m0 + row;
                const long long q0 = FG_T();
                mbar_wait(&tmem_full[acc], ((uint32_t)t >> 1) & 1u);
                tc_fence_after();
                const long long q1 = FG_T();
                const uint32_t d_acc = tmem_base + acc * ACC_STRIDE + lane_addr;
                float headp = 0.f;
                for (int r = 0; r <= L; ++r) {
                    if (r > 0) { mbar_wait(&tail_bars[2], n_out & 1u); ++n_out; tc_fence_after(); }
                    const float* bias = r == 0 ? p.bias1 : tw.b[r - 1];
                    float* hout = r == 0 ? p.h1 : tw.h[r - 1];
                    const long long ldh = r == 0 ? p.ldh1 : (long long)TW_H;
                    for (int c0 = half * 16; c0 < FG_N; c0 += 32) {
                        uint32_t a0[16], a1[16];
                        tmem_ld16(d_acc + (uint32_t)c0, a0);
                        tmem_ld16(d_acc + (uint32_t)(FG_N + c0), a1);
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            v[j] = fmaxf(__uint_as_float(a1[j]) + __uint_as_float(a0[j]) + __ldg(bias + c0 + j), 0.f);
                        if (m < p.M) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                stg_f4(hout + (size_t)m * ldh + c0 + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                        }
                        if (r < L) {
                            uint32_t hi[16], lo[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                hi[j] = __float_as_uint(v[j]) & 0xFFFFE000u;
                                lo[j] = __float_as_uint(v[j] - __uint_as_float(hi[j]));
                            }
                            tmem_st16(tmem_base + TAIL_A + lane_addr + (uint32_t)c0, hi);
                            tmem_st16(tmem_base + TAIL_A + 64u + lane_addr + (uint32_t)c0, lo);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) headp = fmaf(v[j], __ldg(tw.w_out + c0 + j), headp);
                        }
                    }
                    if (r < L) {
                        tmem_st_wait();
                        tc_fence_before();
                        mbar_arrive(&tail_bars[1]);               // 256 arrivals: the operand of tail layer r is complete
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);     // every read of this accumulator buffer is done
                const long long q2 = FG_T();
                mbar_wait(&fm_ready[t % FG_FM_BUF], ((uint32_t)t / FG_FM_BUF) & 1u);
                float* head_part = tw_As;                         // [128] partial row-dots of the half-1 warps
                if (half == 1) head_part[row] = headp;
                epi_sync();
                if (half == 0 && m < p.M) {
                    const float z = headp + head_part[row] + tw_bo + fm_tile[(t % FG_FM_BUF) * TC_BLOCK_M + row];
                    tw.logit[m] = z;
                    if (tw.pred != nullptr) {
                        const float q = 1.f / (1.f + expf(-z));
                        tw.pred[m] = q;
                        if (tw.label != nullptr) {
                            const float y = __ldg(tw.label + m);
                            const float pe = q + tw.eps;
                            const float l1 = fmaxf(logf(pe), -100.f);
                            const float l0 = fmaxf(logf(1.f - pe), -100.f);
                            loss_acc += -(y * l1 + (1.f - y) * l0);
                        }
                    }
                }
                epi_sync();                                       // head_part is free for the next tile
                if (trace && et == 0) {
                    g_fg_trace[8] += (unsigned long long)(q1 - q0); g_fg_trace[9] += (unsigned long long)(q2 - q1);
                    g_fg_trace[10] += (unsigned long long)(FG_T() - q2);
                }
            }
        } else {
        tower_load_weights_t<FG_EPI_WARPS * 32>(tw, tw_Bs, et);
        const float4 tw_wo = ldg_f4(tw.w_out + (et & 15) * 4);
        const float tw_bo = tw.b_out != nullptr ? __ldg(tw.b_out) : 0.f;
        for (int t = 0; t < my_tiles; ++t) {
            const int m0 = ((int)blockIdx.x + t * (int)gridDim.x) * TC_BLOCK_M;
            const uint32_t acc = (uint32_t)t & 1u;
            const int m = m0 + row;
            const long long q0 = FG_T();
            mbar_wait(&tmem_full[acc], ((uint32_t)t >> 1) & 1u);
            tc_fence_after();
            const long long q1 = FG_T();
            epi_sync();                                // previous tile's activations fully consumed (weights visible)
            for (int c0 = half * 16; c0 < FG_N; c0 += 32) {
                uint32_t a0[16], a1[16];
                tmem_ld16(tmem_base + acc * ACC_STRIDE + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, a0);
                tmem_ld16(tmem_base + acc * ACC_STRIDE + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(FG_N + c0), a1);
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    v[j] = fmaxf(__uint_as_float(a1[j]) + __uint_as_float(a0[j]) + __ldg(p.bias1 + c0 + j), 0.f);
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 q4 = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    if (m < p.M) stg_f4(p.h1 + (size_t)m * p.ldh1 + c0 + j, q4);
                    *reinterpret_cast<float4*>(tw_As + row * TW_LDA + c0 + j) = q4;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            mbar_wait(&fm_ready[t % FG_FM_BUF], ((uint32_t)t / FG_FM_BUF) & 1u);
            epi_sync();
            const long long q2 = FG_T();
            tower_tail_tile_fwd<FG_EPI_WARPS * 32>(tw, tw_As, tw_Bs, m0, et, tw_wo, tw_bo, loss_acc, epi_sync,
                                                   fm_tile + (t % FG_FM_BUF) * TC_BLOCK_M);
            if (trace && et == 0) {
                g_fg_trace[8] += (unsigned long long)(q1 - q0); g_fg_trace[9] += (unsigned long long)(q2 - q1);
                g_fg_trace[10] += (unsigned long long)(FG_T() - q2);
            }
        }
        }
        if (tw.loss != nullptr) tw_loss[et] = loss_acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
    if (trace && threadIdx.x == 0) g_fg_trace[0] = (unsigned long long)(clock64() - t_start);
    if (tw.loss != nullptr && warp == 0) {
        // deterministic mean BCE: 256 epilogue partials -> per-CTA partial -> the last CTA adds them in index order
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < FG_EPI_WARPS; ++i) s += tw_loss[lane + 32 * i];
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
}


// =====================================================================================================================
// v2 of the one-kernel forward (default): EIGHT gather/split warps and the feature row stored by TMA.
//
// The per-role trace of the 4-warp kernel above (profiles/r01_experiments.md, BENCH_r01 experiments.tc_fwd) showed the
// four gather/split warps busy for 159 k of 210 k cycles per CTA — one warp per SM sub-partition doing, per k-block and
// row, the row requests, 8 LDS.128, the FM sums, 32 hi/lo splits, two tcgen05.st.x32 and 8 cooperative STG.128 of x,
// every instruction at its full dependent latency — while the MMA thread waited for operands 60 % of the time.  Here
//   * warps 2-9 gather: warps w and w+4 share a TMEM lane quarter (rows) and own ONE FIELD each of a k-block (16 of its
//     32 columns), i.e. half the work per thread and two independent instruction streams per sub-partition;
//   * a warp's rows of a k-block live in a private [32 rows x 64 B] stage written by cp.async in the TMA SWIZZLE_64B
//     pattern (thread = row LDS.128 stays conflict-free without padding), and the feature row x leaves the SM as ONE
//     `cp.async.bulk.tensor.2d.global.shared::cta` per warp and k-block issued by lane 0 — no STG, no second LDS pass;
//   * the FM sums of the two field halves meet once per tile through shared memory (named barrier per lane quarter).
// Warps: 0 = weight TMA producer, 1 = MMA issuer (+TMEM alloc), 2-9 = gather + split, 10-17 = epilogue + tower tail.
constexpr int F8_THREADS = 576;
constexpr int F8_GW = 8;                          // gather warps
constexpr int F8_STAGE_BYTES = F8_GW * 32 * 64;   // one k-block of a 128-row tile: 8 private [32 x 64 B] regions = 16 KiB
constexpr int F8_FMX_LD = 17;                     // floats per row of the FM exchange (sum_f e [16] | sum e^2), odd stride

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }

// Diagnostics as for the kernel above; gather rows describe warp 2 (field half 0, lane quarter 2).
// TCTAIL (default): the 64x64 tail layers run on tcgen05 as well — the epilogue warps turn an accumulator into the activation
// row h = relu(acc + b), store it for backward, split it into (hi, lo) and write it back into TENSOR MEMORY as the A operand
// of the next layer; the MMA thread issues that layer's 8 k-steps x 2 TS-mode MMAs against the resident stacked
// [W hi ; W lo] operand (TMA-loaded once per CTA) into the accumulator buffer the tile just vacated, between the k-blocks of
// the next tile.  Activations never touch shared memory, and the drain after a CTA's last gather shrinks from ~25 k cycles
// (two fp32 64x64 layers on 8 CUDA-core warps) to the latency of the TMEM round trips.
// Tensor memory: 2 x 128 accumulator columns | OPN x 64 layer-1 operand ring | TCTAIL: 128 columns tail operand (hi | lo).
template <int LA, bool SHARDED, bool TCTAIL>
__global__ void __launch_bounds__(F8_THREADS, 1)
deepfm_fwd_fused8_kernel(const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                         const __grid_constant__ CUtensorMap tmThi, const __grid_constant__ CUtensorMap tmTlo,
                         const __grid_constant__ CUtensorMap tmX,
                         const __grid_constant__ FusedFwdParams p, const __grid_constant__ TowerFwdParams tw, int full_rounds, int chunk) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int OPN = TCTAIL ? FT_OP : FG_OP;                               // depth of the tensor-memory operand ring
    uint8_t* b_base = smem;                                                   // FG_LB x 16 KiB, 1 KiB aligned (SWIZZLE_128B)
    uint8_t* t_base = b_base + FG_LB * FG_B_BYTES;                            // TCTAIL: n_tail x 32 KiB resident tail operands
    uint8_t* a_base = t_base + (TCTAIL ? tw.n_tail * FT_TAIL_B_BYTES : 0);    // LA x 16 KiB: [stage][gather warp][32 rows][64 B], SWIZZLE_64B
    long long* id_base = reinterpret_cast<long long*>(a_base + LA * F8_STAGE_BYTES);     // [LA][gather warp][32] ids
    uint64_t* bars = reinterpret_cast<uint64_t*>(id_base + LA * F8_GW * 32);
    uint64_t* full_b = bars;                       // [FG_LB]  weight k-block landed
    uint64_t* empty_b = full_b + FG_LB;            // [FG_LB]  MMAs that read it are done
    uint64_t* ready_op = empty_b + FG_LB;          // [FG_OP]  A hi/lo of a k-block are in tensor memory (one arrival per gather warp)
    uint64_t* empty_op = ready_op + FG_OP;         // [FG_OP]
    uint64_t* tmem_full = empty_op + FG_OP;        // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint64_t* fm_ready = tmem_empty + 2;           // [FG_FM_BUF]  FM values of a tile written (4 arrivals: the half-0 warps)
    uint64_t* tail_bars = fm_ready + FG_FM_BUF;    // [4] TCTAIL: [0] tail weights landed, [1] tail operand written, [2] tail MMAs done
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tail_bars + 4);
    float* fm_tile = reinterpret_cast<float*>(tmem_ptr + 4);                  // [FG_FM_BUF][128]
    float* fm_x = fm_tile + FG_FM_BUF * TC_BLOCK_M;                           // [128][F8_FMX_LD] half-1 -> half-0 FM partials
    float* tw_As = fm_x + TC_BLOCK_M * F8_FMX_LD;                             // 128 * 17 floats: still 16-byte aligned
    float* tw_Bs = tw_As + (TCTAIL ? TC_BLOCK_M : TC_BLOCK_M * TW_LDA);       // TCTAIL: tw_As = 128 head partials only
    float* tw_loss = tw_Bs + (TCTAIL ? 0 : tw.n_tail * TW_H * TW_H);
    static_assert((TC_BLOCK_M * F8_FMX_LD) % 4 == 0, "fm_x must keep tw_As 16-byte aligned");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.nkb;
    // Tile schedule: `full_rounds` rounds of 128-row tiles over all CTAs, then ONE partial round in which the remaining rows
    // are cut into `chunk`-row pieces (a multiple of 32, <= 128) so that every CTA gets a piece: at config 2 (512 tiles on
    // 148 CTAs) 68 CTAs used to run a 4th full tile while 80 idled — 22 % of the kernel; now 136 CTAs run a 64-row piece.
    // A piece is an ordinary tile whose rows >= tile_rows(t) are masked (no row requests, no stores).
    const int part_m0 = full_rounds * (int)gridDim.x * TC_BLOCK_M + (int)blockIdx.x * chunk;
    const int my_tiles = full_rounds + ((chunk > 0 && part_m0 < p.M) ? 1 : 0);
    auto tile_m0 = [&](int t) -> int { return t < full_rounds ? ((int)blockIdx.x + t * (int)gridDim.x) * TC_BLOCK_M : part_m0; };
    auto tile_rows = [&](int t) -> int { return min(t < full_rounds ? TC_BLOCK_M : chunk, p.M - tile_m0(t)); };
    const uint32_t G = (uint32_t)my_tiles * (uint32_t)nkb;                    // k-blocks this CTA walks
    constexpr uint32_t ACC_STRIDE = 2 * FG_N;                                 // stacked accumulator: [a.b_hi | a.b_lo]
    constexpr uint32_t A_COL = 2 * ACC_STRIDE;                                // first TMEM column of the operand ring
    constexpr uint32_t TAIL_A = A_COL + FT_OP * 64u;                          // TCTAIL: tail operand, hi [0,64) | lo [64,128)

    if (threadIdx.x == 0) {
        for (int s = 0; s < FG_LB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < FG_OP; ++s) { mbar_init(&ready_op[s], F8_GW); mbar_init(&empty_op[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], FG_EPI_WARPS); }
        for (int s = 0; s < FG_FM_BUF; ++s) mbar_init(&fm_ready[s], 4);
        mbar_init(&tail_bars[0], 1); mbar_init(&tail_bars[1], FG_EPI_WARPS); mbar_init(&tail_bars[2], 1); mbar_init(&tail_bars[3], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const bool trace = g_fg_trace_on != 0 && blockIdx.x == 0;
    const long long t_start = FG_T();
    if (g_fg_trace_on != 0 && threadIdx.x == 0 && blockIdx.x < 256) {
        g_fg_cta[blockIdx.x * 4 + 0] = fg_smid(); g_fg_cta[blockIdx.x * 4 + 1] = fg_globaltimer(); g_fg_cta[blockIdx.x * 4 + 3] = (unsigned long long)my_tiles;
    }

    if (warp == 0) {
        // ---------------- weight producer: [B hi ; B lo] of every k-block through a FG_LB-deep ring
        if (lane == 0) {
            if constexpr (TCTAIL) {
                // resident tail operands: layer l, k-block kb -> [W_l hi ; W_l lo] columns kb*32 .. kb*32+31 (128 rows x 128 B)
                mbar_arrive_expect_tx(&tail_bars[0], (uint32_t)(tw.n_tail * FT_TAIL_B_BYTES));
                for (int l = 0; l < tw.n_tail; ++l)
                    for (int kb = 0; kb < 2; ++kb) {
                        uint8_t* st = t_base + (size_t)(l * 2 + kb) * FG_B_BYTES;
                        tma_load_2d(st, &tmThi, &tail_bars[0], kb * TC_BLOCK_K, l * TW_H);
                        tma_load_2d(st + FG_B_BYTES / 2, &tmTlo, &tail_bars[0], kb * TC_BLOCK_K, l * TW_H);
                    }
            }
            long long w_b = 0;
            for (uint32_t g = 0; g < G; ++g) {
                const int s = g % FG_LB, kb = g % nkb;
                const long long c0 = FG_T();
                mbar_wait(&empty_b[s], ((g / FG_LB) & 1u) ^ 1u);
                w_b += FG_T() - c0;
                if (trace && g + 1 == G) g_fg_trace[11] = (unsigned long long)w_b;
                uint8_t* st = b_base + (size_t)s * FG_B_BYTES;
                mbar_arrive_expect_tx(&full_b[s], (uint32_t)FG_B_BYTES);
                tma_load_2d(st, &tmBhi, &full_b[s], kb * TC_BLOCK_K, 0);
                tma_load_2d(st + FG_B_BYTES / 2, &tmBlo, &full_b[s], kb * TC_BLOCK_K, 0);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: per k-step two TS-mode MMAs (a_lo, a_hi) against the stacked 128-row weight operand
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TC_BLOCK_M, 2 * FG_N);
            uint32_t g = 0;
            long long w_fb = 0, w_op = 0, w_acc = 0, w_is = 0;
            // TCTAIL: tail steps are issued in (tile, layer) order as soon as the epilogue warps have written the operand
            // (tail_bars[1]); between k-blocks of the running tile without blocking, and blocking before an accumulator
            // buffer is re-used (tile t + 2 needs tile t's tail finished) and after the last tile.
            int tl_t = 0, tl_l = 0; uint32_t tl_n = 0;
            auto tail_step = [&](bool block) -> bool {
                if (!block && !mbar_test(&tail_bars[1], tl_n & 1u)) return false;
                mbar_wait(&tail_bars[1], tl_n & 1u);
                if (tl_n == 0) mbar_wait(&tail_bars[0], 0u);               // resident tail operands have landed
                tc_fence_after();
                const uint32_t d_t = tmem_base + ((uint32_t)tl_t & 1u) * ACC_STRIDE;
                const uint32_t tb = smem_u32(t_base + (size_t)tl_l * FT_TAIL_B_BYTES);
#pragma unroll
                for (int kk = 0; kk < TW_H / TC_UMMA_K; ++kk) {
                    const uint64_t db = make_kmajor_sw128_desc(tb + (uint32_t)(kk >> 2) * FG_B_BYTES + (uint32_t)(kk & 3) * TC_UMMA_K * 4);
                    umma_tf32_ts(d_t, tmem_base + TAIL_A + 64u + kk * TC_UMMA_K, db, idesc, kk > 0 ? 1u : 0u);
                    umma_tf32_ts(d_t, tmem_base + TAIL_A + kk * TC_UMMA_K, db, idesc, 1u);
                }
                umma_commit(&tail_bars[2]);
                ++tl_n;
                if (++tl_l == tw.n_tail) { tl_l = 0; ++tl_t; }
                return true;
            };
            for (int t = 0; t < my_tiles; ++t) {
                const uint32_t acc = (uint32_t)t & 1u;
                const long long c0 = FG_T();
                if constexpr (TCTAIL) { while (tl_t + 2 <= t) tail_step(true); }
                mbar_wait(&tmem_empty[acc], (((uint32_t)t >> 1) & 1u) ^ 1u);
                w_acc += FG_T() - c0;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % FG_LB, o = g % OPN;
                    if constexpr (TCTAIL) { if (tl_t < t) tail_step(false); }
                    const long long c1 = FG_T();
                    mbar_wait(&full_b[s], (g / FG_LB) & 1u);
                    const long long c2 = FG_T();
                    if constexpr (TCTAIL) {
                        // do not sit on the operand barrier while a tail layer of the previous tile becomes ready
                        while (!mbar_test(&ready_op[o], (g / OPN) & 1u)) { if (tl_t < t) tail_step(false); }
                    }
                    mbar_wait(&ready_op[o], (g / OPN) & 1u);
                    const long long c3 = FG_T();
                    w_fb += c2 - c1; w_op += c3 - c2;
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(b_base + (size_t)s * FG_B_BYTES);
                    const uint32_t ta_hi = tmem_base + A_COL + (uint32_t)o * 64u, ta_lo = ta_hi + 32u;
#pragma unroll
                    for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
                        const uint64_t db = make_kmajor_sw128_desc(b_addr + k * TC_UMMA_K * 4);
                        umma_tf32_ts(d_tmem, ta_lo + k * TC_UMMA_K, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_tf32_ts(d_tmem, ta_hi + k * TC_UMMA_K, db, idesc, 1u);
                    }
                    umma_commit(&empty_op[o]);
                    umma_commit(&empty_b[s]);
                    w_is += FG_T() - c3;
                }
                umma_commit(&tmem_full[acc]);
            }
            if (trace) g_fg_cta[804] = (unsigned long long)(clock64() - t_start);      // last layer-1 MMA issued
            if constexpr (TCTAIL) { while (tl_t < my_tiles) tail_step(true); }
            if (trace) { g_fg_trace[4] = (unsigned long long)w_fb; g_fg_trace[5] = (unsigned long long)w_op; g_fg_trace[6] = (unsigned long long)w_acc; g_fg_trace[7] = (unsigned long long)w_is; }
        }
    } else if (warp < 2 + F8_GW) {
        // ---------------- gather + split warps: thread = sample row of the tile (= its TMEM lane), one field per k-block
        const int gw = warp - 2;                         // 0..7
        const int q = warp & 3;                          // TMEM lane quarter this warp may touch (= warp id mod 4)
        const int half = gw >> 2;                        // field 2*kb + half, columns half*16 .. +15 of the k-block
        const int wrow0 = q * 32;                        // first tile row of this warp
        const int r = wrow0 + lane;
        const int K_emb_cols = p.F * 16;
        const int sw = (lane >> 1) & 3;                  // SWIZZLE_64B: 16-byte chunk c of row `lane` lives at chunk c ^ sw
        uint8_t* my_stage = a_base + gw * (32 * 64);     // + slot * F8_STAGE_BYTES
        long long* my_ids = id_base + gw * 32;           // + slot * F8_GW * 32
        // issue(): request the table row pieces (or dense columns) of the next un-requested k-block `gi` into stage gi % LA,
        // and this lane's id of k-block gi + LA - 1 into the id FIFO; one cp.async group per call (empty past the end).
        // Rows are requested cooperatively: 4 consecutive lanes fetch the four 16-byte pieces of one 64-byte row.
        uint32_t gi = 0; int i_kb = 0, i_t = 0;          // next k-block to request: global index, k-block, local tile
        uint32_t gj = (uint32_t)(LA - 1); int j_kb = (LA - 1) % nkb, j_t = (LA - 1) / nkb;     // next ids to request
        auto issue = [&]() {
            if (gi < G) {
                const int mt = tile_m0(i_t), nr = tile_rows(i_t);
                const int slot = (int)(gi % LA);
                uint8_t* stg = my_stage + slot * F8_STAGE_BYTES;
                if (i_kb < p.nkb_emb) {
                    const int piece = lane & 3, f = 2 * i_kb + half;
                    const long long* ids = my_ids + slot * (F8_GW * 32);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = i * 8 + (lane >> 2);              // row inside the warp's 32
                        const bool ok = wrow0 + rr < nr;
                        long long id = ids[rr];
                        if (!ok) id = 0;
                        else if ((unsigned long long)id >= (unsigned long long)p.rows[f]) { if (piece == 0) fg_bad_index(p.err, f, mt + wrow0 + rr, id); id = 0; }
                        const float* src;
                        if constexpr (SHARDED) {
                            const unsigned long long iu = (unsigned long long)id, gg = (unsigned long long)p.G;
                            const float* base = reinterpret_cast<const float*>(__ldg(reinterpret_cast<const unsigned long long*>(p.shard_tab) + (size_t)f * gg + (size_t)(iu % gg)));
                            src = base + (size_t)(iu / gg) * 16 + piece * 4;
                        } else {
                            src = p.tables[f] + (size_t)id * 16 + piece * 4;
                        }
                        fg_cp16(reinterpret_cast<float*>(stg + rr * 64 + ((piece ^ ((rr >> 1) & 3)) << 4)), src, ok);
                    }
                } else {
                    const int m = mt + r;
                    const bool ok = r < nr;
                    const int c0 = i_kb * TC_BLOCK_K - K_emb_cols + half * 16;    // first dense column of this warp's half
#pragma unroll 4
                    for (int j = 0; j < 16; ++j) {
                        const int c = c0 + j;
                        float* dst = reinterpret_cast<float*>(stg + lane * 64 + (((j >> 2) ^ sw) << 4)) + (j & 3);
                        if (c < p.Nd) fg_cp4(dst, p.dense[c] + (ok ? m : 0), ok);
                        else *dst = 0.f;
                    }
                }
            }
            if (gj < G && j_kb < p.nkb_emb) {
                const int m = tile_m0(j_t) + r;
                const bool ok = r < tile_rows(j_t);
                fg_cp8(my_ids + (int)(gj % LA) * (F8_GW * 32) + lane, p.idx[2 * j_kb + half] + (ok ? m : 0), ok);
            }
            fg_commit();
            ++gi; if (++i_kb == nkb) { i_kb = 0; ++i_t; }
            ++gj; if (++j_kb == nkb) { j_kb = 0; ++j_t; }
        };
        // prologue: ids of the first LA-1 k-blocks with plain loads, then LA-1 groups in flight
        {
            int kb = 0, tl = 0;
            for (uint32_t qq = 0; qq + 1 < (uint32_t)LA && qq < G; ++qq) {
                const int m = tile_m0(tl) + r;
                const bool ok = kb < p.nkb_emb && r < tile_rows(tl);
                my_ids[(int)(qq % LA) * (F8_GW * 32) + lane] = ok ? __ldg(p.idx[2 * kb + half] + m) : 0;
                if (++kb == nkb) { kb = 0; ++tl; }
            }
        }
        __syncwarp();
        for (int qq = 0; qq + 1 < LA; ++qq) issue();

        if (trace && gw == 0 && lane == 0) g_fg_cta[800] = (unsigned long long)(clock64() - t_start);     // prologue done
        long long w_te = 0;
        float fs[16];                                   // sum over this warp's fields of e, and of e^2
        float fq = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) fs[j] = 0.f;
        uint32_t g = 0;
        long long w_cp = 0, w_eo = 0, w_wk = 0, w_p0 = 0, w_p1 = 0, w_p2 = 0, w_p3 = 0;
        const bool store_x = p.x != nullptr;
        for (int t = 0; t < my_tiles; ++t) {
            const int mt = tile_m0(t), nr = tile_rows(t);
            const int m = mt + r;
            for (int kb = 0; kb < nkb; ++kb, ++g) {
                const long long c0 = FG_T();
                fg_wait<LA - 2>();                      // group g has landed: this lane's pieces of stage g and its id of k-block g + LA - 1
                const long long c1 = FG_T();
                const uint8_t* stg = my_stage + (g % LA) * F8_STAGE_BYTES;
                if (store_x) fence_proxy_async();       // the TMA store below reads what cp.async / st.shared wrote
                __syncwarp();                           // ... of every lane of this warp
                if (store_x && lane == 0) {
                    const int col = kb * TC_BLOCK_K + half * 16;
                    if (col < (int)p.ldx && wrow0 < nr) tma_store_2d(&tmX, stg, col, mt + wrow0);      // nr is a multiple of 32 or ends at M (rows >= M / columns >= ldx are clipped by the TMA unit)
                    bulk_commit();
                }
                const long long c1a = FG_T();
                float4 v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = *reinterpret_cast<const float4*>(stg + lane * 64 + ((j ^ sw) << 4));
                if (kb < p.nkb_emb) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 a = v[j];
                        fs[4 * j + 0] += a.x; fs[4 * j + 1] += a.y; fs[4 * j + 2] += a.z; fs[4 * j + 3] += a.w;
                        fq = fmaf(a.x, a.x, fq); fq = fmaf(a.y, a.y, fq); fq = fmaf(a.z, a.z, fq); fq = fmaf(a.w, a.w, fq);
                    }
                }
                uint32_t h[16], l[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 qv = v[j];
                    h[4 * j + 0] = __float_as_uint(qv.x) & 0xFFFFE000u; l[4 * j + 0] = __float_as_uint(qv.x - __uint_as_float(h[4 * j + 0]));
                    h[4 * j + 1] = __float_as_uint(qv.y) & 0xFFFFE000u; l[4 * j + 1] = __float_as_uint(qv.y - __uint_as_float(h[4 * j + 1]));
                    h[4 * j + 2] = __float_as_uint(qv.z) & 0xFFFFE000u; l[4 * j + 2] = __float_as_uint(qv.z - __uint_as_float(h[4 * j + 2]));
                    h[4 * j + 3] = __float_as_uint(qv.w) & 0xFFFFE000u; l[4 * j + 3] = __float_as_uint(qv.w - __uint_as_float(h[4 * j + 3]));
                }
                const int o = g % OPN;
                const long long c2 = FG_T();
                mbar_wait(&empty_op[o], ((g / OPN) & 1u) ^ 1u);
                const long long c3 = FG_T();
                tc_fence_after();
                const uint32_t ta = tmem_base + A_COL + (uint32_t)o * 64u + ((uint32_t)wrow0 << 16) + (uint32_t)half * 16u;
                tmem_st16(ta, h);
                tmem_st16(ta + 32u, l);
                tmem_st_wait();
                tc_fence_before();
                const long long c4 = FG_T();
                if (lane == 0) bulk_wait_read<1>();     // the store of k-block g - 1 has finished reading its stage ...
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready_op[o]);
                const long long c5 = FG_T();
                issue();                                // ... which the requests of k-block g + LA - 1 now refill
                const long long c6 = FG_T();
                w_cp += c1 - c0; w_eo += c3 - c2; w_wk += (c2 - c1) + (c6 - c3);
                w_p0 += c1a - c1; w_p1 += c2 - c1a; w_p2 += c4 - c3; w_p3 += c6 - c5;
            }
            // FM second order of this sample: 0.5 * (sum_d s_d^2 - sum_{f,d} e^2); the two field halves meet in shared memory,
            // the half-0 thread finishes and hands the value to the tail
            const long long te0 = FG_T();
            if (half == 1) {
#pragma unroll
                for (int j = 0; j < 16; ++j) fm_x[r * F8_FMX_LD + j] = fs[j];
                fm_x[r * F8_FMX_LD + 16] = fq;
            }
            asm volatile("bar.sync %0, 64;" :: "r"(2 + q) : "memory");
            if (half == 0) {
                float ss = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) { fs[j] += fm_x[r * F8_FMX_LD + j]; ss = fmaf(fs[j], fs[j], ss); }
                const float fmv = 0.5f * (ss - (fq + fm_x[r * F8_FMX_LD + 16]));
                fm_tile[(t % FG_FM_BUF) * TC_BLOCK_M + r] = fmv;
                if (r < nr) {
                    if (p.fm != nullptr) p.fm[m] = fmv;
                    if (p.fm_s != nullptr) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            stg_f4(p.fm_s + (size_t)m * 16 + 4 * j, make_float4(fs[4 * j], fs[4 * j + 1], fs[4 * j + 2], fs[4 * j + 3]));
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&fm_ready[t % FG_FM_BUF]);
            }
            fq = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) fs[j] = 0.f;
            w_te += FG_T() - te0;
        }
        if (trace && gw == 0 && lane == 0) { g_fg_cta[801] = (unsigned long long)w_te; g_fg_cta[802] = (unsigned long long)(clock64() - t_start); }
        fg_wait<0>();
        if (lane == 0) bulk_wait_all<0>();
        if (trace && gw == 0 && lane == 0) {
            g_fg_trace[1] = (unsigned long long)w_cp; g_fg_trace[2] = (unsigned long long)w_eo; g_fg_trace[3] = (unsigned long long)w_wk;
            g_fg_trace[12] = (unsigned long long)w_p0; g_fg_trace[13] = (unsigned long long)w_p1; g_fg_trace[14] = (unsigned long long)w_p2; g_fg_trace[15] = (unsigned long long)w_p3;
        }
    } else {
        // ---------------- epilogue warps: layer-1 epilogue -> h1 (HBM + shared memory) -> tower tail (tower_tile.cuh)
        const int quarter = warp & 3;
        const int half = (warp - (2 + F8_GW)) >> 2;
        const int row = quarter * 32 + lane;
        const int et = threadIdx.x - (2 + F8_GW) * 32;
        auto epi_sync = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
        float loss_acc = 0.f;
        if constexpr (TCTAIL) {
            const float tw_bo = tw.b_out != nullptr ? __ldg(tw.b_out) : 0.f;
            // Round r = 0 .. n_tail of a tile: read the accumulator of layer r (r = 0: layer 1; both stacked halves), add the
            // bias, ReLU, store the activation row piece for backward, and either hand it back to the tensor core as the
            // next layer's operand (hi | lo in tensor memory) or, after the last layer, fold it into the output row-dot.
            // Thread = (row, half): the two warps of a lane quarter own columns {half*16 .. +15} and {32 + half*16 .. +15}.
            const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
            const int L = tw.n_tail;
            uint32_t n_out = 0;                                   // tail_bars[2] phases consumed
            for (int t = 0; t < my_tiles; ++t) {
                const int m0 = tile_m0(t);
                const bool live = row < tile_rows(t);
                const uint32_t acc = (uint32_t)t & 1u;
                const int m = m0 + row;
                const long long q0 = FG_T();
                mbar_wait(&tmem_full[acc], ((uint32_t)t >> 1) & 1u);
                tc_fence_after();
                const long long q1 = FG_T();
                const uint32_t d_acc = tmem_base + acc * ACC_STRIDE + lane_addr;
                float headp = 0.f;
                for (int r = 0; r <= L; ++r) {
                    if (r > 0) { mbar_wait(&tail_bars[2], n_out & 1u); ++n_out; tc_fence_after(); }
                    const float* bias = r == 0 ? p.bias1 : tw.b[r - 1];
                    float* hout = r == 0 ? p.h1 : tw.h[r - 1];
                    const long long ldh = r == 0 ? p.ldh1 : (long long)TW_H;
                    for (int c0 = half * 16; c0 < FG_N; c0 += 32) {
                        uint32_t a0[16], a1[16];
                        tmem_ld16(d_acc + (uint32_t)c0, a0);
                        tmem_ld16(d_acc + (uint32_t)(FG_N + c0), a1);
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            v[j] = fmaxf(__uint_as_float(a1[j]) + __uint_as_float(a0[j]) + __ldg(bias + c0 + j), 0.f);
                        if (live) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                stg_f4(hout + (size_t)m * ldh + c0 + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                        }
                        if (r < L) {
                            uint32_t hi[16], lo[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                hi[j] = __float_as_uint(v[j]) & 0xFFFFE000u;
                                lo[j] = __float_as_uint(v[j] - __uint_as_float(hi[j]));
                            }
                            tmem_st16(tmem_base + TAIL_A + lane_addr + (uint32_t)c0, hi);
                            tmem_st16(tmem_base + TAIL_A + 64u + lane_addr + (uint32_t)c0, lo);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) headp = fmaf(v[j], __ldg(tw.w_out + c0 + j), headp);
                        }
                    }
                    if (r < L) {
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tail_bars[1]);    // 8 arrivals: the operand of tail layer r is complete
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);     // every read of this accumulator buffer is done
                const long long q2 = FG_T();
                mbar_wait(&fm_ready[t % FG_FM_BUF], ((uint32_t)t / FG_FM_BUF) & 1u);
                float* head_part = tw_As;                         // [128] partial row-dots of the half-1 warps
                if (half == 1) head_part[row] = headp;
                epi_sync();
                if (half == 0 && live) {
                    const float z = headp + head_part[row] + tw_bo + fm_tile[(t % FG_FM_BUF) * TC_BLOCK_M + row];
                    tw.logit[m] = z;
                    if (tw.pred != nullptr) {
                        const float qq = 1.f / (1.f + expf(-z));
                        tw.pred[m] = qq;
                        if (tw.label != nullptr) {
                            const float y = __ldg(tw.label + m);
                            const float pe = qq + tw.eps;
                            const float l1 = fmaxf(logf(pe), -100.f);
                            const float l0 = fmaxf(logf(1.f - pe), -100.f);
                            loss_acc += -(y * l1 + (1.f - y) * l0);
                        }
                    }
                }
                epi_sync();                                       // head_part is free for the next tile
                if (trace && et == 0) {
                    g_fg_trace[8] += (unsigned long long)(q1 - q0); g_fg_trace[9] += (unsigned long long)(q2 - q1);
                    g_fg_trace[10] += (unsigned long long)(FG_T() - q2);
                }
            }
        } else {
        tower_load_weights_t<FG_EPI_WARPS * 32>(tw, tw_Bs, et);
        const float4 tw_wo = ldg_f4(tw.w_out + (et & 15) * 4);
        const float tw_bo = tw.b_out != nullptr ? __ldg(tw.b_out) : 0.f;
        for (int t = 0; t < my_tiles; ++t) {
            const int m0 = tile_m0(t);                 // CUDA-core tail: the host passes chunk = 128 (whole tiles only)
            const uint32_t acc = (uint32_t)t & 1u;
            const int m = m0 + row;
            const long long q0 = FG_T();
            mbar_wait(&tmem_full[acc], ((uint32_t)t >> 1) & 1u);
            tc_fence_after();
            const long long q1 = FG_T();
            epi_sync();                                // previous tile's activations fully consumed (weights visible)
            for (int c0 = half * 16; c0 < FG_N; c0 += 32) {
                uint32_t a0[16], a1[16];
                tmem_ld16(tmem_base + acc * ACC_STRIDE + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, a0);
                tmem_ld16(tmem_base + acc * ACC_STRIDE + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(FG_N + c0), a1);
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    v[j] = fmaxf(__uint_as_float(a1[j]) + __uint_as_float(a0[j]) + __ldg(p.bias1 + c0 + j), 0.f);
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 q4 = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    if (m < p.M) stg_f4(p.h1 + (size_t)m * p.ldh1 + c0 + j, q4);
                    *reinterpret_cast<float4*>(tw_As + row * TW_LDA + c0 + j) = q4;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            mbar_wait(&fm_ready[t % FG_FM_BUF], ((uint32_t)t / FG_FM_BUF) & 1u);
            epi_sync();
            const long long q2 = FG_T();
            tower_tail_tile_fwd<FG_EPI_WARPS * 32>(tw, tw_As, tw_Bs, m0, et, tw_wo, tw_bo, loss_acc, epi_sync,
                                                   fm_tile + (t % FG_FM_BUF) * TC_BLOCK_M);
            if (trace && et == 0) {
                g_fg_trace[8] += (unsigned long long)(q1 - q0); g_fg_trace[9] += (unsigned long long)(q2 - q1);
                g_fg_trace[10] += (unsigned long long)(FG_T() - q2);
            }
        }
        }
        if (tw.loss != nullptr) tw_loss[et] = loss_acc;
        if (trace && et == 0) g_fg_cta[803] = (unsigned long long)(clock64() - t_start);          // epilogue of the last tile done
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
    if (trace && threadIdx.x == 0) g_fg_trace[0] = (unsigned long long)(clock64() - t_start);
    if (g_fg_trace_on != 0 && threadIdx.x == 0 && blockIdx.x < 256) g_fg_cta[blockIdx.x * 4 + 2] = fg_globaltimer();
    if (tw.loss != nullptr && warp == 0) {
        // deterministic mean BCE: 256 epilogue partials -> per-CTA partial -> the last CTA adds them in index order
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < FG_EPI_WARPS; ++i) s += tw_loss[lane + 32 * i];
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
}

static size_t f8_smem_bytes(int la, int n_tail, bool tctail) {
    return (size_t)FG_LB * FG_B_BYTES + (tctail ? (size_t)n_tail * FT_TAIL_B_BYTES : 0) + (size_t)la * F8_STAGE_BYTES + (size_t)la * F8_GW * 32 * 8 +
           (2 * FG_LB + 2 * FG_OP + 4 + FG_FM_BUF + 4) * 8 + 16 + FG_FM_BUF * TC_BLOCK_M * 4 + (size_t)(TC_BLOCK_M * F8_FMX_LD) * 4 +
           (tctail ? (size_t)(TC_BLOCK_M + 256) * 4 : (size_t)(TC_BLOCK_M * TW_LDA + n_tail * TW_H * TW_H + 256) * 4) + 1024;
}


// =====================================================================================================================
// v3 of the one-kernel forward (default): DEDICATED FETCH WARPS.
//
// What the 8-gather-warp kernel above taught (profiles/r02_rowfetch.md, tools/exp/exp_rowfetch.cu): 1.7 M random 64-byte row
// reads cost 49.5 us on this HBM however they are requested (LDGSTS, LDG, TMA gather4, cp.async.bulk: all >= 49 us) — a warp
// that requests rows is stalled by the memory system's back-pressure for ~1 k cycles per k-block — and, decisive, work that
// the SAME warp does between its requests is NOT hidden: 2.2 k cycles of ALU work per round added 2.2 k cycles per round,
// whatever the ring depth.  A warp that only requests rows keeps the queue full.  Hence:
//   warps 2-5  FETCH: warp q requests both fields of a k-block for the 32 rows of lane quarter q (8 LDGSTS per lane and
//              round, 4 lanes per 64-byte row, ids prefetched one round ahead in registers) into stage [q][32 rows][128 B]
//              (SWIZZLE_128B pattern) and signals `full_a[slot][q]` through cp.async.mbarrier.arrive.noinc;
//   warps 6-9  SPLIT: thread = row; waits for its quarter, sends the 32-column piece of the feature row x to HBM with one TMA
//              store per warp and round, accumulates the FM sums (both fields: no cross-warp exchange), splits fp32 into
//              (hi, lo) and writes the TS-mode operand into tensor memory; frees the stage one round later (after the TMA
//              store has read it);
//   warp 0 weight TMA producer, warp 1 MMA issuer, 8 epilogue warps with the tower tail on tcgen05 — as above;
//   warp 2  X-STORE: one TMA store per lane quarter and k-block, so no split warp pays for the proxy fence.
// Warps, in warpgroups of 4: [0 weights, 1 MMA, 2 x-store, 3 idle | NF fetch | 4 split | 8 epilogue].  NF = 16 (1024 threads,
// opt-in, measured slower than NF = 8: 103 vs 96 us): the kernel starts with 64 registers per thread and hands them out again
// with setmaxnreg — 40 for the first 20 warps, 104 for the split and epilogue warps.
constexpr int fs_threads(int nf) { return (4 + nf + 4 + FG_EPI_WARPS) * 32; }
constexpr int FS_STAGE_BYTES = TC_BLOCK_M * TC_BLOCK_K * 4;       // 16 KiB: [4 quarters][32 rows][128 B]
constexpr int FS_IDD = 4;                                         // rounds by which the id copies run ahead of the row requests

__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// NF fetch warps (4 or 8): a warp that requests rows is blocked ~250-375 cycles per LDGSTS (8 scattered 64-byte rows each),
// so the request rate of an SM grows with the number of warps that request: 16 warps per SM fetch the 1.7 M rows in 35 us,
// 8 warps in 49.5 us (tools/exp/exp_rowfetch.cu).  With NF = 8 warp fw requests field fw / 4 of lane quarter fw % 4.
template <int LA, bool SHARDED, int NF>
__global__ void __launch_bounds__(fs_threads(NF), 1)
deepfm_fwd_fs_kernel(const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                     const __grid_constant__ CUtensorMap tmThi, const __grid_constant__ CUtensorMap tmTlo,
                     const __grid_constant__ CUtensorMap tmX,
                     const __grid_constant__ FusedFwdParams p, const __grid_constant__ TowerFwdParams tw, int full_rounds, int chunk) {
    constexpr int OPN = FT_OP;                                                // depth of the tensor-memory operand ring
    constexpr int FS_FETCH_WARP0 = 4;                                         // warpgroup 0 = [weights, MMA, x-store, idle]
    constexpr int FS_SPLIT_WARP0 = FS_FETCH_WARP0 + NF, FS_EPI_WARP0 = FS_SPLIT_WARP0 + 4;     // warp & 3 is the TMEM lane quarter
    constexpr int FS_STORE_WARP = 2;                                          // one warp sends the feature row x to HBM (TMA stores)
    constexpr int NI = 32 / NF;                                               // LDGSTS per lane and round of a fetch warp
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* b_base = smem;                                                   // FG_LB x 16 KiB, 1 KiB aligned (SWIZZLE_128B)
    uint8_t* t_base = b_base + FG_LB * FG_B_BYTES;                            // n_tail x 32 KiB resident tail operands
    uint8_t* a_base = t_base + tw.n_tail * FT_TAIL_B_BYTES;                   // LA x 16 KiB row stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_base + LA * FS_STAGE_BYTES);
    uint64_t* full_b = bars;                       // [FG_LB]  weight k-block landed
    uint64_t* empty_b = full_b + FG_LB;            // [FG_LB]  MMAs that read it are done
    uint64_t* ready_op = empty_b + FG_LB;          // [FG_OP]  A hi/lo of a k-block are in tensor memory (one arrival per split warp)
    uint64_t* empty_op = ready_op + FG_OP;         // [FG_OP]
    uint64_t* tmem_full = empty_op + FG_OP;        // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint64_t* fm_ready = tmem_empty + 2;           // [FG_FM_BUF]  FM values of a tile written (4 arrivals: the split warps)
    uint64_t* tail_bars = fm_ready + FG_FM_BUF;    // [4] [0] tail weights landed, [1] tail operand written, [2] tail MMAs done
    uint64_t* full_a = tail_bars + 4;              // [LA][4]  rows of quarter q of stage s have landed (32 noinc arrivals)
    uint64_t* empty_a = full_a + LA * 4;           // [LA][4]  split warp q is done with stage s (1 arrival)
    long long* id_fifo = reinterpret_cast<long long*>(empty_a + LA * 4);      // [FS_IDD + 1][fetch warp][fields it serves][32] ids
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(id_fifo + (FS_IDD + 1) * (NF == 16 ? 16 : 8) * 32);
    float* fm_tile = reinterpret_cast<float*>(tmem_ptr + 4);                  // [FG_FM_BUF][128]
    float* tw_As = fm_tile + FG_FM_BUF * TC_BLOCK_M;                          // 128 head partials
    float* tw_loss = tw_As + TC_BLOCK_M;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.nkb;
    // tile schedule: see deepfm_fwd_fused8_kernel
    const int part_m0 = full_rounds * (int)gridDim.x * TC_BLOCK_M + (int)blockIdx.x * chunk;
    const int my_tiles = full_rounds + ((chunk > 0 && part_m0 < p.M) ? 1 : 0);
    auto tile_m0 = [&](int t) -> int { return t < full_rounds ? ((int)blockIdx.x + t * (int)gridDim.x) * TC_BLOCK_M : part_m0; };
    auto tile_rows = [&](int t) -> int { return min(t < full_rounds ? TC_BLOCK_M : chunk, p.M - tile_m0(t)); };
    const uint32_t G = (uint32_t)my_tiles * (uint32_t)nkb;                    // k-blocks this CTA walks
    constexpr uint32_t ACC_STRIDE = 2 * FG_N;                                 // stacked accumulator: [a.b_hi | a.b_lo]
    constexpr uint32_t A_COL = 2 * ACC_STRIDE;                                // first TMEM column of the operand ring
    constexpr uint32_t TAIL_A = A_COL + FT_OP * 64u;                          // tail operand, hi [0,64) | lo [64,128)

    if (threadIdx.x == 0) {
        for (int s = 0; s < FG_LB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < FG_OP; ++s) { mbar_init(&ready_op[s], 4); mbar_init(&empty_op[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], FG_EPI_WARPS); }
        for (int s = 0; s < FG_FM_BUF; ++s) mbar_init(&fm_ready[s], 4);
        mbar_init(&tail_bars[0], 1); mbar_init(&tail_bars[1], FG_EPI_WARPS); mbar_init(&tail_bars[2], 1); mbar_init(&tail_bars[3], 1);
        for (int s = 0; s < LA * 4; ++s) { mbar_init(&full_a[s], NF * 8); mbar_init(&empty_a[s], p.x != nullptr ? 2 : 1); }
        tmem_ptr[1] = 0u;                          // tile the fetch warps are on (read by the L2 prefetch warp)
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const bool trace = g_fg_trace_on != 0 && blockIdx.x == 0;
    const long long t_start = FG_T();
    if (g_fg_trace_on != 0 && threadIdx.x == 0 && blockIdx.x < 256) {
        g_fg_cta[blockIdx.x * 4 + 0] = fg_smid(); g_fg_cta[blockIdx.x * 4 + 1] = fg_globaltimer(); g_fg_cta[blockIdx.x * 4 + 3] = (unsigned long long)my_tiles;
    }

    // Role dispatch by WARPGROUP (4 warps): setmaxnreg must be executed by all four warps of a warpgroup at the same instruction.
    if (warp < FS_FETCH_WARP0) {
        if constexpr (NF == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
        // ---------------- weight producer: [B hi ; B lo] of every k-block through a FG_LB-deep ring
        if (lane == 0) {
            if constexpr (true) {
                // resident tail operands: layer l, k-block kb -> [W_l hi ; W_l lo] columns kb*32 .. kb*32+31 (128 rows x 128 B)
                mbar_arrive_expect_tx(&tail_bars[0], (uint32_t)(tw.n_tail * FT_TAIL_B_BYTES));
                for (int l = 0; l < tw.n_tail; ++l)
                    for (int kb = 0; kb < 2; ++kb) {
                        uint8_t* st = t_base + (size_t)(l * 2 + kb) * FG_B_BYTES;
                        tma_load_2d(st, &tmThi, &tail_bars[0], kb * TC_BLOCK_K, l * TW_H);
                        tma_load_2d(st + FG_B_BYTES / 2, &tmTlo, &tail_bars[0], kb * TC_BLOCK_K, l * TW_H);
                    }
            }
            long long w_b = 0;
            for (uint32_t g = 0; g < G; ++g) {
                const int s = g % FG_LB, kb = g % nkb;
                const long long c0 = FG_T();
                mbar_wait(&empty_b[s], ((g / FG_LB) & 1u) ^ 1u);
                w_b += FG_T() - c0;
                if (trace && g + 1 == G) g_fg_trace[11] = (unsigned long long)w_b;
                uint8_t* st = b_base + (size_t)s * FG_B_BYTES;
                mbar_arrive_expect_tx(&full_b[s], (uint32_t)FG_B_BYTES);
                tma_load_2d(st, &tmBhi, &full_b[s], kb * TC_BLOCK_K, 0);
                tma_load_2d(st + FG_B_BYTES / 2, &tmBlo, &full_b[s], kb * TC_BLOCK_K, 0);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: per k-step two TS-mode MMAs (a_lo, a_hi) against the stacked 128-row weight operand
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TC_BLOCK_M, 2 * FG_N);
            uint32_t g = 0;
            long long w_fb = 0, w_op = 0, w_acc = 0, w_is = 0;
            // TCTAIL: tail steps are issued in (tile, layer) order as soon as the epilogue warps have written the operand
            // (tail_bars[1]); between k-blocks of the running tile without blocking, and blocking before an accumulator
            // buffer is re-used (tile t + 2 needs tile t's tail finished) and after the last tile.
            int tl_t = 0, tl_l = 0; uint32_t tl_n = 0;
            auto tail_step = [&](bool block) -> bool {
                if (!block && !mbar_test(&tail_bars[1], tl_n & 1u)) return false;
                mbar_wait(&tail_bars[1], tl_n & 1u);
                if (tl_n == 0) mbar_wait(&tail_bars[0], 0u);               // resident tail operands have landed
                tc_fence_after();
                const uint32_t d_t = tmem_base + ((uint32_t)tl_t & 1u) * ACC_STRIDE;
                const uint32_t tb = smem_u32(t_base + (size_t)tl_l * FT_TAIL_B_BYTES);
#pragma unroll
                for (int kk = 0; kk < TW_H / TC_UMMA_K; ++kk) {
                    const uint64_t db = make_kmajor_sw128_desc(tb + (uint32_t)(kk >> 2) * FG_B_BYTES + (uint32_t)(kk & 3) * TC_UMMA_K * 4);
                    umma_tf32_ts(d_t, tmem_base + TAIL_A + 64u + kk * TC_UMMA_K, db, idesc, kk > 0 ? 1u : 0u);
                    umma_tf32_ts(d_t, tmem_base + TAIL_A + kk * TC_UMMA_K, db, idesc, 1u);
                }
                umma_commit(&tail_bars[2]);
                ++tl_n;
                if (++tl_l == tw.n_tail) { tl_l = 0; ++tl_t; }
                return true;
            };
            for (int t = 0; t < my_tiles; ++t) {
                const uint32_t acc = (uint32_t)t & 1u;
                const long long c0 = FG_T();
                if constexpr (true) { while (tl_t + 2 <= t) tail_step(true); }
                mbar_wait(&tmem_empty[acc], (((uint32_t)t >> 1) & 1u) ^ 1u);
                w_acc += FG_T() - c0;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % FG_LB, o = g % OPN;
                    if constexpr (true) { if (tl_t < t) tail_step(false); }
                    const long long c1 = FG_T();
                    mbar_wait(&full_b[s], (g / FG_LB) & 1u);
                    const long long c2 = FG_T();
                    if constexpr (true) {
                        // do not sit on the operand barrier while a tail layer of the previous tile becomes ready
                        while (!mbar_test(&ready_op[o], (g / OPN) & 1u)) { if (tl_t < t) tail_step(false); }
                    }
                    mbar_wait(&ready_op[o], (g / OPN) & 1u);
                    const long long c3 = FG_T();
                    w_fb += c2 - c1; w_op += c3 - c2;
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(b_base + (size_t)s * FG_B_BYTES);
                    const uint32_t ta_hi = tmem_base + A_COL + (uint32_t)o * 64u, ta_lo = ta_hi + 32u;
#pragma unroll
                    for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
                        const uint64_t db = make_kmajor_sw128_desc(b_addr + k * TC_UMMA_K * 4);
                        umma_tf32_ts(d_tmem, ta_lo + k * TC_UMMA_K, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_tf32_ts(d_tmem, ta_hi + k * TC_UMMA_K, db, idesc, 1u);
                    }
                    umma_commit(&empty_op[o]);
                    umma_commit(&empty_b[s]);
                    w_is += FG_T() - c3;
                }
                umma_commit(&tmem_full[acc]);
            }
            if (trace) g_fg_cta[804] = (unsigned long long)(clock64() - t_start);      // last layer-1 MMA issued
            if constexpr (true) { while (tl_t < my_tiles) tail_step(true); }
            if (trace) { g_fg_trace[4] = (unsigned long long)w_fb; g_fg_trace[5] = (unsigned long long)w_op; g_fg_trace[6] = (unsigned long long)w_acc; g_fg_trace[7] = (unsigned long long)w_is; }
        }
    } else if (warp == FS_STORE_WARP) {
        // ---------------- x-store warp: one TMA store per lane quarter and k-block ([32 rows x 32 columns], SWIZZLE_128B = the
        // stage layout).  fence.proxy.async after the mbarrier acquire orders the fetch warps' cp.async writes before the
        // async-proxy read; the stage is released one round later, when its commit group has finished reading.
        if (lane == 0 && p.x != nullptr) {
            uint32_t g = 0;
            long long w_st = 0;
            for (int t = 0; t < my_tiles; ++t) {
                const int mt = tile_m0(t), nr = tile_rows(t);
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int slot = (int)(g % LA);
                    const long long c0 = FG_T();
                    for (int q = 0; q < 4; ++q) {
                        mbar_wait(&full_a[slot * 4 + q], (g / LA) & 1u);
                        if (q == 3) fence_proxy_async();
                    }
                    w_st += FG_T() - c0;
                    if (kb * TC_BLOCK_K < (int)p.ldx) {
                        for (int q = 0; q < 4; ++q)
                            if (q * 32 < nr) tma_store_2d(&tmX, a_base + slot * FS_STAGE_BYTES + q * 4096, kb * TC_BLOCK_K, mt + q * 32);
                    }
                    bulk_commit();
                    bulk_wait_read<1>();                 // the stores of round g - 1 have finished reading their stage
                    if (g > 0) { for (int q = 0; q < 4; ++q) mbar_arrive(&empty_a[(int)((g - 1) % LA) * 4 + q]); }
                }
            }
            bulk_wait_all<0>();
            if (trace) g_fg_trace[12] = (unsigned long long)w_st;
        }
    } else if (warp == 3) {
        // ---------------- L2 prefetch warp (the fourth warp of this warpgroup was idle): asks L2 for the table rows of the tile(s)
        // AHEAD of the one the fetch warps are on, so that their cp.async requests — whose acceptance rate is what bounds the
        // kernel — are answered by L2.  prefetch.global.L2 returns nothing to the SM.  MEASURED SLOWER (132.7 vs 98.9 us, the same
        // for 1, 2 and 4 tiles ahead: profiles/r02_rowfetch.md): the prefetches travel the same SM request path as the row
        // requests, and that path — not DRAM latency — is what the fetch warps wait for.  Off by default (fused_l2_prefetch = 0).
        if constexpr (!SHARDED) {
            if (p.l2_prefetch > 0) {
                volatile uint32_t* prog = reinterpret_cast<volatile uint32_t*>(tmem_ptr + 1);
                for (int t = 0; t < my_tiles; ++t) {
                    while ((int)*prog + p.l2_prefetch < t) __nanosleep(256);
                    const int m0 = tile_m0(t), nr = tile_rows(t);
                    for (int r0 = 0; r0 < nr; r0 += 32) {
                        const int r = r0 + lane;
                        const bool ok = r < nr;
                        for (int f0 = 0; f0 < p.F; f0 += 13) {
                            long long id[13];
#pragma unroll
                            for (int j = 0; j < 13; ++j) id[j] = (ok && f0 + j < p.F) ? __ldg(p.idx[f0 + j] + m0 + r) : -1;
#pragma unroll
                            for (int j = 0; j < 13; ++j) {
                                const int f = f0 + j;
                                if (f < p.F && id[j] >= 0 && (unsigned long long)id[j] < (unsigned long long)p.rows[f]) {
                                    const float* src = p.tables[f] + (size_t)id[j] * 16;
                                    asm volatile("prefetch.global.L2 [%0];" :: "l"(src));
                                    asm volatile("prefetch.global.L2 [%0];" :: "l"(src + 8));
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    } else if (warp >= FS_FETCH_WARP0 && warp < FS_SPLIT_WARP0) {
        if constexpr (NF == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        // ---------------- fetch warps: warp (q, part) requests NI row groups of lane quarter q for every k-block.  The ids
        // it needs travel through a small FIFO in shared memory: lane = row of the quarter copies ITS id of the field it
        // serves FS_IDD rounds ahead (cp.async, one group per round), so no request ever waits for an id load.
        const int fw = warp - FS_FETCH_WARP0;
        const int q = fw & 3, part = fw >> 2;              // request j = part * NI + i of a round: field j / 4, row group j % 4
        const int piece = lane & 3;
        const int K_emb_cols = p.F * 16;
        constexpr int NFLD = NF >= 8 ? 1 : 2;              // fields of a k-block this warp serves (NF >= 8: one, field `my_f`)
        const int my_f = NF == 16 ? part >> 1 : part;
        long long* my_ids = id_fifo + fw * (NFLD * 32);    // + slot * (NF * NFLD * 32): [field][row of the quarter]
        constexpr int IDS = FS_IDD + 1;
        auto request_ids = [&](uint32_t gg, int tt, int kk) {          // ids of round gg (tile tt, k-block kk) -> FIFO slot gg % IDS
            if (gg < G && kk < p.nkb_emb) {
                const int m = tile_m0(tt) + q * 32 + lane;
                const bool ok = q * 32 + lane < tile_rows(tt);
#pragma unroll
                for (int fl = 0; fl < NFLD; ++fl)
                    fg_cp8(my_ids + (int)(gg % IDS) * (NF * NFLD * 32) + fl * 32 + lane, p.idx[2 * kk + (NF >= 8 ? my_f : fl)] + (ok ? m : 0), ok);
            }
        };
        long long w_em = 0, w_is = 0;
        int t = 0, kb = 0;                                  // round g
        int jt = 0, jkb = 0; uint32_t gj = 0;               // round g + FS_IDD (ids to request)
        // prologue: ids of the first FS_IDD rounds with plain loads
        for (; gj < (uint32_t)FS_IDD && gj < G; ++gj) {
            if (jkb < p.nkb_emb) {
                const int m = tile_m0(jt) + q * 32 + lane;
                const bool ok = q * 32 + lane < tile_rows(jt);
#pragma unroll
                for (int fl = 0; fl < NFLD; ++fl)
                    my_ids[(int)(gj % IDS) * (NF * NFLD * 32) + fl * 32 + lane] = ok ? __ldg(p.idx[2 * jkb + (NF >= 8 ? my_f : fl)] + m) : 0;
            }
            if (++jkb == nkb) { jkb = 0; ++jt; }
        }
        gj = (uint32_t)FS_IDD;
        __syncwarp();
        for (uint32_t g = 0; g < G; ++g) {
            const int slot = (int)(g % LA);
            if (kb == 0 && fw == 0 && lane == 0) *reinterpret_cast<volatile uint32_t*>(tmem_ptr + 1) = (uint32_t)t;
            fg_wait<FS_IDD - 1>();                          // the id copies of round g (requested FS_IDD rounds ago) have landed
            __syncwarp();
            const long long c0 = FG_T();
            mbar_wait(&empty_a[slot * 4 + q], ((g / LA) & 1u) ^ 1u);
            const long long c1 = FG_T();
            uint8_t* stg = a_base + slot * FS_STAGE_BYTES + q * 4096;
            if (kb < p.nkb_emb) {
                const int nr = tile_rows(t) - q * 32;
                const long long* ids = my_ids + (int)(g % IDS) * (NF * NFLD * 32);
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const int j = part * NI + i, rr = (j & 3) * 8 + (lane >> 2), fsel = j >> 2, f = 2 * kb + fsel;
                    const bool ok = rr < nr;
                    long long id = ids[(NF >= 8 ? 0 : fsel) * 32 + rr];
                    if (!ok) id = 0;
                    else if ((unsigned long long)id >= (unsigned long long)p.rows[f]) { if (piece == 0) fg_bad_index(p.err, f, tile_m0(t) + q * 32 + rr, id); id = 0; }
                    const float* src;
                    if constexpr (SHARDED) {
                        const unsigned long long iu = (unsigned long long)id, gg = (unsigned long long)p.G;
                        const float* base = reinterpret_cast<const float*>(__ldg(reinterpret_cast<const unsigned long long*>(p.shard_tab) + (size_t)f * gg + (size_t)(iu % gg)));
                        src = base + (size_t)(iu / gg) * 16 + piece * 4;
                    } else {
                        src = p.tables[f] + (size_t)id * 16 + piece * 4;
                    }
                    fg_cp16(reinterpret_cast<float*>(stg + rr * 128 + (((fsel * 4 + piece) ^ (rr & 7)) << 4)), src, ok);
                }
            } else if (part == 0) {
                // dense columns (and the zero padding of the last k-block): thread = row, zero-fill through the same cp.async path
                const int m = tile_m0(t) + q * 32 + lane;
                const bool ok = q * 32 + lane < tile_rows(t);
                const int c0d = kb * TC_BLOCK_K - K_emb_cols;                 // first dense column of this k-block
#pragma unroll 4
                for (int j = 0; j < TC_BLOCK_K; ++j) {
                    const int c = c0d + j;
                    float* dst = reinterpret_cast<float*>(stg + lane * 128 + (((j >> 2) ^ (lane & 7)) << 4)) + (j & 3);
                    const bool has = ok && c < p.Nd;
                    fg_cp4(dst, has ? p.dense[c] + m : p.dense[0], has);
                }
            }
            cp_async_arrive_noinc(&full_a[slot * 4 + q]);   // fires when this lane's copies of rounds <= g have landed
            request_ids(gj, jt, jkb);
            fg_commit();
            if (++kb == nkb) { kb = 0; ++t; }
            ++gj; if (++jkb == nkb) { jkb = 0; ++jt; }
            w_em += c1 - c0; w_is += FG_T() - c1;
        }
        fg_wait<0>();
        if (trace && fw == 0 && lane == 0) { g_fg_trace[1] = (unsigned long long)w_em; g_fg_trace[15] = (unsigned long long)w_is; }
    } else if (warp >= FS_SPLIT_WARP0 && warp < FS_EPI_WARP0) {
        if constexpr (NF == 16) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ---------------- split warps: thread = sample row of the tile (= its TMEM lane), both fields of every k-block
        const int q = warp & 3;                          // TMEM lane quarter this warp may touch (= warp id mod 4)
        const int wrow0 = q * 32;
        const int r = wrow0 + lane;
        const int sw = lane & 7;                         // SWIZZLE_128B: 16-byte chunk c of row `lane` lives at chunk c ^ sw
        if (trace && q == 2 && lane == 0) g_fg_cta[800] = (unsigned long long)(clock64() - t_start);
        float fs[16];                                   // sum_f e of this sample, and the sum of squares
        float fq = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) fs[j] = 0.f;
        uint32_t g = 0;
        long long w_cp = 0, w_eo = 0, w_p0 = 0, w_p1 = 0, w_p2 = 0, w_te = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int mt = tile_m0(t), nr = tile_rows(t);
            const int m = mt + r;
            for (int kb = 0; kb < nkb; ++kb, ++g) {
                const int slot = (int)(g % LA);
                const long long c0 = FG_T();
                mbar_wait(&full_a[slot * 4 + q], (g / LA) & 1u);              // the 32 rows of this quarter have landed
                const long long c1 = FG_T();
                const uint8_t* stg = a_base + slot * FS_STAGE_BYTES + q * 4096;
                const long long c1a = FG_T();
                const int o = g % OPN;
                const uint32_t ta = tmem_base + A_COL + (uint32_t)o * 64u + ((uint32_t)wrow0 << 16);
                long long wait_slot = 0;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {         // field hf of the k-block: columns hf*16 .. +15
                    float4 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = *reinterpret_cast<const float4*>(stg + lane * 128 + (((hf * 4 + j) ^ sw) << 4));
                    if (kb < p.nkb_emb) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 a = v[j];
                            fs[4 * j + 0] += a.x; fs[4 * j + 1] += a.y; fs[4 * j + 2] += a.z; fs[4 * j + 3] += a.w;
                            fq = fmaf(a.x, a.x, fq); fq = fmaf(a.y, a.y, fq); fq = fmaf(a.z, a.z, fq); fq = fmaf(a.w, a.w, fq);
                        }
                    }
                    uint32_t h[16], l[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 qv = v[j];
                        h[4 * j + 0] = __float_as_uint(qv.x) & 0xFFFFE000u; l[4 * j + 0] = __float_as_uint(qv.x - __uint_as_float(h[4 * j + 0]));
                        h[4 * j + 1] = __float_as_uint(qv.y) & 0xFFFFE000u; l[4 * j + 1] = __float_as_uint(qv.y - __uint_as_float(h[4 * j + 1]));
                        h[4 * j + 2] = __float_as_uint(qv.z) & 0xFFFFE000u; l[4 * j + 2] = __float_as_uint(qv.z - __uint_as_float(h[4 * j + 2]));
                        h[4 * j + 3] = __float_as_uint(qv.w) & 0xFFFFE000u; l[4 * j + 3] = __float_as_uint(qv.w - __uint_as_float(h[4 * j + 3]));
                    }
                    if (hf == 0) {
                        const long long c2 = FG_T();
                        mbar_wait(&empty_op[o], ((g / OPN) & 1u) ^ 1u);
                        wait_slot = FG_T() - c2;
                        tc_fence_after();
                    }
                    tmem_st16(ta + (uint32_t)hf * 16u, h);
                    tmem_st16(ta + 32u + (uint32_t)hf * 16u, l);
                }
                __syncwarp();                            // every lane has read its row of the stage
                if (lane == 0) mbar_arrive(&empty_a[slot * 4 + q]);
                const long long c3 = FG_T();
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready_op[o]);
                const long long c4 = FG_T();
                w_cp += c1 - c0; w_eo += wait_slot; w_p0 += c1a - c1; w_p1 += (c3 - c1a) - wait_slot; w_p2 += c4 - c3;
            }
            // FM second order of this sample: 0.5 * (sum_d s_d^2 - sum_{f,d} e^2), handed to the tail through shared memory
            const long long te0 = FG_T();
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) ss = fmaf(fs[j], fs[j], ss);
            const float fmv = 0.5f * (ss - fq);
            fm_tile[(t % FG_FM_BUF) * TC_BLOCK_M + r] = fmv;
            if (r < nr) {
                if (p.fm != nullptr) p.fm[m] = fmv;
                if (p.fm_s != nullptr) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        stg_f4(p.fm_s + (size_t)m * 16 + 4 * j, make_float4(fs[4 * j], fs[4 * j + 1], fs[4 * j + 2], fs[4 * j + 3]));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&fm_ready[t % FG_FM_BUF]);
            fq = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) fs[j] = 0.f;
            w_te += FG_T() - te0;
        }
        if (trace && q == 2 && lane == 0) {
            g_fg_trace[2] = (unsigned long long)w_eo; g_fg_trace[3] = (unsigned long long)(w_p0 + w_p1 + w_p2);
            g_fg_trace[13] = (unsigned long long)w_p1; g_fg_trace[14] = (unsigned long long)w_p2;
            g_fg_cta[801] = (unsigned long long)w_te; g_fg_cta[802] = (unsigned long long)(clock64() - t_start); g_fg_cta[805] = (unsigned long long)w_cp;
        }
    } else if (warp >= FS_EPI_WARP0 && warp < FS_EPI_WARP0 + FG_EPI_WARPS) {
        if constexpr (NF == 16) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ---------------- epilogue warps: layer-1 epilogue -> h1 (HBM + shared memory) -> tower tail (tower_tile.cuh)
        const int quarter = warp & 3;
        const int half = (warp - FS_EPI_WARP0) >> 2;
        const int row = quarter * 32 + lane;
        const int et = threadIdx.x - FS_EPI_WARP0 * 32;
        auto epi_sync = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
        float loss_acc = 0.f;
        {
            const float tw_bo = tw.b_out != nullptr ? __ldg(tw.b_out) : 0.f;
            // Round r = 0 .. n_tail of a tile: read the accumulator of layer r (r = 0: layer 1; both stacked halves), add the
            // bias, ReLU, store the activation row piece for backward, and either hand it back to the tensor core as the
            // next layer's operand (hi | lo in tensor memory) or, after the last layer, fold it into the output row-dot.
            // Thread = (row, half): the two warps of a lane quarter own columns {half*16 .. +15} and {32 + half*16 .. +15}.
            const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
            const int L = tw.n_tail;
            uint32_t n_out = 0;                                   // tail_bars[2] phases consumed
            for (int t = 0; t < my_tiles; ++t) {
                const int m0 = tile_m0(t);
                const bool live = row < tile_rows(t);
                const uint32_t acc = (uint32_t)t & 1u;
                const int m = m0 + row;
                const long long q0 = FG_T();
                mbar_wait(&tmem_full[acc], ((uint32_t)t >> 1) & 1u);
                tc_fence_after();
                const long long q1 = FG_T();
                const uint32_t d_acc = tmem_base + acc * ACC_STRIDE + lane_addr;
                float headp = 0.f;
                for (int r = 0; r <= L; ++r) {
                    if (r > 0) { mbar_wait(&tail_bars[2], n_out & 1u); ++n_out; tc_fence_after(); }
                    const float* bias = r == 0 ? p.bias1 : tw.b[r - 1];
                    float* hout = r == 0 ? p.h1 : tw.h[r - 1];
                    const long long ldh = r == 0 ? p.ldh1 : (long long)TW_H;
                    for (int c0 = half * 16; c0 < FG_N; c0 += 32) {
                        uint32_t a0[16], a1[16];
                        tmem_ld16(d_acc + (uint32_t)c0, a0);
                        tmem_ld16(d_acc + (uint32_t)(FG_N + c0), a1);
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            v[j] = fmaxf(__uint_as_float(a1[j]) + __uint_as_float(a0[j]) + __ldg(bias + c0 + j), 0.f);
                        if (live) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                stg_f4(hout + (size_t)m * ldh + c0 + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                        }
                        if (r < L) {
                            uint32_t hi[16], lo[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                hi[j] = __float_as_uint(v[j]) & 0xFFFFE000u;
                                lo[j] = __float_as_uint(v[j] - __uint_as_float(hi[j]));
                            }
                            tmem_st16(tmem_base + TAIL_A + lane_addr + (uint32_t)c0, hi);
                            tmem_st16(tmem_base + TAIL_A + 64u + lane_addr + (uint32_t)c0, lo);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) headp = fmaf(v[j], __ldg(tw.w_out + c0 + j), headp);
                        }
                    }
                    if (r < L) {
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tail_bars[1]);    // 8 arrivals: the operand of tail layer r is complete
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);     // every read of this accumulator buffer is done
                const long long q2 = FG_T();
                mbar_wait(&fm_ready[t % FG_FM_BUF], ((uint32_t)t / FG_FM_BUF) & 1u);
                float* head_part = tw_As;                         // [128] partial row-dots of the half-1 warps
                if (half == 1) head_part[row] = headp;
                epi_sync();
                if (half == 0 && live) {
                    const float z = headp + head_part[row] + tw_bo + fm_tile[(t % FG_FM_BUF) * TC_BLOCK_M + row];
                    tw.logit[m] = z;
                    if (tw.pred != nullptr) {
                        const float qq = 1.f / (1.f + expf(-z));
                        tw.pred[m] = qq;
                        if (tw.label != nullptr) {
                            const float y = __ldg(tw.label + m);
                            const float pe = qq + tw.eps;
                            const float l1 = fmaxf(logf(pe), -100.f);
                            const float l0 = fmaxf(logf(1.f - pe), -100.f);
                            loss_acc += -(y * l1 + (1.f - y) * l0);
                        }
                    }
                }
                epi_sync();                                       // head_part is free for the next tile
                if (trace && et == 0) {
                    g_fg_trace[8] += (unsigned long long)(q1 - q0); g_fg_trace[9] += (unsigned long long)(q2 - q1);
                    g_fg_trace[10] += (unsigned long long)(FG_T() - q2);
                }
            }
        }
        if (tw.loss != nullptr) tw_loss[et] = loss_acc;
        if (trace && et == 0) g_fg_cta[803] = (unsigned long long)(clock64() - t_start);          // epilogue of the last tile done
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
    if (trace && threadIdx.x == 0) g_fg_trace[0] = (unsigned long long)(clock64() - t_start);
    if (g_fg_trace_on != 0 && threadIdx.x == 0 && blockIdx.x < 256) g_fg_cta[blockIdx.x * 4 + 2] = fg_globaltimer();
    if (tw.loss != nullptr && warp == 0) {
        // deterministic mean BCE: 256 epilogue partials -> per-CTA partial -> the last CTA adds them in index order
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < FG_EPI_WARPS; ++i) s += tw_loss[lane + 32 * i];
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
}


static size_t fs_smem_bytes(int la, int n_tail, int nf = 8) {
    return (size_t)FG_LB * FG_B_BYTES + (size_t)n_tail * FT_TAIL_B_BYTES + (size_t)la * FS_STAGE_BYTES +
           (2 * FG_LB + 2 * FG_OP + 4 + FG_FM_BUF + 4 + 8 * la) * 8 + (size_t)(FS_IDD + 1) * (nf == 16 ? 16 : 8) * 32 * 8 + 16 + FG_FM_BUF * TC_BLOCK_M * 4 +
           (size_t)(TC_BLOCK_M + 256) * 4 + 1024;
}

static size_t fg_smem_bytes(int la, int n_tail, bool tctail = false) {
    return (size_t)FG_LB * FG_B_BYTES + (tctail ? (size_t)n_tail * FT_TAIL_B_BYTES : 0) +
           (size_t)la * TC_BLOCK_M * FG_ROW * 4 + (size_t)la * TC_BLOCK_M * 16 +
           (2 * FG_LB + 2 * FG_OP + 4 + FG_FM_BUF + 4) * 8 + 16 + FG_FM_BUF * TC_BLOCK_M * 4 +
           (tctail ? (size_t)(TC_BLOCK_M + 256) * 4 : (size_t)(TC_BLOCK_M * TW_LDA + n_tail * TW_H * TW_H + 256) * 4) + 1024;
}



}  // namespace rpb

using namespace rpb;

RPB_API int rpb_debug_fused_cta_times(uint64_t* out1024) {
    if (out1024 == nullptr) return RPB_ERR_BAD_ARG;
    return (int)cudaMemcpyFromSymbol(out1024, g_fg_cta, sizeof(unsigned long long) * 1024);
}

RPB_API int rpb_debug_fused_trace(uint64_t* out16, int enable) {
    int on = enable != 0;
    cudaError_t e = cudaSuccess;
    if (out16 != nullptr) {
        e = cudaMemcpyFromSymbol(out16, g_fg_trace, sizeof(unsigned long long) * 16);
        if (e != cudaSuccess) return (int)e;
    }
    static const unsigned long long zeros[16] = {};
    e = cudaMemcpyToSymbol(g_fg_trace, zeros, sizeof(zeros));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_fg_trace_on, &on, sizeof(int));
    return (int)e;
}

RPB_API int rpb_deepfm_fwd_fused(const RpbGatherDesc* g, const float* W1, const float* b1, const RpbTowerFwdDesc* d,
                                 void* stream) {
    if (g == nullptr || d == nullptr || W1 == nullptr || b1 == nullptr) return RPB_ERR_BAD_ARG;
    TowerFwdParams tw{};
    const int prc = tower_fwd_params(d, &tw);
    if (prc != 0) return prc;
    if (g->B != d->M || (g->tables == nullptr && g->G <= 1) || g->rows == nullptr || g->idx == nullptr || (g->Nd > 0 && g->dense == nullptr))
        return RPB_ERR_BAD_ARG;
    // shapes this kernel is built for: 64-byte rows, an even number of fields (one k-block = two fields)
    const bool sharded = g->G > 1;
    if (sharded && g->shard_tab == nullptr) return RPB_ERR_BAD_ARG;
    if (g->D != 16 || g->F < 2 || (g->F & 1) || g->F > RPB_MAX_FIELDS || g->Nd > RPB_MAX_DENSE || g->lr_tables != nullptr ||
        g->lr_in != nullptr || d->n_tail < 1 || d->M < 512 || !g_gemm_v2 || (d->ldh1 & 3) != 0)
        return RPB_ERR_UNSUPPORTED;
    const int K = g->F * 16 + g->Nd;
    if (g->x != nullptr && (g->ldx < ((K + 3) / 4) * 4 || (g->ldx & 3) != 0 || (reinterpret_cast<uintptr_t>(g->x) & 15u))) return RPB_ERR_BAD_ARG;
    if (g->fm_s != nullptr && (reinterpret_cast<uintptr_t>(g->fm_s) & 15u)) return RPB_ERR_UNSUPPORTED;
    FusedFwdParams p{};
    for (int f = 0; f < g->F; ++f) {
        if (g->idx[f] == nullptr || (reinterpret_cast<uintptr_t>(g->idx[f]) & 7u)) return RPB_ERR_UNSUPPORTED;
        if (!sharded && (g->tables[f] == nullptr || (reinterpret_cast<uintptr_t>(g->tables[f]) & 15u))) return RPB_ERR_UNSUPPORTED;
        p.tables[f] = sharded ? nullptr : g->tables[f];
        p.idx[f] = reinterpret_cast<const long long*>(g->idx[f]);
        p.rows[f] = g->rows[f];
    }
    for (int j = 0; j < g->Nd; ++j) {
        if (g->dense[j] == nullptr) return RPB_ERR_BAD_ARG;
        p.dense[j] = g->dense[j];
    }
    p.x = g->x; p.ldx = g->ldx; p.fm = g->fm; p.fm_s = g->fm_s; p.err = reinterpret_cast<long long*>(g->err);
    p.h1 = const_cast<float*>(d->h1); p.ldh1 = d->ldh1; p.bias1 = b1;
    p.M = d->M; p.F = g->F; p.Nd = g->Nd;
    p.G = sharded ? g->G : 1;
    p.shard_tab = sharded ? g->shard_tab : nullptr;
    p.l2_prefetch = sharded ? 0 : g_fused_l2_prefetch;
    p.nkb_emb = g->F / 2;
    p.nkb = p.nkb_emb + ceil_div(g->Nd, TC_BLOCK_K);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUtensorMap tmBhi, tmBlo;
    int rc = tc_prepare_weight(W1, K, FG_N, K, &tmBhi, &tmBlo, st);
    if (rc != 0) return rc;
    const int m_tiles = ceil_div(d->M, TC_BLOCK_M);
    const int grid = min(m_tiles, 148);
    const size_t cap = 227 * 1024;
    if (g_fused_gather_warps == 8 && g_fused_fetch_warps != 0 && g_fused_tc_tail != 0 && fs_smem_bytes(3, d->n_tail) <= cap) {
        // default: dedicated fetch warps + split warps + tcgen05 tower tail — deepfm_fwd_fs_kernel
        CUtensorMap tmX = tmBhi;                       // placeholder when x is not materialised
        if (p.x != nullptr) {
            rc = tc_make_map2d(&tmX, p.x, p.M, p.ldx, p.ldx, 32, 32, 128);
            if (rc != 0) return rc;
        }
        CUtensorMap tmThi, tmTlo;
        rc = tc_prepare_tail_weights(tw.W, d->n_tail, &tmThi, &tmTlo, st, 0, 6);
        if (rc != 0) return rc;
        const int full_rounds = m_tiles / grid;
        const int rem_rows = p.M - full_rounds * grid * TC_BLOCK_M;
        const int chunk = rem_rows > 0 ? min(TC_BLOCK_M, ((rem_rows + grid - 1) / grid + 31) / 32 * 32) : 0;
        auto launch_fs = [&](auto la_tag, auto sh_tag, auto nf_tag) -> int {
            constexpr int LA = decltype(la_tag)::value, NF = decltype(nf_tag)::value;
            constexpr bool SH = decltype(sh_tag)::value;
            const size_t smem = fs_smem_bytes(LA, d->n_tail, NF);
            cudaError_t e = cudaFuncSetAttribute(deepfm_fwd_fs_kernel<LA, SH, NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            deepfm_fwd_fs_kernel<LA, SH, NF><<<grid, fs_threads(NF), smem, st>>>(tmBhi, tmBlo, tmThi, tmTlo, tmX, p, tw, full_rounds, chunk);
            return (int)cudaGetLastError();
        };
        auto pick_fs = [&](auto la_tag) -> int {
            if (g_fused_fetch_warps == 16)
                return sharded ? launch_fs(la_tag, std::true_type{}, std::integral_constant<int, 16>{}) : launch_fs(la_tag, std::false_type{}, std::integral_constant<int, 16>{});
            if (g_fused_fetch_warps == 8)
                return sharded ? launch_fs(la_tag, std::true_type{}, std::integral_constant<int, 8>{}) : launch_fs(la_tag, std::false_type{}, std::integral_constant<int, 8>{});
            return sharded ? launch_fs(la_tag, std::true_type{}, std::integral_constant<int, 4>{}) : launch_fs(la_tag, std::false_type{}, std::integral_constant<int, 4>{});
        };
        const int want = g_fused_ring > 0 ? g_fused_ring : 6;
        for (int la = want; la >= 3; --la) {
            if (fs_smem_bytes(la, d->n_tail, g_fused_fetch_warps) > cap) continue;
            switch (la) {
                case 6: return pick_fs(std::integral_constant<int, 6>{});
                case 5: return pick_fs(std::integral_constant<int, 5>{});
                case 4: return pick_fs(std::integral_constant<int, 4>{});
                default: return pick_fs(std::integral_constant<int, 3>{});
            }
        }
        return RPB_ERR_UNSUPPORTED;
    }
    if (g_fused_gather_warps == 8) {
        // rpb_set_option("fused_fetch_warps", 0): 8 gather warps, x stored by TMA ([32 rows x 16 columns] boxes, SWIZZLE_64B) — deepfm_fwd_fused8_kernel;
        // tower-tail layers on tcgen05 unless rpb_set_option("fused_tc_tail", 0)
        const bool tc8 = g_fused_tc_tail != 0 && f8_smem_bytes(3, d->n_tail, true) <= cap;
        CUtensorMap tmX = tmBhi;                       // placeholder when x is not materialised
        if (p.x != nullptr) {
            rc = tc_make_map2d(&tmX, p.x, p.M, p.ldx, p.ldx, 16, 32, 64);
            if (rc != 0) return rc;
        }
        CUtensorMap tmThi = tmBhi, tmTlo = tmBlo;      // placeholders when the tail runs on the CUDA cores
        if (tc8) {
            rc = tc_prepare_tail_weights(tw.W, d->n_tail, &tmThi, &tmTlo, st, 0, 6);
            if (rc != 0) return rc;
        }
        auto launch8 = [&](auto la_tag, auto sh_tag, auto tc_tag) -> int {
            constexpr int LA = decltype(la_tag)::value;
            constexpr bool SH = decltype(sh_tag)::value, TC = decltype(tc_tag)::value;
            const size_t smem = f8_smem_bytes(LA, d->n_tail, TC);
            cudaError_t e = cudaFuncSetAttribute(deepfm_fwd_fused8_kernel<LA, SH, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            // schedule: full rounds of 128-row tiles, then the rest cut into per-CTA pieces (tcgen05 tail only: the CUDA-core
            // tail routine works on whole tiles)
            const int full_rounds = m_tiles / grid;
            const int rem_rows = p.M - full_rounds * grid * TC_BLOCK_M;
            int chunk = 0;
            if (rem_rows > 0) chunk = TC ? min(TC_BLOCK_M, ((rem_rows + grid - 1) / grid + 31) / 32 * 32) : TC_BLOCK_M;
            deepfm_fwd_fused8_kernel<LA, SH, TC><<<grid, F8_THREADS, smem, st>>>(tmBhi, tmBlo, tmThi, tmTlo, tmX, p, tw, full_rounds, chunk);
            return (int)cudaGetLastError();
        };
        auto pick8 = [&](auto la_tag) -> int {
            if (tc8) return sharded ? launch8(la_tag, std::true_type{}, std::true_type{}) : launch8(la_tag, std::false_type{}, std::true_type{});
            return sharded ? launch8(la_tag, std::true_type{}, std::false_type{}) : launch8(la_tag, std::false_type{}, std::false_type{});
        };
        const int want = g_fused_ring > 0 ? g_fused_ring : 6;
        for (int la = want; la >= 3; --la) {
            if (f8_smem_bytes(la, d->n_tail, tc8) > cap) continue;
            switch (la) {
                case 6: return pick8(std::integral_constant<int, 6>{});
                case 5: return pick8(std::integral_constant<int, 5>{});
                case 4: return pick8(std::integral_constant<int, 4>{});
                default: return pick8(std::integral_constant<int, 3>{});
            }
        }
        return RPB_ERR_UNSUPPORTED;
    }
    // round-1 kernel (rpb_set_option("fused_gather_warps", 4)): tail layers on tcgen05 opt-in, see FT_OP
    const bool tctail = g_fused_tc_tail != 0 && fg_smem_bytes(3, d->n_tail, true) <= cap;
    CUtensorMap tmThi = tmBhi, tmTlo = tmBlo;          // placeholders when the tail runs on the CUDA cores
    if (tctail) {
        rc = tc_prepare_tail_weights(tw.W, d->n_tail, &tmThi, &tmTlo, st, 0, 6);
        if (rc != 0) return rc;
    }
    auto launch = [&](auto la_tag, auto sh_tag, auto tc_tag) -> int {
        constexpr int LA = decltype(la_tag)::value;
        constexpr bool SH = decltype(sh_tag)::value, TC = decltype(tc_tag)::value;
        const size_t smem = fg_smem_bytes(LA, d->n_tail, TC);
        cudaError_t e = cudaFuncSetAttribute(deepfm_fwd_fused_kernel<LA, SH, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        if (g_l2_persist && p.x != nullptr)        // opt-in: keep the feature row in the L2 set-aside for the backward kernels
            return (int)launch_windowed(deepfm_fwd_fused_kernel<LA, SH, TC>, dim3(grid), dim3(FG_THREADS), smem, st, p.x,
                                        (size_t)p.M * (size_t)p.ldx * sizeof(float), tmBhi, tmBlo, tmThi, tmTlo, p, tw, m_tiles);
        deepfm_fwd_fused_kernel<LA, SH, TC><<<grid, FG_THREADS, smem, st>>>(tmBhi, tmBlo, tmThi, tmTlo, p, tw, m_tiles);
        return (int)cudaGetLastError();
    };
    auto pick = [&](auto la_tag) -> int {
        if (tctail) return sharded ? launch(la_tag, std::true_type{}, std::true_type{}) : launch(la_tag, std::false_type{}, std::true_type{});
        return sharded ? launch(la_tag, std::true_type{}, std::false_type{}) : launch(la_tag, std::false_type{}, std::false_type{});
    };
    if (fg_smem_bytes(5, d->n_tail, tctail) <= cap) return pick(std::integral_constant<int, 5>{});
    if (fg_smem_bytes(4, d->n_tail, tctail) <= cap) return pick(std::integral_constant<int, 4>{});
    if (fg_smem_bytes(3, d->n_tail, tctail) <= cap) return pick(std::integral_constant<int, 3>{});
    return RPB_ERR_UNSUPPORTED;
}
