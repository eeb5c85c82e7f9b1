/* rec_pangu_b200.h — C ABI of librec_pangu_b200.so (sm_100a kernels for the rec_pangu ranking hot path).
 *
 * The reference (HaSai666/rec_pangu) has no FFI: its hot path is a chain of ATen calls made from Python
 * nn.Modules.  Each entry point below replaces one such chain; the cited file:line
 * is the reference code whose results it must reproduce.  All pointers are raw device pointers unless a
 * parameter is documented as a host array; `stream` is a cudaStream_t passed as void*; every function
 * returns 0 on success, a cudaError_t (>0) from the launch, or a negative RPB_ERR_* code.  Nothing here
 * depends on PyTorch.  There is no CPU implementation behind any of these symbols.
 *
 * Layout conventions
 *   x  : [B, ldx] fp32 row-major "feature row": columns [0, F*D) = the F gathered embedding rows of a sample
 *        (the reference's [B,F,D] tensor, models/layers/embedding.py:63), columns [F*D, F*D+Nd) = the dense
 *        features (models/utils.py:122-137), columns up to ldx zero.  ldx % 4 == 0.  This is at once
 *        `sparse_embedding` (strided view) and `dnn_input` (ranking/deepfm.py:52-58) — the cat is never made.
 *   W  : nn.Linear layout [N, K] row-major (K contiguous).
 */
#ifndef REC_PANGU_B200_H
#define REC_PANGU_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RPB_MAX_FIELDS 64
#define RPB_MAX_DENSE 64

#define RPB_ERR_UNSUPPORTED (-1)   /* shape outside what the kernels were built for */
#define RPB_ERR_BAD_ARG (-2)
#define RPB_ERR_NO_DRIVER (-3)     /* cuTensorMapEncodeTiled could not be resolved */

/* ABI version; bumped whenever a signature changes. */
int rpb_version(void);
/* Tuning knobs (process-wide): "gather_load_policy" 0 = L1 no-allocate row loads, 1 = + L2 64-byte fetch cap (default),
 * 2 = cached read-only loads, 3 = measurement-only (no x write);  "gather_kernel" 0 = shared-memory tile + bulk
 * store (default), 1 = direct stores;  "l2_fetch_granularity" = 32|64|128 (cudaLimitMaxL2FetchGranularity);
 * "gemm_v2" 1 = persistent tcgen05 GEMM (default), 0 = one tile per CTA;  "gemm_a_tmem" 1 = the split A operand lives in
 * tensor memory (TS-mode MMA, default), 0 = in shared memory;  "gemm_stack_n" 1 = layers of <= 64 outputs multiply by
 * [B hi ; B lo] as one operand (default);  "wgrad_tc" 1 = tcgen05 weight gradient (default). */
int rpb_set_option(const char* name, int64_t value);
/* Diagnostics for the persistent tcgen05 GEMM: per-role stall cycles of CTA 0 of the last launch (see linear_tc.cu);
 * out16 may be NULL; `enable` switches the counters on/off for subsequent launches (synchronous, not graph-safe). */
int rpb_debug_tc_trace(uint64_t* out16, int enable);
/* Last error text for negative codes (static string). */
const char* rpb_error_string(int code);

/* ------------------------------------------------------------------------------------------------
 * Multi-table embedding gather (+ dense pack, + FM second order, + LR wide inputs) — one launch.
 * Replaces EmbeddingLayer.forward (models/layers/embedding.py:49-63: F x aten::embedding + stack),
 * get_linear_input (models/utils.py:122-137), FM_Layer / InnerProductLayer product_sum_pooling
 * (models/layers/interaction.py:36-44,225-235) and LR_Layer's second D=1 gather
 * (models/layers/shallow.py:22-26).
 * Index semantics are the reference's: int64 row ids, table f has rows[f] = vocab_size+1 rows
 * (embedding.py:32), id == vocab_size is the OOV row; id < 0 or id >= rows[f] is an error: the kernel
 * records {1, field, sample, value} into err[0..3] (first error wins) and reads row 0 instead.
 * ---------------------------------------------------------------------------------------------- */
typedef struct RpbGatherDesc {
    int32_t B, F, D, Nd;
    int32_t ldx;                    /* row stride of x in floats, >= F*D+Nd, % 4 == 0 */
    int32_t ld_lr;                  /* row stride of lr_in in floats (>= F+Nd), 0 if lr_in == NULL */
    const float* const* tables;     /* host array [F] of device ptrs: table f = float[rows[f]][D] */
    const int64_t* rows;            /* host array [F] */
    const int64_t* const* idx;      /* host array [F] of device ptrs int64[B] */
    const float* const* dense;      /* host array [Nd] of device ptrs float[B]; NULL when Nd == 0 */
    const float* const* lr_tables;  /* host array [F] of device ptrs float[rows[f]] (D=1 tables) or NULL */
    float* x;                       /* out [B, ldx]; may be NULL when only fm / lr_in are consumed (FM inference) */
    float* fm;                      /* out [B]: 0.5*sum_d((sum_f e)^2 - sum_f e^2), or NULL */
    float* fm_s;                    /* out [B, D]: sum_f e (saved for backward), or NULL */
    float* lr_in;                   /* out [B, ld_lr]: [lr_table_f[idx_f] (F) | dense (Nd)], or NULL */
    int64_t* err;                   /* device-visible int64[4] error record (see above) */
    /* Row-sharded tables over G GPUs of one NVLink domain (BASELINE.json config 5): owner = id mod G, local row =
     * id div G.  shard_tab is a DEVICE array [F*G] of shard base pointers, entry f*G+g = rank g's shard of table f
     * (float[ceil(rows[f]/G)][D]) mapped into this process (NVLink peer memory); the kernel reads remote rows
     * directly, so the lookup all-to-all is fused into the gather.  G <= 1: unsharded, `tables` is used.
     * rows[] stay the GLOBAL row counts (bounds check).  LR tables are not supported together with sharding. */
    int32_t G;
    const float* const* shard_tab;
} RpbGatherDesc;
int rpb_gather_fwd(const RpbGatherDesc* d, void* stream);

/* Backward of the gather: scatter-add of per-sample row gradients into per-table gradients
 * (autograd of aten::embedding = embedding_dense_backward), fused with the FM and LR
 * backward terms:  gE[b,f,:] = dx[b, f*D:(f+1)*D] + dfm[b]*(fm_s[b,:] - x[b,f,:]);
 *                  grads[f][idx[f][b], :] += gE[b,f,:];   lr_grads[f][idx[f][b]] += dlr_in[b,f].
 * grads / lr_grads must be zero-initialised (dense mode) by the caller; additions are fp32 atomics. */
typedef struct RpbScatterDesc {
    int32_t B, F, D;
    int32_t lddx, ldx, ld_dlr;
    float* const* grads;            /* host array [F] of device ptrs float[rows[f]][D]; entry may be NULL (frozen table) */
    float* const* lr_grads;         /* host array [F] of device ptrs float[rows[f]] or NULL */
    const int64_t* rows;            /* host array [F] */
    const int64_t* const* idx;      /* host array [F] of device ptrs int64[B] */
    const float* dx;                /* [B, lddx] grad wrt x, or NULL */
    const float* x;                 /* [B, ldx] forward output (needed iff dfm) */
    const float* dfm;               /* [B] grad wrt fm, or NULL */
    const float* fm_s;              /* [B, D] saved sum_f e (needed iff dfm) */
    const float* dlr_in;            /* [B, ld_dlr] grad wrt lr_in (first F columns used), or NULL */
    /* Row-sharded gradient buffers (see RpbGatherDesc.G): DEVICE array [F*G] of grad-shard base pointers; remote
     * shards receive fp32 vector reductions over NVLink.  G <= 1: `grads` is used. */
    int32_t G;
    float* const* grad_shard_tab;
} RpbScatterDesc;
int rpb_gather_bwd(const RpbScatterDesc* d, void* stream);
/* Sparse zero_grad for persistent dense-grad buffers: grads[f][idx[f][b], :] = 0 and lr_grads[f][idx[f][b]] = 0
 * for every (b, f) of a previous backward (uses B, F, D, grads, lr_grads, rows, idx of the descriptor only).
 * Leaves a [rows, D] buffer all-zero again at O(batch) cost instead of the reference's O(vocabulary) zero-fill. */
int rpb_rows_zero(const RpbScatterDesc* d, void* stream);

/* Hashed-id encoder for tables too large for the reference's sorted-unique vocabulary map (dataset/base_dataset.py:57-61,92),
 * BASELINE.json config 5: out[i] = splitmix64(raw[i]) mod vocab_size, int64 in [0, vocab_size) — row vocab_size stays
 * the OOV slot.  Integer contract restated in oracle/index_routing.py::hash_to_row (bit-exact).  In place allowed. */
int rpb_hash_to_row(const int64_t* raw, int64_t* out, int64_t n, int64_t vocab_size, void* stream);

/* Standalone FM second-order term on any [B,F,D] tensor with row stride lde (floats) between samples:
 * InnerProductLayer (interaction.py:36-44).  out_sum [B] (product_sum_pooling) and/or out_bi [B,D]
 * (Bi_interaction_pooling) may be NULL. */
int rpb_fm_fwd(const float* e, int64_t lde, int B, int F, int D, float* out_sum, float* out_bi, void* stream);
/* de[b,f,:] (+)= (dsum[b] + dbi[b,:]) * (s[b,:] - e[b,f,:]); accumulate != 0 adds into de. */
int rpb_fm_bwd(const float* e, int64_t lde, int B, int F, int D, const float* dsum, const float* dbi,
               float* de, int64_t ldde, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense layers.  nn.Linear + activation of MLP (models/layers/deep.py:62-70).
 * y[M,N] = act(x[M,K] @ W[N,K]^T + bias[N]);  act: 0 none, 1 relu.
 * impl: 0 = auto (tcgen05 3xTF32 when the shape qualifies), 1 = SIMT fp32, 2 = tcgen05 3xTF32.
 * ---------------------------------------------------------------------------------------------- */
int rpb_linear_fwd(const float* x, int64_t ldx, const float* W, const float* bias, float* y, int64_t ldy,
                   int M, int N, int K, int act, int impl, void* stream);
/* dx[M,K] = (dy[M,N] @ W[N,K]) * (mask ? (mask[m,k] > 0) : 1)   — mask = saved post-ReLU input of this
 * layer, fusing the previous layer's ReLU backward.  dx may be NULL.
 * dW[N,K] += dy^T @ x,  db[N] += colsum(dy)   (dW/db must be zero-initialised or hold the running sum). */
int rpb_linear_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* W,
                   const float* mask, int64_t ldmask, float* dx, int64_t lddx, float* dW, float* db,
                   int M, int N, int K, int impl, void* stream);

/* First-layer dx GEMM fused with the gradient scatter: computes dx = dy[M,N] @ W[N,K] tile by tile on tcgen05 and, in
 * the epilogue, adds dx[m, f*D:(f+1)*D] + dfm[m]*(fm_s[m,:] - x[m, f*D:(f+1)*D]) straight into grads[f][idx[f][m], :]
 * (fields of RpbScatterDesc: B (= M), F, D, grads, rows, idx, x, ldx, dfm, fm_s, G, grad_shard_tab; dx / lddx / LR
 * fields are ignored).  With G > 1 the reductions go to the owners' gradient shards (local HBM or NVLink peer memory)
 * and grads[f] != NULL only marks table f as trainable.  dx is never written to HBM.  Returns RPB_ERR_UNSUPPORTED when the persistent tcgen05
 * kernel cannot take the shape (the caller then runs rpb_linear_bwd + rpb_gather_bwd). */
int rpb_linear_dx_scatter(const float* dy, int64_t lddy, const float* W, int M, int N, int K,
                          const RpbScatterDesc* d, void* stream);

/* Row dot (N=1 linear): out[m] = x[m,:K].w + bias[0] + add0[m] + add1[m] + add2[m]  (NULL addends skipped).
 * Final Linear(->1) of the MLP plus the logit sum of ranking/deepfm.py:61, xdeepfm.py:69, autoint.py:72-81. */
int rpb_rowdot_fwd(const float* x, int64_t ldx, const float* w, const float* bias, const float* add0,
                   const float* add1, const float* add2, float* out, int M, int K, void* stream);
/* dx[m,k] = dout[m]*w[k] * (mask ? mask[m,k]>0 : 1) (dx may be NULL); dw[k] += sum_m dout[m]*x[m,k]; db[0] += sum dout. */
int rpb_rowdot_bwd(const float* dout, const float* x, int64_t ldx, const float* w, const float* mask,
                   int64_t ldmask, float* dx, int64_t lddx, float* dw, float* db, int M, int K, void* stream);

/* pred = sigmoid(logit); loss = mean BCE(pred, label) with log clamped at -100 like ATen
 * (torch.nn.BCELoss, e.g. ranking/deepfm.py:31,61-63).  eps is added to pred before the BCE
 * (multi_task/mmoe.py:127-128 uses 1e-6; ranking models 0).  loss_out[0] = scale * mean.  label/loss_out may
 * be NULL (inference).  work: device int32[2 + 1024*2] scratch owned by the caller, zero-initialised once. */
int rpb_sigmoid_bce_fwd(const float* logit, const float* label, float* pred, float* loss_out, float eps,
                        float scale, int M, void* work, void* stream);
/* dlogit[m] = gloss[0]*scale/M * dBCE/dp * p(1-p), ATen formulas (binary_cross_entropy_backward: denominator
 * clamped at 1e-12). */
int rpb_sigmoid_bce_bwd(const float* pred, const float* label, const float* gloss, float eps, float scale,
                        float* dlogit, int M, void* stream);

/* ESSM head (multi_task/essm.py:50-75): click = sigmoid(z1), conv = sigmoid(z2),
 * loss[0] = mean BCE(click*conv, y2) + w_ctr * mean BCE(click, y1)  (the reference feeds the PRODUCT to its loss while
 * reporting conv as task2_pred; w_ctr = 0.5).  y1 / y2 / loss_out may be NULL (inference).  work as rpb_sigmoid_bce_fwd.
 * bwd: dz1, dz2 = gloss[0] * dloss/dz (gloss NULL = 1), ATen clamps. */
int rpb_essm_head_fwd(const float* z1, const float* z2, const float* y1, const float* y2, float* click, float* conv,
                      float* loss_out, float w_ctr, int M, void* work, void* stream);
int rpb_essm_head_bwd(const float* click, const float* conv, const float* y1, const float* y2, const float* gloss,
                      float w_ctr, float* dz1, float* dz2, int M, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MLP tower tail: the n_tail (0..RPB_TOWER_MAX_TAIL) square H x H hidden layers that follow the first layer, the
 * Linear(H -> 1) output layer, the logit sum, the sigmoid and the mean BCE in ONE launch (models/layers/deep.py:62-84
 * with ReLU after every hidden layer and dropout inactive; ranking/deepfm.py:57-63 for `fm + dnn`, sigmoid, BCELoss).
 * Exact fp32 FMAs on a shared-memory resident activation tile; H must be 64.
 *   h[l] = relu(h[l-1] @ W[l]^T + b[l])   (h[-1] = h1, the post-ReLU output of layer 1, row stride ldh1)
 *   logit = h[n_tail-1] . w_out + b_out[0] (+ addend);  pred = sigmoid(logit);  loss = scale * mean BCE(pred + eps, label)
 * W, b, h are HOST arrays of n_tail device pointers; h[l] is [M, H] contiguous (saved for backward).  pred / label /
 * loss may be NULL (loss needs label, pred and work: the same int32[2 + 2048] scratch as rpb_sigmoid_bce_fwd). */
#define RPB_TOWER_MAX_TAIL 4
typedef struct RpbTowerFwdDesc {
    int32_t M, H, n_tail;
    const float* h1; int64_t ldh1;
    const float* const* W; const float* const* b; float* const* h;
    const float* w_out; const float* b_out; const float* addend;
    float* logit; const float* label; float* pred; float* loss;
    float eps, scale;
    void* work;
} RpbTowerFwdDesc;
int rpb_tower_tail_fwd(const RpbTowerFwdDesc* d, void* stream);
/* The first layer and the tail in ONE launch: h1 = relu(x[M,K] @ W1[64,K]^T + b1) on the persistent tcgen05 GEMM
 * (3xTF32), whose epilogue warps then run rpb_tower_tail_fwd's tile routine on every finished 128-sample tile while
 * the tensor pipe works on the next ones — h1 goes to HBM once (d->h1, an OUTPUT here, saved for backward) and is never
 * read back.  Needs n_tail >= 1 and M >= 512; returns RPB_ERR_UNSUPPORTED otherwise (the caller then launches
 * rpb_linear_fwd + rpb_tower_tail_fwd). */
int rpb_linear_tower_fwd(const float* x, int64_t ldx, const float* W1, const float* b1, int K,
                         const RpbTowerFwdDesc* d, void* stream);
/* DeepFM forward in ONE launch (ranking/deepfm.py:41-67): the gather + dense pack + FM second order of rpb_gather_fwd, the
 * layer-1 GEMM and rpb_tower_tail_fwd, overlapped inside one persistent tcgen05 kernel — the GEMM's operand warps fetch
 * their own table rows with cp.async (the random-row DRAM latency hides behind the tensor / tail work of the previous
 * k-blocks) and the FM term is formed from the registers that feed the MMAs.  Uses of `g`: B, F, D (= 16), Nd, tables,
 * rows, idx, dense, err; x / fm / fm_s are OPTIONAL outputs (x and fm_s only when backward needs them; ldx as for
 * rpb_gather_fwd).  `d` as for rpb_linear_tower_fwd (d->h1 is an output, d->addend is ignored: the FM term is added inside).
 * Row-sharded tables (g->G > 1, g->shard_tab): the same row requests go to the owner's shard, local or over NVLink.
 * Needs D == 16, an even F, n_tail >= 1, M >= 512; RPB_ERR_UNSUPPORTED otherwise (the caller then runs
 * rpb_gather_fwd + rpb_linear_tower_fwd). */
int rpb_deepfm_fwd_fused(const RpbGatherDesc* g, const float* W1, const float* b1, const RpbTowerFwdDesc* d, void* stream);
/* Per-role stall cycles of CTA 0 of the last rpb_deepfm_fwd_fused launch (same protocol as rpb_debug_tc_trace; see
 * deepfm_fused.cu for the 12 counters). */
int rpb_debug_fused_trace(uint64_t* out16, int enable);
/* per-CTA record of the last traced launch of the 8-gather-warp kernel: out1024[cta*4 + {0: SM id, 1: start ns, 2: end ns,
 * 3: tiles}] (globaltimer), 256 CTAs. */
int rpb_debug_fused_cta_times(uint64_t* out1024);
/* Backward of rpb_tower_tail_fwd.  hin: HOST array of n_tail+1 device pointers, hin[0] = h1 (row stride ldh1),
 * hin[j] = h[j-1] ([M, H] contiguous).  dlogit[m] = dlogit_in[m] when given, else gloss[0]*scale/M * dBCE/dp * p(1-p)
 * from (pred, label) with ATen's clamps (gloss NULL = 1); it is written to dlogit_out when non-NULL.
 * dz: HOST array of n_tail+1 device pointers, out [M, H]: dz[j] = gradient wrt the pre-activation that produced hin[j]
 * (dz[0] feeds the first layer's weight-gradient and dx kernels, dz[j] with hin[j-1] gives dW[j-1]).
 * db (HOST array of n_tail+1 device pointers or NULL, entries may be NULL): db[j] += colsum(dz[j]);
 * dw_out[H] += sum_m dlogit[m] * hin[n_tail][m, :];  db_out[0] += sum_m dlogit[m]. */
typedef struct RpbTowerBwdDesc {
    int32_t M, H, n_tail;
    const float* const* hin; int64_t ldh1;
    const float* const* W; const float* w_out;
    float* const* dz; float* const* db;
    float* dw_out; float* db_out;
    const float* pred; const float* label; const float* gloss; float eps, scale;
    const float* dlogit_in; float* dlogit_out;
} RpbTowerBwdDesc;
int rpb_tower_tail_bwd(const RpbTowerBwdDesc* d, void* stream);

/* torch.nn.LayerNorm over the last dimension of x [M, N] (row stride ldx), N <= 1024: y = (x - mean) * rstd * gamma + beta with
 * the biased variance and eps inside the square root; mean / rstd [M] are saved for backward.  Used by the MaskBlock of
 * MaskNet (models/layers/interaction.py:254-283).  bwd: dx, and dgamma / dbeta ACCUMULATED into zero-initialised [N]
 * buffers (column sums over the batch, fp32 atomics). */
int rpb_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, float* y, int64_t ldy,
                      float* mean, float* rstd, int32_t M, int32_t N, void* stream);
int rpb_layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* gamma, const float* mean,
                      const float* rstd, float* dx, int64_t lddx, float* dgamma, float* dbeta, int32_t M, int32_t N, void* stream);

/* nn.Dropout of the MLP (models/layers/deep.py:71-72, default p=0.1 for xDeepFM/AutoInt) on a contiguous
 * buffer of n floats.  keep(i) comes from a counter-based generator keyed by (seed, i), so backward recomputes
 * the mask instead of storing it.  Train-mode equivalence with torch's Philox stream is statistical
 * (SURVEY.md §7 hard-part 4).  bwd: dx = dy * keep/(1-p) * (relu_out ? relu_out > 0 : 1).
 * `epoch` (device pointer or NULL): a step counter in device memory that is mixed into the seed when the kernel RUNS, so a
 * training step captured once as a CUDA graph draws a fresh mask at every replay (the host-drawn `seed` is baked into the
 * graph); forward and backward of one step must see the same *epoch. */
int rpb_dropout_fwd(const float* x, float* y, int64_t n, float p, uint64_t seed, const uint64_t* epoch, void* stream);
int rpb_dropout_bwd(const float* dy, const float* relu_out, float* dx, int64_t n, float p, uint64_t seed,
                    const uint64_t* epoch, void* stream);

/* ------------------------------------------------------------------------------------------------
 * DCN CrossNet, all L (<= 8) layers in one kernel (models/layers/interaction.py:119-141):
 *   x_{l+1} = x_l + (w_l . x_l) x_0 + b_l.   x0: [B, ldx] (K valid columns), out: [B, ldo] (columns >= K zeroed),
 * w/bias: host arrays of L device pointers float[K] (CrossInteractionLayer.weight.weight / .bias),
 * S: [B, L] out, s_l = w_l . x_l saved for backward (may be NULL for inference).  K <= 1024.
 * bwd: dx0 [B, lddx] may be NULL; dw[l] / db[l] (float[K], may be NULL) are accumulated into (+=). */
int rpb_crossnet_fwd(const float* x0, int64_t ldx, int K, int L, const float* const* w, const float* const* bias,
                     float* out, int64_t ldo, float* S, int B, void* stream);
int rpb_crossnet_bwd(const float* x0, int64_t ldx, int K, int L, const float* const* w, const float* const* bias,
                     const float* S, const float* dout, int64_t lddo, float* dx0, int64_t lddx,
                     float* const* dw, float* const* db, int B, void* stream);

/* ------------------------------------------------------------------------------------------------
 * xDeepFM CIN (models/layers/interaction.py:144-171).  e: [B,F,D] with sample stride lde; units[L];
 * W[k]: Conv1d weight float[U_k][F*M_k] (M_0 = F, M_k = U_{k-1}), bias[k]: float[U_k];
 * pooled: [B, ldp] out, columns [sum_{j<k} U_j, +U_k) = sum_d X_{k+1}[b,:,d] (the tensor fed to cin.fc).
 * Limits: F <= 32, U_k <= 32, L <= 8, D in {8,16,32}, all layers' weights must fit in shared memory.
 * bwd: de[b,f,:] (+)= dL/dE (accumulate != 0 adds), dW[k]/db[k] are accumulated into (+=). */
int rpb_cin_fwd(const float* e, int64_t lde, int B, int F, int D, int L, const int32_t* units,
                const float* const* W, const float* const* bias, float* pooled, int64_t ldp, void* stream);
int rpb_cin_bwd(const float* e, int64_t lde, int B, int F, int D, int L, const int32_t* units,
                const float* const* W, const float* const* bias, const float* dpooled, int64_t lddp,
                float* de, int64_t ldde, int accumulate, float* const* dW, float* const* db, void* stream);
/* Training variants (tensor-core shapes only: F = 26, D = 16, U_k = 16; RPB_ERR_UNSUPPORTED otherwise — fall back to the pair
 * above): the forward keeps X_1 .. X_{L-1} in xsave [B, ldx] (layer k+1 at column (U_0 + .. + U_{k-1}) * D, ldx >= that sum over
 * the first L-1 layers, 16-byte aligned) — what autograd keeps alive in the reference (interaction.py:160-168) — and the
 * backward reads them instead of recomputing the forward. */
int rpb_cin_fwd_save(const float* e, int64_t lde, int B, int F, int D, int L, const int32_t* units,
                     const float* const* W, const float* const* bias, float* pooled, int64_t ldp,
                     float* xsave, int64_t ldx, void* stream);
int rpb_cin_bwd_saved(const float* e, int64_t lde, int B, int F, int D, int L, const int32_t* units,
                      const float* const* W, const float* const* bias, const float* dpooled, int64_t lddp,
                      float* de, int64_t ldde, int accumulate, float* const* dW, float* const* db,
                      const float* xsaved, int64_t ldx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * AutoInt interacting-layer core (models/layers/attention.py:12-32,63-95) on projected inputs.
 * qkvr: [B*F, ldq] = X @ [W_q;W_k;W_v;W_res]^T  (Q at column 0, K at H*d, V at 2H*d, R at 3H*d).  When the layer
 * has no W_res (D == H*d) pass res = X [B*F, ldres] and only Q|K|V in qkvr.
 * out: [B, F, H*d] = relu(attention(raw-view head regroup, no scale) + residual).   F <= 32, d in {4,8,16}.
 * bwd: dqkvr [B*F, lddq] receives dQ|dK|dV|dR (dR = ReLU-masked dout, always written). */
int rpb_autoint_attn_fwd(const float* qkvr, int64_t ldq, const float* res, int64_t ldres, float* out, int B, int F,
                         int H, int d, void* stream);
int rpb_autoint_attn_bwd(const float* qkvr, int64_t ldq, int has_res_proj, const float* out, const float* dout,
                         float* dqkvr, int64_t lddq, int B, int F, int H, int d, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MMOE (models/multi_task/mmoe.py:86-112).
 * rpb_matmul_kn_*: y[M,N] = x[M,K] @ Wkn[K,N] + bias[N] with the weight stored [K,N] (row stride ldw) like the
 * reference's `experts` [hid, Hh, E] viewed as [hid, Hh*E] and `gates[t]` [hid, E]; einsum('ij,jkl->ikl') and
 * einsum('ab,bc->ac') of mmoe.py:86,92 become one GEMM over the column-concatenated weight.
 * bwd: dx = dy @ Wkn^T (may be NULL), dWkn += x^T dy, db += colsum(dy) (each may be NULL). */
int rpb_matmul_kn_fwd(const float* x, int64_t ldx, const float* Wkn, int64_t ldw, const float* bias, float* y,
                      int64_t ldy, int M, int N, int K, int impl, void* stream);
int rpb_matmul_kn_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* Wkn, int64_t ldw,
                      float* dx, int64_t lddx, float* dWkn, float* db, int M, int N, int K, int impl, void* stream);
/* eo: [B, ld] = [experts_out (Hh*E, column k*E+l) | gate logits (T*E)].  gate[b, t*E+l] = softmax_l(logits_t)
 * (mmoe.py:95), out[t, b, k] = sum_l eo[b, k*E+l] * gate[b, t*E+l] (mmoe.py:99-104).  E <= 32.
 * bwd writes deo [B, ldd] for all Hh*E + T*E columns (pad columns zeroed). */
int rpb_mmoe_combine_fwd(const float* eo, int64_t ld, int B, int Hh, int E, int T, float* out, float* gate, void* stream);
int rpb_mmoe_combine_bwd(const float* eo, int64_t ld, const float* gate, const float* dout, int B, int Hh, int E, int T,
                         float* deo, int64_t ldd, void* stream);
/* BatchNorm1d of the task towers (mmoe.py:54-56) on a contiguous [M,N] activation.
 * rpb_bn_stats: sum[n] += sum_m x, sumsq[n] += sum_m x^2 (zero-initialised by the caller; mean/var formed on [N]).
 * rpb_bn_apply: y = (x - mean) * invstd * gamma + beta.
 * rpb_bn_bwd: dbeta += colsum(dy), dgamma += colsum(dy * xhat) (zero-initialised by the caller);
 *             dx = gamma*invstd*(dy - dbeta/M - xhat*dgamma/M) when use_batch_stats, else gamma*invstd*dy. */
int rpb_bn_stats(const float* x, int M, int N, float* sum, float* sumsq, void* stream);
int rpb_bn_apply(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                 float* y, int M, int N, void* stream);
int rpb_bn_bwd(const float* dy, const float* x, const float* mean, const float* invstd, const float* gamma,
               float* dx, float* dgamma, float* dbeta, int M, int N, int use_batch_stats, void* stream);
/* The two halves of rpb_bn_bwd, for batch statistics that span several GPUs (BatchNorm1d needs cross-GPU
 * batch statistics for parity with the single-process reference): rpb_bn_bwd_stats accumulates the LOCAL column sums
 * (these are also the parameter gradients of this rank's samples), the caller all-reduces them, rpb_bn_bwd_dx applies
 * dx = gamma*invstd*(dy - dbeta_sum*inv_count - xhat*dgamma_sum*inv_count) with inv_count = 1 / (global sample count). */
int rpb_bn_bwd_stats(const float* dy, const float* x, const float* mean, const float* invstd, float* dgamma,
                     float* dbeta, int M, int N, void* stream);
int rpb_bn_bwd_dx(const float* dy, const float* x, const float* mean, const float* invstd, const float* gamma,
                  const float* dgamma_sum, const float* dbeta_sum, float* dx, int M, int N, float inv_count,
                  int use_batch_stats, void* stream);

/* ------------------------------------------------------------------------------------------------
 * FiBiNet interaction (ranking/fibinet.py:59-66): SENET_Layer (interaction.py:238-251) + BilinearInteractionLayer
 * 'field_interaction' (interaction.py:55-81) applied to the raw and the SENET-reweighted embeddings with the SAME
 * weights, written straight into the MLP input row
 *   comb[b] = [ bilinear(E) (P*D) | bilinear(SENET(E)) (P*D) | dense (Nd) | 0-pad ],  P = F(F-1)/2 (combinations order).
 * x: feature row [B, ldx] (embeddings then dense);  W1: [R, F], W2: [F, R] (excitation.0/.2 weights, no bias);
 * Wb: [P, D, D] stacked bilinear_layer.<p>.weight;  A: [B, F] out — SENET weights saved for backward.
 * Limits: 2 <= F <= 32, D in {8,16,32}, R <= D.
 * bwd: dx [B, lddx] (embedding columns; the rest zero) is written; dW1/dW2/dWb are accumulated into (+=). */
int rpb_fibinet_fwd(const float* x, int64_t ldx, int B, int F, int D, int Nd, const float* W1, int R,
                    const float* W2, const float* Wb, float* comb, int64_t ldc, float* A, void* stream);
int rpb_fibinet_bwd(const float* x, int64_t ldx, int B, int F, int D, const float* W1, int R, const float* W2,
                    const float* Wb, const float* A, const float* dcomb, int64_t lddc, float* dx, int64_t lddx,
                    float* dW1, float* dW2, float* dWb, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer step (reference: torch.optim.Adam created in rec_pangu/trainer.py:75, stepped in model_pipeline.py:57).
 * rpb_adam_dense: element-wise Adam (torch semantics: bias-corrected, eps added after the sqrt(v)/sqrt(bc2)).
 * rpb_sparse_adam: row-sparse ("lazy") Adam for embedding tables in 'persistent' grad mode — only rows named by
 * idx are updated (once per step even when several samples hit them: claimed through stamps[f][row] = step), and the
 * row of the persistent gradient buffer is re-zeroed in the same pass.  Rows without gradient keep value and moments
 * (torch.optim.SparseAdam semantics; the reference's dense Adam would keep moving them by momentum). */
int rpb_adam_dense(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, int step, void* stream);
/* The same update for up to RPB_ADAM_MAX_TENSORS dense parameters in ONE launch (host arrays of `count` device
 * pointers / element counts).  step_dev: optional DEVICE int32 holding the step number; when non-NULL it overrides
 * `step`, so a training step captured into a CUDA graph keeps the right bias correction across replays. */
#define RPB_ADAM_MAX_TENSORS 32
typedef struct RpbAdamMultiDesc {
    int32_t count, step;
    float lr, beta1, beta2, eps;
    float* const* params; const float* const* grads; float* const* exp_avg; float* const* exp_avg_sq;
    const int64_t* numel;
    const int32_t* step_dev;
} RpbAdamMultiDesc;
int rpb_adam_multi(const RpbAdamMultiDesc* d, void* stream);
typedef struct RpbSparseAdamDesc {
    int32_t B, F, D, step;          /* step >= 1: global optimizer step (bias correction) */
    float lr, beta1, beta2, eps;
    float* const* weights;          /* host arrays [F] of device pointers; weights[f] == NULL skips table f */
    float* const* grads;            /* persistent dense gradient buffers [rows, D] */
    float* const* exp_avg;
    float* const* exp_avg_sq;
    int32_t* const* stamps;         /* int32[rows[f]], zero-initialised once */
    const int64_t* rows;
    const int64_t* const* idx;      /* int64[B] per field: the batch whose backward filled `grads` */
    const int32_t* step_dev;        /* optional DEVICE int32 step number (overrides `step`; see rpb_adam_multi) */
} RpbSparseAdamDesc;
int rpb_sparse_adam(const RpbSparseAdamDesc* d, void* stream);
/* Exact-lazy Adam (results equal the reference's DENSE torch.optim.Adam, trainer.py:75, although only touched rows are ever
 * accessed): stamps[f][row] = last optimizer step applied to the row.  rpb_sparse_adam_catchup replays, for the rows of the
 * batch `idx` that is about to be READ, the zero-gradient steps they missed (moments decay, weight walks by its momentum) up
 * to the current step (*step_dev, or `step`); `grads` is ignored.  rpb_sparse_adam_flush does the same for every row of one
 * table (before a state_dict is taken).  rpb_sparse_adam then applies the step with gradients as before. */
int rpb_sparse_adam_catchup(const RpbSparseAdamDesc* d, void* stream);
int rpb_sparse_adam_flush(float* w, float* m, float* v, int32_t* stamp, int64_t rows, int32_t D, float lr, float beta1,
                          float beta2, float eps, const int32_t* step_dev, int32_t step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REC_PANGU_B200_H */
