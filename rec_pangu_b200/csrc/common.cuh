// rec_pangu_b200 — common device/host helpers for the sm_100a kernel library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rec_pangu_b200.h"

#define RPB_API extern "C" __attribute__((visibility("default")))

#define RPB_LAUNCH_CHECK()                                  \
    do {                                                    \
        cudaError_t _e = cudaGetLastError();                \
        if (_e != cudaSuccess) return (int)_e;              \
    } while (0)

namespace rpb {

constexpr int kWarp = 32;

void* workspace(int slot, size_t bytes, int* err);   // api.cu
extern int g_gather_kernel;                          // api.cu: 0 = smem-tile + bulk-store kernel, 1 = direct-store kernel
extern int g_gemm_v2;                                // api.cu: 1 = persistent GEMM kernel (v2), 0 = one tile per CTA (v1)
extern int g_gemm_a_tmem;                            // api.cu: 1 = v2 GEMM keeps the split A operand in tensor memory (TS-mode MMA)
extern int g_gemm_stack_n;                           // api.cu: 1 = narrow layers run [B hi ; B lo] as one N = 2*block_n operand
extern int g_tf32_raw_hi;                            // api.cu: 1 = feed unmasked fp32 as the tf32 'hi' operand (the tensor core ignores the low 13 mantissa bits)
extern int g_scatter_reverse;                        // api.cu: 1 = fused dx+scatter GEMM walks the sample tiles back to front
extern int g_wgrad_stages;                           // api.cu: pipeline depth of the tcgen05 weight-gradient kernel (2..4)
extern int g_wgrad_tc;                               // api.cu: 1 = tcgen05 weight-gradient kernel in auto mode
extern int g_fused_tc_tail;                          // api.cu: 1 = one-kernel DeepFM forward runs its 64x64 tail layers on tcgen05 (deepfm_fused.cu, TCTAIL)
extern int g_fused_gather_warps;                     // api.cu: 8 = one-kernel DeepFM forward with eight gather warps + TMA x store (default), 4 = the round-1 kernel
extern int g_fused_l2_prefetch;                     // api.cu: tiles by which the FS kernel's L2 prefetch warp runs ahead (0 = off, the default: measured 133 vs 99 us)
extern int g_fused_fetch_warps;                      // api.cu: 4 = one-kernel DeepFM forward with dedicated fetch warps (default), 0 = every gather warp fetches its own rows
extern int g_fused_ring;                             // api.cu: gather ring depth of the 8-warp forward (0 = deepest that fits)
extern int g_tower_bwd_tc;                           // api.cu: 1 = tower-tail backward runs its dz chain on tcgen05 (tower_tc.cu)
extern int g_rows_zero_blocks;                      // api.cu: > 0 = sparse re-zero launched as that many 128-thread blocks (co-resident with the forward kernel)
extern int g_cin_tc;                                 // api.cu: 1 = CIN layers on tcgen05 (cin_tc.cu) where the shape is instantiated
extern int g_autoint_vec;                            // api.cu: 1 = AutoInt attention kernels move a lane's outputs as float4 (autoint.cu, VEC)
extern int g_l2_persist;                             // api.cu: 1 = launches that write / re-read the feature row x carry an L2 persisting access-policy window on it
extern size_t g_l2_aside, g_l2_max_window;           // api.cu: persisting L2 set-aside / largest policy window of the device (set by rpb_set_option("l2_persist", 1))
extern int g_gather_policy;                          // api.cu: 0 = L1 no-allocate, 1 = + L2::64B, 2 = __ldg

// Fused scatter epilogue of the layer-1 dx GEMM: instead of writing dx[M, F*D+Nd] the epilogue adds every sample's
// per-field gradient row straight into the table gradients,
//   grads[f][idx[f][m], :] += dx[m, f*D:(f+1)*D] + dfm[m] * (fm_s[m, :] - x[m, f*D:(f+1)*D]),
// i.e. rpb_gather_bwd runs inside the GEMM epilogue and the 113 MB dx round trip through HBM disappears.
struct TcScatter {
    float* grads[RPB_MAX_FIELDS];
    const long long* idx[RPB_MAX_FIELDS];
    long long rows[RPB_MAX_FIELDS];
    const float* x; long long ldx;       // forward feature rows (FM term), may be null when dfm is null
    const float* dfm;                    // [M] or null
    const float* fm_s;                   // [M, D] or null
    int F, D, enabled;
    // row-sharded gradient buffers (RpbScatterDesc.G / grad_shard_tab): owner = id mod G, local row = id div G, entry
    // f*G+g of the DEVICE array = rank g's gradient shard of table f; remote shards take the reductions over NVLink.
    // G <= 1: `grads` holds the buffers.  With G > 1 grads[f] is only the "table f is trainable" flag.
    int G;
    float* const* grad_shard_tab;
};

struct TowerFwdParams;   // tower_tile.cuh

// epilogue descriptor of the tcgen05 GEMM (linear_tc.cu)
struct TcEpilogue {
    float* C; long long ldc;
    const float* bias;
    const float* mask; long long ldmask;
    int M, N;            // valid extents
    int relu;
    const TcScatter* sc; // host pointer (copied into the kernel parameter when enabled), else null
    // host pointer or null: run the rest of the MLP tower (tower_tile.cuh) on every finished 64-wide tile from the epilogue
    // warps; C then receives h1 as usual and the tail's outputs (h[l], logit, pred, loss) come from the same launch
    const TowerFwdParams* tail;
};

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Launch with an L2 access-policy window on [wptr, wptr + wbytes): the selected fraction of those lines is kept in the
// persisting set-aside of the L2, the rest streams.  A pure hint (opt-in g_l2_persist): the feature row x (113 MB at config 2)
// is written by the forward kernel and read again by the layer-1 weight gradient and by the scatter epilogue, and the L2
// holds 126 MB.  Launch attributes travel into CUDA-graph kernel nodes, unlike a stream attribute set during capture.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_windowed(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                          const void* wptr, size_t wbytes, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    int na = 0;
    if (wptr != nullptr && wbytes > 0 && rpb::g_l2_aside > 0 && rpb::g_l2_max_window > 0) {
        const size_t win = wbytes < rpb::g_l2_max_window ? wbytes : rpb::g_l2_max_window;
        const float ratio = (float)rpb::g_l2_aside / (float)win;
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = const_cast<void*>(wptr);
        attr[0].val.accessPolicyWindow.num_bytes = win;
        attr[0].val.accessPolicyWindow.hitRatio = ratio < 1.f ? ratio : 1.f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        na = 1;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float4 ldg_f4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
// streaming 128-bit load that does not allocate in L1 (rows are touched once)
__device__ __forceinline__ float4 ldg_f4_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
// same, with the L2 sector-promotion capped at 64 B: a 64 B embedding row must not drag its 128 B line in
__device__ __forceinline__ float4 ldg_f4_stream64(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_f4(float* p, const float4& v) {
    *reinterpret_cast<float4*>(p) = v;
}
// vector reduction: one 16-byte fp32x4 add, no return value (sm_90+)
__device__ __forceinline__ void red_add_f4(float* p, const float4& v) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red_add_f1(float* p, float v) {
    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int LANES>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum for blockDim.x <= 1024, result valid in thread 0
__device__ __forceinline__ float block_sum(float v, float* smem32) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) smem32[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? smem32[threadIdx.x] : 0.f;
    if (w == 0) v = warp_sum(v);
    __syncthreads();
    return v;
}

}  // namespace rpb
