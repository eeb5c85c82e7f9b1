// Library-level entry points.
#include <string>

#include "common.cuh"

namespace rpb {

int g_gather_policy = 1;
int g_gather_kernel = 0;
int g_wgrad_tc = 1;
int g_scatter_reverse = 1;
int g_wgrad_stages = 2;     // 2 stages = 2 CTAs per SM: measured faster than 4 stages x 1 CTA (16.9 vs 22.6 us on the 64x64 layers)
int g_gemm_v2 = 1;
int g_gemm_a_tmem = 1;
int g_gemm_stack_n = 1;
int g_tf32_raw_hi = 1;
int g_fused_gather_warps = 8;
int g_fused_fetch_warps = 8;
int g_fused_l2_prefetch = 0;
int g_fused_ring = 0;
int g_fused_tc_tail = 1;    // tower-tail layers of the one-kernel forward on tcgen05 (0 = fp32 CUDA-core tail)
int g_tower_bwd_tc = 0;     // opt-in until measured on hardware
int g_cin_tc = 1;
int g_rows_zero_blocks = 0;
int g_autoint_vec = 1;      // float4 lane I/O: bit-identical, 3.60 -> 3.18 ms per config-4 step (BENCH_r01 experiments.safe.autoint_vec)
int g_l2_persist = 0;       // opt-in until measured on hardware
size_t g_l2_aside = 0, g_l2_max_window = 0;

// Grow-only per-device scratch buffers (slot = call site).  Kernels of one stream that share a slot are ordered by
// the stream, so reuse is safe for the single-stream execution model of the reference's training loop.  Growth uses
// cudaMalloc and therefore must not happen inside CUDA-graph capture: run one eager warm-up step first.
void* workspace(int slot, size_t bytes, int* err) {
    constexpr int kMaxDev = 16, kSlots = 12;
    static void* ptr[kMaxDev][kSlots] = {};
    static size_t cap[kMaxDev][kSlots] = {};
    int dev = 0;
    *err = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev >= kMaxDev || slot >= kSlots) { *err = e != cudaSuccess ? (int)e : RPB_ERR_UNSUPPORTED; return nullptr; }
    if (cap[dev][slot] < bytes) {
        if (ptr[dev][slot] != nullptr) {
            cudaDeviceSynchronize();           // previous users of the old buffer must be done before it is freed
            cudaFree(ptr[dev][slot]);
            ptr[dev][slot] = nullptr;
            cap[dev][slot] = 0;
        }
        const size_t want = bytes + bytes / 4 + 4096;
        e = cudaMalloc(&ptr[dev][slot], want);
        if (e != cudaSuccess) { *err = (int)e; return nullptr; }
        cap[dev][slot] = want;
    }
    return ptr[dev][slot];
}

}  // namespace rpb

RPB_API int rpb_version(void) { return 8; }

RPB_API const char* rpb_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case RPB_ERR_UNSUPPORTED: return "shape outside what the sm_100a kernels support";
        case RPB_ERR_BAD_ARG: return "bad argument";
        case RPB_ERR_NO_DRIVER: return "cuTensorMapEncodeTiled not available from the CUDA driver";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

RPB_API int rpb_set_option(const char* name, int64_t value) {
    if (name == nullptr) return RPB_ERR_BAD_ARG;
    const std::string n(name);
    if (n == "gather_load_policy") {
        if (value < 0 || value > 3) return RPB_ERR_BAD_ARG;   // 3 = measurement-only variant without the x write
        rpb::g_gather_policy = (int)value;
        return 0;
    }
    if (n == "gemm_v2") { rpb::g_gemm_v2 = value != 0; return 0; }
    if (n == "gemm_a_tmem") { rpb::g_gemm_a_tmem = value != 0; return 0; }
    if (n == "gemm_stack_n") { rpb::g_gemm_stack_n = value != 0; return 0; }
    if (n == "tf32_raw_hi") { rpb::g_tf32_raw_hi = value != 0; return 0; }
    if (n == "l2_persist") {                    // enable / disable the persisting L2 set-aside of the current device (not during capture)
        int dev = 0, aside = 0, win = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&aside, cudaDevAttrMaxPersistingL2CacheSize, dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        if (e != cudaSuccess) return (int)e;
        if (value != 0 && (aside <= 0 || win <= 0)) return RPB_ERR_UNSUPPORTED;
        e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, value != 0 ? (size_t)aside : 0);
        if (e != cudaSuccess) return (int)e;
        if (value == 0) cudaCtxResetPersistingL2Cache();
        rpb::g_l2_aside = value != 0 ? (size_t)aside : 0;
        rpb::g_l2_max_window = value != 0 ? (size_t)win : 0;
        rpb::g_l2_persist = value != 0;
        return 0;
    }
    if (n == "rows_zero_blocks") { if (value < 0 || value > 65535) return RPB_ERR_BAD_ARG; rpb::g_rows_zero_blocks = (int)value; return 0; }
    if (n == "cin_tc") { rpb::g_cin_tc = value != 0; return 0; }
    if (n == "autoint_vec") { rpb::g_autoint_vec = value != 0; return 0; }
    if (n == "tower_bwd_tc") { rpb::g_tower_bwd_tc = value != 0; return 0; }
    if (n == "fused_gather_warps") { if (value != 4 && value != 8) return RPB_ERR_BAD_ARG; rpb::g_fused_gather_warps = (int)value; return 0; }
    if (n == "fused_l2_prefetch") { if (value < 0 || value > 8) return RPB_ERR_BAD_ARG; rpb::g_fused_l2_prefetch = (int)value; return 0; }
    if (n == "fused_fetch_warps") { if (value != 0 && value != 4 && value != 8 && value != 16) return RPB_ERR_BAD_ARG; rpb::g_fused_fetch_warps = (int)value; return 0; }
    if (n == "fused_ring") { if (value != 0 && (value < 3 || value > 6)) return RPB_ERR_BAD_ARG; rpb::g_fused_ring = (int)value; return 0; }
    if (n == "fused_tc_tail") { rpb::g_fused_tc_tail = value != 0; return 0; }
    if (n == "wgrad_tc") { rpb::g_wgrad_tc = value != 0; return 0; }
    if (n == "scatter_reverse") { rpb::g_scatter_reverse = value != 0; return 0; }
    if (n == "wgrad_stages") { if (value < 2 || value > 4) return RPB_ERR_BAD_ARG; rpb::g_wgrad_stages = (int)value; return 0; }
    if (n == "gather_kernel") {
        if (value < 0 || value > 1) return RPB_ERR_BAD_ARG;
        rpb::g_gather_kernel = (int)value;
        return 0;
    }
    if (n == "l2_fetch_granularity") {          // bytes: 32, 64 or 128 (cudaLimitMaxL2FetchGranularity, device-wide hint)
        return (int)cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value);
    }
    return RPB_ERR_BAD_ARG;
}
