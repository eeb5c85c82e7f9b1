// Multi-table embedding gather / scatter-add for sm_100a.
//
// Forward replaces F x aten::embedding + stack + the dense-feature stack + the FM second-order reduction
// + LR_Layer's D=1 gather (reference: models/layers/embedding.py:49-63, models/utils.py:122-137,
// models/layers/interaction.py:36-44, models/layers/shallow.py:22-26) with ONE launch.
//
// Mapping: a group of LPR lanes owns one sample; lane l of the group owns VEC consecutive floats of every
// embedding row (16 B per lane when D % 4 == 0), so a row load is one fully-used 16 B x (D/4) request and
// the FM sums (sum_f e, sum_f e^2) live in that lane's registers for the whole field loop; only the final
// sum over d crosses lanes (log2(LPR) shuffles).  Fields are processed U at a time: U index loads, then U
// independent row loads in flight per lane (HBM random-row latency is hidden by MLP, not occupancy).
#include <type_traits>

#include "common.cuh"

namespace rpb {

struct GatherFwdParams {
    const float* tables[RPB_MAX_FIELDS];
    const long long* idx[RPB_MAX_FIELDS];
    long long rows[RPB_MAX_FIELDS];
    const float* lr_tables[RPB_MAX_FIELDS];
    const float* dense[RPB_MAX_DENSE];
    float* x;
    float* fm;
    float* fm_s;
    float* lr_in;
    long long* err;
    const float* const* shard_tab;     // device array [F*G] of shard base pointers (row-sharded tables), or nullptr
    int B, F, D, Nd, ldx, ld_lr, G;
};

// Row-sharded tables (owner = id mod G, local row = id div G; oracle/index_routing.py::shard_route): the shard of
// another GPU is read directly through its NVLink peer mapping — the lookup "all-to-all" is fused into the gather.
__device__ __forceinline__ const float* table_row(const GatherFwdParams& p, int f, long long ix, int D) {
    if (p.G > 1) {
        const unsigned u = (unsigned)ix, g = (unsigned)p.G;
        const float* base = reinterpret_cast<const float*>(__ldg(reinterpret_cast<const unsigned long long*>(p.shard_tab) + (size_t)f * g + (u % g)));
        return base + (size_t)(u / g) * D;
    }
    return p.tables[f] + (size_t)ix * D;
}

template <int VEC> struct Vec;
template <> struct Vec<4> {
    float4 v;
    __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ void load_stream(const float* p) { v = ldg_f4_stream(p); }
    template <int POLICY> __device__ __forceinline__ void load_row(const float* p) {
        if constexpr (POLICY == 0) v = ldg_f4_stream(p);
        else if constexpr (POLICY == 1 || POLICY == 3) v = ldg_f4_stream64(p);
        else v = ldg_f4(p);
    }
    __device__ __forceinline__ void load(const float* p) { v = ldg_f4(p); }
    __device__ __forceinline__ void store(float* p) const { stg_f4(p, v); }
    __device__ __forceinline__ void red(float* p) const { red_add_f4(p, v); }
    __device__ __forceinline__ void add(const Vec& o) { v.x += o.v.x; v.y += o.v.y; v.z += o.v.z; v.w += o.v.w; }
    __device__ __forceinline__ void add_sq(const Vec& o) {
        v.x = fmaf(o.v.x, o.v.x, v.x); v.y = fmaf(o.v.y, o.v.y, v.y);
        v.z = fmaf(o.v.z, o.v.z, v.z); v.w = fmaf(o.v.w, o.v.w, v.w);
    }
    // this = a + c * (s - e)
    __device__ __forceinline__ void fm_grad(const Vec& a, float c, const Vec& s, const Vec& e) {
        v.x = fmaf(c, s.v.x - e.v.x, a.v.x); v.y = fmaf(c, s.v.y - e.v.y, a.v.y);
        v.z = fmaf(c, s.v.z - e.v.z, a.v.z); v.w = fmaf(c, s.v.w - e.v.w, a.v.w);
    }
    __device__ __forceinline__ float sq_minus(const Vec& q) const {   // sum_c (s_c^2 - q_c)
        return (v.x * v.x - q.v.x) + (v.y * v.y - q.v.y) + (v.z * v.z - q.v.z) + (v.w * v.w - q.v.w);
    }
};
template <> struct Vec<1> {
    float v;
    __device__ __forceinline__ void zero() { v = 0.f; }
    __device__ __forceinline__ void load_stream(const float* p) { v = __ldg(p); }
    template <int POLICY> __device__ __forceinline__ void load_row(const float* p) { v = __ldg(p); }
    __device__ __forceinline__ void load(const float* p) { v = __ldg(p); }
    __device__ __forceinline__ void store(float* p) const { *p = v; }
    __device__ __forceinline__ void red(float* p) const { red_add_f1(p, v); }
    __device__ __forceinline__ void add(const Vec& o) { v += o.v; }
    __device__ __forceinline__ void add_sq(const Vec& o) { v = fmaf(o.v, o.v, v); }
    __device__ __forceinline__ void fm_grad(const Vec& a, float c, const Vec& s, const Vec& e) {
        v = fmaf(c, s.v - e.v, a.v);
    }
    __device__ __forceinline__ float sq_minus(const Vec& q) const { return v * v - q.v; }
};

__device__ __forceinline__ void report_bad_index(long long* err, int f, int b, long long ix) {
    if (err != nullptr) {
        unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(err), 0ull, 1ull);
        if (old == 0ull) { err[1] = f; err[2] = b; err[3] = ix; __threadfence_system(); }
    }
}

template <int VEC, int LPR, int U, bool HAS_LR, int POLICY>
__global__ void __launch_bounds__(256)
gather_fwd_kernel(const __grid_constant__ GatherFwdParams p) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b_raw = (int)(gt / LPR);
    const int l = (int)(gt % LPR);
    const bool valid = b_raw < p.B;
    const int b = valid ? b_raw : p.B - 1;
    const int DV = p.D / VEC;
    const bool lane_on = l < DV;
    float* __restrict__ xrow = p.x + (size_t)b * p.ldx;

    Vec<VEC> s, q;
    s.zero(); q.zero();

    for (int f0 = 0; f0 < p.F; f0 += U) {
        long long ix[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int f = f0 + u;
            ix[u] = 0;
            if (f < p.F) {
                long long v = __ldg(p.idx[f] + b);
                if ((unsigned long long)v >= (unsigned long long)p.rows[f]) {
                    if (valid && l == 0) report_bad_index(p.err, f, b, v);
                    v = 0;
                }
                ix[u] = v;
            }
        }
        Vec<VEC> e[U];
        float lrv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int f = f0 + u;
            e[u].zero();
            lrv[u] = 0.f;
            if (f < p.F) {
                if (lane_on) e[u].template load_row<POLICY>(table_row(p, f, ix[u], p.D) + l * VEC);
                if (HAS_LR && (f % LPR) == l) lrv[u] = __ldg(p.lr_tables[f] + ix[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int f = f0 + u;
            if (f < p.F) {
                if (lane_on && valid && POLICY != 3 && p.x != nullptr) e[u].store(xrow + f * p.D + l * VEC);   // x == NULL: FM-only consumers
                s.add(e[u]);
                q.add_sq(e[u]);
                if (HAS_LR && valid && (f % LPR) == l) p.lr_in[(size_t)b * p.ld_lr + f] = lrv[u];
            }
        }
    }

    if (valid) {
        const int FD = p.F * p.D;
        for (int j = l; j < p.Nd; j += LPR) {
            const float dv = __ldg(p.dense[j] + b);
            if (p.x != nullptr) xrow[FD + j] = dv;
            if (HAS_LR) p.lr_in[(size_t)b * p.ld_lr + p.F + j] = dv;
        }
        if (p.x != nullptr) for (int j = FD + p.Nd + l; j < p.ldx; j += LPR) xrow[j] = 0.f;
        if (HAS_LR) for (int j = p.F + p.Nd + l; j < p.ld_lr; j += LPR) p.lr_in[(size_t)b * p.ld_lr + j] = 0.f;
        if (p.fm_s != nullptr && lane_on) s.store(p.fm_s + (size_t)b * p.D + l * VEC);
    }
    if (p.fm != nullptr) {
        float t = lane_on ? s.sq_minus(q) : 0.f;
        t = group_sum<LPR>(t);
        if (valid && l == 0) p.fm[b] = 0.5f * t;
    }
}


// ---------------------------------------------------------------- tile variant (default): staged through shared memory
// CTA = 128 threads = S samples (S = 128/LPR).  (1) the [F x S] index tile is fetched with fully coalesced 8-byte
// loads into shared memory (one exposed latency instead of one per field chunk), (2) every lane then has U independent
// 16-byte row loads in flight per chunk, rows are written into a shared-memory copy of the output tile, (3) the tile —
// S complete feature rows, contiguous in x — leaves with ONE asynchronous bulk copy (cp.async.bulk, the TMA engine),
// so the 113 MB of x writes cost no LSU issue slots and no partial-line traffic.
__device__ __forceinline__ void bulk_store_tile(float* gdst, const float* ssrc, unsigned bytes) {
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(saddr), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <int LPR, int U, bool HAS_LR>
__global__ void __launch_bounds__(128)
gather_fwd_tile_kernel(const __grid_constant__ GatherFwdParams p) {
    constexpr int S = 128 / LPR;
    extern __shared__ __align__(128) unsigned char tile_raw[];
    float* xs = reinterpret_cast<float*>(tile_raw);                         // [S][ldx]
    long long* is = reinterpret_cast<long long*>(tile_raw + (size_t)S * p.ldx * sizeof(float));   // [F][S]
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * S;
    const int ns = min(S, p.B - b0);

    for (int i = tid; i < p.F * S; i += 128) {
        const int f = i / S, s = i % S;
        long long v = 0;
        if (s < ns) {
            v = __ldg(p.idx[f] + b0 + s);
            if ((unsigned long long)v >= (unsigned long long)p.rows[f]) { report_bad_index(p.err, f, b0 + s, v); v = 0; }
        }
        is[i] = v;
    }
    const int FD = p.F * p.D;
    for (int i = tid; i < p.Nd * S; i += 128) {
        const int j = i / S, s = i % S;
        const float dv = (s < ns) ? __ldg(p.dense[j] + b0 + s) : 0.f;
        if (p.x != nullptr) xs[(size_t)s * p.ldx + FD + j] = dv;
        if (HAS_LR && s < ns) p.lr_in[(size_t)(b0 + s) * p.ld_lr + p.F + j] = dv;
    }
    if (p.x != nullptr) {
        for (int i = tid; i < (p.ldx - FD - p.Nd) * S; i += 128) {
            const int j = i / S, s = i % S;
            xs[(size_t)s * p.ldx + FD + p.Nd + j] = 0.f;
        }
    }
    __syncthreads();

    const int s = tid / LPR, l = tid % LPR;
    const bool valid = s < ns;
    const int b = b0 + (valid ? s : 0);
    const bool lane_on = l < p.D / 4;
    float* xrow = xs + (size_t)s * p.ldx;
    Vec<4> sum, sq;
    sum.zero(); sq.zero();
    for (int f0 = 0; f0 < p.F; f0 += U) {
        Vec<4> e[U];
        float lrv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int f = f0 + u;
            e[u].zero();
            lrv[u] = 0.f;
            if (f < p.F) {
                const long long ix = is[f * S + s];
                if (lane_on) e[u].template load_row<1>(table_row(p, f, ix, p.D) + l * 4);
                if (HAS_LR && (f % LPR) == l) lrv[u] = __ldg(p.lr_tables[f] + ix);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int f = f0 + u;
            if (f < p.F) {
                if (lane_on && p.x != nullptr) *reinterpret_cast<float4*>(xrow + f * p.D + l * 4) = e[u].v;
                sum.add(e[u]);
                sq.add_sq(e[u]);
                if (HAS_LR && valid && (f % LPR) == l) p.lr_in[(size_t)b * p.ld_lr + f] = lrv[u];
            }
        }
    }
    if (valid) {
        if (HAS_LR) for (int j = p.F + p.Nd + l; j < p.ld_lr; j += LPR) p.lr_in[(size_t)b * p.ld_lr + j] = 0.f;
        if (p.fm_s != nullptr && lane_on) sum.store(p.fm_s + (size_t)b * p.D + l * 4);
    }
    if (p.fm != nullptr) {
        float t = lane_on ? sum.sq_minus(sq) : 0.f;
        t = group_sum<LPR>(t);
        if (valid && l == 0) p.fm[b] = 0.5f * t;
    }
    // publish the tile: generic-proxy smem writes -> async proxy, then one bulk copy of ns complete rows
    if (p.x == nullptr) return;                 // FM-only consumers: nothing to materialise
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) bulk_store_tile(p.x + (size_t)b0 * p.ldx, xs, (unsigned)((size_t)ns * p.ldx * sizeof(float)));
}

struct ScatterParams {
    float* grads[RPB_MAX_FIELDS];
    float* lr_grads[RPB_MAX_FIELDS];
    const long long* idx[RPB_MAX_FIELDS];
    long long rows[RPB_MAX_FIELDS];
    const float* dx;
    const float* x;
    const float* dfm;
    const float* fm_s;
    const float* dlr_in;
    float* const* grad_shard_tab;      // device array [F*G] of grad-shard base pointers, or nullptr
    int B, F, D, lddx, ldx, ld_dlr, G;
};

__device__ __forceinline__ float* grad_row(const ScatterParams& p, int f, long long ix, int D) {
    if (p.G > 1) {
        const unsigned u = (unsigned)ix, g = (unsigned)p.G;
        float* base = reinterpret_cast<float*>(__ldg(reinterpret_cast<const unsigned long long*>(p.grad_shard_tab) + (size_t)f * g + (u % g)));
        return base + (size_t)(u / g) * D;
    }
    return p.grads[f] != nullptr ? p.grads[f] + (size_t)ix * D : nullptr;
}

template <int VEC, int LPR, int U>
__global__ void __launch_bounds__(256)
gather_bwd_kernel(const __grid_constant__ ScatterParams p) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = (int)(gt / LPR);
    const int l = (int)(gt % LPR);
    if (b >= p.B) return;                      // no warp-collective below
    const int DV = p.D / VEC;
    const bool lane_on = l < DV;
    const bool has_fm = p.dfm != nullptr;
    const bool has_dx = p.dx != nullptr;

    Vec<VEC> s;
    s.zero();
    float c = 0.f;
    if (has_fm) {
        c = __ldg(p.dfm + b);
        if (lane_on) s.load(p.fm_s + (size_t)b * p.D + l * VEC);
    }
    const float* __restrict__ dxrow = has_dx ? p.dx + (size_t)b * p.lddx : nullptr;
    const float* __restrict__ xrow = has_fm ? p.x + (size_t)b * p.ldx : nullptr;

    for (int f0 = 0; f0 < p.F; f0 += U) {
        long long ix[U];
        Vec<VEC> g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int f = f0 + u;
            ix[u] = 0;
            g[u].zero();
            if (f < p.F) {
                long long v = __ldg(p.idx[f] + b);
                if ((unsigned long long)v >= (unsigned long long)p.rows[f]) v = 0;
                ix[u] = v;
                if (lane_on) {
                    Vec<VEC> a, e;
                    a.zero(); e.zero();
                    if (has_dx) a.load_stream(dxrow + f * p.D + l * VEC);
                    if (has_fm) { e.load_stream(xrow + f * p.D + l * VEC); g[u].fm_grad(a, c, s, e); }
                    else g[u] = a;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int f = f0 + u;
            if (f < p.F) {
                if (lane_on) {
                    float* gr = grad_row(p, f, ix[u], p.D);
                    if (gr != nullptr) g[u].red(gr + l * VEC);
                }
                if (p.dlr_in != nullptr && (f % LPR) == l && p.lr_grads[f] != nullptr)
                    red_add_f1(p.lr_grads[f] + ix[u], __ldg(p.dlr_in + (size_t)b * p.ld_dlr + f));
            }
        }
    }
}

// Sparse zero_grad: clear exactly the rows touched by a previous backward (persistent dense-grad buffers).
template <int VEC, int LPR>
__global__ void __launch_bounds__(256)
rows_zero_kernel(const __grid_constant__ ScatterParams p) {
    // grid-stride: the default launch covers every (sample, lane) once; the co-resident launch (option rows_zero_blocks: a few
    // 128-thread blocks per SM, 24 registers) is sized to fit NEXT TO the one-CTA-per-SM forward kernel it overlaps with
    const long long total = (long long)p.B * LPR, stride = (long long)gridDim.x * blockDim.x;
    Vec<VEC> z;
    z.zero();
    for (long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x; gt < total; gt += stride) {
        const int b = (int)(gt / LPR), l = (int)(gt % LPR);
        const bool lane_on = l < p.D / VEC;
        for (int f0 = 0; f0 < p.F; f0 += 8) {                             // eight index loads in flight, then their stores
            long long v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = f0 + j < p.F ? __ldg(p.idx[f0 + j] + b) : 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int f = f0 + j;
                if (f < p.F) {
                    if ((unsigned long long)v[j] >= (unsigned long long)p.rows[f]) v[j] = 0;
                    if (lane_on) {
                        float* gr = grad_row(p, f, v[j], p.D);
                        if (gr != nullptr) z.store(gr + l * VEC);
                    }
                    if ((f % LPR) == l && p.lr_grads[f] != nullptr) p.lr_grads[f][v[j]] = 0.f;
                }
            }
        }
    }
}

// ---------------------------------------------------------------- standalone FM on a [B,F,D] tensor
template <int VEC, int LPR>
__global__ void __launch_bounds__(256)
fm_fwd_kernel(const float* __restrict__ e, long long lde, int B, int F, int D, float* out_sum, float* out_bi) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b_raw = (int)(gt / LPR), l = (int)(gt % LPR);
    const bool valid = b_raw < B;
    const int b = valid ? b_raw : B - 1;
    const bool lane_on = l < D / VEC;
    Vec<VEC> s, q;
    s.zero(); q.zero();
    if (lane_on) {
        const float* row = e + (size_t)b * lde + l * VEC;
#pragma unroll 4
        for (int f = 0; f < F; ++f) { Vec<VEC> v; v.load(row + f * D); s.add(v); q.add_sq(v); }
    }
    if (out_bi != nullptr && valid && lane_on) {
        // 0.5 * (s^2 - q) per component
        Vec<VEC> r;
        if constexpr (VEC == 4) {
            r.v = make_float4(0.5f * (s.v.x * s.v.x - q.v.x), 0.5f * (s.v.y * s.v.y - q.v.y),
                              0.5f * (s.v.z * s.v.z - q.v.z), 0.5f * (s.v.w * s.v.w - q.v.w));
        } else {
            r.v = 0.5f * (s.v * s.v - q.v);
        }
        r.store(out_bi + (size_t)b * D + l * VEC);
    }
    if (out_sum != nullptr) {
        float t = lane_on ? s.sq_minus(q) : 0.f;
        t = group_sum<LPR>(t);
        if (valid && l == 0) out_sum[b] = 0.5f * t;
    }
}

template <int VEC, int LPR>
__global__ void __launch_bounds__(256)
fm_bwd_kernel(const float* __restrict__ e, long long lde, int B, int F, int D, const float* dsum,
              const float* dbi, float* de, long long ldde, int accumulate) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = (int)(gt / LPR), l = (int)(gt % LPR);
    if (b >= B || l >= D / VEC) return;
    const float* row = e + (size_t)b * lde + l * VEC;
    float* drow = de + (size_t)b * ldde + l * VEC;
    Vec<VEC> s;
    s.zero();
    for (int f = 0; f < F; ++f) { Vec<VEC> v; v.load(row + f * D); s.add(v); }
    const float cs = dsum != nullptr ? __ldg(dsum + b) : 0.f;
    Vec<VEC> cb;
    cb.zero();
    if (dbi != nullptr) cb.load(dbi + (size_t)b * D + l * VEC);
    for (int f = 0; f < F; ++f) {
        Vec<VEC> v, g;
        v.load(row + f * D);
        if constexpr (VEC == 4) {
            g.v = make_float4((cs + cb.v.x) * (s.v.x - v.v.x), (cs + cb.v.y) * (s.v.y - v.v.y),
                              (cs + cb.v.z) * (s.v.z - v.v.z), (cs + cb.v.w) * (s.v.w - v.v.w));
        } else {
            g.v = (cs + cb.v) * (s.v - v.v);
        }
        if (accumulate) { Vec<VEC> o; o.load(drow + f * D); g.add(o); }
        g.store(drow + f * D);
    }
}

template <typename F>
static int dispatch_shape(int D, bool aligned16, F&& fn) {
    // fn(vec_tag, lpr_tag)
    if (D % 4 == 0 && aligned16) {
        const int dv = D / 4;
        if (dv <= 1) return fn(std::integral_constant<int, 4>{}, std::integral_constant<int, 1>{});
        if (dv <= 2) return fn(std::integral_constant<int, 4>{}, std::integral_constant<int, 2>{});
        if (dv <= 4) return fn(std::integral_constant<int, 4>{}, std::integral_constant<int, 4>{});
        if (dv <= 8) return fn(std::integral_constant<int, 4>{}, std::integral_constant<int, 8>{});
        if (dv <= 16) return fn(std::integral_constant<int, 4>{}, std::integral_constant<int, 16>{});
        if (dv <= 32) return fn(std::integral_constant<int, 4>{}, std::integral_constant<int, 32>{});
        return RPB_ERR_UNSUPPORTED;
    }
    if (D <= 1) return fn(std::integral_constant<int, 1>{}, std::integral_constant<int, 1>{});
    if (D <= 2) return fn(std::integral_constant<int, 1>{}, std::integral_constant<int, 2>{});
    if (D <= 4) return fn(std::integral_constant<int, 1>{}, std::integral_constant<int, 4>{});
    if (D <= 8) return fn(std::integral_constant<int, 1>{}, std::integral_constant<int, 8>{});
    if (D <= 16) return fn(std::integral_constant<int, 1>{}, std::integral_constant<int, 16>{});
    if (D <= 32) return fn(std::integral_constant<int, 1>{}, std::integral_constant<int, 32>{});
    return RPB_ERR_UNSUPPORTED;
}

static inline bool is_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace rpb

using namespace rpb;

RPB_API int rpb_gather_fwd(const RpbGatherDesc* d, void* stream) {
    if (d == nullptr || d->B <= 0 || d->F <= 0 || d->D <= 0) return RPB_ERR_BAD_ARG;
    if (d->x == nullptr && d->fm == nullptr && d->lr_in == nullptr) return RPB_ERR_BAD_ARG;      // nothing to produce
    if (d->F > RPB_MAX_FIELDS || d->Nd > RPB_MAX_DENSE || d->Nd < 0) return RPB_ERR_UNSUPPORTED;
    if (d->x != nullptr && d->ldx < d->F * d->D + d->Nd) return RPB_ERR_BAD_ARG;
    GatherFwdParams p{};
    bool aligned = (d->x == nullptr || ((d->ldx % 4 == 0) && is_aligned16(d->x))) && (d->fm_s == nullptr || is_aligned16(d->fm_s));
    for (int f = 0; f < d->F; ++f) {
        p.tables[f] = d->tables ? d->tables[f] : nullptr;
        p.idx[f] = reinterpret_cast<const long long*>(d->idx[f]);
        p.rows[f] = d->rows[f];
        p.lr_tables[f] = d->lr_tables ? d->lr_tables[f] : nullptr;
        if (d->G <= 1) aligned = aligned && is_aligned16(d->tables[f]);
        if (d->G > 1 && d->rows[f] > 0xFFFFFFFFll) return RPB_ERR_UNSUPPORTED;
    }
    for (int j = 0; j < d->Nd; ++j) p.dense[j] = d->dense[j];
    p.x = d->x; p.fm = d->fm; p.fm_s = d->fm_s; p.lr_in = d->lr_tables ? d->lr_in : nullptr;
    p.err = reinterpret_cast<long long*>(d->err);
    p.B = d->B; p.F = d->F; p.D = d->D; p.Nd = d->Nd; p.ldx = d->ldx; p.ld_lr = d->ld_lr;
    if (d->x == nullptr) p.ldx = 0;            // no tile rows to stage
    p.G = d->G > 1 ? d->G : 1;
    p.shard_tab = d->shard_tab;
    if (p.G > 1 && (d->shard_tab == nullptr || d->lr_tables != nullptr)) return RPB_ERR_BAD_ARG;
    if (d->lr_tables && (d->lr_in == nullptr || d->ld_lr < d->F + d->Nd)) return RPB_ERR_BAD_ARG;
    const bool has_lr = d->lr_tables != nullptr;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return dispatch_shape(d->D, aligned, [&](auto vec, auto lpr) -> int {
        constexpr int VEC = decltype(vec)::value, LPR = decltype(lpr)::value;
        const int pol = g_gather_policy;
        if constexpr (VEC == 4 && LPR <= 16) {
            constexpr int S = 128 / LPR;
            const size_t smem = (size_t)S * p.ldx * sizeof(float) + (size_t)p.F * S * sizeof(long long);
            if (g_gather_kernel == 0 && pol == 1 && smem <= 100 * 1024) {
                auto launch = [&](auto kern) -> int {
                    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    if (e != cudaSuccess) return (int)e;
                    kern<<<ceil_div(p.B, S), 128, smem, st>>>(p);
                    return (int)cudaGetLastError();
                };
                if (p.F % 13 == 0) return has_lr ? launch(gather_fwd_tile_kernel<LPR, 13, true>) : launch(gather_fwd_tile_kernel<LPR, 13, false>);
                return has_lr ? launch(gather_fwd_tile_kernel<LPR, 8, true>) : launch(gather_fwd_tile_kernel<LPR, 8, false>);
            }
        }
        const int grid = ceil_div((long long)p.B * LPR, 256);
        if (has_lr) {
            if (pol == 0) gather_fwd_kernel<VEC, LPR, 8, true, 0><<<grid, 256, 0, st>>>(p);
            else if (pol == 1) gather_fwd_kernel<VEC, LPR, 8, true, 1><<<grid, 256, 0, st>>>(p);
            else gather_fwd_kernel<VEC, LPR, 8, true, 2><<<grid, 256, 0, st>>>(p);
        } else {
            if (pol == 0) gather_fwd_kernel<VEC, LPR, 8, false, 0><<<grid, 256, 0, st>>>(p);
            else if (pol == 1) gather_fwd_kernel<VEC, LPR, 8, false, 1><<<grid, 256, 0, st>>>(p);
            else if (pol == 2) gather_fwd_kernel<VEC, LPR, 8, false, 2><<<grid, 256, 0, st>>>(p);
            else gather_fwd_kernel<VEC, LPR, 8, false, 3><<<grid, 256, 0, st>>>(p);
        }
        RPB_LAUNCH_CHECK();
        return 0;
    });
}

RPB_API int rpb_gather_bwd(const RpbScatterDesc* d, void* stream) {
    if (d == nullptr || d->B <= 0 || d->F <= 0 || d->D <= 0) return RPB_ERR_BAD_ARG;
    if (d->F > RPB_MAX_FIELDS) return RPB_ERR_UNSUPPORTED;
    if (d->dfm != nullptr && (d->x == nullptr || d->fm_s == nullptr)) return RPB_ERR_BAD_ARG;
    ScatterParams p{};
    bool aligned = true;
    if (d->dx) aligned = aligned && is_aligned16(d->dx) && d->lddx % 4 == 0;
    if (d->dfm) aligned = aligned && is_aligned16(d->x) && d->ldx % 4 == 0 && is_aligned16(d->fm_s);
    for (int f = 0; f < d->F; ++f) {
        p.grads[f] = d->grads ? d->grads[f] : nullptr;
        p.lr_grads[f] = d->lr_grads ? d->lr_grads[f] : nullptr;
        p.idx[f] = reinterpret_cast<const long long*>(d->idx[f]);
        p.rows[f] = d->rows[f];
        if (p.grads[f]) aligned = aligned && is_aligned16(p.grads[f]);
    }
    p.dx = d->dx; p.x = d->x; p.dfm = d->dfm; p.fm_s = d->fm_s;
    p.dlr_in = d->lr_grads ? d->dlr_in : nullptr;
    p.B = d->B; p.F = d->F; p.D = d->D; p.lddx = d->lddx; p.ldx = d->ldx; p.ld_dlr = d->ld_dlr;
    p.G = d->G > 1 ? d->G : 1;
    p.grad_shard_tab = d->grad_shard_tab;
    if (p.G > 1 && d->grad_shard_tab == nullptr) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return dispatch_shape(d->D, aligned, [&](auto vec, auto lpr) -> int {
        constexpr int VEC = decltype(vec)::value, LPR = decltype(lpr)::value;
        const int grid = ceil_div((long long)p.B * LPR, 256);
        gather_bwd_kernel<VEC, LPR, 8><<<grid, 256, 0, st>>>(p);
        RPB_LAUNCH_CHECK();
        return 0;
    });
}

RPB_API int rpb_rows_zero(const RpbScatterDesc* d, void* stream) {
    if (d == nullptr || d->B <= 0 || d->F <= 0 || d->D <= 0) return RPB_ERR_BAD_ARG;
    if (d->F > RPB_MAX_FIELDS) return RPB_ERR_UNSUPPORTED;
    ScatterParams p{};
    bool aligned = true;
    for (int f = 0; f < d->F; ++f) {
        p.grads[f] = d->grads ? d->grads[f] : nullptr;
        p.lr_grads[f] = d->lr_grads ? d->lr_grads[f] : nullptr;
        p.idx[f] = reinterpret_cast<const long long*>(d->idx[f]);
        p.rows[f] = d->rows[f];
        if (p.grads[f]) aligned = aligned && is_aligned16(p.grads[f]);
    }
    p.B = d->B; p.F = d->F; p.D = d->D;
    p.G = d->G > 1 ? d->G : 1;
    p.grad_shard_tab = d->grad_shard_tab;
    if (p.G > 1 && d->grad_shard_tab == nullptr) return RPB_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return dispatch_shape(d->D, aligned, [&](auto vec, auto lpr) -> int {
        constexpr int VEC = decltype(vec)::value, LPR = decltype(lpr)::value;
        if (g_rows_zero_blocks > 0) {
            // same L1 / shared-memory split as the forward kernel it is meant to run next to (an SM is not re-configured while busy)
            cudaFuncSetAttribute(rows_zero_kernel<VEC, LPR>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            rows_zero_kernel<VEC, LPR><<<g_rows_zero_blocks, 128, 0, st>>>(p);
        }
        else rows_zero_kernel<VEC, LPR><<<ceil_div((long long)p.B * LPR, 256), 256, 0, st>>>(p);
        RPB_LAUNCH_CHECK();
        return 0;
    });
}

RPB_API int rpb_fm_fwd(const float* e, int64_t lde, int B, int F, int D, float* out_sum, float* out_bi,
                       void* stream) {
    if (e == nullptr || B <= 0 || F <= 0 || D <= 0) return RPB_ERR_BAD_ARG;
    const bool aligned = is_aligned16(e) && lde % 4 == 0 && (out_bi == nullptr || is_aligned16(out_bi));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return dispatch_shape(D, aligned, [&](auto vec, auto lpr) -> int {
        constexpr int VEC = decltype(vec)::value, LPR = decltype(lpr)::value;
        const int grid = ceil_div((long long)B * LPR, 256);
        fm_fwd_kernel<VEC, LPR><<<grid, 256, 0, st>>>(e, lde, B, F, D, out_sum, out_bi);
        RPB_LAUNCH_CHECK();
        return 0;
    });
}

RPB_API int rpb_fm_bwd(const float* e, int64_t lde, int B, int F, int D, const float* dsum, const float* dbi,
                       float* de, int64_t ldde, int accumulate, void* stream) {
    if (e == nullptr || de == nullptr || B <= 0 || F <= 0 || D <= 0) return RPB_ERR_BAD_ARG;
    const bool aligned = is_aligned16(e) && lde % 4 == 0 && is_aligned16(de) && ldde % 4 == 0 &&
                         (dbi == nullptr || is_aligned16(dbi));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return dispatch_shape(D, aligned, [&](auto vec, auto lpr) -> int {
        constexpr int VEC = decltype(vec)::value, LPR = decltype(lpr)::value;
        const int grid = ceil_div((long long)B * LPR, 256);
        fm_bwd_kernel<VEC, LPR><<<grid, 256, 0, st>>>(e, lde, B, F, D, dsum, dbi, de, ldde, accumulate);
        RPB_LAUNCH_CHECK();
        return 0;
    });
}
