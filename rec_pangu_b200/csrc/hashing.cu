// Hashed-id encoder for tables too large for a sorted-unique vocabulary (BASELINE.json config 5: 100 M-row hashed tables).
// The reference encodes ids with a per-column dict built from the data (dataset/base_dataset.py:57-61,92: sorted unique
// values -> 1..N, unseen -> vocab_size); at 1e8 rows that map is replaced by row = splitmix64(raw) mod V, the integer
// contract restated in oracle/index_routing.py::hash_to_row and checked bit-exactly.  Row V remains the OOV slot (never
// produced by the hash).  Pure 64-bit integer work, one 8-byte load and one 8-byte store per id: HBM bound.
#include "common.cuh"

namespace rpb {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// two ids per thread (16-byte accesses) when the pointers allow it
__global__ void __launch_bounds__(256)
hash_rows_kernel(const long long* __restrict__ raw, long long* __restrict__ out, long long n, unsigned long long vocab,
                 int vec2) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec2) {
        const long long i = 2 * t;
        if (i + 1 < n) {
            const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(raw) + t);
            longlong2 r;
            r.x = (long long)(splitmix64((unsigned long long)v.x) % vocab);
            r.y = (long long)(splitmix64((unsigned long long)v.y) % vocab);
            reinterpret_cast<longlong2*>(out)[t] = r;
        } else if (i < n) {
            out[i] = (long long)(splitmix64((unsigned long long)__ldg(raw + i)) % vocab);
        }
    } else if (t < n) {
        out[t] = (long long)(splitmix64((unsigned long long)__ldg(raw + t)) % vocab);
    }
}

}  // namespace rpb

using namespace rpb;

RPB_API int rpb_hash_to_row(const int64_t* raw, int64_t* out, int64_t n, int64_t vocab_size, void* stream) {
    if (raw == nullptr || out == nullptr || n <= 0 || vocab_size <= 0) return RPB_ERR_BAD_ARG;
    const int vec2 = ((reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0 ? 1 : 0;
    const long long threads = vec2 ? (n + 1) / 2 : n;
    hash_rows_kernel<<<ceil_div(threads, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long*>(raw), reinterpret_cast<long long*>(out), n, (unsigned long long)vocab_size, vec2);
    RPB_LAUNCH_CHECK();
    return 0;
}
