// Device-side synthetic inputs for the BASELINE.json configurations.  Counter based:
//   u(seed, col, i) = splitmix64(seed + 0x9E3779B97F4A7C15 * (i + (col << 40)))
// so any sub-range can be regenerated on the CPU (tests/synth.py mirrors this file) and every
// range-partitioned shard generates exactly its slice of the global data set.
#include "../../include/bowgpu.h"
#include "kernels.h"

namespace bowgpu {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ uint64_t synth_u(uint64_t seed, uint64_t col, uint64_t i) {
    return splitmix64(seed + 0x9E3779B97F4A7C15ull * (i + (col << 40)));
}

namespace {

// time: t[i] = t0 + (row0 + i) * step
__global__ void gen_time_regular(int64_t *t, int64_t n, int64_t row0, int64_t t0, int64_t step) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        t[i] = t0 + (row0 + i) * step;
}

// value column c (1-based column id `col`): float64 (u >> 11) * 2^-53 in [0,1) or int64 u % 2^20;
// validity word built with ballot (32 rows per warp pass): null iff u(seed+1, col, i) % null_mod == 0
__global__ void gen_values(uint64_t *v, uint32_t *validity, int64_t n, int64_t row0, uint64_t seed, uint64_t col,
                           int is_int, uint32_t null_mod) {
    const int64_t n32 = (n + 31) & ~(int64_t)31;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += (int64_t)gridDim.x * blockDim.x) {
        bool valid = false;
        if (i < n) {
            const uint64_t gi = (uint64_t)(row0 + i);
            const uint64_t u = synth_u(seed, col, gi);
            v[i] = is_int ? (u & 0xFFFFFull) : f64_as_bits((double)(u >> 11) * 0x1.0p-53);
            valid = true;
            if (validity && null_mod) valid = (synth_u(seed + 1, col, gi) % null_mod) != 0;
        }
        if (validity) {
            const uint32_t ball = __ballot_sync(0xffffffffu, valid);
            if ((threadIdx.x & 31) == 0) validity[i >> 5] = ball;
        }
    }
}

// ---- BURSTY (BASELINE configs[3]: window sizes 0 .. ~1e6 rows, load-balance stress) -------------------------
// Window k of the lattice S_k = t0 + k*I holds c_k rows, c_k drawn from u(seed, 0, k):
//   50 %: 0 rows;  45 %: 1 + (u >> 8) % 100 rows;  5 %: (1000 + (u >> 8) % 1000) << ((u >> 32) % 10) rows
// (integer-only so that the numpy mirror is bit-exact).  Row j of the window sits at
//   S_k + floor(j * I / c_k)            when bit 40 of u is clear (the window has a row exactly at S_k)
//   S_k + floor((2j + 1) * I / (2 c_k)) otherwise (no row at S_k: Interpolate must insert one).
__host__ __device__ __forceinline__ int64_t bursty_count(uint64_t u) {
    const uint64_t r = u % 100;
    if (r < 50) return 0;
    if (r < 95) return 1 + (int64_t)((u >> 8) % 100);
    return (int64_t)((1000 + (u >> 8) % 1000) << ((u >> 32) % 10));
}

__global__ void gen_bursty_counts(int64_t *cnt, int64_t nw, uint64_t seed) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nw; k += (int64_t)gridDim.x * blockDim.x)
        cnt[k] = bursty_count(synth_u(seed, 0, (uint64_t)k));
}

// off = exclusive scan of the counts (off[nw] = total rows); row gi lies in the window k with off[k] <= gi < off[k+1]
__global__ void gen_time_bursty(int64_t *t, int64_t n, int64_t row0, int64_t t0, int64_t interval, uint64_t seed,
                                const int64_t *off, int64_t nw) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gi = row0 + i;
        int64_t lo = 0, hi = nw;  // largest k with off[k] <= gi
        while (hi - lo > 1) {
            const int64_t mid = lo + ((hi - lo) >> 1);
            if (off[mid] <= gi)
                lo = mid;
            else
                hi = mid;
        }
        const int64_t k = lo, c = off[k + 1] - off[k], j = gi - off[k];
        const bool shifted = (synth_u(seed, 0, (uint64_t)k) >> 40) & 1u;
        const int64_t dt = shifted ? ((2 * j + 1) * interval) / (2 * c) : (j * interval) / c;
        t[i] = t0 + k * interval + dt;
    }
}

}  // namespace

int launch_gen_bursty_counts(int64_t *cnt, int64_t nw, uint64_t seed, cudaStream_t stream) {
    if (nw == 0) return 0;
    gen_bursty_counts<<<592, 256, 0, stream>>>(cnt, nw, seed);
    return (int)cudaGetLastError();
}

int launch_gen_bursty_time(int64_t *time, int64_t n, int64_t row0, int64_t t0, int64_t interval, uint64_t seed,
                           const int64_t *off, int64_t nw, cudaStream_t stream) {
    if (n == 0) return 0;
    gen_time_bursty<<<1184, 256, 0, stream>>>(time, n, row0, t0, interval, seed, off, nw);
    return (int)cudaGetLastError();
}

int launch_gen_regular(int64_t *time, int64_t n, int64_t row0, int64_t t0, int64_t step, cudaStream_t stream) {
    if (n == 0) return 0;
    gen_time_regular<<<1184, 256, 0, stream>>>(time, n, row0, t0, step);
    return (int)cudaGetLastError();
}

int launch_gen_values(uint64_t *v, uint8_t *validity, int64_t n, int64_t row0, uint64_t seed, uint64_t col, int is_int,
                      uint32_t null_mod, cudaStream_t stream) {
    if (n == 0) return 0;
    gen_values<<<1184, 256, 0, stream>>>(v, (uint32_t *)validity, n, row0, seed, col, is_int, null_mod);
    return (int)cudaGetLastError();
}

}  // namespace bowgpu
