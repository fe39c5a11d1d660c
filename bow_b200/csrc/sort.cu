// sort — Bow.SortByCol (bowsort.go:10-47): rows reordered by ascending values of one nil-free column.
//
// The reference sorts (value, row index) pairs with sort.Sort over Buffer.Less (`<` on int64 / float64,
// bowbuffer.go:126-139) and gathers every other column through the permuted indices (bowsort.go:27-41).  sort.Sort is
// not stable: the order among equal keys is whatever Go's algorithm leaves behind.  This build sorts STABLY (equal keys
// keep their input order) — one of the outcomes the reference's contract allows, the one all of its own golden vectors
// show (bowsort_test.go:133-157) and the only one that is fully determined when keys are unique.
//
// Least-significant-digit radix sort, 8-bit digits, one streaming pass per digit that is not constant over the column:
//   sort_prepare_kernel   keys -> order-preserving unsigned keys, histograms of all eight digits, sort.IsSorted and NaN flags
//   sort_pass_kernel      one digit: tile-local stable ranks (warp match + per-warp digit counters), tile offsets by
//                         decoupled look-back over per-tile digit counts (tiles take tickets in start order), scatter
//   sort_gather_kernel    every column of the frame through the final permutation, validity bits rebuilt with ballots
#include "../../include/bowgpu.h"
#include "common.cuh"
#include "kernels.h"

namespace bowgpu {

namespace {

constexpr int SORT_NT = 256, SORT_NW = SORT_NT / 32;
#ifndef SORT_CFG_ITEMS
#define SORT_CFG_ITEMS 12
#endif
#ifndef SORT_CFG_MINB
#define SORT_CFG_MINB 3
#endif
constexpr int SORT_ITEMS = SORT_CFG_ITEMS;
constexpr int SORT_TILE = SORT_NT * SORT_ITEMS;
constexpr uint64_t SORT_AGG = 1ull << 62, SORT_INCL = 1ull << 63, SORT_VAL = SORT_AGG - 1;
constexpr uint64_t SIGN = 0x8000000000000000ull;

// order-preserving map onto unsigned keys; float64: -0.0 joins +0.0 (equal under `<`), negative values reverse
__device__ __forceinline__ uint64_t sort_key(uint64_t raw, bool is_int) {
    if (is_int) return raw ^ SIGN;
    if (raw == SIGN) raw = 0;
    return (raw & SIGN) ? ~raw : raw | SIGN;
}
__device__ __forceinline__ bool sort_less(uint64_t a, uint64_t b, bool is_int) {  // Buffer.Less, bowbuffer.go:126-131
    return is_int ? (int64_t)a < (int64_t)b : bits_as_f64(a) < bits_as_f64(b);
}

// hist[8][256] (global, zeroed), flags: bit 0 = some row is less than its predecessor (not sort.IsSorted), bit 1 = NaN key
__global__ void __launch_bounds__(SORT_NT) sort_prepare_kernel(const uint64_t *values, const int is_int, const int64_t n,
                                                               uint64_t *keys, unsigned long long *hist, int32_t *flags) {
    __shared__ uint32_t h[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += SORT_NT) (&h[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int32_t f = 0;
    // whole warps walk the column so that the warp-wide digit test below sees 32 consecutive rows
    const int64_t nwarps = (int64_t)gridDim.x * SORT_NW, wid = (int64_t)blockIdx.x * SORT_NW + (threadIdx.x >> 5);
    for (int64_t base = wid * 32; base < n; base += nwarps * 32) {
        const int64_t i = base + lane;
        const bool in = i < n;
        const uint64_t raw = in ? values[i] : 0;
        if (in) {
            if (i > 0 && sort_less(raw, values[i - 1], is_int)) f |= 1;
            if (!is_int && (raw & ~SIGN) > 0x7ff0000000000000ull) f |= 2;
        }
        const uint64_t key = sort_key(raw, is_int);
        if (in) keys[i] = key;
        const uint32_t act = __ballot_sync(0xffffffffu, in);
        const uint64_t key0 = __shfl_sync(0xffffffffu, key, 0);
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const uint32_t d = (uint32_t)(key >> (8 * p)) & 255u, d0 = (uint32_t)(key0 >> (8 * p)) & 255u;
            if (__all_sync(0xffffffffu, !in || d == d0)) {  // a digit shared by the whole warp: one add
                if (lane == 0) atomicAdd(&h[p][d0], (uint32_t)__popc(act));
            } else if (in) {
                atomicAdd(&h[p][d], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * 256; i += SORT_NT) {
        const uint32_t c = (&h[0][0])[i];
        if (c) atomicAdd(&hist[i], (unsigned long long)c);
    }
    f = __reduce_or_sync(0xffffffffu, f);
    if (lane == 0 && f) atomicOr(flags, f);
}

// One digit.  kin/vin -> kout/vout; vin == null means the identity permutation (first pass).  digit_base[256] is the
// exclusive prefix of this digit's histogram; status[ntiles][256] (zeroed) carries the per-tile digit counts.
// The tile is put in digit order in shared memory first, so that a warp's stores cover a few contiguous runs of the
// output instead of 32 different buckets (and pages).
__global__ void __launch_bounds__(SORT_NT, SORT_CFG_MINB) sort_pass_kernel(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin,
                                                            uint64_t *__restrict__ kout, uint32_t *__restrict__ vout,
                                                            const int64_t n, const int shift,
                                                            const unsigned long long *__restrict__ digit_base,
                                                            volatile unsigned long long *status, uint32_t *ticket) {
    __shared__ uint32_t cnt[SORT_NW][257];  // (slot 256: rows beyond the end of the column)
    __shared__ uint64_t gadj[256];          // output position of tile slot s holding digit d = gadj[d] + s
    __shared__ uint32_t tile_off[256];      // first tile slot of digit d
    __shared__ uint32_t wsum[SORT_NW];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t xch[SORT_TILE];     // the tile in digit order: keys, then the permutation entries
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < SORT_NW * 257; i += SORT_NT) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t base = tile * SORT_TILE + (int64_t)warp * (32 * SORT_ITEMS) + lane;
    const int64_t left = n - tile * SORT_TILE;
    const int nvalid = left < SORT_TILE ? (int)left : SORT_TILE;

    // keys and the permutation entries that travel with them: all loads of the tile in flight before the ranking starts
    uint64_t key[SORT_ITEMS];
    uint32_t val[SORT_ITEMS];
    uint32_t rank[SORT_ITEMS];  // the warp peers of the row (same digit), then its rank among the tile's rows of that digit, then its slot
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const int64_t idx = base + i * 32;
        key[i] = idx < n ? kin[idx] : ~0ull;
        val[i] = vin ? (idx < n ? vin[idx] : 0u) : (uint32_t)idx;
    }
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {  // (independent of each other: the matches pipeline)
        const int64_t idx = base + i * 32;
        const uint32_t d = idx < n ? (uint32_t)(key[i] >> shift) & 255u : 256u;
        rank[i] = __match_any_sync(0xffffffffu, d);
    }
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const int64_t idx = base + i * 32;
        const uint32_t d = idx < n ? (uint32_t)(key[i] >> shift) & 255u : 256u;
        const uint32_t peers = rank[i];
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader) {
            old = cnt[warp][d];
            cnt[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[i] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();

    // thread t owns digit t: exclusive scan over the warps (-> cnt), over the digits (-> tile_off), publish the tile's count
    const int t = tid;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < SORT_NW; ++w) {
        const uint32_t c = cnt[w][t];
        cnt[w][t] = run;
        run += c;
    }
    volatile unsigned long long *mine = status + tile * 256 + t;
    if (tile > 0) *mine = SORT_AGG | run;
    uint32_t incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t toff = incl - run;
#pragma unroll
    for (int w = 0; w < SORT_NW; ++w)
        if (w < warp) toff += wsum[w];
    tile_off[t] = toff;
    __syncthreads();

#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const int64_t idx = base + i * 32;
        if (idx < n) {
            const uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
            rank[i] += tile_off[d] + cnt[warp][d];
            xch[rank[i]] = key[i];
        }
    }

    {   // the tile's place among the tiles: decoupled look-back over the counts of digit t
        uint64_t excl = 0;
        if (tile > 0) {
            for (int64_t p = tile - 1;; --p) {
                volatile unsigned long long *q = status + p * 256 + t;
                unsigned long long sv;
                while ((sv = *q) == 0) {
                }
                excl += sv & SORT_VAL;
                if (sv & SORT_INCL) break;
            }
        }
        *mine = SORT_INCL | (excl + run);
        gadj[t] = digit_base[t] + excl - toff;
    }
    __syncthreads();

#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const int sl = tid + SORT_NT * i;
        if (sl < nvalid) {
            key[i] = xch[sl];
            kout[gadj[(uint32_t)(key[i] >> shift) & 255u] + sl] = key[i];
        }
    }
    __syncthreads();
    uint32_t *xv = reinterpret_cast<uint32_t *>(xch);
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i)
        if (base + i * 32 < n) xv[rank[i]] = val[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const int sl = tid + SORT_NT * i;
        if (sl < nvalid) vout[gadj[(uint32_t)(key[i] >> shift) & 255u] + sl] = xv[sl];
    }
}

// 8-byte load of a randomly placed element; SORT_CFG_GATHER_HINT selects the L2 prefetch-size qualifier (0 = plain load)
#ifndef SORT_CFG_GATHER_HINT
#define SORT_CFG_GATHER_HINT 64
#endif
__device__ __forceinline__ uint64_t gather_u64(const uint64_t *p) {
#if SORT_CFG_GATHER_HINT == 64
    uint64_t v;
    asm volatile("ld.global.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
#elif SORT_CFG_GATHER_HINT == 1
    return __ldcs(p);
#elif SORT_CFG_GATHER_HINT == 2
    return __ldg(p);
#else
    return *p;
#endif
}

// Every column through the permutation.  A warp owns 64 consecutive output rows (two per lane), so each ballot is one
// whole 32-bit word of an output bitmap; the permutation entry is loaded once for all columns.
__global__ void __launch_bounds__(256) sort_gather_kernel(const __grid_constant__ SortGather G) {
    const int lane = threadIdx.x & 31;
    const int64_t wbase = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 64;
    if (wbase >= G.n) return;
    uint32_t src[2];
    bool in[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int64_t j = wbase + 32 * u + lane;
        in[u] = j < G.n;
        src[u] = in[u] ? G.idx[j] : 0u;
    }
    for (int c = 0; c < G.ncols; ++c) {
        const uint64_t *sv = G.values[c];
        const uint8_t *sm = G.validity[c];
        uint64_t *dv = G.out_values[c];
        const bool from_keys = G.key_col == c && G.key_is_int;  // int64 keys: the sorted keys themselves (no gather)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t j = wbase + 32 * u + lane;
            bool ok = in[u];
            if (in[u]) {
                dv[j] = from_keys ? G.sorted_keys[j] ^ SIGN : gather_u64(sv + src[u]);
                if (sm) ok = (sm[src[u] >> 3] >> (src[u] & 7)) & 1;
            }
            if (G.out_validity[c]) {  // padded device bitmap: whole words
                const uint32_t ball = __ballot_sync(0xffffffffu, ok);
                if (lane == 0 && wbase + 32 * u < G.n) *reinterpret_cast<uint32_t *>(G.out_validity[c] + ((wbase + 32 * u) >> 3)) = ball;
            }
        }
    }
}

struct SortScratch {
    uint64_t *keys[2];
    uint32_t *idx[2];
    unsigned long long *hist;        // [8][256]
    unsigned long long *digit_base;  // [8][256]
    unsigned long long *status;      // [ntiles][256]
    uint32_t *tickets;               // [8]
    int32_t *flags;
};
inline int64_t sort_ntiles(int64_t n) { return (n + SORT_TILE - 1) / SORT_TILE; }
inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }
inline SortScratch sort_carve(void *scratch, int64_t n) {
    uint8_t *p = (uint8_t *)scratch;
    SortScratch S;
    for (int i = 0; i < 2; ++i) {
        S.keys[i] = (uint64_t *)p;
        p += up256((size_t)n * 8);
    }
    for (int i = 0; i < 2; ++i) {
        S.idx[i] = (uint32_t *)p;
        p += up256((size_t)n * 4);
    }
    S.hist = (unsigned long long *)p;
    p += 8 * 256 * 8;
    S.digit_base = (unsigned long long *)p;
    p += 8 * 256 * 8;
    S.tickets = (uint32_t *)p;
    p += 256;
    S.flags = (int32_t *)p;
    p += 256;
    S.status = (unsigned long long *)p;
    return S;
}

}  // namespace

size_t sort_scratch_bytes(int64_t n) {
    return 2 * up256((size_t)n * 8) + 2 * up256((size_t)n * 4) + 2 * 8 * 256 * 8 + 512 + (size_t)sort_ntiles(n) * 256 * 8;
}

// keys, digit histograms and flags; *flags_host / hist_host[8*256] are valid after the stream is synchronised
int launch_sort_prepare(const uint64_t *values, int is_int, int64_t n, void *scratch, cudaStream_t stream, int32_t *flags_host,
                        unsigned long long *hist_host) {
    SortScratch S = sort_carve(scratch, n);
    cudaError_t e = cudaMemsetAsync(S.hist, 0, 2 * 8 * 256 * 8 + 512, stream);  // hist, digit_base, tickets, flags
    if (e != cudaSuccess) return (int)e;
    int64_t blocks = (n + SORT_NT * 8 - 1) / (SORT_NT * 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    sort_prepare_kernel<<<(unsigned)blocks, SORT_NT, 0, stream>>>(values, is_int, n, S.keys[0], S.hist, S.flags);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpyAsync(hist_host, S.hist, 8 * 256 * 8, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return (int)e;
    return (int)cudaMemcpyAsync(flags_host, S.flags, 4, cudaMemcpyDeviceToHost, stream);
}

// the digit passes; on return *idx / *sorted_keys point into the scratch.  hist_host as filled by launch_sort_prepare.
int launch_sort_passes(int64_t n, void *scratch, const unsigned long long *hist_host, cudaStream_t stream,
                       const uint32_t **idx, const uint64_t **sorted_keys, int *npasses) {
    SortScratch S = sort_carve(scratch, n);
    unsigned long long base[8 * 256];
    bool skip[8];
    for (int p = 0; p < 8; ++p) {
        unsigned long long run = 0;
        skip[p] = false;
        for (int d = 0; d < 256; ++d) {
            base[p * 256 + d] = run;
            run += hist_host[p * 256 + d];
            skip[p] |= hist_host[p * 256 + d] == (unsigned long long)n;  // constant digit: the pass is the identity
        }
    }
    cudaError_t e = cudaMemcpyAsync(S.digit_base, base, sizeof base, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return (int)e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return (int)e;  // `base` lives on this stack frame
    const int64_t ntiles = sort_ntiles(n);
    int cur = 0, done = 0;
    for (int p = 0; p < 8; ++p) {
        if (skip[p]) continue;
        if ((e = cudaMemsetAsync(S.status, 0, (size_t)ntiles * 256 * 8, stream)) != cudaSuccess) return (int)e;
        sort_pass_kernel<<<(unsigned)ntiles, SORT_NT, 0, stream>>>(S.keys[cur], done ? S.idx[cur] : nullptr, S.keys[cur ^ 1],
                                                                  S.idx[cur ^ 1], n, 8 * p, S.digit_base + p * 256, S.status,
                                                                  S.tickets + p);
        if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
        cur ^= 1;
        ++done;
    }
    if (!done) return (int)cudaErrorInvalidValue;  // (a column of equal keys is sorted: the caller never gets here)
    *idx = S.idx[cur];
    *sorted_keys = S.keys[cur];
    *npasses = done;
    return 0;
}

int launch_sort_gather(const SortGather &G, cudaStream_t stream) {
    if (G.n <= 0) return 0;
    sort_gather_kernel<<<(unsigned)((G.n + 511) / 512), 256, 0, stream>>>(G);
    return (int)cudaGetLastError();
}

}  // namespace bowgpu
