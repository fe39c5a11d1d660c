// bounds — window-boundary computation for rolling.IntervalRolling.
//
// Replaces the reference's row-at-a-time iterator (HasNext/Next, rolling/rolling.go:162-239) by one
// coalesced streaming pass over the sorted int64 time column: first[k] = Window.FirstIndex of
// window k (= lower bound of S_k; 0 for k == 0), first[W] = n.  The same pass verifies the GPU
// path's precondition (time sorted ascending).  Row tiles are staged by TMA bulk copies
// (tile_pipe.cuh); thread t owns R consecutive rows and walks them with a running window end, so
// there is one exact 64-bit division per thread and tile, not per row.  The thread that owns the
// last row of a window writes the start index of every window up to the next non-empty one, so
// empty windows need no second pass.
#include <atomic>
#include "kernels.h"
#include "tile_pipe.cuh"

namespace bowgpu {

namespace {

constexpr int BND_NT = 128;
constexpr int BND_R = 17;
using BndG = TileGeom<BND_NT, BND_R>;
constexpr int BND_STAGE_BYTES = BndG::TIME_BYTES;
constexpr int BND_HEADER_BYTES = 128;
constexpr int BND_STAGES = 4;

// Rolled, out-of-line walk over the rows of ONE thread that holds at least one window start (or row 0, the
// last row, or rows before s0).  Kept out of the unrolled fast path: most threads of most tiles see no window
// start at all (long windows), and the fast path must stay small enough to live in the instruction cache.
__device__ __noinline__ void bounds_walk(const BoundsLaunch &P, const int64_t *trow /* my first row in smem */,
                                          int64_t row0 /* its global index */, int nmine /* rows I own */,
                                          bool has_next /* the row after my last one exists */) {
    const WindowGeom &g = P.g;
    const uint64_t d = g.div.d, Wu = (uint64_t)g.W;
    auto xrel = [&](int j) -> uint64_t {  // rows before s0 (left halo / negative timestamps) collapse onto window 0
        return row0 + j < g.early_rows ? 0 : (uint64_t)trow[j] - (uint64_t)g.s0;
    };
    uint64_t kcur = div_u64(xrel(0), g.div);
    uint64_t erel = (kcur + 1) * d;
    if (row0 == 0)  // owner of row 0 (kcur > 0 only for a shard with leading empty windows)
        for (uint64_t k = 0; k <= kcur && k <= Wu; ++k) P.first[k] = g.shard ? g.early_rows : 0;
    const int nsteps = has_next ? nmine : nmine - 1;
    for (int j = 0; j < nsteps; ++j) {
        const uint64_t xn = xrel(j + 1);
        if (xn >= erel) {  // row j+1 starts window knew; windows kcur+1..knew begin there
            const uint64_t knew = xn - erel < d ? kcur + 1 : div_u64(xn, g.div);
            const int64_t row = row0 + j + 1;
            for (uint64_t k = kcur + 1; k <= knew && k <= Wu; ++k) P.first[k] = row;
            kcur = knew;
            erel = (kcur + 1) * d;
        }
    }
    if (!has_next)  // owner of the last row: trailing (empty) windows end at n
        for (uint64_t k = kcur + 1; k <= Wu; ++k) P.first[k] = g.n;
}

__global__ void __launch_bounds__(BND_NT, 4) bounds_kernel(const __grid_constant__ BoundsLaunch P, const int64_t ntiles) {
    using G = BndG;
    constexpr int R = G::R;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint8_t *stages = smem_raw + BND_HEADER_BYTES;
    const int tid = threadIdx.x;
    const WindowGeom &g = P.g;
    TileSrc src{P.time, nullptr, nullptr, g.n};

    if (tid == 0) {
        for (int s = 0; s < BND_STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();
    if (tid == 0) {
        int64_t tl = blockIdx.x;
        for (int s = 0; s < BND_STAGES && tl < ntiles; ++s, tl += gridDim.x)
            issue_tile<G, false>(src, tl, stages + (size_t)s * BND_STAGE_BYTES, &full[s]);
    }
    int stage = 0;
    uint32_t phase = 0;
    bool bad = false;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(&full[stage], phase);
        const int64_t *tsm = reinterpret_cast<const int64_t *>(stages + (size_t)stage * BND_STAGE_BYTES);
        const int64_t r0 = tile * G::T;
        const int64_t nrem = g.n - r0;
        const int ti0 = tid * R;
        if (ti0 < nrem) {
            const bool whole = ti0 + R < nrem;  // my R rows and the row after them exist
            const int nmine = whole ? R : (int)(nrem - ti0);
            int64_t x[R + 1];
#pragma unroll
            for (int j = 0; j <= R; ++j) x[j] = tsm[ti0 + G::HALO + j];
            if (r0 + ti0 > 0) bad |= x[0] < tsm[ti0 + G::HALO - 1];
#pragma unroll
            for (int j = 1; j < R; ++j)
                if (whole || j < nmine) bad |= x[j] < x[j - 1];
            // Sorted rows: a window starts inside (x[0], x[R]] iff x[R] lies at or beyond the end of x[0]'s window.
            // The common cases - no start, or exactly one start of the very next window - are handled branch-free
            // here; empty windows in between, several starts, tile / column edges and halo rows take the walk.
            bool walk = !whole || r0 + ti0 == 0 || r0 + ti0 < g.early_rows;
            if (!walk) {
                const uint64_t rel0 = (uint64_t)x[0] - (uint64_t)g.s0;
                const uint64_t kf = div_u64(rel0, g.div);
                const uint64_t erel = (kf + 1) * g.div.d;
                const uint64_t relR = (uint64_t)x[R] - (uint64_t)g.s0;
                if (relR >= erel) {
                    walk = relR - erel >= g.div.d;  // x[R] is beyond window kf+1: more than one start (or empty windows)
                    if (!walk) {
                        int before = 0;  // rows among x[1..R] that still belong to window kf (monotone predicate)
#pragma unroll
                        for (int j = 1; j <= R; ++j) before += (uint64_t)x[j] - (uint64_t)g.s0 < erel;
                        if (kf + 1 <= (uint64_t)g.W) P.first[kf + 1] = r0 + ti0 + before + 1;
                    }
                }
            }
            if (walk) bounds_walk(P, tsm + ti0 + G::HALO, r0 + ti0, nmine, whole);
        }
        __syncthreads();
        if (tid == 0) {
            const int64_t nxt = tile + (int64_t)BND_STAGES * gridDim.x;
            if (nxt < ntiles) issue_tile<G, false>(src, nxt, stages + (size_t)stage * BND_STAGE_BYTES, &full[stage]);
        }
        if (++stage == BND_STAGES) {
            stage = 0;
            phase ^= 1u;
        }
    }
    if (bad) atomicOr(P.status, ST_UNSORTED);
}

// inc[k] = first[k+1] < n && t[first[k+1]] == S_{k+1}  (Window.IsInclusive, rolling.go:201-209)
__global__ void inclusive_bitmap_kernel(const int64_t *time, const int64_t *first, WindowGeom g, uint8_t *bitmap) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool inc = false;
    if (k < g.W) {
        const int64_t b = first[k + 1];
        const uint64_t erel = ((uint64_t)k + 1) * g.div.d;
        if (b < g.n && b >= g.early_rows) inc = ((uint64_t)time[b] - (uint64_t)g.s0) == erel;
    }
    const uint32_t ball = __ballot_sync(0xffffffffu, inc);
    const int lane = threadIdx.x & 31;
    if ((lane & 7) == 0 && k < g.W) bitmap[k >> 3] = (uint8_t)(ball >> lane);
}

// first[k] by one binary search per window instead of one pass over the time column: the same values as bounds_kernel
// (first row at or beyond S_k among the rows from s0 on; first[0] = 0 unless the rows before s0 are a shard's halo) at
// (W + 1) * log2(n) scattered loads instead of 8 n streamed bytes — the better deal when windows hold hundreds of rows
// (configs[2]: 1.1e6 windows over 1e9 rows, 1.28 ms -> see DESIGN 3.2).  No order check: callers use it only where a
// streaming kernel that checks every row follows (the fused Interpolate -> Aggregate path).
__global__ void bounds_search_kernel(const __grid_constant__ BoundsLaunch P) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const WindowGeom &g = P.g;
    if (k > g.W) return;
    if (k == 0) {
        P.first[0] = g.shard ? g.early_rows : 0;
        return;
    }
    const uint64_t target = (uint64_t)k * g.div.d;  // S_k - s0
    int64_t lo = g.early_rows, hi = g.n;
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if ((uint64_t)P.time[mid] - (uint64_t)g.s0 < target)
            lo = mid + 1;
        else
            hi = mid;
    }
    P.first[k] = lo;
}

__global__ void lower_bound_kernel(const int64_t *time, int64_t n, int64_t x, int64_t *out) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (time[mid] < x)
            lo = mid + 1;
        else
            hi = mid;
    }
    *out = lo;
}

// dst word w holds bits [32w, 32w+32) of the logical bitmap = src bits [off+32w, ...)
__global__ void bitmap_realign_kernel(const uint8_t *src, int64_t off, int64_t nbits, uint32_t *dst, int64_t dst_words) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= dst_words) return;
    const int64_t b0 = w * 32;
    uint32_t v = 0;
    if (b0 < nbits) {
        const int64_t sbit = off + b0;
        const int64_t sbyte = sbit >> 3;
        const int sh = (int)(sbit & 7);
        const int64_t last_byte = (off + nbits - 1) >> 3;
        uint64_t acc = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i)
            if (sbyte + i <= last_byte) acc |= (uint64_t)src[sbyte + i] << (8 * i);
        v = (uint32_t)(acc >> sh);
        const int64_t rem = nbits - b0;
        if (rem < 32) v &= (1u << rem) - 1u;
    }
    dst[w] = v;
}

__global__ void bitmap_popcount_kernel(const uint32_t *bm, int64_t nwords, unsigned long long *out) {
    unsigned long long c = 0;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (int64_t)gridDim.x * blockDim.x)
        c += __popc(bm[w]);
    for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

}  // namespace

int launch_bounds(const BoundsLaunch &L, int sm_count, cudaStream_t stream, cudaEvent_t e0, cudaEvent_t e1) {
    const int64_t ntiles = (L.g.n + BndG::T - 1) / BndG::T;
    if (ntiles == 0) return 0;
    const int smem = BND_HEADER_BYTES + BND_STAGES * BND_STAGE_BYTES;
    static std::atomic<bool> configured[64];  // function attributes are per device (one ctx per GPU may live in one process)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63].load(std::memory_order_acquire)) {  // (host threads may launch concurrently)
        cudaError_t e = cudaFuncSetAttribute(bounds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured[dev & 63].store(true, std::memory_order_release);
    }
    int64_t grid = (int64_t)sm_count * 3;
    if (grid > ntiles) grid = ntiles;
    if (e0) cudaEventRecord(e0, stream);
    bounds_kernel<<<(unsigned)grid, BND_NT, smem, stream>>>(L, ntiles);
    if (e1) cudaEventRecord(e1, stream);
    return (int)cudaGetLastError();
}

int launch_bounds_search(const BoundsLaunch &L, cudaStream_t stream) {
    const int nt = 128;
    bounds_search_kernel<<<(unsigned)((L.g.W + 1 + nt - 1) / nt), nt, 0, stream>>>(L);
    return (int)cudaGetLastError();
}

int launch_inclusive_bitmap(const int64_t *time, const int64_t *first, WindowGeom g, uint8_t *bitmap,
                            cudaStream_t stream) {
    if (g.W <= 0) return 0;
    const int nt = 256;
    inclusive_bitmap_kernel<<<(unsigned)((g.W + nt - 1) / nt), nt, 0, stream>>>(time, first, g, bitmap);
    return (int)cudaGetLastError();
}

int launch_lower_bound(const int64_t *time, int64_t n, int64_t x, int64_t *out, cudaStream_t stream) {
    lower_bound_kernel<<<1, 1, 0, stream>>>(time, n, x, out);
    return (int)cudaGetLastError();
}

int launch_bitmap_realign(const uint8_t *src, int64_t off, int64_t nbits, uint8_t *dst, int64_t dst_bytes,
                          cudaStream_t stream) {
    const int64_t words = dst_bytes / 4;
    if (words == 0) return 0;
    const int nt = 256;
    bitmap_realign_kernel<<<(unsigned)((words + nt - 1) / nt), nt, 0, stream>>>(src, off, nbits, (uint32_t *)dst, words);
    return (int)cudaGetLastError();
}

int launch_bitmap_popcount(const uint8_t *bm, int64_t nbits, unsigned long long *out, cudaStream_t stream) {
    const int64_t words = (nbits + 31) / 32;
    if (words == 0) return 0;
    const int nt = 256;
    int64_t grid = (words + nt - 1) / nt;
    if (grid > 1184) grid = 1184;
    bitmap_popcount_kernel<<<(unsigned)grid, nt, 0, stream>>>((const uint32_t *)bm, words, out);
    return (int)cudaGetLastError();
}

}  // namespace bowgpu
