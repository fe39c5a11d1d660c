// interp — Rolling.Interpolate on the device (reference rolling/interpolation.go:30-161 and the
// closures of rolling/interpolation/{windowstart,linear,stepprevious,none}.go).
//
// Output Bow = for every window k: [one synthetic row at S_k iff the window has no row whose time
// equals S_k] ++ the rows of the window (interpolation.go:118-161).  Three steps:
//   1. interp_window_kernel (W threads): slice [lo, hi) of the window from the bounds array,
//      "missing start" test through the reference's float64 round trip (interpolation.go:121-128),
//      the synthetic row's value for every column (prev / next valid row found by word-wise scans of
//      the validity bitmap), and the window's output row count;
//   2. an exclusive scan of the counts (block-level scan + partial sums) -> output offset per window;
//   3. interp_gather_kernel: row tiles of the OUTPUT frame; each CTA loads the offsets of the windows
//      overlapping its tile into shared memory, every thread binary-searches its row's window there and
//      copies all columns; validity words are built with ballot, so the bit-granular shift caused by
//      inserted rows costs nothing.  Balanced by rows, independent of window sizes.
#include <algorithm>
#include "../../include/bowgpu.h"
#include "kernels.h"

namespace bowgpu {

namespace {

// ---- validity lookups (GetPrev* / GetNext* of the reference, bowgetters.go:65-107,282-311) -------------
// Word-wise on the bitmap; when the word of the row holds no answer the lookup climbs the summary pyramid (one word
// per level) to the nearest non-empty word and descends again: O(levels) however long the null run is.  Without a
// pyramid (summary == null) the words are walked one by one.
struct Bits {
    const uint32_t *bm;       // level 0
    const uint32_t *summary;  // levels 1.. or null
};
__device__ __forceinline__ const uint32_t *pyr_level(const Bits &b, const ValidityPyramid &py, int l) {  // l >= 1
    return b.summary + py.off[l - 1];
}
__device__ __noinline__ int64_t prev_valid_far(const Bits &b, const ValidityPyramid &py, int64_t w) {  // last non-zero word < w
    if (!b.summary) {
        while (--w >= 0)
            if (b.bm[w]) return w * 32 + 31 - __clz(b.bm[w]);
        return -1;
    }
    int64_t idx = w;
    for (int l = 1; l <= py.nlev; ++l) {
        const int64_t ww = idx >> 5;
        const int bit = (int)(idx & 31);
        const uint32_t m = bit ? pyr_level(b, py, l)[ww] & ((1u << bit) - 1u) : 0u;
        if (m) {
            int64_t j = ww * 32 + 31 - __clz(m);
            for (int ll = l - 1; ll >= 1; --ll) j = j * 32 + 31 - __clz(pyr_level(b, py, ll)[j]);
            return j * 32 + 31 - __clz(b.bm[j]);
        }
        idx = ww;
    }
    return -1;
}
__device__ __noinline__ int64_t next_valid_far(const Bits &b, const ValidityPyramid &py, int64_t w, int64_t nw) {  // first non-zero word > w
    if (!b.summary) {
        while (++w < nw)
            if (b.bm[w]) return w * 32 + __ffs(b.bm[w]) - 1;
        return -1;
    }
    int64_t idx = w;
    for (int l = 1; l <= py.nlev; ++l) {
        const int64_t ww = idx >> 5;
        const int bit = (int)(idx & 31);
        const uint32_t m = bit < 31 ? pyr_level(b, py, l)[ww] & (0xFFFFFFFFu << (bit + 1)) : 0u;
        if (m) {
            int64_t j = ww * 32 + __ffs(m) - 1;
            for (int ll = l - 1; ll >= 1; --ll) j = j * 32 + __ffs(pyr_level(b, py, ll)[j]) - 1;
            return j * 32 + __ffs(b.bm[j]) - 1;
        }
        idx = ww;
    }
    return -1;
}
__device__ __forceinline__ int64_t prev_valid(const Bits &b, const ValidityPyramid &py, int64_t i) {  // last valid row <= i, or -1
    if (i < 0) return -1;
    if (!b.bm) return i;
    const int64_t w = i >> 5;
    const uint32_t m = b.bm[w] & (0xFFFFFFFFu >> (31 - (int)(i & 31)));
    if (m) return w * 32 + 31 - __clz(m);
    return prev_valid_far(b, py, w);
}
__device__ __forceinline__ int64_t next_valid(const Bits &b, const ValidityPyramid &py, int64_t i, int64_t n) {  // first valid row >= i
    if (i >= n) return -1;
    if (!b.bm) return i;
    const int64_t w = i >> 5;
    const uint32_t m = b.bm[w] & (0xFFFFFFFFu << (int)(i & 31));
    const int64_t r = m ? w * 32 + __ffs(m) - 1 : next_valid_far(b, py, w, (n + 31) >> 5);
    return r < n ? r : -1;
}

// level l+1 from level l: one ballot per 32 input words
__global__ void pyramid_level_kernel(const uint32_t *in, const int64_t nin, uint32_t *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t ball = __ballot_sync(0xffffffffu, i < nin && in[i] != 0);
    if ((threadIdx.x & 31) == 0 && i < nin) out[i >> 5] = ball;
}

// the first level reads the validity bitmap itself (n / 8 bytes): four words per thread (one 16-byte load), eight threads
// share an output word.  The bitmap is padded with zero bytes to a multiple of 16 (DevCol).
__global__ void pyramid_level4_kernel(const uint4 *in, const int64_t nin /* words */, uint32_t *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t nib = 0;
    if (4 * i < nin) {
        const uint4 w = in[i];
        nib = (w.x != 0 ? 1u : 0u) | (w.y != 0 ? 2u : 0u) | (w.z != 0 ? 4u : 0u) | (w.w != 0 ? 8u : 0u);
        const int64_t left = nin - 4 * i;  // (only the words of the bitmap count, whatever the padding holds)
        if (left < 4) nib &= (1u << left) - 1u;
    }
    uint32_t v = nib << (4 * (lane & 7));
    v |= __shfl_xor_sync(0xffffffffu, v, 1);
    v |= __shfl_xor_sync(0xffffffffu, v, 2);
    v |= __shfl_xor_sync(0xffffffffu, v, 4);
    if ((lane & 7) == 0 && 4 * i < nin) out[i >> 3] = v;
}

// One thread per (window, column): blockIdx.y is the column.  The chain of dependent loads of a window (first row, its
// time, validity words, summary levels, the two points of Linear) is the whole cost of this kernel; one thread walking it
// for every column in turn took 551 us for 1.1e6 windows x 5 columns (ncu: 44 long-scoreboard stall cycles per issued
// instruction), a thread per column recomputes the window's header and walks one chain.  The per-window outputs are
// written by the thread of column 0.
__global__ void interp_window_kernel(const __grid_constant__ InterpLaunch P) {  // (grid constant: P.cols[j] is read in place; by value the 2.8 KB struct was copied to every thread's stack)
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int jcol = blockIdx.y;
    const WindowGeom &g = P.g;
    if (k >= g.W) return;
    const int64_t *t = P.time;
    const uint64_t d = g.div.d;
    const int64_t Sk = (int64_t)((uint64_t)g.s0 + (uint64_t)k * d);
    int64_t lo = P.first[k];
    const int64_t b = P.first[k + 1];
    bool inc = false;  // Window.IsInclusive, rolling.go:201-209
    if (P.inclusive && b < g.n && b >= g.early_rows) inc = ((uint64_t)t[b] - (uint64_t)g.s0) == ((uint64_t)k + 1) * d;
    int64_t hi = b + (inc ? 1 : 0);
    if (k == 0 && g.early_rows > 0 && !g.early_keep && !g.shard) lo = hi = 0;  // rows before s0 dropped with an empty window 0
    const int64_t first_index = lo;
    // interpolation.go:119-128: the comparison goes through float64
    const int64_t fcv = hi > lo ? f64_to_i64_go((double)t[lo]) : -1;
    const int missing = fcv != Sk;
    if (jcol == 0) {
        if (P.off) {
            P.off[k] = (int64_t)missing + (hi - lo);
            P.wsrc[k] = (lo - missing) * 2 + missing;  // fixed up to "source row - output row" after the scan
        }
        if (P.missing) P.missing[k] = (uint8_t)missing;
        // beyond 2^53 a first row NEAR S_k passes the float64 test without sitting exactly on it: the fused
        // Interpolate -> Aggregate path cannot express that frame and falls back to the materialising one
        if (!missing && hi > lo && t[lo] != Sk && P.status) atomicOr(P.status, ST_INEXACT_START);
    }
    if (!missing || jcol >= P.ncols) return;
    {
        const int j = jcol;
        const InterpCol &c = P.cols[j];
        uint64_t bits = 0;
        bool valid = false;
        switch (c.op) {
        case BOWGPU_INTERP_WINDOW_START:  // interpolation/windowstart.go:8-14
            bits = (uint64_t)Sk;
            valid = true;
            break;
        case BOWGPU_INTERP_NONE:  // interpolation/none.go:8-14
            break;
        case BOWGPU_INTERP_STEP_PREVIOUS: {  // interpolation/stepprevious.go:8-26
            const int64_t p = prev_valid(Bits{c.validity, c.summary}, P.pyr, first_index - 1);
            if (p >= 0) {
                bits = c.values[p];
                valid = true;
            } else if (c.prev_valid) {
                bits = c.prev_bits;
                valid = true;
            }
            break;
        }
        case BOWGPU_INTERP_STEP_NEXT: {  // not upstream: StepPrevious mirrored over GetNextValues (bowgetters.go:111-123)
            const int64_t nx = next_valid(Bits{c.validity, c.summary}, P.pyr, first_index, g.n);
            if (nx >= 0) {
                bits = c.values[nx];
                valid = true;
            }
            break;
        }
        case BOWGPU_INTERP_LINEAR: {  // interpolation/linear.go:8-38
            double t0, v0;
            const int64_t p = prev_valid(Bits{c.validity, c.summary}, P.pyr, first_index - 1);
            if (p >= 0) {
                t0 = (double)t[p];
                v0 = c.is_int ? (double)(int64_t)c.values[p] : bits_as_f64(c.values[p]);
            } else if (c.prev_valid && P.prev_time_valid) {
                t0 = (double)P.prev_time;
                v0 = c.is_int ? (double)(int64_t)c.prev_bits : bits_as_f64(c.prev_bits);
            } else {
                break;
            }
            const int64_t nx = next_valid(Bits{c.validity, c.summary}, P.pyr, first_index, g.n);
            if (nx < 0) break;
            const double t2 = (double)t[nx];
            const double v2 = c.is_int ? (double)(int64_t)c.values[nx] : bits_as_f64(c.values[nx]);
            const double coef = __ddiv_rn(__dsub_rn((double)Sk, t0), __dsub_rn(t2, t0));  // linear.go:34
            const double res = __dadd_rn(__dmul_rn(__dsub_rn(v2, v0), coef), v0);          // linear.go:35
            bits = c.is_int ? (uint64_t)f64_to_i64_go(res) : f64_as_bits(res);  // SetOrDrop into the column type
            valid = true;
            break;
        }
        }
        c.syn_val[k] = valid ? bits : 0;
        c.syn_ok[k] = valid;
    }
}

// ---- exclusive scan of int64 counts, in place; data[n] receives the total ---------------------------------
constexpr int SCAN_NT = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_NT * SCAN_ITEMS;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t *sh /*[32+1]*/, int64_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int64_t w = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0;
        int64_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t u = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += u;
        }
        sh[lane] = wi - w;
        if (lane == 31) sh[32] = wi;
    }
    __syncthreads();
    const int64_t res = sh[warp] + incl - v;
    total = sh[32];
    __syncthreads();
    return res;
}

__global__ void scan_block_sums(const int64_t *data, int64_t n, int64_t *bsum) {
    __shared__ int64_t sh[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) s += data[base + i];
    int64_t total;
    block_exclusive_scan(s, sh, total);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

__global__ void scan_partials(int64_t *bsum, int64_t nb) {  // one block
    __shared__ int64_t sh[33];
    int64_t carry = 0;
    for (int64_t base = 0; base < nb; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = i < nb ? bsum[i] : 0;
        int64_t total;
        const int64_t ex = block_exclusive_scan(v, sh, total);
        if (i < nb) bsum[i] = carry + ex;
        carry += total;
    }
}

__global__ void scan_apply(int64_t *data, int64_t n, const int64_t *bsum) {
    __shared__ int64_t sh[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t v[SCAN_ITEMS];
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = base + i < n ? data[base + i] : 0;
        s += v[i];
    }
    int64_t total;
    int64_t run = block_exclusive_scan(s, sh, total) + bsum[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) data[base + i] = run;
        run += v[i];
        if (base + i == n - 1) data[n] = run;
    }
}

// after the scan: wsrc[k] = ((source row of the window's first copied row - its output row) << 1) | missing
__global__ void interp_fix_wsrc(const int64_t *off, int64_t *wsrc, int64_t W) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= W) return;
    const int64_t v = wsrc[k];
    const int64_t missing = v & 1;
    wsrc[k] = ((v >> 1) - off[k]) * 2 + missing;
}

// ---- gather -----------------------------------------------------------------------------------------------
constexpr int GA_NT = 256, GA_ROWS = 8, GA_TILE = GA_NT * GA_ROWS;  // output rows per CTA
constexpr int GA_CAP = GA_TILE + 4;                                   // windows overlapping one tile

// tile_k[t] = last window k with off[k] <= t * GA_TILE  (t < ntiles); tile_k[ntiles] = W - 1.
// One thread per tile: the searches run in parallel here instead of serialising at the head of every gather CTA.
__global__ void interp_tile_windows(const int64_t *off, int64_t W, int64_t ntiles, int64_t *tile_k) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    if (t == ntiles) {
        tile_k[t] = W - 1;
        return;
    }
    const int64_t o = t * GA_TILE;
    int64_t lo = 0, hi = W;  // off[0] = 0 <= o
    while (hi - lo > 1) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (off[mid] <= o)
            lo = mid;
        else
            hi = mid;
    }
    tile_k[t] = lo;
}

// One CTA per tile of GA_TILE output rows.  Row o of the output comes from input row o + shift(window of o), or is
// the synthetic start row of its window.  Loads of a column (GA_ROWS per thread, coalesced across the warp) are all
// issued before the first store, 4 CTAs per SM keep ~64 KB per SM in flight.
__global__ void __launch_bounds__(GA_NT, 4)
    interp_gather_kernel(const __grid_constant__ InterpLaunch P, const int64_t n_out, const int64_t *__restrict__ tile_k) {
    __shared__ int64_t s_off[GA_CAP];
    __shared__ int64_t s_src[GA_CAP];
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t tile = blockIdx.x;
    const int64_t o0 = tile * GA_TILE;
    const int64_t k_lo = tile_k[tile], k_hi = tile_k[tile + 1];
    // m <= GA_TILE + 2: every window but one (an empty window starting at -1, interpolation.go:119,128)
    // emits at least one row; clamp anyway so shared memory can never be overrun
    const int m = (int)((k_hi - k_lo + 1) < (int64_t)GA_CAP ? (k_hi - k_lo + 1) : (int64_t)GA_CAP);
    for (int i = tid; i < m; i += GA_NT) {
        s_off[i] = P.off[k_lo + i];
        s_src[i] = P.wsrc[k_lo + i];
    }
    __syncthreads();
    int64_t src[GA_ROWS];  // >= 0: input row; -1 - k: synthetic row of window k; INT64_MIN: past the end
    bool any_syn = false;
#pragma unroll
    for (int i = 0; i < GA_ROWS; ++i) {
        const int64_t o = o0 + tid + (int64_t)i * GA_NT;
        src[i] = INT64_MIN;
        if (o < n_out) {
            int lo = 0, hi = m;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_off[mid] <= o)
                    lo = mid;
                else
                    hi = mid;
            }
            const int64_t v = s_src[lo];
            const bool syn = (v & 1) && o == s_off[lo];
            src[i] = syn ? -1 - (k_lo + lo) : o + (v >> 1);
            any_syn |= syn;
        }
    }
    const bool full = o0 + GA_TILE <= n_out;
    const int ncols = P.ncols;
    for (int j = 0; j < ncols; ++j) {
        const uint64_t *__restrict__ vin = P.cols[j].values;
        const uint32_t *__restrict__ bin = P.cols[j].validity;
        uint64_t *__restrict__ vout = P.cols[j].out_values;
        uint32_t *__restrict__ bout = P.cols[j].out_validity;
        uint64_t val[GA_ROWS];
        uint32_t okm = 0;
#pragma unroll
        for (int i = 0; i < GA_ROWS; ++i) {
            const int64_t sr = src[i];
            val[i] = 0;
            if (sr >= 0) {
                val[i] = vin[sr];
                const uint32_t w = bin ? bin[sr >> 5] : 0xFFFFFFFFu;
                okm |= ((w >> (sr & 31)) & 1u) << i;
            }
        }
        if (any_syn) {  // at most one row per window; out of the streaming path
#pragma unroll
            for (int i = 0; i < GA_ROWS; ++i) {
                const int64_t sr = src[i];
                if (sr < 0 && sr != INT64_MIN) {
                    const int64_t k = -1 - sr;
                    val[i] = P.cols[j].syn_val[k];
                    okm |= (uint32_t)(P.cols[j].syn_ok[k] != 0) << i;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < GA_ROWS; ++i) {
            const int64_t o = o0 + tid + (int64_t)i * GA_NT;
            if (full || o < n_out) vout[o] = val[i];
            if (bout) {
                const uint32_t ball = __ballot_sync(0xffffffffu, (okm >> i) & 1u);
                if (lane == 0 && (full || o < n_out)) bout[o >> 5] = ball;
            }
        }
    }
}

}  // namespace

ValidityPyramid make_pyramid(int64_t n) {
    ValidityPyramid py;
    memset(&py, 0, sizeof py);
    int64_t words = (n + 31) / 32, off = 0;
    while (words > 1 && py.nlev < PYR_MAXLEV) {
        words = (words + 31) / 32;
        py.off[py.nlev] = off;
        py.words[py.nlev] = words;
        off += (words + 3) / 4 * 4;  // levels stay 16-byte aligned
        ++py.nlev;
    }
    return py;
}

size_t interp_pyramid_bytes(int64_t n) {
    const ValidityPyramid py = make_pyramid(n);
    return py.nlev ? (size_t)(py.off[py.nlev - 1] + (py.words[py.nlev - 1] + 3) / 4 * 4) * 4 : 16;
}

int launch_interp_windows(const InterpLaunch &L0, cudaStream_t stream) {
    if (L0.g.W <= 0) return 0;
    InterpLaunch L = L0;
    L.pyr = make_pyramid(L.g.n);
    for (int j = 0; j < L.ncols; ++j) {
        InterpCol &c = L.cols[j];
        const bool looks = c.op == BOWGPU_INTERP_STEP_PREVIOUS || c.op == BOWGPU_INTERP_LINEAR || c.op == BOWGPU_INTERP_STEP_NEXT;
        if (!looks || !c.validity || !c.summary || L.pyr.nlev == 0) {
            c.summary = nullptr;
            continue;
        }
        const uint32_t *in = c.validity;
        int64_t nin = (L.g.n + 31) / 32;
        for (int l = 0; l < L.pyr.nlev; ++l) {
            uint32_t *out = c.summary + L.pyr.off[l];
            if (l == 0 && ((uintptr_t)in & 15) == 0) {
                const int64_t groups = (nin + 3) / 4;
                pyramid_level4_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const uint4 *>(in), nin, out);
            } else {
                pyramid_level_kernel<<<(unsigned)((nin + 255) / 256), 256, 0, stream>>>(in, nin, out);
            }
            in = out;
            nin = L.pyr.words[l];
        }
    }
    const int nt = 128;
    interp_window_kernel<<<dim3((unsigned)((L.g.W + nt - 1) / nt), (unsigned)std::max(1, (int)L.ncols)), nt, 0, stream>>>(L);
    return (int)cudaGetLastError();
}

size_t scan_scratch_bytes(int64_t n) { return (size_t)((n + SCAN_TILE - 1) / SCAN_TILE + 1) * 8; }

int launch_exclusive_scan(int64_t *data, int64_t n, int64_t *scratch, cudaStream_t stream) {
    if (n <= 0) return 0;
    const int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_block_sums<<<(unsigned)nb, SCAN_NT, 0, stream>>>(data, n, scratch);
    scan_partials<<<1, 1024, 0, stream>>>(scratch, nb);
    scan_apply<<<(unsigned)nb, SCAN_NT, 0, stream>>>(data, n, scratch);
    return (int)cudaGetLastError();
}

int launch_interp_gather(const InterpLaunch &L, int64_t n_out, int64_t *tile_k, cudaStream_t stream, cudaEvent_t e0,
                         cudaEvent_t e1) {
    if (L.g.W > 0) {
        const int nt = 256;
        interp_fix_wsrc<<<(unsigned)((L.g.W + nt - 1) / nt), nt, 0, stream>>>(L.off, L.wsrc, L.g.W);
    }
    if (n_out <= 0) return (int)cudaGetLastError();
    const int64_t ntiles = interp_gather_tiles(n_out);
    interp_tile_windows<<<(unsigned)((ntiles + 1 + 255) / 256), 256, 0, stream>>>(L.off, L.g.W, ntiles, tile_k);
    if (e0) cudaEventRecord(e0, stream);
    interp_gather_kernel<<<(unsigned)ntiles, GA_NT, 0, stream>>>(L, n_out, tile_k);
    if (e1) cudaEventRecord(e1, stream);
    return (int)cudaGetLastError();
}

int64_t interp_gather_tiles(int64_t n_out) { return (n_out + GA_TILE - 1) / GA_TILE; }

}  // namespace bowgpu
