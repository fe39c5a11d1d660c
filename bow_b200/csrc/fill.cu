// fill — whole-column null filling on the device: FillPrevious / FillNext / FillMean / FillLinear
// (reference bowfill.go:14-288) and the sortedness test FillLinear applies to its reference column
// (IsColSorted, bowassertion.go:15-81).
//
// Every null row needs the nearest valid row before and after it IN THE ORIGINAL column (the reference reads the
// immutable Bow, not the buffer being filled).  Three passes over the validity bitmap (1 bit / row) and one over
// the values (8 bytes / row):
//   1. fill_block_edges : per block of 8192 rows, its last and first valid row;
//   2. fill_carry_scan  : exclusive max-scan (previous valid row before each block) and reverse exclusive min-scan
//                         (next valid row after it) over the per-block edges;
//   3. fill_apply       : per block, the same two scans at 32-row word granularity in shared memory, then every warp
//                         walks its rows coalesced: valid rows are copied, null rows fetch their neighbours' values
//                         (nearby, cache resident) and apply the method; output validity words come from ballots.
// Algorithmic bytes: 16 B/row (read + write values) + bitmaps; HBM bound.
#include <math_constants.h>

#include <cstring>

#include "../../include/bowgpu.h"
#include "kernels.h"

namespace bowgpu {

namespace {

constexpr int FILL_NT = 256;               // threads = 32-row words per block
constexpr int FILL_ROWS = FILL_NT * 32;    // rows per block
constexpr int64_t NONE_NEXT = INT64_MAX;   // "no valid row after"

__device__ __forceinline__ uint32_t load_word(const uint32_t *bm, int64_t w, int64_t n) {
    const int64_t r0 = w * 32;
    if (r0 >= n) return 0u;
    uint32_t x = bm ? bm[w] : 0xFFFFFFFFu;
    if (n - r0 < 32) x &= (1u << (int)(n - r0)) - 1u;
    return x;
}

// block-wide inclusive scans of one int64 per thread (blockDim.x <= 1024), max and min flavours
template <bool IS_MAX>
__device__ __forceinline__ int64_t comb(int64_t a, int64_t b) {
    return IS_MAX ? (a > b ? a : b) : (a < b ? a : b);
}
template <bool IS_MAX>
__device__ __forceinline__ int64_t block_scan_incl(int64_t v, int64_t *sh /*[33]*/, bool reverse, int64_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int64_t ident = IS_MAX ? INT64_MIN : INT64_MAX;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t u = reverse ? __shfl_down_sync(0xffffffffu, v, o) : __shfl_up_sync(0xffffffffu, v, o);
        if (reverse ? lane + o < 32 : lane >= o) v = comb<IS_MAX>(v, u);
    }
    if (lane == (reverse ? 0 : 31)) sh[warp] = v;  // the warp's total
    __syncthreads();
    if (warp == 0) {
        int64_t w = lane < nw ? sh[lane] : ident;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t u = reverse ? __shfl_down_sync(0xffffffffu, w, o) : __shfl_up_sync(0xffffffffu, w, o);
            if (reverse ? lane + o < 32 : lane >= o) w = comb<IS_MAX>(w, u);
        }
        int64_t ex = reverse ? __shfl_down_sync(0xffffffffu, w, 1) : __shfl_up_sync(0xffffffffu, w, 1);
        if (reverse ? lane == 31 : lane == 0) ex = ident;
        __syncwarp();
        if (lane < nw) sh[lane] = ex;  // combination of the warps before (after, when reversed) this one
        if (lane == (reverse ? 0 : 31)) sh[32] = w;
    }
    __syncthreads();
    const int64_t res = comb<IS_MAX>(v, sh[warp]);
    total = sh[32];
    __syncthreads();
    return res;
}

__global__ void fill_block_edges(const uint32_t *bm, int64_t n, int64_t *blk_last, int64_t *blk_first) {
    __shared__ int64_t sh[33];
    const int64_t w = (int64_t)blockIdx.x * FILL_NT + threadIdx.x;
    const uint32_t word = load_word(bm, w, n);
    const int64_t last = word ? w * 32 + 31 - __clz(word) : -1;
    const int64_t first = word ? w * 32 + __ffs(word) - 1 : NONE_NEXT;
    int64_t tot_last, tot_first;
    block_scan_incl<true>(last, sh, false, tot_last);
    block_scan_incl<false>(first, sh, false, tot_first);
    if (threadIdx.x == 0) {
        blk_last[blockIdx.x] = tot_last;
        blk_first[blockIdx.x] = tot_first;
    }
}

// one block: carry_prev[b] = max(blk_last[0..b)), carry_next[b] = min(blk_first(b..nb)) (both exclusive)
__global__ void __launch_bounds__(1024) fill_carry_scan(const int64_t *blk_last, const int64_t *blk_first, int64_t nb,
                                                        int64_t *carry_prev, int64_t *carry_next) {
    __shared__ int64_t sh[33];
    __shared__ int64_t tmp[1024];
    const int tid = threadIdx.x, nt = blockDim.x;
    int64_t run = -1;
    for (int64_t base = 0; base < nb; base += nt) {
        const int64_t i = base + tid;
        int64_t total;
        tmp[tid] = block_scan_incl<true>(i < nb ? blk_last[i] : -1, sh, false, total);
        __syncthreads();
        const int64_t ex = tid ? tmp[tid - 1] : -1;
        if (i < nb) carry_prev[i] = ex > run ? ex : run;
        run = total > run ? total : run;
        __syncthreads();
    }
    run = NONE_NEXT;
    for (int64_t c = (nb + nt - 1) / nt - 1; c >= 0; --c) {
        const int64_t i = c * nt + tid;
        int64_t total;
        tmp[tid] = block_scan_incl<false>(i < nb ? blk_first[i] : NONE_NEXT, sh, true, total);
        __syncthreads();
        const int64_t ex = tid + 1 < nt ? tmp[tid + 1] : NONE_NEXT;
        if (i < nb) carry_next[i] = ex < run ? ex : run;
        run = total < run ? total : run;
        __syncthreads();
    }
}

struct FillArgs {
    const uint64_t *values;
    const uint32_t *bm;        // validity of the column to fill (never null here: a column without nulls is not filled)
    const uint64_t *ref_values;  // FillLinear: reference column
    const uint32_t *ref_bm;      // its validity or null
    uint64_t *out_values;
    uint32_t *out_bm;
    const int64_t *carry_prev, *carry_next;
    int64_t n;
    int32_t ref_is_int;
    int32_t _pad;
};

__device__ __forceinline__ double as_f64(uint64_t raw, bool is_int) {
    return is_int ? (double)(int64_t)raw : bits_as_f64(raw);
}
__device__ __forceinline__ bool bit_at(const uint32_t *bm, int64_t i) { return !bm || ((bm[i >> 5] >> (i & 31)) & 1u); }

// METHOD: BOWGPU_FILL_*
template <int METHOD, bool IS_INT>
__global__ void __launch_bounds__(FILL_NT) fill_apply(const FillArgs A) {
    __shared__ int64_t sh[33];
    __shared__ uint32_t s_word[FILL_NT];
    __shared__ int64_t s_prev[FILL_NT], s_next[FILL_NT];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t w = (int64_t)blockIdx.x * FILL_NT + tid;
    const uint32_t word = load_word(A.bm, w, A.n);
    s_word[tid] = word;
    {
        const int64_t last = word ? w * 32 + 31 - __clz(word) : -1;
        const int64_t first = word ? w * 32 + __ffs(word) - 1 : NONE_NEXT;
        int64_t total;
        // exclusive scans at word granularity: shift the inclusive results by one thread
        const int64_t incl = block_scan_incl<true>(last, sh, false, total);
        __shared__ int64_t tmp[FILL_NT];
        tmp[tid] = incl;
        __syncthreads();
        const int64_t cp = A.carry_prev[blockIdx.x];
        int64_t ex = tid ? tmp[tid - 1] : -1;
        s_prev[tid] = ex > cp ? ex : cp;
        __syncthreads();
        const int64_t incl2 = block_scan_incl<false>(first, sh, true, total);
        tmp[tid] = incl2;
        __syncthreads();
        const int64_t cn = A.carry_next[blockIdx.x];
        ex = tid + 1 < FILL_NT ? tmp[tid + 1] : NONE_NEXT;
        s_next[tid] = ex < cn ? ex : cn;
        __syncthreads();
    }
    // each warp walks its 32 words; lane = row inside the word (coalesced value accesses)
    for (int k = 0; k < 32; ++k) {
        const int wi = warp * 32 + k;
        const int64_t wg = (int64_t)blockIdx.x * FILL_NT + wi;
        const int64_t row = wg * 32 + lane;
        if (wg * 32 >= A.n) break;  // warp-uniform
        const uint32_t wd = s_word[wi];
        const bool in = row < A.n;
        bool ok = (wd >> lane) & 1u;
        uint64_t out = 0;
        if (in && ok) {
            out = A.values[row];
        } else if (in) {
            const uint32_t below = wd & ((1u << lane) - 1u);
            const uint32_t above = lane == 31 ? 0u : wd & ~((2u << lane) - 1u);
            const int64_t p = below ? wg * 32 + 31 - __clz(below) : s_prev[wi];
            const int64_t q = above ? wg * 32 + __ffs(above) - 1 : s_next[wi];
            const bool hp = p >= 0, hq = q != NONE_NEXT;
            if (METHOD == BOWGPU_FILL_PREVIOUS) {  // bowfill.go:160-164
                if (hp) {
                    out = A.values[p];
                    ok = true;
                }
            } else if (METHOD == BOWGPU_FILL_NEXT) {  // bowfill.go:154-158
                if (hq) {
                    out = A.values[q];
                    ok = true;
                }
            } else if (METHOD == BOWGPU_FILL_MEAN) {  // bowfill.go:136-146
                if (hp && hq) {
                    const double m = __dmul_rn(__dadd_rn(as_f64(A.values[p], IS_INT), as_f64(A.values[q], IS_INT)), 0.5);
                    out = IS_INT ? (uint64_t)f64_to_i64_go(round(m)) : f64_as_bits(m);  // math.Round: half away from zero
                    ok = true;
                }
            } else {  // FillLinear, bowfill.go:66-95
                if (hp && hq && bit_at(A.ref_bm, row) && bit_at(A.ref_bm, p) && bit_at(A.ref_bm, q)) {
                    const bool ri = A.ref_is_int != 0;
                    const double prev_to_fill = as_f64(A.values[p], IS_INT), next_to_fill = as_f64(A.values[q], IS_INT);
                    const double row_ref = as_f64(A.ref_values[row], ri), prev_ref = as_f64(A.ref_values[p], ri),
                                 next_ref = as_f64(A.ref_values[q], ri);
                    double tmp = __dsub_rn(row_ref, prev_ref);
                    tmp = __ddiv_rn(tmp, __dsub_rn(next_ref, prev_ref));
                    tmp = __dmul_rn(tmp, __dsub_rn(next_to_fill, prev_to_fill));
                    tmp = __dadd_rn(tmp, prev_to_fill);
                    out = IS_INT ? (uint64_t)f64_to_i64_go(round(tmp)) : f64_as_bits(tmp);
                    ok = true;
                }
            }
        }
        if (in) A.out_values[row] = out;
        const uint32_t ball = __ballot_sync(0xffffffffu, in && ok);
        if (lane == 0) A.out_bm[wg] = ball;
    }
}

// IsColSorted over the VALID rows of a column (bowassertion.go:15-81): flags bit 0 = some valid row is smaller
// than the previous valid one, bit 1 = some is larger; sorted (ascending or descending, ties allowed) iff not both.
template <bool IS_INT>
__global__ void __launch_bounds__(FILL_NT) sorted_flags_kernel(const uint64_t *values, const uint32_t *bm, int64_t n,
                                                               const int64_t *carry_prev, int32_t *flags) {
    __shared__ int64_t sh[33];
    __shared__ uint32_t s_word[FILL_NT];
    __shared__ int64_t s_prev[FILL_NT], tmp[FILL_NT];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t w = (int64_t)blockIdx.x * FILL_NT + tid;
    const uint32_t word = load_word(bm, w, n);
    s_word[tid] = word;
    const int64_t last = word ? w * 32 + 31 - __clz(word) : -1;
    int64_t total;
    tmp[tid] = block_scan_incl<true>(last, sh, false, total);
    __syncthreads();
    const int64_t cp = carry_prev[blockIdx.x];
    const int64_t ex = tid ? tmp[tid - 1] : -1;
    s_prev[tid] = ex > cp ? ex : cp;
    __syncthreads();
    int f = 0;
    for (int k = 0; k < 32; ++k) {
        const int wi = warp * 32 + k;
        const int64_t wg = (int64_t)blockIdx.x * FILL_NT + wi;
        const int64_t row = wg * 32 + lane;
        if (wg * 32 >= n) break;
        const uint32_t wd = s_word[wi];
        if (row < n && ((wd >> lane) & 1u)) {
            const uint32_t below = wd & ((1u << lane) - 1u);
            const int64_t p = below ? wg * 32 + 31 - __clz(below) : s_prev[wi];
            if (p >= 0) {
                if (IS_INT) {
                    const int64_t a = (int64_t)values[p], b = (int64_t)values[row];
                    f |= (b < a ? 1 : 0) | (b > a ? 2 : 0);
                } else {
                    const double a = bits_as_f64(values[p]), b = bits_as_f64(values[row]);
                    f |= (b < a ? 1 : 0) | (b > a ? 2 : 0);
                }
            }
        }
    }
    f = __reduce_or_sync(0xffffffffu, f);
    if (lane == 0 && f) atomicOr(flags, f);
}

// ---- DropNils (bow.go:188-224): stream compaction of the rows that are valid in every selected column ----------
constexpr int DROP_MAX_COLS = 32;
struct DropArgs {
    const uint32_t *sel_bm[DROP_MAX_COLS];  // validity of the selected columns that have nulls
    int32_t nsel;
    int32_t ncols;
    const uint64_t *values[DROP_MAX_COLS];
    const uint32_t *bm[DROP_MAX_COLS];      // input validity of columns whose nulls SURVIVE (not selected), else null
    uint64_t *out_values[DROP_MAX_COLS];
    uint32_t *out_bm[DROP_MAX_COLS];        // zero-initialised, or null
    uint32_t *keep;                          // [words] rows kept
    int64_t *blk;                            // [nblocks + 1] kept rows per block, scanned in place
    int64_t n;
};

__global__ void __launch_bounds__(FILL_NT) drop_mark_kernel(const DropArgs A) {
    __shared__ int64_t sh[32];
    const int64_t w = (int64_t)blockIdx.x * FILL_NT + threadIdx.x;
    uint32_t keep = load_word(nullptr, w, A.n);  // all rows that exist
    for (int j = 0; j < A.nsel; ++j) keep &= load_word(A.sel_bm[j], w, A.n);
    if (w * 32 < A.n) A.keep[w] = keep;
    int64_t c = __popc(keep);
    for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        for (int i = 0; i < FILL_NT / 32; ++i) t += sh[i];
        A.blk[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(FILL_NT) drop_compact_kernel(const DropArgs A) {
    __shared__ uint32_t s_word[FILL_NT];
    __shared__ int32_t s_off[FILL_NT];
    __shared__ int32_t s_warp[FILL_NT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t w = (int64_t)blockIdx.x * FILL_NT + tid;
    const uint32_t word = w * 32 < A.n ? A.keep[w] : 0u;
    s_word[tid] = word;
    int c = __popc(word), incl = c;  // exclusive prefix of the popcounts inside the block
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int i = 0; i < warp; ++i) base += s_warp[i];
    s_off[tid] = base + incl - c;
    __syncthreads();
    const int64_t blk_base = A.blk[blockIdx.x];
    for (int k = 0; k < 32; ++k) {
        const int wi = warp * 32 + k;
        const int64_t wg = (int64_t)blockIdx.x * FILL_NT + wi;
        if (wg * 32 >= A.n) break;
        const uint32_t wd = s_word[wi];
        if (!((wd >> lane) & 1u)) continue;
        const int64_t row = wg * 32 + lane;
        const int64_t dst = blk_base + s_off[wi] + __popc(wd & ((1u << lane) - 1u));
        for (int j = 0; j < A.ncols; ++j) {
            A.out_values[j][dst] = A.values[j][row];
            if (A.out_bm[j] && ((A.bm[j][row >> 5] >> (row & 31)) & 1u)) atomicOr(&A.out_bm[j][dst >> 5], 1u << (dst & 31));
        }
    }
}

}  // namespace

size_t drop_scratch_bytes(int64_t n) {
    const int64_t nb = (n + FILL_ROWS - 1) / FILL_ROWS, words = (n + 31) / 32;
    return (size_t)words * 4 + 256 + (size_t)(nb + 2) * 8 + scan_scratch_bytes(nb) + 256;
}

// pass 1: keep mask + per-block counts scanned in place; *d_total (device int64) = rows kept
int launch_drop_mark(const DropLaunch &L, void *scratch, cudaStream_t stream, int64_t **d_total) {
    const int64_t nb = (L.n + FILL_ROWS - 1) / FILL_ROWS, words = (L.n + 31) / 32;
    if (L.n <= 0 || L.ncols > DROP_MAX_COLS) return (int)cudaErrorInvalidValue;
    DropArgs A;
    memset(&A, 0, sizeof A);
    uint8_t *p = (uint8_t *)scratch;
    A.keep = (uint32_t *)p;
    p += ((size_t)words * 4 + 255) / 256 * 256;
    A.blk = (int64_t *)p;
    p += (size_t)(nb + 2) * 8;
    int64_t *scan_tmp = (int64_t *)(((uintptr_t)p + 255) / 256 * 256);
    A.n = L.n;
    for (int j = 0; j < L.ncols; ++j)
        if (L.selected[j] && L.validity[j]) A.sel_bm[A.nsel++] = (const uint32_t *)L.validity[j];
    drop_mark_kernel<<<(unsigned)nb, FILL_NT, 0, stream>>>(A);
    int e = launch_exclusive_scan(A.blk, nb, scan_tmp, stream);
    *d_total = A.blk + nb;
    return e ? e : (int)cudaGetLastError();
}

// pass 2: compaction into the output columns (out_validity[j] zero-initialised, or null when the column cannot hold nulls)
int launch_drop_compact(const DropLaunch &L, void *scratch, cudaStream_t stream) {
    const int64_t nb = (L.n + FILL_ROWS - 1) / FILL_ROWS, words = (L.n + 31) / 32;
    DropArgs A;
    memset(&A, 0, sizeof A);
    uint8_t *p = (uint8_t *)scratch;
    A.keep = (uint32_t *)p;
    p += ((size_t)words * 4 + 255) / 256 * 256;
    A.blk = (int64_t *)p;
    A.n = L.n;
    A.ncols = L.ncols;
    for (int j = 0; j < L.ncols; ++j) {
        A.values[j] = L.values[j];
        A.out_values[j] = L.out_values[j];
        A.out_bm[j] = (uint32_t *)L.out_validity[j];
        A.bm[j] = (const uint32_t *)L.validity[j];
    }
    drop_compact_kernel<<<(unsigned)nb, FILL_NT, 0, stream>>>(A);
    return (int)cudaGetLastError();
}

size_t fill_scratch_bytes(int64_t n) {
    const int64_t nb = (n + FILL_ROWS - 1) / FILL_ROWS;
    return (size_t)(4 * nb + 8) * 8;
}

// scratch: 4*nb int64 (blk_last, blk_first, carry_prev, carry_next); returns cudaError_t as int
static int fill_prepare(const uint32_t *bm, int64_t n, int64_t *scratch, cudaStream_t stream) {
    const int64_t nb = (n + FILL_ROWS - 1) / FILL_ROWS;
    fill_block_edges<<<(unsigned)nb, FILL_NT, 0, stream>>>(bm, n, scratch, scratch + nb);
    fill_carry_scan<<<1, 1024, 0, stream>>>(scratch, scratch + nb, nb, scratch + 2 * nb, scratch + 3 * nb);
    return (int)cudaGetLastError();
}

int launch_fill(int method, const FillLaunch &L, int64_t *scratch, cudaStream_t stream) {
    if (L.n <= 0) return 0;
    const int64_t nb = (L.n + FILL_ROWS - 1) / FILL_ROWS;
    int e = fill_prepare((const uint32_t *)L.validity, L.n, scratch, stream);
    if (e) return e;
    FillArgs A;
    A.values = L.values;
    A.bm = (const uint32_t *)L.validity;
    A.ref_values = L.ref_values;
    A.ref_bm = (const uint32_t *)L.ref_validity;
    A.out_values = L.out_values;
    A.out_bm = (uint32_t *)L.out_validity;
    A.carry_prev = scratch + 2 * nb;
    A.carry_next = scratch + 3 * nb;
    A.n = L.n;
    A.ref_is_int = L.ref_is_int;
    A._pad = 0;
    const unsigned grid = (unsigned)nb;
#define FILL_CASE(M)                                                       \
    case M:                                                                \
        if (L.is_int)                                                      \
            fill_apply<M, true><<<grid, FILL_NT, 0, stream>>>(A);          \
        else                                                               \
            fill_apply<M, false><<<grid, FILL_NT, 0, stream>>>(A);         \
        break;
    switch (method) {
        FILL_CASE(BOWGPU_FILL_PREVIOUS)
        FILL_CASE(BOWGPU_FILL_NEXT)
        FILL_CASE(BOWGPU_FILL_MEAN)
        FILL_CASE(BOWGPU_FILL_LINEAR)
    default: return (int)cudaErrorInvalidValue;
    }
#undef FILL_CASE
    return (int)cudaGetLastError();
}

int launch_sorted_flags(const uint64_t *values, const uint8_t *validity, int is_int, int64_t n, int64_t *scratch,
                        int32_t *flags, cudaStream_t stream) {
    if (n <= 0) return 0;
    const int64_t nb = (n + FILL_ROWS - 1) / FILL_ROWS;
    int e = fill_prepare((const uint32_t *)validity, n, scratch, stream);
    if (e) return e;
    if (is_int)
        sorted_flags_kernel<true><<<(unsigned)nb, FILL_NT, 0, stream>>>(values, (const uint32_t *)validity, n, scratch + 2 * nb, flags);
    else
        sorted_flags_kernel<false><<<(unsigned)nb, FILL_NT, 0, stream>>>(values, (const uint32_t *)validity, n, scratch + 2 * nb, flags);
    return (int)cudaGetLastError();
}

}  // namespace bowgpu
