// Host-callable launchers of the libbowgpu kernels (internal; the public boundary is include/bowgpu.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace bowgpu {

// Window geometry shared by every kernel (SURVEY 8 notation):
//   S_k = s0 + k*I, window of row i: k(i) = floor((t[i]-s0)/I); rows [0, early_rows) have t < s0
//   (negative timestamps, rolling.go:96-99) and belong to window 0 iff early_keep.  On a range-partitioned
//   shard (`shard` set) the early rows are the LEFT HALO: they belong to no window (early_keep = 0,
//   first[0] = early_rows) and only serve the prev-valid searches of Rolling.Interpolate.
struct WindowGeom {
    int64_t n;           // rows
    int64_t s0;          // first window start (rolling.go:96-99)
    int64_t W;           // number of windows (rolling.go:143-154)
    DivU64 div;          // interval
    int64_t early_rows;
    int32_t early_keep;
    int32_t shard;
    // aggregation.Aggregate over the whole Bow (rolling/aggregation/whole.go): ONE window holding every row, whose
    // FirstValue / LastValue are the first / last times (through float64) instead of lattice values
    int64_t whole_first, whole_last;
    int32_t whole;
    int32_t _pad;
};
__host__ __device__ __forceinline__ int64_t window_first_value(const WindowGeom &g, int64_t k) {
    return g.whole ? g.whole_first : (int64_t)((uint64_t)g.s0 + (uint64_t)k * g.div.d);
}
__host__ __device__ __forceinline__ int64_t window_last_value(const WindowGeom &g, int64_t k) {
    return g.whole ? g.whole_last : (int64_t)((uint64_t)g.s0 + ((uint64_t)k + 1) * g.div.d);
}

// status word bits written by the kernels
enum { ST_UNSORTED = 1, ST_INEXACT_START = 2, ST_PARQUET = 4 /* malformed page data */ };

// ---- carry record of one tile edge (segreduce) ------------------------------------------------
struct alignas(16) BasicCarry {
    int64_t key;     // window index, -1 = none
    int64_t cnt;     // valid rows; bit 62 set on head records = window closed inside the tile
    double sum;
    double mn, mx;
    uint64_t first, last;  // raw bits of first / last valid value
    int64_t edge_t;        // head records: time of the tile's first row; tail records: of its last row
};

struct BasicOut {  // per-window outputs of one input column (any pointer may be null)
    int64_t *cnt;
    double *sum;   // also the numerator of ArithmeticMean (the epilogue divides by cnt)
    double *mn;
    double *mx;
    uint64_t *first;
    uint64_t *last;
};

// ops bit mask of the basic family
enum { OPS_SUMCNT = 1, OPS_MINMAX = 2, OPS_FIRSTLAST = 4 };

// Fused Rolling.Interpolate -> Rolling.Aggregate: the synthetic window-start rows of the interpolated frame are not
// materialised; the reduction injects them at window boundaries from these per-window arrays (interp_window_kernel).
struct FusedSyn {
    const uint8_t *missing;  // [W] 1 = the interpolated frame holds a synthetic row at S_k (null = not fused)
    const uint64_t *val;     // [W] its value in this column (raw bits)
    const uint8_t *ok;       // [W] its validity
    int64_t len;             // entries in the three arrays: W, or W + 1 on a shard (start row of the next shard's window)
    const int64_t *first;    // [len + 1] window boundaries (bounds kernel): first row of window k (segmc: empty windows)
};

struct SegLaunch {
    const int64_t *time;
    const uint64_t *values;
    const uint8_t *validity;  // null = all valid
    int32_t is_int;           // values are int64 (converted with (double)) else float64 bits
    uint32_t ops;             // OPS_* mask
    WindowGeom g;
    BasicOut out;
    BasicCarry *carry_head;   // [ntiles]
    BasicCarry *carry_tail;   // [ntiles]
    BasicCarry *skip;         // [seg_skip_bytes(n) / sizeof] combined records of long runs of tiles, or null
    int32_t *status;
    FusedSyn syn;
    int32_t *gate;            // side-by-side lanes (api.cu): every lane's first CTA checks in here and all CTAs wait
    int32_t gate_lanes;       // until gate_lanes lanes have (bounded wait), so that the lanes walk the rows abreast; null = none
};

int64_t seg_num_tiles(int64_t n);
size_t seg_carry_bytes(int64_t n);  // bytes for carry_head + carry_tail
size_t seg_skip_bytes(int64_t n);   // bytes of the skip records (windows spanning very many tiles)
// launches main + fixup kernels on `stream`; returns cudaError_t as int
int launch_segreduce_basic(const SegLaunch &L, int sm_count, cudaStream_t stream, cudaEvent_t ev_main0,
                           cudaEvent_t ev_main1);

// ---- integral family (IntegralStep / IntegralTrapezoid / WeightedAverage*) ------------------------
struct IntegralOut {
    int64_t *n_step;  // [W] zeroed by the caller: number of valid points (IntegralStep is valid iff > 0)
    int64_t *n_trap;  // [W] zeroed by the caller: 1 iff the trapezoid integral is valid (>= 2 points)
    double *step;     // [W] or null
    double *trap;     // [W] or null
};
struct IntLaunch {
    const int64_t *time;
    const uint64_t *values;
    const uint8_t *validity;
    int32_t is_int;
    WindowGeom g;
    IntegralOut out;
    void *carry_head;  // [ntiles] of integral_carry_bytes(n) / 2
    void *carry_tail;
    void *skip;        // integral_skip_bytes(n) or null
    int32_t *status;
    FusedSyn syn;
    int32_t *gate;            // side-by-side lanes (api.cu): every lane's first CTA checks in here and all CTAs wait
    int32_t gate_lanes;       // until gate_lanes lanes have (bounded wait), so that the lanes walk the rows abreast; null = none
};
size_t integral_carry_bytes(int64_t n);
size_t integral_skip_bytes(int64_t n);

// ---- both families of one column in one launch (seg_full.cu) -------------------------------------------
struct FullLaunch {
    const int64_t *time;
    const uint64_t *values;
    const uint8_t *validity;
    int32_t is_int;
    WindowGeom g;
    BasicOut out_basic;        // cnt is required (zeroed by the caller); the other pointers may be null
    IntegralOut out_integral;  // all four arrays required (n_step / n_trap zeroed by the caller)
    void *carry_head;          // [ntiles] of full_carry_bytes(n) / 2
    void *carry_tail;
    void *skip;                // full_skip_bytes(n) or null
    int32_t *status;
    FusedSyn syn;
    int32_t *gate;
    int32_t gate_lanes;
};
size_t full_carry_bytes(int64_t n);
size_t full_skip_bytes(int64_t n);
int launch_segreduce_full(const FullLaunch &L, int sm_count, cudaStream_t stream, cudaEvent_t ev_main0, cudaEvent_t ev_main1);
int launch_segreduce_integral(const IntLaunch &L, int sm_count, cudaStream_t stream, cudaEvent_t ev_main0,
                              cudaEvent_t ev_main1);

// ---- segmc: multi-column, multi-family streaming segmented reduction (segmc.cuh) ----------------------------------
// One launch stages the time column ONCE per tile and streams up to MC_MAXC value columns past it; every aggregation of
// a column (basic and integral family together) comes out of that one pass, and the FINAL output values are written by
// whoever completes a window (no per-window intermediate arrays, no epilogue pass over values).
constexpr int MC_MAXC = 8;
enum { MC_SUM = 1, MC_MINMAX = 2, MC_FIRSTLAST = 4 };  // basic family (the valid-row count is always kept)
enum { MC_STEP = 1, MC_TRAP = 2 };                      // integral family
struct McColOut {  // final outputs of one input column; null = not requested
    int64_t *cnt;
    double *sum, *mean, *mn, *mx;
    uint64_t *first, *last;
    double *step, *trap, *wstep, *wlin;
    uint32_t *vb_cnt;   // zero-initialised validity words: bit k set iff window k holds a valid row (atomicOr)
    uint32_t *vb_trap;  // ... iff the trapezoid integral of window k is defined (>= 2 points, integral.go:33-35)
};
struct alignas(16) McCarry {  // one edge record of a chunk of tiles (joined by mc_fixup_kernel)
    int64_t key;              // window index, -1 = none
    int64_t cnt;              // valid rows; bit 62 = the window closed inside the chunk (head records)
    double sum, mn, mx;
    uint64_t first, last;     // raw bits of the first / last valid value
    double fT, lT, lV, sS, sT;
    double incV, incT;
    int64_t inc_has;
    int64_t edge_t;           // head records: first row of the chunk (time, raw value bits, validity);
    uint64_t edge_raw;        // tail records: edge_t = time of the chunk's last row
    int64_t edge_valid;
    int64_t _pad;
};
static_assert(sizeof(McCarry) == 160, "record layout");
struct McColArgs {
    const uint64_t *values;
    const uint8_t *validity;  // null = all valid (a launch is homogeneous: all columns with, or all without, nulls)
    McColOut out;
    FusedSyn syn;
    McCarry *rec;             // [2 * nchunks]: head, tail of every chunk
};
struct McLaunch {
    const int64_t *time;
    WindowGeom g;
    int32_t ncols;
    int32_t is_int;    // value columns are int64 (a launch is homogeneous)
    uint32_t bops;     // MC_SUM | MC_MINMAX | MC_FIRSTLAST needed by some column (0 = none)
    uint32_t iops;     // MC_STEP | MC_TRAP
    uint32_t *touched; // zero-initialised bitmap words: bit k set iff window k holds at least one row
    int32_t *status;
    McColArgs col[MC_MAXC];
};
int mc_max_chunks(int sm_count);              // upper bound of chunks (records per column = 2 * this)
int launch_segmc(const McLaunch &L, int sm_count, cudaStream_t stream, cudaEvent_t ev_main0, cudaEvent_t ev_main1);

// Per-window finishing pass (finish.cu): validity bitmaps of every output from the few bitmaps the streaming kernel
// set, WindowStart columns, and value 0 in the slots of windows that hold no row at all.
enum { FIN_SRC = 0, FIN_ALWAYS = 1, FIN_WINDOW_START = 2 };
struct FinishDst {
    uint64_t *values;     // [W]
    uint8_t *validity;    // [ceil(W/8)] bytes
    const uint32_t *src;  // FIN_SRC: validity words to copy
    int32_t kind;
    int32_t zero_empty;   // write value 0 for windows without rows (not on the fused path: those hold a synthetic row)
};
int launch_finish(const FinishDst *dst_host, int ndst, const uint32_t *touched, WindowGeom g, cudaStream_t stream,
                  int *launches);
// transformation.Factor (factor.go:7-20) applied in place to the valid slots of a finished output column
int launch_factor(uint64_t *values, const uint8_t *validity, int64_t W, int out_is_int, int nfactors,
                  const double *factors, cudaStream_t stream);

// ---- bounds ------------------------------------------------------------------------------------
struct BoundsLaunch {
    const int64_t *time;
    WindowGeom g;
    int64_t *first;     // [W+1] device
    int32_t *status;
};
int launch_bounds(const BoundsLaunch &L, int sm_count, cudaStream_t stream, cudaEvent_t ev0, cudaEvent_t ev1);
// the same first[] by a binary search per window (no order check): for windows of hundreds of rows
int launch_bounds_search(const BoundsLaunch &L, cudaStream_t stream);
// inclusive flags: inc[k] = first[k+1] < n && t[first[k+1]] == S_{k+1}; bitmap of ceil(W/8) bytes
int launch_inclusive_bitmap(const int64_t *time, const int64_t *first, WindowGeom g, uint8_t *bitmap,
                            cudaStream_t stream);

// ---- interpolate (interp.cu) ------------------------------------------------------------------------
constexpr int INTERP_MAX_COLS = 32;
// Summary pyramid over a validity bitmap: bit i of level l (l >= 1) = word i of level l-1 is non-zero (level 0 = the
// bitmap).  Bounds the prev-valid / next-valid lookups of StepPrevious / Linear / StepNext to one word per level even
// when a column holds null runs of millions of rows.  All levels of one column live in one buffer.
constexpr int PYR_MAXLEV = 8;
struct ValidityPyramid {
    int32_t nlev;               // summary levels (0 = none: bitmaps of at most one word)
    int32_t _pad;
    int64_t off[PYR_MAXLEV];    // word offset of level l+1 inside the column's summary buffer
    int64_t words[PYR_MAXLEV];  // its length in words
};
ValidityPyramid make_pyramid(int64_t n);
size_t interp_pyramid_bytes(int64_t n);  // summary buffer of one column
struct InterpCol {
    const uint64_t *values;    // input column
    const uint32_t *validity;  // input validity (bit offset 0) or null
    uint32_t *summary;         // pyramid levels 1.. of `validity` (scratch of interp_pyramid_bytes(n), filled by
                               // launch_interp_windows) or null = plain word-by-word scans
    uint64_t *syn_val;         // [W] value of the synthetic window-start row (scratch)
    uint8_t *syn_ok;           // [W] its validity
    uint64_t *out_values;      // [n_out]
    uint32_t *out_validity;    // [ceil(n_out/32)] words or null
    uint64_t prev_bits;        // Options.PrevRow cell of this column
    int32_t prev_valid;
    int32_t op;                // BOWGPU_INTERP_*
    int32_t is_int;
    int32_t _pad;
};
struct InterpLaunch {
    const int64_t *time;
    const int64_t *first;  // [W+1] from the bounds kernel
    int64_t *off;          // [W+1] output rows per window, scanned in place to output offsets (null: fused path)
    int64_t *wsrc;         // [W] (null: fused path)
    uint8_t *missing;      // [W] 1 = window k gets a synthetic start row (may be null)
    int32_t *status;       // ST_INEXACT_START is raised when a window's "has start" row is not exactly at S_k
    WindowGeom g;
    int64_t prev_time;     // Options.PrevRow time cell
    int32_t prev_time_valid;
    int32_t inclusive;
    int32_t ncols;
    int32_t _pad;
    ValidityPyramid pyr;   // (filled by launch_interp_windows)
    InterpCol cols[INTERP_MAX_COLS];
};
int launch_interp_windows(const InterpLaunch &L, cudaStream_t stream);
size_t scan_scratch_bytes(int64_t n);
// exclusive scan in place; data[n] receives the total
int launch_exclusive_scan(int64_t *data, int64_t n, int64_t *scratch, cudaStream_t stream);
int64_t interp_gather_tiles(int64_t n_out);
// tile_k: scratch of interp_gather_tiles(n_out) + 1 int64
int launch_interp_gather(const InterpLaunch &L, int64_t n_out, int64_t *tile_k, cudaStream_t stream, cudaEvent_t e0,
                         cudaEvent_t e1);

// ---- per-window epilogue (validity bitmaps, defaults of empty windows, WindowStart, Factor) -----
struct EpilogueSpec {
    int32_t op;          // BOWGPU_AGG_*
    int32_t out_is_int;  // output dtype is int64
    const int64_t *cnt;  // valid-row count of the input column per window (null for WindowStart)
    const uint8_t *ok;   // optional per-window validity bytes overriding cnt > 0 (integral family)
    const double *sum_src;  // ArithmeticMean: per-window sums; WeightedAverage*: integrals (may alias values)
    void *values;        // [W]
    uint8_t *validity;   // [ceil(W/8)] bytes
    int32_t nfactors;
    int32_t _pad;
    double factors[4];
};
int launch_epilogue(const EpilogueSpec *specs_host, int nspecs, WindowGeom g, cudaStream_t stream);

// Fast path for the outputs of one input column that share a per-window count (and carry no Factor): one read of
// the count drives every validity bitmap, the zeros of null slots, the mean / weighted-average divisions and the
// WindowStart column.
constexpr int EPIG_NULL = 8, EPIG_BM = 12, EPIG_ALL = 4, EPIG_DIV = 4, EPIG_WS = 4;
struct EpiGroup {
    const int64_t *cnt;            // null: a group of always-valid outputs only (WindowStart)
    int32_t n_null, n_bm_cnt, n_bm_all, n_div, n_ws, _pad;
    uint64_t *null_vals[EPIG_NULL];  // final values are already there; slots of windows with cnt == 0 are zeroed
    uint8_t *bm_cnt[EPIG_BM];        // validity bitmaps: valid iff cnt > 0
    uint8_t *bm_all[EPIG_ALL];       // validity bitmaps: always valid (Count, Sum, WindowStart)
    double *div_dst[EPIG_DIV];       // dst = src / (by_cnt ? float64(cnt) : float64(interval)), 0 where cnt == 0
    const double *div_src[EPIG_DIV];
    int32_t div_by_cnt[EPIG_DIV];
    int64_t *ws[EPIG_WS];            // WindowStart outputs
};
int launch_epilogue_group(const EpiGroup &G, WindowGeom g, cudaStream_t stream);

// ---- utilities ------------------------------------------------------------------------------------
// dst bitmap (bit offset 0, padded bytes zeroed up to dst_bytes) from src bitmap at bit offset `off`
int launch_bitmap_realign(const uint8_t *src, int64_t off, int64_t nbits, uint8_t *dst, int64_t dst_bytes,
                          cudaStream_t stream);
// number of set bits among the first nbits of a bit-offset-0 bitmap -> *out (device int64, must be zeroed)
int launch_bitmap_popcount(const uint8_t *bm, int64_t nbits, unsigned long long *out, cudaStream_t stream);
// lower bound of `x` in sorted time[0..n) -> *out (device)
int launch_lower_bound(const int64_t *time, int64_t n, int64_t x, int64_t *out, cudaStream_t stream);

// ---- whole-column fills (fill.cu): bowfill.go ----------------------------------------------------------------------
struct FillLaunch {
    const uint64_t *values;      // column to fill
    const uint8_t *validity;     // its bitmap (bit offset 0), never null
    const uint64_t *ref_values;  // FillLinear: reference column values / validity (null = all valid) / dtype
    const uint8_t *ref_validity;
    uint64_t *out_values;        // n rows
    uint8_t *out_validity;       // padded device bitmap (written whole words)
    int64_t n;
    int32_t is_int, ref_is_int;
};
size_t fill_scratch_bytes(int64_t n);
int launch_fill(int method, const FillLaunch &L, int64_t *scratch, cudaStream_t stream);
// IsColSorted over the valid rows (bowassertion.go:15-81): *flags |= 1 if some valid row < previous valid, 2 if >
int launch_sorted_flags(const uint64_t *values, const uint8_t *validity, int is_int, int64_t n, int64_t *scratch,
                        int32_t *flags, cudaStream_t stream);

// ---- DropNils (fill.cu): bow.go:188-224 ----------------------------------------------------------------------------------
struct DropLaunch {
    int32_t ncols;
    int32_t _pad;
    int64_t n;
    const uint64_t *values[32];
    const uint8_t *validity[32];  // input bitmaps (null = no nulls)
    uint8_t selected[32];         // rows with a null in a selected column are dropped
    uint64_t *out_values[32];     // (second pass)
    uint8_t *out_validity[32];    // zero-initialised bitmaps of the columns whose nulls survive, else null
};
size_t drop_scratch_bytes(int64_t n);
int launch_drop_mark(const DropLaunch &L, void *scratch, cudaStream_t stream, int64_t **d_total);
int launch_drop_compact(const DropLaunch &L, void *scratch, cudaStream_t stream);

// ---- SortByCol (sort.cu): bowsort.go:10-47 --------------------------------------------------------------------------------
struct SortGather {
    int32_t ncols, key_col, key_is_int, _pad;
    int64_t n;
    const uint32_t *idx;          // final permutation: output row j = input row idx[j]
    const uint64_t *sorted_keys;  // order-preserving unsigned keys in output order
    const uint64_t *values[32];
    const uint8_t *validity[32];  // input bitmaps (null = no nulls)
    uint64_t *out_values[32];
    uint8_t *out_validity[32];    // padded device bitmaps of the columns that carry nulls, else null
};
size_t sort_scratch_bytes(int64_t n);
int launch_sort_prepare(const uint64_t *values, int is_int, int64_t n, void *scratch, cudaStream_t stream, int32_t *flags_host,
                        unsigned long long *hist_host);
int launch_sort_passes(int64_t n, void *scratch, const unsigned long long *hist_host, cudaStream_t stream,
                       const uint32_t **idx, const uint64_t **sorted_keys, int *npasses);
int launch_sort_gather(const SortGather &G, cudaStream_t stream);

// ---- synthetic generators (generate.cu) ----------------------------------------------------------
int launch_gen_regular(int64_t *time, int64_t n, int64_t row0, int64_t t0, int64_t step, cudaStream_t stream);
// BURSTY: per-window row counts (to be scanned in place with launch_exclusive_scan) and the time column
int launch_gen_bursty_counts(int64_t *cnt, int64_t nw, uint64_t seed, cudaStream_t stream);
int launch_gen_bursty_time(int64_t *time, int64_t n, int64_t row0, int64_t t0, int64_t interval, uint64_t seed,
                           const int64_t *off, int64_t nw, cudaStream_t stream);
int launch_gen_values(uint64_t *v, uint8_t *validity, int64_t n, int64_t row0, uint64_t seed, uint64_t col, int is_int,
                      uint32_t null_mod, cudaStream_t stream);

}  // namespace bowgpu
