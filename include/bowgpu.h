/*
 * bowgpu.h — C ABI of libbowgpu.so: Bow's interval-rolling path on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary.  The Go side of Metronlab/bow keeps its public API
 * (rolling.IntervalRolling -> Rolling.Interpolate -> Rolling.Aggregate, reference
 * rolling/rolling.go:14-29,60) and binds these entry points through cgo (see
 * INTEGRATION.md).  Every function is `extern "C"`, takes plain pointers and sizes,
 * returns an int32 status (0 == BOWGPU_OK) and never throws or aborts across the ABI.
 *
 * Data hand-off follows the Arrow columnar layout the reference already holds in memory
 * (arrow/go v8 array.Data: validity bitmap LSB-first, 8-byte little-endian values, element
 * offset for zero-copy slices — reference bow.go:279-283, bowbuffer.go:22-40).
 *
 * Threading: a ctx (and everything created from it) may be used by one thread at a time
 * (same contract as the reference, rolling/rolling.go:32,160,175).  Calls may arrive on any
 * OS thread: every entry point selects its CUDA device itself.
 *
 * There is NO CPU fallback anywhere behind this header.
 */
#ifndef BOWGPU_H
#define BOWGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BOWGPU_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------------------- */
enum {
    BOWGPU_OK = 0,
    BOWGPU_EINVAL = 1,         /* bad argument (interval <= 0: rolling.go:115-117; bad index ...) */
    BOWGPU_ETYPE = 2,          /* interval column is not Int64 (rolling.go:70-73); interpolation input type (interpolation.go:84-93) */
    BOWGPU_EFIRSTNULL = 3,     /* first time value is null (rolling.go:91-94) */
    BOWGPU_EPREVROW = 4,       /* Options.PrevRow must have exactly one row (rolling.go:130-141) */
    BOWGPU_ENOINTERVALCOL = 5, /* "must keep interval column" (aggregation.go:163-166, interpolation.go:49-51) */
    BOWGPU_ECAPACITY = 6,      /* caller-provided output too small */
    BOWGPU_EUNSORTED = 7,      /* time column is not sorted ascending (GPU path precondition) */
    BOWGPU_ENULLTIME = 8,      /* time column holds nulls (GPU path precondition) */
    BOWGPU_ECUDA = 9,          /* CUDA runtime / launch failure; see bowgpu_last_error */
    BOWGPU_ENOMEM = 10,        /* device or pinned host allocation failed */
    BOWGPU_EUNSUPPORTED = 11,  /* operator not executable on the device (custom Go closure, Bool/String column) */
    BOWGPU_EIO = 12            /* file cannot be read or is not well-formed Parquet (bowparquet.go:45-56 error paths) */
};

/* ---- column types: bow.Float64 / bow.Int64 (bowtypes.go:21-22) ---------------------- */
#define BOWGPU_FLOAT64 1
#define BOWGPU_INT64 2

/* ---- aggregation opcodes: one per built-in constructor of rolling/aggregation/ ------ */
enum {
    BOWGPU_AGG_WINDOW_START = 0,       /* aggregation.WindowStart           windowstart.go:8-13   */
    BOWGPU_AGG_COUNT = 1,              /* aggregation.Count                 count.go:8-20         */
    BOWGPU_AGG_SUM = 2,                /* aggregation.Sum                   sum.go:8-25           */
    BOWGPU_AGG_MEAN = 3,               /* aggregation.ArithmeticMean        arithmeticmean.go:8-30 */
    BOWGPU_AGG_MIN = 4,                /* aggregation.Min                   minmax.go:8-31        */
    BOWGPU_AGG_MAX = 5,                /* aggregation.Max                   minmax.go:33-56       */
    BOWGPU_AGG_FIRST = 6,              /* aggregation.First                 firstlast.go:8-21     */
    BOWGPU_AGG_LAST = 7,               /* aggregation.Last                  firstlast.go:23-36    */
    BOWGPU_AGG_INTEGRAL_STEP = 8,      /* aggregation.IntegralStep          integral.go:40-69     */
    BOWGPU_AGG_INTEGRAL_TRAPEZOID = 9, /* aggregation.IntegralTrapezoid     integral.go:8-38      */
    BOWGPU_AGG_WAVG_STEP = 10,         /* aggregation.WeightedAverageStep   weightedmean.go:8-20  */
    BOWGPU_AGG_WAVG_LINEAR = 11,       /* aggregation.WeightedAverageLinear weightedmean.go:22-34 */
    BOWGPU_AGG__COUNT = 12
};

/* ---- interpolation opcodes: rolling/interpolation/ ---------------------------------- */
enum {
    BOWGPU_INTERP_WINDOW_START = 0,  /* interpolation.WindowStart   windowstart.go:8-14   */
    BOWGPU_INTERP_LINEAR = 1,        /* interpolation.Linear        linear.go:8-38        */
    BOWGPU_INTERP_STEP_PREVIOUS = 2, /* interpolation.StepPrevious  stepprevious.go:8-26  */
    BOWGPU_INTERP_NONE = 3,          /* interpolation.None          none.go:8-14          */
    BOWGPU_INTERP_STEP_NEXT = 4      /* named by the north-star, NOT in the reference: the mirror image of StepPrevious over
                                      * the reference's own getter Bow.GetNextValues (bowgetters.go:111-123) — the value of
                                      * the first row at or after the window's first row whose time and value are both valid,
                                      * else nil.  Parity unpinned (no upstream implementation, no golden vectors). */
};

/* ---- memory spaces of the pointers inside a bowgpu_col / bowgpu_out_col --------------- */
#define BOWGPU_MEM_HOST 0   /* pageable or pinned host memory (Go heap, Arrow buffers) */
#define BOWGPU_MEM_DEVICE 1 /* device memory of the ctx's GPU (zero-copy, must outlive the frame) */

/* One Arrow array, borrowed for the duration of the call (host) or of the frame (device).
 * Replaces the reference's per-cell getters over array.Data (bowgetters.go:46-311). */
typedef struct bowgpu_col {
    const void *values;      /* 8-byte elements (int64 or float64), `offset + length` of them */
    const uint8_t *validity; /* LSB-first bitmap covering `offset + length` bits, or NULL = all valid */
    int64_t offset;          /* element offset into both buffers (Arrow slice, bow.go:279-283) */
    int64_t length;          /* number of rows */
    int64_t null_count;      /* < 0 = unknown */
    int32_t dtype;           /* BOWGPU_FLOAT64 | BOWGPU_INT64 */
    int32_t _pad;
} bowgpu_col;

/* One aggregation, i.e. one output column of Rolling.Aggregate (aggregation.go:40-50). */
typedef struct bowgpu_agg_spec {
    int32_t op;        /* BOWGPU_AGG_* */
    int32_t col;       /* input column index (ColAggregation.InputIndex, aggregation.go:70-76) */
    int32_t nfactors;  /* number of transformation.Factor applied in order (factor.go:7-20), <= 4 */
    int32_t _pad;
    double factors[4];
} bowgpu_agg_spec;

/* Caller-allocated destination of one output column: what bow.NewBuffer(W, typ) holds in the
 * reference (aggregation.go:198, bowbuffer.go:22-40).  values: 8*W bytes, validity: ceil(W/8)
 * bytes, unused trailing bits 0, null slots hold value 0. */
typedef struct bowgpu_out_col {
    void *values;
    uint8_t *validity;
    int32_t dtype; /* filled by the callee: ColAggregation.GetReturnType (aggregation.go:110-121) */
    int32_t _pad;
} bowgpu_out_col;

typedef struct bowgpu_ctx bowgpu_ctx;         /* one GPU + stream + scratch arena */
typedef struct bowgpu_frame bowgpu_frame;     /* device-resident Bow: columns of equal length */
typedef struct bowgpu_rolling bowgpu_rolling; /* intervalRolling (rolling.go:31-43) on a frame */

/* Timing record of the last aggregate / interpolate / bounds call (CUDA events on the ctx stream). */
typedef struct bowgpu_timing {
    float total_ms;      /* whole device-side chain of the call */
    float main_ms;       /* dominant streaming kernel only (segreduce / scatter / bounds) */
    int32_t launches;    /* kernels launched by the call */
    int32_t main_launches;
} bowgpu_timing;

/* ---- context -------------------------------------------------------------------------- */
int32_t bowgpu_abi_version(void);
/* stream: a cudaStream_t owned by the caller (e.g. torch's current stream), or NULL to let the ctx
 * create its own non-blocking stream. */
int32_t bowgpu_ctx_create(int32_t device, void *stream, bowgpu_ctx **out);
void bowgpu_ctx_destroy(bowgpu_ctx *ctx);
const char *bowgpu_last_error(const bowgpu_ctx *ctx);
const char *bowgpu_status_string(int32_t status);
int32_t bowgpu_ctx_synchronize(bowgpu_ctx *ctx);
/* enable: 0 off, 1 = events of the last call (total_ms, main_ms over all its main launches), 2 = accumulate over calls
 * until bowgpu_ctx_last_timing reads and resets: `launches` counts every kernel, main_ms / main_launches cover every
 * 4th launch of the dominant kernel only (events inside a timed loop are not free), total_ms is 0 */
int32_t bowgpu_ctx_enable_timing(bowgpu_ctx *ctx, int32_t enable);
int32_t bowgpu_ctx_last_timing(bowgpu_ctx *ctx, bowgpu_timing *out); /* synchronizes */
int32_t bowgpu_ctx_sm_count(const bowgpu_ctx *ctx);
/* Column buffers of frames come from a stream-ordered pool owned by the ctx; destroyed frames leave their
 * blocks cached for the next Interpolate / upload.  Returns the cached blocks to the driver (synchronizes). */
int32_t bowgpu_ctx_trim(bowgpu_ctx *ctx);

/* ---- frames ----------------------------------------------------------------------------- */
/* Copies `ncols` Arrow arrays into device memory (mem == BOWGPU_MEM_HOST: chunked pinned staging,
 * complete before return, so Go may release the buffers) or wraps device pointers zero-copy
 * (mem == BOWGPU_MEM_DEVICE; values must be 16-byte aligned at element `offset`).  Validity
 * bitmaps are re-aligned to bit offset 0 on the device.  Replaces the bow.Bow input of
 * rolling.IntervalRolling (rolling.go:60). */
int32_t bowgpu_frame_create(bowgpu_ctx *ctx, const bowgpu_col *cols, int32_t ncols, int32_t mem, bowgpu_frame **out);
void bowgpu_frame_destroy(bowgpu_frame *frame);
int64_t bowgpu_frame_num_rows(const bowgpu_frame *frame);
int32_t bowgpu_frame_num_cols(const bowgpu_frame *frame);
int32_t bowgpu_frame_col_dtype(const bowgpu_frame *frame, int32_t col);
/* has_validity: 1 if the column carries a bitmap on the device */
int32_t bowgpu_frame_col_has_validity(const bowgpu_frame *frame, int32_t col);
/* device addresses of a column (for device-side consumers such as the synthetic generators) */
int32_t bowgpu_frame_col_device_ptrs(const bowgpu_frame *frame, int32_t col, void **values, uint8_t **validity);
/* Copies the frame back into caller-allocated host buffers (values 8*n bytes, validity ceil(n/8)
 * bytes, bit offset 0).  outs[j].dtype is filled.  Counterpart of Rolling.Bow() (rolling.go:241). */
int32_t bowgpu_frame_download(const bowgpu_frame *frame, bowgpu_out_col *outs, int32_t ncols);
/* Same for rows [row0, row0 + nrows): the counterpart of Bow.NewSlice (bow.go:279-283) + download. */
int32_t bowgpu_frame_download_range(const bowgpu_frame *frame, int64_t row0, int64_t nrows, bowgpu_out_col *outs,
                                    int32_t ncols);
/* ---- Parquet ingest: bow.NewBowFromParquet (bowparquet.go:44-155) ------------------------
 * The host walks the footer and the page headers only; the bytes of the chosen column chunks go to the
 * device as they are and are decompressed (Snappy) and decoded there (definition levels -> validity
 * bitmap, PLAIN / dictionary values -> bow.NewBuffer layout).  Flat schemas, INT64 / DOUBLE leaves
 * (mapParquetToBowTypes, bowparquet.go:20-25; Boolean and String columns have no GPU type: dtype 0),
 * UNCOMPRESSED / SNAPPY, data pages v1 and v2.  `err` (may be null) receives the message of a failed open. */
typedef struct bowgpu_parquet bowgpu_parquet;
int32_t bowgpu_parquet_open(const char *path, bowgpu_parquet **out, char *err, int32_t err_cap);
void bowgpu_parquet_close(bowgpu_parquet *pq);
int64_t bowgpu_parquet_num_rows(const bowgpu_parquet *pq);
int32_t bowgpu_parquet_num_cols(const bowgpu_parquet *pq);                 /* leaf columns, schema order */
const char *bowgpu_parquet_col_name(const bowgpu_parquet *pq, int32_t col);
int32_t bowgpu_parquet_col_dtype(const bowgpu_parquet *pq, int32_t col);   /* BOWGPU_INT64 / BOWGPU_FLOAT64 / 0 */
int32_t bowgpu_parquet_col_physical_type(const bowgpu_parquet *pq, int32_t col); /* parquet.Type */
/* Reads the leaf columns cols[0..ncols) (every one must have a GPU type: else BOWGPU_ETYPE) into a new
 * device-resident frame; column j of the frame = cols[j].  Malformed page data -> BOWGPU_EIO. */
int32_t bowgpu_parquet_read(bowgpu_ctx *ctx, const bowgpu_parquet *pq, const int32_t *cols, int32_t ncols,
                            bowgpu_frame **out);
/* The page walk alone (host only, no device work): out4 = {pages, bytes uploaded, bytes of uncompressed
 * scratch, dictionary-index entries} for the chosen columns.  Diagnostic; also what sizes the device buffers. */
int32_t bowgpu_parquet_plan(const bowgpu_parquet *pq, const int32_t *cols, int32_t ncols, int64_t *out4, char *err,
                            int32_t err_cap);
/* null count of a device column (-1: bad index) */
int64_t bowgpu_frame_col_null_count(const bowgpu_frame *frame, int32_t col);

/* Device-side synthetic generators for the BASELINE.json configs.  Deterministic in
 * (seed, column, row): any sub-range can be regenerated.  kind: see BOWGPU_GEN_*. */
#define BOWGPU_GEN_REGULAR 0 /* t[i] = t0 + (row0+i)*step ; float64 v in [0,1) ; optional nulls */
#define BOWGPU_GEN_BURSTY 1  /* config 4: window sizes 0..1e6 rows, see DESIGN.md */
typedef struct bowgpu_gen_spec {
    int32_t kind;
    int32_t ncols;        /* value columns (time column is added as column 0) */
    int64_t nrows;
    int64_t row0;         /* global index of the first generated row (range-partitioned shards) */
    int64_t t0;           /* time of global row 0 */
    int64_t step;         /* REGULAR: time step; BURSTY: window interval used for the pattern */
    uint64_t seed;
    uint32_t null_mask;   /* bit c set: value column c has nulls */
    uint32_t int_mask;    /* bit c set: value column c is int64 (u % 2^20) instead of float64 */
    uint32_t null_mod;    /* value (c, i) is null iff u(seed+1, c, i) % null_mod == 0 (10 -> 10 % nulls) */
    uint32_t _pad;
} bowgpu_gen_spec;
int32_t bowgpu_frame_generate(bowgpu_ctx *ctx, const bowgpu_gen_spec *spec, bowgpu_frame **out);

/* ---- rolling ------------------------------------------------------------------------------- */
/* newIntervalRolling (rolling.go:69-112): validates, normalises the offset (rolling.go:114-128),
 * computes the first window start and the number of windows (rolling.go:143-154).
 * prev_row: NULL or `ncols` one-row HOST columns (Options.PrevRow, rolling.go:52). */
int32_t bowgpu_rolling_create(bowgpu_frame *frame, int32_t time_col, int64_t interval, int64_t offset,
                              int32_t inclusive, const bowgpu_col *prev_row, bowgpu_rolling **out);
/* Range-partitioned variant (multi-GPU): the frame is one shard whose first row lies on or
 * after the window start `s0` of its first owned window; exactly `num_windows` windows of the GLOBAL
 * lattice S_k = s0 + k*interval are produced (trailing empty windows included), and rows at or after
 * the end of the last one (the halo) only serve as the inclusive row of that window. */
int32_t bowgpu_rolling_create_shard(bowgpu_frame *frame, int32_t time_col, int64_t interval, int64_t s0,
                                    int64_t num_windows, int32_t inclusive, const bowgpu_col *prev_row,
                                    bowgpu_rolling **out);
void bowgpu_rolling_destroy(bowgpu_rolling *r);
int64_t bowgpu_rolling_num_windows(const bowgpu_rolling *r);        /* Rolling.NumWindows, rolling.go:156 */
int64_t bowgpu_rolling_first_window_start(const bowgpu_rolling *r); /* s0 */
int32_t bowgpu_rolling_inclusive(const bowgpu_rolling *r);
/* Rows whose time lies before the first window start (only with negative timestamps, rolling.go:96-99:
 * Go's truncating division can leave s0 > t[0]).  The reference keeps them inside window 0's slice iff
 * that window holds a row of its own (rolling.go:189-229); *kept says which. */
int64_t bowgpu_rolling_early_rows(const bowgpu_rolling *r, int32_t *kept);

/* Window boundaries of every window in one pass (replaces the HasNext/Next loop,
 * rolling.go:162-239).  first[k] = Window.FirstIndex for k < W (lower bound of S_k; 0 for k == 0),
 * first[W] = n;  window k holds rows [first[k], first[k+1]) plus, when inclusive[k] is set, the row
 * first[k+1] (Window.IsInclusive).  Both outputs are host buffers: first = W+1 int64,
 * inclusive_bitmap = ceil(W/8) bytes (may be NULL).  Also verifies sortedness / non-null time. */
int32_t bowgpu_rolling_bounds(bowgpu_rolling *r, int64_t *first, uint8_t *inclusive_bitmap);

/* Rolling.Aggregate (aggregation.go:123-238) for built-in aggregations.  outs[j] receives output
 * column j (W entries).  mem says where outs[j].values / validity live. */
int32_t bowgpu_rolling_aggregate(bowgpu_rolling *r, const bowgpu_agg_spec *specs, int32_t nspecs,
                                 bowgpu_out_col *outs, int32_t mem);
/* GetReturnType (aggregation.go:110-121) of a built-in aggregation for an input column type */
int32_t bowgpu_agg_return_type(int32_t op, int32_t input_dtype);
int32_t bowgpu_agg_needs_inclusive(int32_t op); /* integral.go:9, weightedmean.go:24 */

/* aggregation.Aggregate(b, intervalColName, aggrs...) — every aggregation over ONE window holding the whole Bow
 * (rolling/aggregation/whole.go:12-93).  outs[j] receives ONE entry (none when the frame has no rows): 8 value
 * bytes + 1 validity byte.  Window.FirstValue / LastValue are the first / last times taken through float64
 * (whole.go:54-70); the return type uses the INPUT column as iterator type (whole.go:44-46) and values are stored
 * with SetOrDropStrict (whole.go:87), so WindowStart of a Float64 column is null.  The interval column need not be
 * among the aggregated columns.  Same GPU-path precondition: sorted, non-null Int64 interval column. */
int32_t bowgpu_frame_aggregate_whole(bowgpu_frame *frame, int32_t time_col, const bowgpu_agg_spec *specs,
                                     int32_t nspecs, bowgpu_out_col *outs, int32_t mem);

/* ---- whole-column fills (bowfill.go) ------------------------------------------------------------------------------ */
enum {
    BOWGPU_FILL_PREVIOUS = 0, /* Bow.FillPrevious  bowfill.go:160-164 (LOCF) */
    BOWGPU_FILL_NEXT = 1,     /* Bow.FillNext      bowfill.go:154-158 (NOCB) */
    BOWGPU_FILL_MEAN = 2,     /* Bow.FillMean      bowfill.go:104-152 (Int64 results rounded half away from zero) */
    BOWGPU_FILL_LINEAR = 3    /* Bow.FillLinear    bowfill.go:14-102 */
};
/* FillPrevious / FillNext / FillMean of the columns cols[0..ncols) (ncols == 0: every column, selectCols
 * bowfill.go:266-288).  Every null row looks at the nearest valid rows of the ORIGINAL column.  The result is a new
 * device-resident frame (columns that are not filled are copied). */
int32_t bowgpu_frame_fill(bowgpu_frame *frame, int32_t method, const int32_t *cols, int32_t ncols, bowgpu_frame **out);
/* Bow.FillLinear(refColIndex, toFillColIndex): the reference column must be sorted (ascending or descending, nulls
 * skipped: IsColSorted, bowassertion.go:15-81) else BOWGPU_EUNSORTED; an all-null reference column or a column
 * without nulls returns a copy. */
int32_t bowgpu_frame_fill_linear(bowgpu_frame *frame, int32_t ref_col, int32_t tofill_col, bowgpu_frame **out);

/* Bow.DropNils(colIndices...) (bow.go:188-224): drops every row holding a nil in one of cols[0..ncols) (ncols == 0:
 * any column).  What a caller does to a time column with nils before handing it to the rolling path. */
int32_t bowgpu_frame_drop_nils(bowgpu_frame *frame, const int32_t *cols, int32_t ncols, bowgpu_frame **out);
/* Bow.IsColSorted(colIndex) (bowassertion.go:15-81): *sorted = 1 iff the non-nil values are ascending or descending
 * (ties allowed); 0 for an empty column. */
int32_t bowgpu_frame_is_col_sorted(bowgpu_frame *frame, int32_t col, int32_t *sorted);
/* Bow.SortByCol(colIndex) (bowsort.go:10-47): a new frame with the rows in ascending order of column `col` (int64 or
 * float64, compared with `<` like Buffer.Less, bowbuffer.go:126-131).  *out = NULL with BOWGPU_OK when the column is
 * already sorted (sort.IsSorted; the reference returns the same Bow, bowsort.go:18-21) — keep using `frame`.  A sort column
 * with nils is BOWGPU_EINVAL (bowsort.go:11-15).  Equal keys keep their input order (the reference's sort.Sort leaves it
 * unspecified; its golden vectors, bowsort_test.go:133-157, show this order).  A NaN in an unsorted float64 sort column is
 * BOWGPU_EUNSUPPORTED (no order under `<`; undefined upstream). */
int32_t bowgpu_frame_sort_by_col(bowgpu_frame *frame, int32_t col, bowgpu_frame **out);

/* Rolling.Interpolate (interpolation.go:30-161).  ops[j] is the interpolation of column j (the
 * reference matches columns by position, bowappend.go:28-47, so nops must equal the number of
 * columns).  The result is a new device-resident frame with n_out rows. */
int32_t bowgpu_rolling_interpolate(bowgpu_rolling *r, const int32_t *ops, int32_t nops, bowgpu_frame **out_frame,
                                   int64_t *n_out);

/* Rolling.Interpolate(ops...).Aggregate(specs...) in one call (interpolation.go:30-161 then aggregation.go:123-238):
 * the result equals bowgpu_rolling_interpolate followed by bowgpu_rolling_aggregate on the interpolated frame, but
 * the interpolated frame is not materialised (its synthetic window-start rows are injected into the reduction).
 * ops / nops as in bowgpu_rolling_interpolate, specs / outs / mem as in bowgpu_rolling_aggregate; outs hold
 * bowgpu_rolling_num_windows(r) entries.  This is what the Go shim calls when Aggregate follows Interpolate and the
 * interpolated Bow itself is never asked for. */
int32_t bowgpu_rolling_interpolate_aggregate(bowgpu_rolling *r, const int32_t *ops, int32_t nops,
                                             const bowgpu_agg_spec *specs, int32_t nspecs, bowgpu_out_col *outs,
                                             int32_t mem);

/* rolling.IntervalRolling(b, col, interval, Options{Offset, Inclusive}).Aggregate(aggrs...) for a Bow in HOST memory, in
 * one call (rolling.go:60 + aggregation.go:123-238): the window range is processed in chunks by a few worker contexts
 * whose uploads, kernels and downloads overlap, and only the columns the aggregations read cross the bus.  outs[j] are
 * host buffers with room for `out_capacity` windows (values 8 bytes each, validity ceil(capacity / 8) bytes);
 * *num_windows receives W (BOWGPU_ECAPACITY when it exceeds the capacity; the Go side knows W from countWindows,
 * rolling.go:143-154).  Same results as bowgpu_frame_create + bowgpu_rolling_create + bowgpu_rolling_aggregate. */
int32_t bowgpu_aggregate_host(bowgpu_ctx *ctx, const bowgpu_col *cols, int32_t ncols, int32_t time_col, int64_t interval,
                              int64_t offset, int32_t inclusive, const bowgpu_agg_spec *specs, int32_t nspecs,
                              bowgpu_out_col *outs, int64_t out_capacity, int64_t *num_windows);

/* Options of the one-shot host calls; zero-initialise for the defaults.  ONE process drives every GPU it lists: this is
 * what a Go program calling rolling.IntervalRolling(...) (rolling.go:60 - one call, one process) binds for multi-GPU runs;
 * the window range is range-partitioned over the devices chunk by chunk, results land in the caller's buffers, no
 * collective anywhere. */
typedef struct bowgpu_host_opts {
    const int32_t *devices;     /* GPUs to spread the chunks over; NULL / ndevices == 0: the ctx's own device */
    int32_t ndevices;
    int32_t workers_per_device; /* concurrent upload -> kernels -> download pipelines per GPU; 0 = default (3) */
    int64_t chunk_rows;         /* rows per chunk; 0 = default (~12 M).  Inputs below two chunks take the plain path */
    int32_t shard;              /* 1: the rows are ONE range-partitioned shard of a larger Bow (one process per GPU): */
    int32_t _pad;
    int64_t s0;                 /*    start of the first window it owns on the global lattice S_k = s0 + k*interval ...  */
    int64_t num_windows;        /*    ... and how many windows it owns (see bowgpu_rolling_create_shard) */
} bowgpu_host_opts;
int32_t bowgpu_aggregate_host_ex(bowgpu_ctx *ctx, const bowgpu_col *cols, int32_t ncols, int32_t time_col, int64_t interval,
                                 int64_t offset, int32_t inclusive, const bowgpu_agg_spec *specs, int32_t nspecs,
                                 bowgpu_out_col *outs, int64_t out_capacity, int64_t *num_windows,
                                 const bowgpu_host_opts *opts);

/* rolling.IntervalRolling(b, col, interval, Options{Offset, PrevRow}).Interpolate(ops...).Aggregate(aggrs...) for a Bow in
 * HOST memory, in one call (rolling.go:60, interpolation.go:30-69, aggregation.go:123-145): the pipelined counterpart of
 * bowgpu_rolling_interpolate_aggregate.  Every chunk of the window range ships its halo - back to the last valid row of
 * each interpolated column, one extra window and forward to the next valid row (what interpolation.Linear / StepPrevious
 * look at, linear.go:14-27; for the first chunk Options.PrevRow plays the left part) - so the chunks are independent.
 * prev_row: NULL or `ncols` one-row host columns.  ops / nops as in bowgpu_rolling_interpolate. */
int32_t bowgpu_interpolate_aggregate_host(bowgpu_ctx *ctx, const bowgpu_col *cols, int32_t ncols, int32_t time_col,
                                          int64_t interval, int64_t offset, const bowgpu_col *prev_row, const int32_t *ops,
                                          int32_t nops, const bowgpu_agg_spec *specs, int32_t nspecs, bowgpu_out_col *outs,
                                          int64_t out_capacity, int64_t *num_windows, const bowgpu_host_opts *opts);

#ifdef __cplusplus
}
#endif
#endif /* BOWGPU_H */
