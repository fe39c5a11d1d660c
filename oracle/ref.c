/*
 * TEST INFRASTRUCTURE ONLY — sequential C restatement of Bow's interval-rolling path.
 *
 * This file restates, in plain C over raw Arrow buffers, the algorithm of the
 * reference (Metronlab/bow, pure Go): the window iterator, the Aggregate and
 * Interpolate drivers and every in-scope aggregation / interpolation closure.
 * Each function cites the reference file:line it follows.  It keeps the
 * reference's structure (one full window iteration per aggregation column, one
 * sequential closure loop per window, left-to-right float accumulation) so it can
 * serve both as the parity checker for the CUDA path and as the timed CPU baseline
 * ("C restatement of the reference algorithm, 1 core").
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (bow_b200/) never does.
 *
 * Parity status: PINNED — validated against oracle/literal.py on randomised
 * inputs and against every in-scope golden vector of the reference's own tests
 * (tests/test_oracle_ref_c.py).  The Go reference itself cannot be built here
 * (no Go toolchain in this image).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared (see oracle/Makefile).
 * -ffp-contract=off matters: Go on amd64 never fuses multiply-add.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <math.h>

#define BOWREF_FLOAT64 1 /* bow.Float64, bowtypes.go:21 */
#define BOWREF_INT64 2   /* bow.Int64,   bowtypes.go:22 */

/* aggregation opcodes (shared numbering with include/bowgpu.h) */
enum {
    BOWREF_AGG_WINDOW_START = 0,
    BOWREF_AGG_COUNT = 1,
    BOWREF_AGG_SUM = 2,
    BOWREF_AGG_MEAN = 3,
    BOWREF_AGG_MIN = 4,
    BOWREF_AGG_MAX = 5,
    BOWREF_AGG_FIRST = 6,
    BOWREF_AGG_LAST = 7,
    BOWREF_AGG_INTEGRAL_STEP = 8,
    BOWREF_AGG_INTEGRAL_TRAPEZOID = 9,
    BOWREF_AGG_WAVG_STEP = 10,
    BOWREF_AGG_WAVG_LINEAR = 11,
};
enum {
    BOWREF_INTERP_WINDOW_START = 0,
    BOWREF_INTERP_LINEAR = 1,
    BOWREF_INTERP_STEP_PREVIOUS = 2,
    BOWREF_INTERP_NONE = 3,
    BOWREF_INTERP_STEP_NEXT = 4, /* not upstream (north-star): StepPrevious mirrored over GetNextValues */
};
enum {
    BOWREF_OK = 0,
    BOWREF_EINVAL = 1,
    BOWREF_ETYPE = 2,
    BOWREF_EFIRSTNULL = 3,
    BOWREF_EPREVROW = 4,
    BOWREF_ENOINTERVALCOL = 5,
    BOWREF_ECAPACITY = 6,
};

typedef struct {
    const void *values;      /* 8-byte elements */
    const uint8_t *validity; /* LSB-first bitmap, may be NULL (= all valid) */
    int64_t offset;          /* element offset into both buffers (Arrow slice) */
    int64_t length;
    int32_t dtype;
    int32_t _pad;
} bowref_col;

typedef struct {
    const bowref_col *cols;
    int32_t ncols;
    int32_t time_col;
    int64_t nrows;
    int64_t interval;
    int64_t offset;
    int32_t inclusive;
    int32_t _pad;
    const bowref_col *prev_row; /* NULL or ncols one-row columns (Options.PrevRow) */
    int64_t num_windows;
    /* iterator cursor, rolling.go:39-41 */
    int64_t curr_window_first_value;
    int64_t curr_row_index;
    int64_t curr_window_index;
} bowref_rolling;

typedef struct {
    int64_t lo, hi;       /* window.Bow == full.NewSlice(lo, hi); lo == hi: empty slice */
    int64_t first_index;  /* Window.FirstIndex */
    int64_t first_value;  /* Window.FirstValue */
    int64_t last_value;   /* Window.LastValue  */
    int32_t is_inclusive; /* Window.IsInclusive */
    int32_t _pad;
} bowref_window;

typedef struct {
    int32_t op;
    int32_t col;
    int32_t nfactors; /* number of transformation.Factor applied, in order */
    int32_t _pad;
    double factors[4];
} bowref_agg_spec;

typedef struct {
    void *values;      /* int64/float64 [num_windows] */
    uint8_t *validity; /* ceil(num_windows/8) bytes */
    int32_t dtype;     /* filled by the callee */
    int32_t _pad;
} bowref_out_col;

/* ---- cell getters (bowgetters.go:46-311) -------------------------------- */
static inline int col_valid(const bowref_col *c, int64_t i) {
    if (!c->validity) return 1;
    int64_t b = c->offset + i;
    return (c->validity[b >> 3] >> (b & 7)) & 1;
}
static inline int64_t col_raw(const bowref_col *c, int64_t i) { return ((const int64_t *)c->values)[c->offset + i]; }
static inline int64_t f64_to_i64(double x) { /* Go int64(float64) on amd64 (CVTTSD2SI) */
    if (!(x >= -9223372036854775808.0 && x < 9223372036854775808.0)) return INT64_MIN;
    return (int64_t)x;
}
static inline double raw_as_f64(int64_t raw) {
    double d;
    memcpy(&d, &raw, 8);
    return d;
}
static inline int64_t f64_as_raw(double d) {
    int64_t r;
    memcpy(&r, &d, 8);
    return r;
}
/* GetFloat64, bowgetters.go:218-247 (value only; validity checked by caller) */
static inline double col_f64(const bowref_col *c, int64_t i) {
    int64_t raw = col_raw(c, i);
    return c->dtype == BOWREF_INT64 ? (double)raw : raw_as_f64(raw);
}
/* GetInt64, bowgetters.go:155-184 */
static inline int64_t col_i64(const bowref_col *c, int64_t i) {
    int64_t raw = col_raw(c, i);
    return c->dtype == BOWREF_INT64 ? raw : f64_to_i64(raw_as_f64(raw));
}

/* ---- rolling.go:114-154 ---------------------------------------------------- */
static int enforce_interval_and_offset(int64_t interval, int64_t *offset) {
    if (interval <= 0) return BOWREF_EINVAL;
    if (*offset >= interval || *offset <= -interval) *offset = *offset % interval; /* C99 '%' == Go '%' */
    if (*offset < 0) *offset += interval;
    return BOWREF_OK;
}

static int64_t count_windows(const bowref_rolling *r, int64_t first_window_start) {
    if (r->nrows == 0) return 0;
    const bowref_col *t = &r->cols[r->time_col];
    int64_t row = r->nrows - 1; /* GetPrevInt64, bowgetters.go:189-199 */
    while (row >= 0 && !col_valid(t, row)) row--;
    if (row < 0) return 0;
    int64_t last = col_i64(t, row);
    if (first_window_start > last) return 0;
    return (int64_t)(((uint64_t)last - (uint64_t)first_window_start) / (uint64_t)r->interval) + 1;
}

/* newIntervalRolling, rolling.go:69-112 */
int bowref_rolling_init(bowref_rolling *r, const bowref_col *cols, int32_t ncols, int32_t time_col, int64_t interval,
                        int64_t offset, int32_t inclusive, const bowref_col *prev_row) {
    memset(r, 0, sizeof(*r));
    if (ncols <= 0 || time_col < 0 || time_col >= ncols) return BOWREF_EINVAL;
    if (cols[time_col].dtype != BOWREF_INT64) return BOWREF_ETYPE;
    int rc = enforce_interval_and_offset(interval, &offset);
    if (rc) return rc;
    if (prev_row) { /* enforcePrevRow, rolling.go:130-141 */
        if (prev_row[0].length == 0)
            prev_row = NULL;
        else if (prev_row[0].length != 1)
            return BOWREF_EPREVROW;
    }
    r->cols = cols;
    r->ncols = ncols;
    r->time_col = time_col;
    r->nrows = cols[time_col].length;
    r->interval = interval;
    r->offset = offset;
    r->inclusive = inclusive;
    r->prev_row = prev_row;
    int64_t first = 0;
    if (r->nrows > 0) {
        const bowref_col *t = &cols[time_col];
        if (!col_valid(t, 0)) return BOWREF_EFIRSTNULL;
        int64_t v = col_i64(t, 0);
        first = (int64_t)((uint64_t)((v / interval) * interval) + (uint64_t)offset);
        if (first > v) first = (int64_t)((uint64_t)first - (uint64_t)interval);
    }
    r->curr_window_first_value = first;
    r->num_windows = count_windows(r, first);
    return BOWREF_OK;
}

/* HasNext, rolling.go:162-173 */
int bowref_has_next(const bowref_rolling *r) {
    if (r->curr_row_index >= r->nrows) return 0;
    const bowref_col *t = &r->cols[r->time_col];
    if (!col_valid(t, r->nrows - 1)) return 0;
    return r->curr_window_first_value <= col_i64(t, r->nrows - 1);
}

/* Next, rolling.go:177-239.  Returns 1 and fills *w, or 0 when exhausted. */
int bowref_next(bowref_rolling *r, bowref_window *w, int64_t *window_index) {
    if (!bowref_has_next(r)) {
        if (window_index) *window_index = r->curr_window_index;
        return 0;
    }
    const bowref_col *t = &r->cols[r->time_col];
    int64_t first_value = r->curr_window_first_value;
    int64_t last_value = (int64_t)((uint64_t)first_value + (uint64_t)r->interval);
    int is_inclusive = 0;
    int64_t first_row = r->curr_row_index, last_row = -1, row;
    for (row = first_row; row < r->nrows; row++) {
        if (!col_valid(t, row)) continue;
        int64_t val = col_i64(t, row);
        if (val < first_value) continue;
        if (val > last_value) break;
        if (val == last_value) {
            if (is_inclusive) break;
            if (!r->inclusive) break;
            is_inclusive = 1;
        }
        last_row = row;
    }
    r->curr_row_index = is_inclusive ? row - 1 : row;
    r->curr_window_first_value = last_value;
    if (window_index) *window_index = r->curr_window_index;
    r->curr_window_index++;
    if (last_row == -1) {
        w->lo = w->hi = first_row; /* NewEmptySlice */
    } else {
        w->lo = first_row;
        w->hi = last_row + 1;
    }
    w->first_index = first_row;
    w->first_value = first_value;
    w->last_value = last_value;
    w->is_inclusive = is_inclusive;
    return 1;
}

/* ---- aggregation closures (rolling/aggregation/ *.go) ---------------------- */
typedef struct {
    int is_nil;
    int is_int; /* dynamic type of the interface{} value: int64 or float64 */
    int64_t i;
    double f;
} ref_value;

static ref_value nil_value(void) {
    ref_value v = {1, 0, 0, 0.0};
    return v;
}
static ref_value float_value(double f) {
    ref_value v = {0, 0, 0, f};
    return v;
}
static ref_value int_value(int64_t i) {
    ref_value v = {0, 1, i, 0.0};
    return v;
}

/* GetNextFloat64s(timeCol, col, row) restricted to the slice [lo, hi): next row >= row valid in both */
static int64_t next_both_valid(const bowref_col *t, const bowref_col *c, int64_t row, int64_t hi) {
    for (; row < hi; row++)
        if (col_valid(t, row) && col_valid(c, row)) return row;
    return -1;
}

/* integral.go:8-38 */
static ref_value integral_trapezoid(const bowref_rolling *r, int col, const bowref_window *w) {
    if (w->hi == w->lo) return nil_value();
    const bowref_col *t = &r->cols[r->time_col], *c = &r->cols[col];
    double sum = 0.0;
    int ok = 0;
    int64_t row = next_both_valid(t, c, w->lo, w->hi);
    if (row < 0) return nil_value();
    double t0 = col_f64(t, row), v0 = col_f64(c, row);
    while (row >= 0) {
        int64_t nxt = next_both_valid(t, c, row + 1, w->hi);
        if (nxt < 0) break;
        double t1 = col_f64(t, nxt), v1 = col_f64(c, nxt);
        sum += (v0 + v1) / 2 * (t1 - t0);
        ok = 1;
        t0 = t1;
        v0 = v1;
        row = nxt;
    }
    return ok ? float_value(sum) : nil_value();
}

/* integral.go:40-69 */
static ref_value integral_step(const bowref_rolling *r, int col, const bowref_window *w) {
    if (w->hi == w->lo) return nil_value();
    const bowref_col *t = &r->cols[r->time_col], *c = &r->cols[col];
    double sum = 0.0;
    int ok = 0;
    int64_t row = next_both_valid(t, c, w->lo, w->hi);
    double t0 = 0, v0 = 0;
    if (row >= 0) {
        t0 = col_f64(t, row);
        v0 = col_f64(c, row);
    }
    while (row >= 0) {
        int64_t nxt = next_both_valid(t, c, row + 1, w->hi);
        double t1, v1 = 0;
        if (nxt < 0) {
            t1 = (double)w->last_value;
        } else {
            t1 = col_f64(t, nxt);
            v1 = col_f64(c, nxt);
        }
        sum += v0 * (t1 - t0);
        ok = 1;
        if (nxt < 0) break;
        t0 = t1;
        v0 = v1;
        row = nxt;
    }
    return ok ? float_value(sum) : nil_value();
}

static int agg_needs_inclusive(int op) { /* integral.go:9, weightedmean.go:24 */
    return op == BOWREF_AGG_INTEGRAL_TRAPEZOID || op == BOWREF_AGG_WAVG_LINEAR;
}

static ref_value agg_closure(const bowref_rolling *r, int op, int col, const bowref_window *w) {
    const bowref_col *c = &r->cols[col];
    int64_t n = w->hi - w->lo;
    switch (op) {
    case BOWREF_AGG_WINDOW_START: /* windowstart.go:8-13 */
        return int_value(w->first_value);
    case BOWREF_AGG_COUNT: { /* count.go:8-20 */
        int64_t count = 0;
        for (int64_t i = w->lo; i < w->hi; i++)
            if (col_valid(c, i)) count++;
        return int_value(count);
    }
    case BOWREF_AGG_SUM: { /* sum.go:8-25 */
        if (n == 0) return float_value(0.0);
        double sum = 0.0;
        for (int64_t i = w->lo; i < w->hi; i++) {
            if (!col_valid(c, i)) continue;
            sum += col_f64(c, i);
        }
        return float_value(sum);
    }
    case BOWREF_AGG_MEAN: { /* arithmeticmean.go:8-30 */
        if (n == 0) return nil_value();
        double sum = 0.0;
        int64_t count = 0;
        for (int64_t i = w->lo; i < w->hi; i++) {
            if (!col_valid(c, i)) continue;
            sum += col_f64(c, i);
            count++;
        }
        if (count == 0) return nil_value();
        return float_value(sum / (double)count);
    }
    case BOWREF_AGG_MIN:   /* minmax.go:8-31 */
    case BOWREF_AGG_MAX: { /* minmax.go:33-56 */
        if (n == 0) return nil_value();
        int have = 0;
        double m = 0.0;
        for (int64_t i = w->lo; i < w->hi; i++) {
            if (!col_valid(c, i)) continue;
            double v = col_f64(c, i);
            if (have) {
                if (op == BOWREF_AGG_MIN ? (v < m) : (v > m)) m = v;
                continue;
            }
            m = v;
            have = 1;
        }
        return have ? float_value(m) : nil_value();
    }
    case BOWREF_AGG_FIRST: { /* firstlast.go:8-21 */
        if (n == 0) return nil_value();
        for (int64_t i = w->lo; i < w->hi; i++)
            if (col_valid(c, i))
                return c->dtype == BOWREF_INT64 ? int_value(col_raw(c, i)) : float_value(raw_as_f64(col_raw(c, i)));
        return nil_value();
    }
    case BOWREF_AGG_LAST: { /* firstlast.go:23-36 */
        if (n == 0) return nil_value();
        for (int64_t i = w->hi - 1; i >= w->lo; i--)
            if (col_valid(c, i))
                return c->dtype == BOWREF_INT64 ? int_value(col_raw(c, i)) : float_value(raw_as_f64(col_raw(c, i)));
        return nil_value();
    }
    case BOWREF_AGG_INTEGRAL_STEP:
        return integral_step(r, col, w);
    case BOWREF_AGG_INTEGRAL_TRAPEZOID:
        return integral_trapezoid(r, col, w);
    case BOWREF_AGG_WAVG_STEP:     /* weightedmean.go:8-20 */
    case BOWREF_AGG_WAVG_LINEAR: { /* weightedmean.go:22-34 */
        ref_value v = op == BOWREF_AGG_WAVG_STEP ? integral_step(r, col, w) : integral_trapezoid(r, col, w);
        if (v.is_nil) return v;
        double wide = (double)(int64_t)((uint64_t)w->last_value - (uint64_t)w->first_value);
        return float_value(v.f / wide);
    }
    }
    return nil_value();
}

/* colAggregation.GetReturnType, aggregation.go:110-121 + constructor types */
int32_t bowref_agg_return_type(int op, int32_t input_type) {
    switch (op) {
    case BOWREF_AGG_WINDOW_START: return BOWREF_INT64; /* IteratorDependent; iterator column is Int64 */
    case BOWREF_AGG_COUNT: return BOWREF_INT64;
    case BOWREF_AGG_FIRST:
    case BOWREF_AGG_LAST: return input_type; /* InputDependent */
    default: return BOWREF_FLOAT64;
    }
}

/* transformation.Factor, factor.go:7-20 */
static ref_value apply_factor(ref_value v, double n) {
    if (v.is_nil) return v;
    if (v.is_int) return int_value(f64_to_i64((double)v.i * n));
    return float_value(v.f * n);
}

/* Buffer.SetOrDrop, bowbuffer.go:60-80 with Type.Convert (bowconvert.go:11-73) */
static void set_or_drop(void *values, uint8_t *validity, int32_t dtype, int64_t i, ref_value v) {
    if (v.is_nil) return; /* aggregation.go:223-225: slot stays zero / null */
    if (dtype == BOWREF_INT64)
        ((int64_t *)values)[i] = v.is_int ? v.i : f64_to_i64(v.f);
    else
        ((double *)values)[i] = v.is_int ? (double)v.i : v.f;
    validity[i >> 3] |= (uint8_t)(1u << (i & 7));
}

/* Aggregate / aggregateWindows, aggregation.go:123-238.
 * One full window iteration per aggregation, exactly like the reference.
 * outs[j].values / validity are caller allocated for r->num_windows entries. */
int bowref_aggregate(const bowref_rolling *r, const bowref_agg_spec *specs, int32_t nspecs, bowref_out_col *outs) {
    if (nspecs <= 0) return BOWREF_EINVAL;
    bowref_rolling base = *r;
    int keeps_interval = 0;
    for (int j = 0; j < nspecs; j++) { /* validateAggregation :171-188 */
        if (specs[j].col < 0 || specs[j].col >= r->ncols) return BOWREF_EINVAL;
        if (agg_needs_inclusive(specs[j].op)) base.inclusive = 1;
        if (specs[j].col == r->time_col) keeps_interval = 1;
    }
    if (!keeps_interval) return BOWREF_ENOINTERVALCOL;
    for (int j = 0; j < nspecs; j++) {
        bowref_rolling rc = base;
        const bowref_agg_spec *a = &specs[j];
        int32_t typ = bowref_agg_return_type(a->op, r->cols[a->col].dtype);
        outs[j].dtype = typ;
        memset(outs[j].values, 0, (size_t)rc.num_windows * 8); /* bow.NewBuffer, bowbuffer.go:22-40 */
        memset(outs[j].validity, 0, (size_t)((rc.num_windows + 7) / 8));
        bowref_window w;
        int64_t wi;
        while (bowref_next(&rc, &w, &wi)) {
            if (!agg_needs_inclusive(a->op) && w.is_inclusive) { /* UnsetInclusive, window.go:23-31 */
                w.is_inclusive = 0;
                w.hi -= 1;
            }
            ref_value v = agg_closure(&rc, a->op, a->col, &w);
            for (int k = 0; k < a->nfactors; k++) v = apply_factor(v, a->factors[k]);
            set_or_drop(outs[j].values, outs[j].validity, typ, wi, v);
        }
    }
    return BOWREF_OK;
}

/* aggregation.Aggregate (whole frame, ONE window), rolling/aggregation/whole.go:12-93.
 * The window holds every row of the Bow; FirstValue / LastValue are the first / last non-null times taken
 * through float64 (GetNextFloat64 / GetPrevFloat64, whole.go:54-62; -1 when the column has none); the return
 * type is resolved with the INPUT column as iterator type (whole.go:44-46) and the value is stored with
 * SetOrDropStrict (whole.go:87, bowbuffer.go:82-104): a dynamic type different from the buffer type gives null.
 * outs[j].values / validity hold 1 entry (0 entries when the Bow has no rows). */
static int32_t whole_return_type(int op, int32_t input_type) {
    switch (op) {
    case BOWREF_AGG_WINDOW_START: return input_type; /* IteratorDependent, iterator := input column */
    case BOWREF_AGG_COUNT: return BOWREF_INT64;
    case BOWREF_AGG_FIRST:
    case BOWREF_AGG_LAST: return input_type;
    default: return BOWREF_FLOAT64;
    }
}
int bowref_aggregate_whole(const bowref_col *cols, int32_t ncols, int32_t time_col, const bowref_agg_spec *specs,
                           int32_t nspecs, bowref_out_col *outs) {
    if (nspecs <= 0 || time_col < 0 || time_col >= ncols) return BOWREF_EINVAL;
    const int64_t n = ncols ? cols[0].length : 0;
    bowref_rolling r;
    memset(&r, 0, sizeof r);
    r.cols = cols;
    r.ncols = ncols;
    r.time_col = time_col;
    r.nrows = n;
    for (int j = 0; j < nspecs; j++) {
        const bowref_agg_spec *a = &specs[j];
        if (a->col < 0 || a->col >= ncols) return BOWREF_EINVAL;
        const int32_t typ = whole_return_type(a->op, cols[a->col].dtype);
        outs[j].dtype = typ;
        if (n == 0) continue;
        memset(outs[j].values, 0, 8);
        outs[j].validity[0] = 0;
        const bowref_col *tc = &cols[time_col];
        double first_value = -1, last_value = -1;
        for (int64_t i = 0; i < n; i++)
            if (col_valid(tc, i)) { first_value = col_f64(tc, i); break; }
        for (int64_t i = n - 1; i >= 0; i--)
            if (col_valid(tc, i)) { last_value = col_f64(tc, i); break; }
        bowref_window w;
        memset(&w, 0, sizeof w);
        w.lo = 0;
        w.hi = n;
        w.first_index = 0;
        w.first_value = f64_to_i64(first_value);
        w.last_value = f64_to_i64(last_value);
        w.is_inclusive = 1;
        ref_value v = agg_closure(&r, a->op, a->col, &w);
        for (int k = 0; k < a->nfactors; k++) v = apply_factor(v, a->factors[k]);
        if (v.is_nil) continue;
        if (typ == BOWREF_INT64 && v.is_int) { /* SetOrDropStrict: value.(int64) */
            ((int64_t *)outs[j].values)[0] = v.i;
            outs[j].validity[0] = 1;
        } else if (typ == BOWREF_FLOAT64 && !v.is_int) { /* value.(float64) */
            ((double *)outs[j].values)[0] = v.f;
            outs[j].validity[0] = 1;
        }
    }
    return BOWREF_OK;
}

/* ---- Window iteration export: first[k], end[k] (exclusive, incl. the inclusive row), inclusive flag */
int64_t bowref_windows(const bowref_rolling *r, int64_t *first_index, int64_t *lo, int64_t *hi, int64_t *first_value,
                       uint8_t *is_inclusive) {
    bowref_rolling rc = *r;
    bowref_window w;
    int64_t wi, n = 0;
    while (bowref_next(&rc, &w, &wi)) {
        if (first_index) first_index[wi] = w.first_index;
        if (lo) lo[wi] = w.lo;
        if (hi) hi[wi] = w.hi;
        if (first_value) first_value[wi] = w.first_value;
        if (is_inclusive) is_inclusive[wi] = (uint8_t)w.is_inclusive;
        n++;
    }
    return n;
}

/* ---- interpolation closures (rolling/interpolation/ *.go) ------------------ */
/* GetPrevFloat64s / GetPrevValues over the FULL bow: last row <= row valid in both time and col */
static int64_t prev_both_valid(const bowref_col *t, const bowref_col *c, int64_t row) {
    for (; row >= 0; row--)
        if (col_valid(t, row) && col_valid(c, row)) return row;
    return -1;
}

static ref_value interp_closure(const bowref_rolling *r, int op, int col, const bowref_window *w) {
    const bowref_col *t = &r->cols[r->time_col], *c = &r->cols[col];
    switch (op) {
    case BOWREF_INTERP_WINDOW_START: /* interpolation/windowstart.go:8-14 */
        return int_value(w->first_value);
    case BOWREF_INTERP_NONE: /* interpolation/none.go:8-14 */
        return nil_value();
    case BOWREF_INTERP_STEP_PREVIOUS: { /* interpolation/stepprevious.go:8-26 (stateless equivalent) */
        int64_t p = prev_both_valid(t, c, w->first_index - 1);
        if (p >= 0) return c->dtype == BOWREF_INT64 ? int_value(col_raw(c, p)) : float_value(raw_as_f64(col_raw(c, p)));
        if (r->prev_row && col_valid(&r->prev_row[col], 0)) {
            const bowref_col *pc = &r->prev_row[col];
            return pc->dtype == BOWREF_INT64 ? int_value(col_raw(pc, 0)) : float_value(raw_as_f64(col_raw(pc, 0)));
        }
        return nil_value();
    }
    case BOWREF_INTERP_STEP_NEXT: { /* no upstream closure: Bow.GetNextValues(intervalCol, col, w.FirstIndex), bowgetters.go:111-123 */
        int64_t nx = next_both_valid(t, c, w->first_index, r->nrows);
        if (nx < 0) return nil_value();
        return c->dtype == BOWREF_INT64 ? int_value(col_raw(c, nx)) : float_value(raw_as_f64(col_raw(c, nx)));
    }
    case BOWREF_INTERP_LINEAR: { /* interpolation/linear.go:8-38 (stateless equivalent) */
        double t0, v0;
        int64_t p = prev_both_valid(t, c, w->first_index - 1);
        if (p >= 0) {
            t0 = col_f64(t, p);
            v0 = col_f64(c, p);
        } else {
            if (!r->prev_row) return nil_value();
            const bowref_col *pt = &r->prev_row[r->time_col], *pc = &r->prev_row[col];
            if (!col_valid(pt, 0) || !col_valid(pc, 0)) return nil_value();
            t0 = col_f64(pt, 0);
            v0 = col_f64(pc, 0);
        }
        int64_t nx = next_both_valid(t, c, w->first_index, r->nrows);
        if (nx < 0) return nil_value();
        double t2 = col_f64(t, nx), v2 = col_f64(c, nx);
        double coef = ((double)w->first_value - t0) / (t2 - t0);
        return float_value(((v2 - v0) * coef) + v0);
    }
    }
    return nil_value();
}

/* Interpolate / interpolateWindows / interpolateWindow, interpolation.go:30-161.
 * ops[j] is the interpolation applied to column j (the reference matches columns
 * by POSITION in bow.AppendBows, bowappend.go:28-47, so interps must name every
 * column in schema order).  With out_values == NULL only counts the output rows.
 * Returns the number of output rows, or -(error code). */
int64_t bowref_interpolate(const bowref_rolling *r, const int32_t *ops, int32_t nops, void **out_values,
                           uint8_t **out_validity, int64_t capacity) {
    if (nops != r->ncols) return -BOWREF_EINVAL;
    for (int j = 0; j < nops; j++) { /* validateInterpolation :71-96: accepted input types */
        if (ops[j] == BOWREF_INTERP_WINDOW_START && r->cols[j].dtype != BOWREF_INT64) return -BOWREF_ETYPE;
    }
    bowref_rolling rc = *r;
    bowref_window w;
    int64_t wi, n_out = 0;
    const bowref_col *t = &r->cols[r->time_col];
    if (out_values)
        for (int j = 0; j < nops; j++) memset(out_validity[j], 0, (size_t)((capacity + 7) / 8));
    while (bowref_next(&rc, &w, &wi)) {
        int64_t first_col_value = -1; /* interpolation.go:119-125 */
        if (w.hi > w.lo) {
            int64_t i = w.lo;
            while (i < w.hi && !col_valid(t, i)) i++; /* GetNextFloat64 */
            if (i < w.hi) first_col_value = f64_to_i64(col_f64(t, i));
        }
        if (first_col_value != w.first_value) { /* missing start: one synthetic row, :140-160 */
            if (out_values) {
                if (n_out >= capacity) return -BOWREF_ECAPACITY;
                for (int j = 0; j < nops; j++) {
                    ref_value v = interp_closure(&rc, ops[j], j, &w);
                    if (v.is_nil) {
                        ((int64_t *)out_values[j])[n_out] = 0;
                    } else if (r->cols[j].dtype == BOWREF_INT64) { /* SetOrDrop into a 1-row buffer of the column type */
                        ((int64_t *)out_values[j])[n_out] = v.is_int ? v.i : f64_to_i64(v.f);
                        out_validity[j][n_out >> 3] |= (uint8_t)(1u << (n_out & 7));
                    } else {
                        ((double *)out_values[j])[n_out] = v.is_int ? (double)v.i : v.f;
                        out_validity[j][n_out >> 3] |= (uint8_t)(1u << (n_out & 7));
                    }
                }
            }
            n_out++;
        }
        if (out_values) { /* AppendBows copies the window rows, bowappend.go:31-63 */
            if (n_out + (w.hi - w.lo) > capacity) return -BOWREF_ECAPACITY;
            for (int j = 0; j < nops; j++) {
                const bowref_col *c = &r->cols[j];
                for (int64_t i = w.lo; i < w.hi; i++) {
                    int64_t o = n_out + (i - w.lo);
                    ((int64_t *)out_values[j])[o] = col_raw(c, i);
                    if (col_valid(c, i)) out_validity[j][o >> 3] |= (uint8_t)(1u << (o & 7));
                }
            }
        }
        n_out += w.hi - w.lo;
    }
    return n_out;
}

/* ---- whole-column fills, bowfill.go:14-288 (Int64 / Float64 columns) --------------------------------------
 * method: 0 FillPrevious, 1 FillNext, 2 FillMean, 3 FillLinear (ref_col = sorted reference column).
 * out_values: n 8-byte slots, out_validity: ceil(n/8) bytes (bit offset 0).  Returns 0, or BOWREF_EINVAL. */
static double go_round(double x) { /* math.Round: half away from zero */
    if (x != x || x - x != 0.0) return x;
    double a = x < 0 ? -x : x;
    if (a >= 4503599627370496.0) return x;
    double f = (double)(int64_t)a; /* floor for 0 <= a < 2^52, exact */
    if (a - f >= 0.5) f += 1.0;    /* a - f is exact */
    return x < 0 ? -f : (x == 0 ? x : f);
}
int bowref_fill(const bowref_col *cols, int32_t ncols, int32_t method, int32_t col, int32_t ref_col, void *out_values,
                uint8_t *out_validity) {
    if (col < 0 || col >= ncols || method < 0 || method > 3) return BOWREF_EINVAL;
    if (method == 3 && (ref_col < 0 || ref_col >= ncols || ref_col == col)) return BOWREF_EINVAL;
    const bowref_col *c = &cols[col];
    const int64_t n = c->length;
    const int is_int = c->dtype == BOWREF_INT64;
    int64_t *ov = (int64_t *)out_values;
    memset(out_validity, 0, (size_t)((n + 7) / 8));
    for (int64_t r = 0; r < n; r++) {
        if (col_valid(c, r)) {
            ov[r] = col_raw(c, r);
            out_validity[r >> 3] |= (uint8_t)(1u << (r & 7));
            continue;
        }
        ov[r] = 0;
        int64_t p = r - 1, q = r + 1;
        while (p >= 0 && !col_valid(c, p)) p--;
        while (q < n && !col_valid(c, q)) q++;
        if (q >= n) q = -1;
        int ok = 0;
        int64_t raw = 0;
        if (method == 0) { /* bowfill.go:160-164, 255-264 */
            if (p >= 0) { raw = col_raw(c, p); ok = 1; }
        } else if (method == 1) { /* bowfill.go:154-158 */
            if (q >= 0) { raw = col_raw(c, q); ok = 1; }
        } else if (method == 2) { /* bowfill.go:136-146 */
            if (p >= 0 && q >= 0) {
                double m = (col_f64(c, p) + col_f64(c, q)) / 2;
                raw = is_int ? f64_to_i64(go_round(m)) : f64_as_raw(m);
                ok = 1;
            }
        } else { /* bowfill.go:66-95 */
            const bowref_col *rc = &cols[ref_col];
            if (p >= 0 && q >= 0 && col_valid(rc, r) && col_valid(rc, p) && col_valid(rc, q)) {
                double prev_to_fill = col_f64(c, p), next_to_fill = col_f64(c, q);
                double tmp = col_f64(rc, r) - col_f64(rc, p);
                tmp /= col_f64(rc, q) - col_f64(rc, p);
                tmp *= next_to_fill - prev_to_fill;
                tmp += prev_to_fill;
                raw = is_int ? f64_to_i64(go_round(tmp)) : f64_as_raw(tmp);
                ok = 1;
            }
        }
        if (ok) {
            ov[r] = raw;
            out_validity[r >> 3] |= (uint8_t)(1u << (r & 7));
        }
    }
    return BOWREF_OK;
}

int32_t bowref_sizeof_rolling(void) { return (int32_t)sizeof(bowref_rolling); }
