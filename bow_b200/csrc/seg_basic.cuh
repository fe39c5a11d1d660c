// seg_basic — Count / Sum / ArithmeticMean / Min / Max / First / Last on the streaming segmented
// reduction (segreduce.cuh).  Replaces the per-window closures of the reference
// (rolling/aggregation/{count,sum,arithmeticmean,minmax,firstlast}.go).
//
// Window state (a monoid; combine(L, R) keeps the left operand on ties, which reproduces the
// reference's sequential semantics bit-exactly for Count/Min/Max/First/Last):
//   cnt   valid rows                                           count.go:8-20
//   sum   sum of float64(v) over valid rows                    sum.go:8-25, arithmeticmean.go:8-30
//   mn/mx min / max over the non-NaN valid values, earliest wins on ties (minmax.go:20, `v < m`)
//   fi/li index of the first / last valid row                  firstlast.go:8-36
// cnt / fi / li never cost anything per row: they are popcount / ffs / fls of the thread's validity
// bits masked to the segment.  Min/Max of the reference start from the FIRST valid value and only
// replace on a strict compare, so a leading NaN is sticky and later NaNs are ignored:
// result = isnan(first) ? first : mn.  ArithmeticMean = sum / float64(cnt) is formed by the epilogue.
#pragma once
#include <math_constants.h>

#include "segreduce.cuh"

namespace bowgpu {

namespace {

constexpr int64_t CLOSED_BIT = (int64_t)1 << 62;

// state of a run of rows
struct BState {
    double sum, mn, mx;
    uint64_t first, last;  // raw bits of the first / last valid value (maintained only when the ops need them)
    uint32_t cnt;          // valid rows
};

// Final per-window write (values only; validity bitmaps, the mean division and empty-window defaults
// are produced by the epilogue from cnt).  first/last are raw value bits.
template <bool IS_INT>
__device__ __forceinline__ void write_window(const BasicOut &o, int64_t W, int64_t k, int64_t cnt, double sum,
                                             double mn, double mx, uint64_t first, uint64_t last) {
    if ((uint64_t)k >= (uint64_t)W || cnt == 0) return;
    o.cnt[k] = cnt;
    if (o.sum) o.sum[k] = sum;
    if (o.mn || o.mx) {
        const double f = bits_as_f64(first);
        const bool sticky = !IS_INT && (f != f);  // first valid value is NaN (minmax.go:14-24)
        if (o.mn) o.mn[k] = sticky ? f : mn;
        if (o.mx) o.mx[k] = sticky ? f : mx;
    }
    if (o.first) o.first[k] = first;
    if (o.last) o.last[k] = last;
}

template <uint32_t OPS, bool IS_INT>
struct BasicPol {
    static constexpr bool LINEAR_PHASE = false;  // (the basic family runs at the DRAM rate with the generic row loop)
    struct Lin {};
    using State = BState;
    using Carry = BasicCarry;
    using Out = BasicOut;
    struct Inc {};
    static constexpr bool NEXT_VALUE = false;
    // Min/Max of a float column start from the FIRST valid value (a leading NaN is sticky): keep its bits
    static constexpr bool NEED_FIRST = (OPS & OPS_FIRSTLAST) || ((OPS & OPS_MINMAX) && !IS_INT);
    static constexpr bool NEED_LAST = (OPS & OPS_FIRSTLAST) != 0;

    static __device__ __forceinline__ Inc make_inc(bool, bool, uint64_t, int64_t) { return Inc(); }

    static __device__ __forceinline__ State identity() {
        State s;
        s.sum = 0.0;
        s.mn = CUDART_INF;
        s.mx = -CUDART_INF;
        s.first = s.last = 0;
        s.cnt = 0;
        return s;
    }
    static __device__ __forceinline__ void accumulate(State &s, int64_t, uint64_t raw) {
        const double v = IS_INT ? (double)(int64_t)raw : bits_as_f64(raw);  // GetFloat64, bowgetters.go:218-229
        s.sum += v;
        if (OPS & OPS_MINMAX) {
            if (v < s.mn) s.mn = v;
            if (v > s.mx) s.mx = v;
        }
    }
    // cnt / first / last cost nothing per row: popcount / ffs / fls of the validity bits of the rows of one phase
    // that joined the run (bit j of mask = row j of the thread's phase, vrow = its values in shared memory)
    static __device__ __forceinline__ void note(State &s, uint32_t mask, const int64_t *, const uint64_t *vrow, const int swz) {
        if (mask) {
            if (NEED_FIRST && s.cnt == 0) s.first = vrow[(__ffs(mask) - 1) ^ swz];
            if (NEED_LAST) s.last = vrow[(31 - __clz(mask)) ^ swz];
            s.cnt += __popc(mask);
        }
    }
    // one synthetic row of the interpolated frame joins a run as its FIRST row (fused Interpolate -> Aggregate)
    static __device__ __forceinline__ void inject(State &s, int64_t t, uint64_t raw, bool valid) {
        if (!valid) return;
        accumulate(s, t, raw);
        if (s.cnt == 0) s.first = raw;
        s.last = raw;
        s.cnt += 1;
    }
    static __device__ __forceinline__ State combine(const State &L, const State &R) {
        State o;
        o.sum = L.sum + R.sum;
        o.mn = CUDART_INF;  // fields the instantiated ops do not maintain stay at their identity values
        o.mx = -CUDART_INF;
        o.first = o.last = 0;
        if (OPS & OPS_MINMAX) {
            o.mn = (R.mn < L.mn) ? R.mn : L.mn;
            o.mx = (R.mx > L.mx) ? R.mx : L.mx;
        }
        if (NEED_FIRST) o.first = L.cnt ? L.first : R.first;
        if (NEED_LAST) o.last = R.cnt ? R.last : L.last;
        o.cnt = L.cnt + R.cnt;
        return o;
    }
    static __device__ __forceinline__ State shfl_up(const State &s, int d) {
        State o;
        o.sum = __shfl_up_sync(0xffffffffu, s.sum, d);
        o.mn = CUDART_INF;
        o.mx = -CUDART_INF;
        o.first = o.last = 0;
        if (OPS & OPS_MINMAX) {
            o.mn = __shfl_up_sync(0xffffffffu, s.mn, d);
            o.mx = __shfl_up_sync(0xffffffffu, s.mx, d);
        }
        if (NEED_FIRST) o.first = __shfl_up_sync(0xffffffffu, (unsigned long long)s.first, d);
        if (NEED_LAST) o.last = __shfl_up_sync(0xffffffffu, (unsigned long long)s.last, d);
        o.cnt = __shfl_up_sync(0xffffffffu, s.cnt, d);
        return o;
    }
    static __device__ __forceinline__ void write(const Out &o, const WindowGeom &g, int64_t k, const State &s,
                                                 const Inc &) {
        write_window<IS_INT>(o, g.W, k, s.cnt, s.sum, s.mn, s.mx, s.first, s.last);
    }
    static __device__ __forceinline__ Carry make_carry(const State &s, const Inc &, int64_t key, bool closed) {
        Carry r;
        r.key = key;
        r.cnt = (int64_t)s.cnt | (closed ? CLOSED_BIT : 0);
        r.sum = s.sum;
        r.mn = s.mn;
        r.mx = s.mx;
        r.first = s.first;
        r.last = s.last;
        r.edge_t = 0;
        return r;
    }
    static __device__ __forceinline__ void carry_set_key(Carry &c, int64_t key) { c.key = key; }
    static __device__ __forceinline__ void carry_set_edge(Carry &c, int64_t t, uint64_t, bool) { c.edge_t = t; }
    static __device__ __forceinline__ int64_t carry_edge_t(const Carry &c) { return c.edge_t; }
    static __device__ __forceinline__ int64_t carry_key(const Carry &c) { return c.key; }
    static __device__ __forceinline__ bool carry_closed(const Carry &c) { return (c.cnt & CLOSED_BIT) != 0; }
    static __device__ __forceinline__ void carry_inc_from_edge(Carry &, const Carry &, int64_t) {}
    static __device__ __forceinline__ void carry_set_inc(Carry &, const Inc &) {}
    static __device__ __forceinline__ uint64_t carry_edge_raw(const Carry &) { return 0; }
    static __device__ __forceinline__ bool carry_edge_valid(const Carry &) { return false; }
    static __device__ __forceinline__ void carry_prepend_point(Carry &a, int64_t t, uint64_t raw, bool valid) {
        if (!valid) return;
        State p = identity();
        inject(p, t, raw, true);
        const int64_t ac = a.cnt & ~CLOSED_BIT;
        a.sum = p.sum + a.sum;
        a.mn = (a.mn < p.mn) ? a.mn : p.mn;
        a.mx = (a.mx > p.mx) ? a.mx : p.mx;
        a.first = raw;
        a.last = ac ? a.last : raw;
        a.cnt += 1;
    }
    static __device__ __forceinline__ void carry_clear_inc(Carry &) {}
    static __device__ __forceinline__ void carry_combine(Carry &a, const Carry &h) {
        const int64_t ac = a.cnt & ~CLOSED_BIT, hc = h.cnt & ~CLOSED_BIT;
        a.sum = a.sum + h.sum;
        a.mn = (h.mn < a.mn) ? h.mn : a.mn;
        a.mx = (h.mx > a.mx) ? h.mx : a.mx;
        a.first = ac ? a.first : h.first;
        a.last = hc ? h.last : a.last;
        a.cnt = ac + hc;
    }
    static __device__ __forceinline__ void write_carry(const Out &o, const WindowGeom &g, int64_t k, const Carry &a) {
        write_window<IS_INT>(o, g.W, k, a.cnt & ~CLOSED_BIT, a.sum, a.mn, a.mx, a.first, a.last);
    }
};

}  // namespace

}  // namespace bowgpu
