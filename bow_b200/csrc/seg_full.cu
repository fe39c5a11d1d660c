// seg_full.cu — BOTH aggregation families of one input column in ONE streaming launch.
//
// A call that asks for basic aggregations (Count / Sum / Mean / Min / Max / First / Last) AND integrals (IntegralStep /
// IntegralTrapezoid / WeightedAverage*) of the same column used to stream the time column and the value column twice, once
// per family (api.cu).  BASELINE configs[4] asks for all eleven on each of 16 columns: 32 launches of 16 B per row each.
// FullPol is the product of the two monoids — the state of a run is (BState, IState), every operation of the policy
// interface applies to both components, and the straight-line phase (segreduce.cuh seg_phase_linear) carries the basic
// sums / minima / maxima of the open and of the next window next to the integral chain — so the pair costs one pass of
// 16 B per row.  Results are those of the two separate launches (same left-to-right order inside a thread's rows, same
// combine at thread / tile edges).
#include "seg_basic.cuh"
#include "seg_integral.cuh"

namespace bowgpu {

namespace {

struct FullOut {
    BasicOut b;
    IntegralOut i;
};
struct alignas(16) FullCarry {
    BasicCarry b;
    ICarry i;
};

template <bool IS_INT>
struct FullPol {
    using B = BasicPol<OPS_SUMCNT | OPS_MINMAX | OPS_FIRSTLAST, IS_INT>;
    using I = IntegralPol<true, true, IS_INT>;
    struct State {
        typename B::State b;
        typename I::State i;
    };
    using Carry = FullCarry;
    using Out = FullOut;
    using Inc = typename I::Inc;
    static constexpr bool NEXT_VALUE = true;
    static constexpr bool LINEAR_PHASE = true;

    static __device__ __forceinline__ double val(uint64_t raw) { return I::val(raw); }
    static __device__ __forceinline__ Inc make_inc(bool at_end, bool valid_next, uint64_t raw_next, int64_t t_next) {
        return I::make_inc(at_end, valid_next, raw_next, t_next);
    }
    static __device__ __forceinline__ State identity() {
        State s;
        s.b = B::identity();
        s.i = I::identity();
        return s;
    }
    static __device__ __forceinline__ void accumulate(State &s, int64_t t, uint64_t raw) {
        B::accumulate(s.b, t, raw);
        I::accumulate(s.i, t, raw);
    }
    static __device__ __forceinline__ void note(State &s, uint32_t mask, const int64_t *trow, const uint64_t *vrow, const int swz) {
        B::note(s.b, mask, trow, vrow, swz);
        I::note(s.i, mask, trow, vrow, swz);
    }
    static __device__ __forceinline__ void inject(State &s, int64_t t, uint64_t raw, bool valid) {
        B::inject(s.b, t, raw, valid);
        I::inject(s.i, t, raw, valid);
    }
    static __device__ __forceinline__ State combine(const State &L, const State &R) {
        State o;
        o.b = B::combine(L.b, R.b);
        o.i = I::combine(L.i, R.i);
        return o;
    }
    static __device__ __forceinline__ State shfl_up(const State &s, int d) {
        State o;
        o.b = B::shfl_up(s.b, d);
        o.i = I::shfl_up(s.i, d);
        return o;
    }
    static __device__ __forceinline__ void write(const Out &o, const WindowGeom &g, int64_t k, const State &s, const Inc &inc) {
        B::write(o.b, g, k, s.b, typename B::Inc());
        I::write(o.i, g, k, s.i, inc);
    }
    // ---- straight-line phase: the integral chain plus the basic sums / minima / maxima of the two windows ---------------
    struct Lin {
        typename I::Lin i;
        double sum_a, mn_a, mx_a, sum_b, mn_b, mx_b;
    };
    static __device__ __forceinline__ Lin lin_begin(const State &s) {
        Lin L;
        L.i = I::lin_begin(s.i);
        L.sum_a = s.b.sum;
        L.mn_a = s.b.mn;
        L.mx_a = s.b.mx;
        L.sum_b = 0.0;
        L.mn_b = CUDART_INF;
        L.mx_b = -CUDART_INF;
        return L;
    }
    static __device__ __forceinline__ void lin_row(Lin &L, const int j, const int b, const int64_t t, const uint64_t raw,
                                                   const bool valid) {
        I::lin_row(L.i, j, b, t, raw, valid);
        const double v = val(raw);
        const bool in_b = j >= b;
        const bool pa = valid && !in_b, pb = valid && in_b;
        if (pa) L.sum_a += v;
        if (pb) L.sum_b += v;
        if (pa && v < L.mn_a) L.mn_a = v;  // minmax.go:20 `v < m`
        if (pb && v < L.mn_b) L.mn_b = v;
        if (pa && v > L.mx_a) L.mx_a = v;
        if (pb && v > L.mx_b) L.mx_b = v;
    }
    static __device__ __forceinline__ void lin_end_open(State &s, const Lin &L) {
        I::lin_end_open(s.i, L.i);
        s.b.sum = L.sum_a;
        s.b.mn = L.mn_a;
        s.b.mx = L.mx_a;
    }
    static __device__ __forceinline__ void lin_end_split(State &sa, State &sb, const Lin &L, const uint32_t ma, const uint32_t mb,
                                                         const int64_t *trow, const uint64_t *vrow, const int swz) {
        I::lin_end_split(sa.i, sb.i, L.i, ma, mb, trow, vrow, swz);
        sa.b.sum = L.sum_a;
        sa.b.mn = L.mn_a;
        sa.b.mx = L.mx_a;
        sb.b.sum = L.sum_b;
        sb.b.mn = L.mn_b;
        sb.b.mx = L.mx_b;
    }
    // ---- tile records ------------------------------------------------------------------------------------------------------
    static __device__ __forceinline__ Carry make_carry(const State &s, const Inc &inc, int64_t key, bool closed) {
        Carry c;
        c.b = B::make_carry(s.b, typename B::Inc(), key, closed);
        c.i = I::make_carry(s.i, inc, key, closed);
        return c;
    }
    static __device__ __forceinline__ void carry_set_key(Carry &c, int64_t key) {
        B::carry_set_key(c.b, key);
        I::carry_set_key(c.i, key);
    }
    static __device__ __forceinline__ void carry_set_edge(Carry &c, int64_t t, uint64_t raw, bool valid) {
        B::carry_set_edge(c.b, t, raw, valid);
        I::carry_set_edge(c.i, t, raw, valid);
    }
    static __device__ __forceinline__ int64_t carry_edge_t(const Carry &c) { return I::carry_edge_t(c.i); }
    static __device__ __forceinline__ int64_t carry_key(const Carry &c) { return I::carry_key(c.i); }
    static __device__ __forceinline__ bool carry_closed(const Carry &c) { return I::carry_closed(c.i); }
    static __device__ __forceinline__ void carry_inc_from_edge(Carry &a, const Carry &h, int64_t E) { I::carry_inc_from_edge(a.i, h.i, E); }
    static __device__ __forceinline__ void carry_clear_inc(Carry &a) { I::carry_clear_inc(a.i); }
    static __device__ __forceinline__ void carry_set_inc(Carry &a, const Inc &inc) { I::carry_set_inc(a.i, inc); }
    static __device__ __forceinline__ uint64_t carry_edge_raw(const Carry &c) { return I::carry_edge_raw(c.i); }
    static __device__ __forceinline__ bool carry_edge_valid(const Carry &c) { return I::carry_edge_valid(c.i); }
    static __device__ __forceinline__ void carry_prepend_point(Carry &a, int64_t t, uint64_t raw, bool valid) {
        B::carry_prepend_point(a.b, t, raw, valid);
        I::carry_prepend_point(a.i, t, raw, valid);
    }
    static __device__ __forceinline__ void carry_combine(Carry &a, const Carry &h) {
        B::carry_combine(a.b, h.b);
        I::carry_combine(a.i, h.i);
    }
    static __device__ __forceinline__ void write_carry(const Out &o, const WindowGeom &g, int64_t k, const Carry &a) {
        B::write_carry(o.b, g, k, a.b);
        I::write_carry(o.i, g, k, a.i);
    }
};

#ifndef SEG_FULL_CTAS
#define SEG_FULL_CTAS 3  // (ptxas fits the two states and the straight-line phase into 168 registers without spills; 2 CTAs per SM: 43.8 instead of 39.3 ms on configs[4])
#endif

template <bool IS_INT>
int launch_full_t(const FullLaunch &L, int sm, cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1) {
    SegArgs<FullPol<IS_INT>> A;
    memset(&A, 0, sizeof A);
    A.time = L.time;
    A.values = L.values;
    A.validity = L.validity;
    A.g = L.g;
    A.out.b = L.out_basic;
    A.out.i = L.out_integral;
    A.carry_head = (FullCarry *)L.carry_head;
    A.carry_tail = (FullCarry *)L.carry_tail;
    A.skip = (FullCarry *)L.skip;
    A.status = L.status;
    A.syn = L.syn;
    A.gate = L.gate;
    A.gate_lanes = L.gate_lanes;
    return L.validity ? seg_launch<FullPol<IS_INT>, true, SEG_FULL_CTAS>(A, sm, s, e0, e1)
                      : seg_launch<FullPol<IS_INT>, false, SEG_FULL_CTAS>(A, sm, s, e0, e1);
}

}  // namespace

size_t full_carry_bytes(int64_t n) { return (size_t)((n + SEG_T - 1) / SEG_T) * 2 * sizeof(FullCarry); }
size_t full_skip_bytes(int64_t n) { return (size_t)seg_skip_records((n + SEG_T - 1) / SEG_T) * sizeof(FullCarry); }

int launch_segreduce_full(const FullLaunch &L, int sm_count, cudaStream_t stream, cudaEvent_t e0, cudaEvent_t e1) {
    return L.is_int ? launch_full_t<true>(L, sm_count, stream, e0, e1) : launch_full_t<false>(L, sm_count, stream, e0, e1);
}

}  // namespace bowgpu
