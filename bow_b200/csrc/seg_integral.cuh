// seg_integral — IntegralStep / IntegralTrapezoid (and, through the epilogue's division by the window
// width, WeightedAverageStep / WeightedAverageLinear) on the streaming segmented reduction.
// Replaces rolling/aggregation/integral.go:8-69 and weightedmean.go:8-34.
//
// A "point" is a row whose value is valid (the time column is non-null on the GPU path).  With
// T = float64(t) converted BEFORE subtracting, exactly like the reference (integral.go:48-55):
//   step       sum_j v_j * (T_{j+1} - T_j)  +  v_last * (float64(E_k) - T_last)         integral.go:40-69
//   trapezoid  sum_j (v_j + v_{j+1}) / 2 * (T_{j+1} - T_j)  over rows_inc(k)             integral.go:8-38
// State of a run of rows: {n, firstT, firstV, lastT, lastV, sumStep, sumTrap}; two adjacent runs
// combine by adding the joint term between L's last and R's first point, so the state is a monoid
// and the generic kernel can split windows at thread and tile boundaries.  The trapezoid needs
// inclusive windows (integral.go:9): the row after a window's last row joins it when its time equals
// the window end; the closing thread sees that row (Inc) and appends it at finalisation.
// No FMA contraction (-fmad=false): Go on amd64 rounds the product and the sum separately.
#pragma once
#include "segreduce.cuh"

namespace bowgpu {

namespace {

#ifndef SEG_INT_CTAS
#define SEG_INT_CTAS 3
#endif

struct IState {
    double fT, fV, lT, lV, sS, sT;
    uint32_t n;    // points (updated when rows are noted, i.e. at window boundaries and phase ends)
    uint32_t has;  // a point has been accumulated since the run began (per-row flag; n > 0 once noted)
};

struct alignas(16) ICarry {
    int64_t key;       // window index, -1 = none
    int64_t n;         // points; bit 62 = window closed inside the tile (head records)
    double fT, fV, lT, lV, sS, sT;
    double incV, incT;
    int64_t inc_has;
    int64_t edge_t;       // head records: first row of the tile (time, raw value bits, validity);
    uint64_t edge_raw;    // tail records: edge_t = time of the tile's last row
    int64_t edge_valid;
    int64_t _pad[2];
};
static_assert(sizeof(ICarry) == 128, "carry record layout");
constexpr int64_t I_CLOSED_BIT = (int64_t)1 << 62;

template <bool STEP, bool TRAP, bool IS_INT>
struct IntegralPol {
    using State = IState;
    using Carry = ICarry;
    using Out = IntegralOut;
    struct Inc {
        double v, T;
        bool has;
    };
    static constexpr bool NEXT_VALUE = TRAP;

    static __device__ __forceinline__ double val(uint64_t raw) {
        return IS_INT ? (double)(int64_t)raw : bits_as_f64(raw);  // GetFloat64, bowgetters.go:218-229
    }
    static __device__ __forceinline__ Inc make_inc(bool at_end, bool valid_next, uint64_t raw_next, int64_t t_next) {
        Inc i;
        i.has = at_end && valid_next;
        i.v = val(raw_next);
        i.T = (double)t_next;
        return i;
    }
    static __device__ __forceinline__ State identity() {
        State s;
        s.fT = s.fV = s.lT = s.lV = 0.0;
        s.sS = s.sT = 0.0;
        s.n = 0;
        s.has = 0;
        return s;
    }
    // One valid point, branch free: the joint term with the previous point is formed unconditionally and selected
    // away for the first point of a run (adding +0.0 never changes a sum that started at +0.0).
    static __device__ __forceinline__ void accumulate(State &s, int64_t t, uint64_t raw) {
        const double T = (double)t, v = val(raw);
        const double dt = T - s.lT;
        if (STEP) s.sS += s.has ? s.lV * dt : 0.0;                // integral.go:57
        if (TRAP) s.sT += s.has ? (s.lV + v) / 2 * dt : 0.0;      // integral.go:28
        s.lT = T;
        s.lV = v;
        s.has = 1;
    }
    // ---- straight-line phase (segreduce.cuh seg_phase_linear) ---------------------------------------------------------------
    static constexpr bool LINEAR_PHASE = true;
    struct Lin {
        double lT, lV;             // the previous valid point of the chain
        double sSa, sTa, sSb, sTb; // sums of the open window (a) and of the window that begins at the boundary row (b)
        bool has;                  // the chain has a previous point inside the same window
    };
    static __device__ __forceinline__ Lin lin_begin(const State &s) {
        Lin L;
        L.lT = s.lT;
        L.lV = s.lV;
        L.sSa = s.sS;
        L.sTa = s.sT;
        L.sSb = L.sTb = 0.0;
        L.has = s.has != 0;
        return L;
    }
    // row j of the phase (j is a constant after unrolling); rows from b on belong to the next window
    static __device__ __forceinline__ void lin_row(Lin &L, const int j, const int b, const int64_t t, const uint64_t raw,
                                                   const bool valid) {
        const double T = (double)t, v = val(raw);
        const bool in_b = j >= b;
        const bool has = L.has && j != b;  // the boundary row has no predecessor in its window
        const double dt = T - L.lT;
        const bool join = valid && has;
        if (STEP) {
            const double ts = L.lV * dt;  // integral.go:57
            if (join && !in_b) L.sSa += ts;
            if (join && in_b) L.sSb += ts;
        }
        if (TRAP) {
            const double tt = (L.lV + v) / 2 * dt;  // integral.go:28
            if (join && !in_b) L.sTa += tt;
            if (join && in_b) L.sTb += tt;
        }
        if (valid) {
            L.lT = T;
            L.lV = v;
        }
        L.has = has || valid;
    }
    static __device__ __forceinline__ void lin_end_open(State &s, const Lin &L) {  // no boundary in the phase
        s.lT = L.lT;
        s.lV = L.lV;
        s.sS = L.sSa;
        s.sT = L.sTa;
        s.has = L.has;
    }
    // one boundary: sa = the open window after its rows (mask ma), sb = the next window's rows of this phase (mask mb)
    static __device__ __forceinline__ void lin_end_split(State &sa, State &sb, const Lin &L, const uint32_t ma, const uint32_t mb,
                                                         const int64_t *trow, const uint64_t *vrow, const int swz) {
        sa.sS = L.sSa;
        sa.sT = L.sTa;
        if (ma) {  // its last point: the last valid row before the boundary
            const int j = (31 - __clz(ma)) ^ swz;
            sa.lT = (double)trow[j];
            sa.lV = val(vrow[j]);
            sa.has = 1;
        }
        sb.sS = L.sSb;
        sb.sT = L.sTb;
        if (mb) {
            sb.lT = L.lT;
            sb.lV = L.lV;
            sb.has = 1;
        }
    }
    // the number of points and the first point of a run come from the validity bits of the rows that joined it
    static __device__ __forceinline__ void note(State &s, uint32_t mask, const int64_t *trow, const uint64_t *vrow, const int swz) {
        if (mask) {
            if (s.n == 0) {
                const int j = (__ffs(mask) - 1) ^ swz;
                s.fT = (double)trow[j];
                s.fV = val(vrow[j]);
            }
            s.n += __popc(mask);
        }
    }
    // one synthetic row of the interpolated frame joins a run as its FIRST point (fused Interpolate -> Aggregate)
    static __device__ __forceinline__ void inject(State &s, int64_t t, uint64_t raw, bool valid) {
        if (!valid) return;
        accumulate(s, t, raw);
        if (s.n == 0) {
            s.fT = (double)t;
            s.fV = val(raw);
        }
        s.n += 1;
    }
    static __device__ __forceinline__ State combine(const State &L, const State &R) {
        State o;
        o.sS = o.sT = 0.0;  // (the sum this instantiation does not maintain)
        const bool l = L.n != 0, r = R.n != 0;
        const double dt = R.fT - L.lT;
        if (STEP) {
            const double j = (L.sS + L.lV * dt) + R.sS;
            o.sS = l ? (r ? j : L.sS) : R.sS;
        }
        if (TRAP) {
            const double j = (L.sT + (L.lV + R.fV) / 2 * dt) + R.sT;
            o.sT = l ? (r ? j : L.sT) : R.sT;
        }
        o.fT = l ? L.fT : R.fT;
        o.fV = l ? L.fV : R.fV;
        o.lT = r ? R.lT : L.lT;
        o.lV = r ? R.lV : L.lV;
        o.n = L.n + R.n;
        o.has = o.n != 0;
        return o;
    }
    static __device__ __forceinline__ State shfl_up(const State &s, int d) {
        State o;
        o.fT = __shfl_up_sync(0xffffffffu, s.fT, d);
        o.fV = __shfl_up_sync(0xffffffffu, s.fV, d);
        o.lT = __shfl_up_sync(0xffffffffu, s.lT, d);
        o.lV = __shfl_up_sync(0xffffffffu, s.lV, d);
        o.sS = o.sT = 0.0;
        if (STEP) o.sS = __shfl_up_sync(0xffffffffu, s.sS, d);
        if (TRAP) o.sT = __shfl_up_sync(0xffffffffu, s.sT, d);
        o.n = __shfl_up_sync(0xffffffffu, s.n, d);
        o.has = o.n != 0;
        return o;
    }
    static __device__ __forceinline__ void finish(const Out &o, const WindowGeom &g, int64_t k, int64_t n,
                                                  double lT, double lV, double sS, double sT, bool inc_has,
                                                  double incV, double incT) {
        if ((uint64_t)k >= (uint64_t)g.W) return;
        if (STEP && n > 0) {
            const double E = (double)window_last_value(g, k);  // float64(w.LastValue)
            o.step[k] = sS + lV * (E - lT);                                             // integral.go:53-57
            o.n_step[k] = n;
        }
        if (TRAP && n + (inc_has ? 1 : 0) >= 2) {  // integral.go:33-35: fewer than two points -> nil
            double s = sT;
            if (inc_has) s += (lV + incV) / 2 * (incT - lT);
            o.trap[k] = s;
            o.n_trap[k] = 1;
        }
    }
    static __device__ __forceinline__ void write(const Out &o, const WindowGeom &g, int64_t k, const State &s,
                                                 const Inc &inc) {
        finish(o, g, k, s.n, s.lT, s.lV, s.sS, s.sT, TRAP && inc.has, inc.v, inc.T);
    }
    static __device__ __forceinline__ Carry make_carry(const State &s, const Inc &inc, int64_t key, bool closed) {
        Carry c;
        c.key = key;
        c.n = (int64_t)s.n | (closed ? I_CLOSED_BIT : 0);
        c.fT = s.fT;
        c.fV = s.fV;
        c.lT = s.lT;
        c.lV = s.lV;
        c.sS = STEP ? s.sS : 0.0;
        c.sT = TRAP ? s.sT : 0.0;
        c.incV = inc.v;
        c.incT = inc.T;
        c.inc_has = TRAP && inc.has;
        c.edge_t = 0;
        c.edge_raw = 0;
        c.edge_valid = 0;
        c._pad[0] = c._pad[1] = 0;
        return c;
    }
    static __device__ __forceinline__ void carry_set_edge(Carry &c, int64_t t, uint64_t raw, bool valid) {
        c.edge_t = t;
        c.edge_raw = raw;
        c.edge_valid = valid;
    }
    static __device__ __forceinline__ int64_t carry_edge_t(const Carry &c) { return c.edge_t; }
    // a window that ends exactly at a tile boundary: its inclusive row, if any, is the next tile's first row
    static __device__ __forceinline__ void carry_inc_from_edge(Carry &a, const Carry &h, int64_t E) {
        a.inc_has = TRAP && h.edge_t == E && h.edge_valid;
        a.incV = val(h.edge_raw);
        a.incT = (double)h.edge_t;
    }
    static __device__ __forceinline__ void carry_clear_inc(Carry &a) { a.inc_has = 0; }
    static __device__ __forceinline__ void carry_set_inc(Carry &a, const Inc &inc) {
        a.inc_has = TRAP && inc.has;
        a.incV = inc.v;
        a.incT = inc.T;
    }
    static __device__ __forceinline__ uint64_t carry_edge_raw(const Carry &c) { return c.edge_raw; }
    static __device__ __forceinline__ bool carry_edge_valid(const Carry &c) { return c.edge_valid != 0; }
    static __device__ __forceinline__ void carry_prepend_point(Carry &a, int64_t t, uint64_t raw, bool valid) {
        if (!valid) return;
        const double T = (double)t, v = val(raw);
        const int64_t an = a.n & ~I_CLOSED_BIT;
        if (an) {
            const double dt = a.fT - T;
            a.sS = (0.0 + v * dt) + a.sS;
            a.sT = (0.0 + (v + a.fV) / 2 * dt) + a.sT;
        } else {
            a.sS = a.sT = 0.0;
            a.lT = T;
            a.lV = v;
        }
        a.fT = T;
        a.fV = v;
        a.n += 1;
    }
    static __device__ __forceinline__ void carry_set_key(Carry &c, int64_t key) { c.key = key; }
    static __device__ __forceinline__ int64_t carry_key(const Carry &c) { return c.key; }
    static __device__ __forceinline__ bool carry_closed(const Carry &c) { return (c.n & I_CLOSED_BIT) != 0; }
    static __device__ __forceinline__ void carry_combine(Carry &a, const Carry &h) {
        const int64_t hn = h.n & ~I_CLOSED_BIT;
        if (hn) {
            if (a.n) {
                const double dt = h.fT - a.lT;
                a.sS = (a.sS + a.lV * dt) + h.sS;
                a.sT = (a.sT + (a.lV + h.fV) / 2 * dt) + h.sT;
            } else {
                a.sS = h.sS;
                a.sT = h.sT;
                a.fT = h.fT;
                a.fV = h.fV;
            }
            a.lT = h.lT;
            a.lV = h.lV;
            a.n += hn;
        }
        a.inc_has = h.inc_has;
        a.incV = h.incV;
        a.incT = h.incT;
    }
    static __device__ __forceinline__ void write_carry(const Out &o, const WindowGeom &g, int64_t k, const Carry &a) {
        finish(o, g, k, a.n & ~I_CLOSED_BIT, a.lT, a.lV, a.sS, a.sT, a.inc_has != 0, a.incV, a.incT);
    }
};

}  // namespace

}  // namespace bowgpu
