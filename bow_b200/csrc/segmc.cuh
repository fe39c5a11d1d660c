// segmc.cuh — multi-column, multi-family streaming segmented reduction behind Rolling.Aggregate.
//
// Replaces aggregateWindows (rolling/aggregation.go:190-238), which iterates every window once per aggregation, and the
// closures of rolling/aggregation/*.go.  One launch reads the time column ONCE and up to MC_MAXC value columns once and
// produces every aggregation of those columns — Count / Sum / ArithmeticMean / Min / Max / First / Last and
// IntegralStep / IntegralTrapezoid / WeightedAverageStep / WeightedAverageLinear — as FINAL output values.
//
// Work decomposition (load balance independent of window sizes):
//   * fixed-size ROW tiles (T = NT * RE rows); each CTA owns a CONTIGUOUS chunk of tiles.  Consumer thread t owns the RE
//     consecutive rows [t*RE, (t+1)*RE) of a tile and walks them in NP phases of P = 16 rows.
//   * a producer warp feeds shared memory with 2-D TMA boxes [NT][P] (cp.async.bulk.tensor.2d, 128-byte swizzle: a
//     quarter warp's LDS.128 touches all 32 banks once).  The NP time boxes of a tile stay resident while the value
//     columns stream past them through a ring of S boxes; full/empty mbarriers per box, no CTA-wide barrier per phase.
//   * TIME PASS, once per tile: sortedness check, the 64-bit mask `bm` of rows that start a later window than their
//     predecessor, the window of the thread's first row (one exact division per thread and tile), the bitmap of
//     windows that hold a row at all (`touched`).
//   * VALUE PASS, once per tile and column: per row one test of a `bm` bit and the accumulation; the window state is a
//     monoid (McPol) so runs split at thread / tile / chunk edges combine exactly as in segreduce.cuh: thread tails by
//     a warp segmented scan (only the steps the flag pattern needs) + one cross-warp pass, the open window of a tile
//     is carried to the CTA's next tile in shared memory, chunks leave a head and a tail record that mc_fixup_kernel
//     joins left to right.  No atomics on values: results are deterministic run to run.
//   * whoever completes a window writes its final values (mean and weighted averages divided on the spot, value 0 in
//     null slots) and sets its validity bit (atomicOr into zeroed words); finish.cu derives every output bitmap from
//     those words and takes care of windows without rows.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <math_constants.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "kernels.h"

namespace bowgpu {

#ifndef MC_CFG_NP
#define MC_CFG_NP 4
#endif
#ifndef MC_CFG_SLOTS
#define MC_CFG_SLOTS 6
#endif
#ifndef MC_CFG_CTAS
#define MC_CFG_CTAS 2
#endif
#ifndef MC_CFG_UNROLL
#define MC_CFG_UNROLL 2
#endif
constexpr int MC_NT = 128;                  // consumer threads of a CTA
constexpr int MC_NW = MC_NT / 32;
constexpr int MC_THREADS = MC_NT + 32;      // + the producer warp
constexpr int MC_P = 16;                    // rows per phase and thread (one 128-byte swizzle row)
constexpr int MC_NP = MC_CFG_NP;            // phases per tile
constexpr int MC_RE = MC_P * MC_NP;         // consecutive rows owned by one thread in a tile
constexpr int MC_T = MC_NT * MC_RE;         // rows per tile
constexpr int MC_BOX = MC_NT * MC_P * 8;    // bytes of one box
constexpr int MC_R = MC_CFG_SLOTS;          // boxes in shared memory
constexpr int MC_UNROLL = MC_CFG_UNROLL;      // row pairs per trip of the streaming loop
constexpr int MC_HEADER = 4096;
constexpr int MC_SMEM = MC_HEADER + MC_R * MC_BOX;
constexpr int64_t MC_CLOSED = (int64_t)1 << 62;
static_assert(MC_RE <= 64, "the boundary mask of a thread is one 64-bit word");
static_assert(MC_R <= 12 && MC_NP <= 4, "barrier arrays");

__device__ __forceinline__ int mc_swz(int t) { return (t & 7) << 1; }  // element j of thread t's row sits at j ^ mc_swz(t)
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// producer side wait: poll with a pause (a tight try_wait loop of the producer warp competes with the consumers for
// issue slots and for the shared-memory pipe)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
#ifndef MC_CFG_BACKOFF_NS
#define MC_CFG_BACKOFF_NS 100
#endif
        __nanosleep(MC_CFG_BACKOFF_NS);
    }
}
__device__ __forceinline__ void mc_bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(MC_NT) : "memory"); }
__device__ __noinline__ static uint64_t mc_div_slow(uint64_t x, uint64_t d, double inv_rd) {
    DivU64 dv{d, inv_rd};
    return div_u64(x, dv);
}
__device__ __forceinline__ uint32_t mc_below(int j) { return (1u << j) - 1u; }  // j in [0, 31]
// acc += v in the lanes whose bit of `mask` is set: ONE predicated DADD (the compiler's own rendering of `if (p) acc += v`
// is an unconditional add plus two selects)
__device__ __forceinline__ void mc_padd(double &acc, const double v, const uint32_t mask, const uint32_t bit) {
    asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 q, t, 0;\n\t@q add.rn.f64 %0, %0, %1;\n\t}"
        : "+d"(acc)
        : "d"(v), "r"(mask), "r"(bit));
}

// ---- window state ---------------------------------------------------------------------------------------------------
struct McState {
    double sum, mn, mx;    // sum.go, minmax.go (min / max over the non-NaN values, earliest wins on ties)
    uint64_t first, last;  // raw bits of the first / last valid value (firstlast.go; a leading NaN is sticky in Min / Max)
    double fT, lT, lV;     // float64(time) of the first / last point, value of the last point (integral.go:48-55)
    double sS, sT;         // step / trapezoid sums between the points of the run
    uint32_t cnt;          // valid rows noted so far
    uint32_t has;          // a point has been accumulated since the run began (ahead of cnt inside a phase)
};
struct McInc {  // the inclusive row of a window (rolling.go:201-209): the row right after it when it sits exactly on the end
    double v, T;
    uint32_t has;
};

template <uint32_t BOPS, uint32_t IOPS, bool IS_INT_>
struct McPol {
    using State = McState;
    using Inc = McInc;
    static constexpr bool IS_INT = IS_INT_;
    static constexpr bool SUMS = BOPS != 0;
    static constexpr bool MINMAX = (BOPS & MC_MINMAX) != 0;
    static constexpr bool FIRSTLAST = (BOPS & MC_FIRSTLAST) != 0;
    static constexpr bool STEP = (IOPS & MC_STEP) != 0;
    static constexpr bool TRAP = (IOPS & MC_TRAP) != 0;
    static constexpr bool INTEG = IOPS != 0;
    // Min / Max of a float column start from the FIRST valid value (minmax.go:14-24); the trapezoid joins runs through it
    static constexpr bool NEED_FIRST = FIRSTLAST || (MINMAX && !IS_INT) || TRAP;
    static constexpr bool NEED_LAST = FIRSTLAST;
    static constexpr bool NEED_TIME = INTEG;   // the accumulation reads the row's time
    static constexpr bool NEXT_VALUE = TRAP;   // the value of the inclusive row matters

    static __device__ __forceinline__ double val(uint64_t raw) {
        return IS_INT ? (double)(int64_t)raw : bits_as_f64(raw);  // GetFloat64, bowgetters.go:218-229
    }
    static __device__ __forceinline__ Inc no_inc() {
        Inc i;
        i.v = i.T = 0.0;
        i.has = 0;
        return i;
    }
    static __device__ __forceinline__ Inc make_inc(bool at_end, bool valid_next, uint64_t raw_next, int64_t t_next) {
        Inc i = no_inc();
        if (TRAP) {
            i.has = at_end && valid_next;
            i.v = val(raw_next);
            i.T = (double)t_next;
        }
        return i;
    }
    static __device__ __forceinline__ State identity() {
        State s;
        s.sum = 0.0;
        s.mn = CUDART_INF;
        s.mx = -CUDART_INF;
        s.first = s.last = 0;
        s.fT = s.lT = s.lV = 0.0;
        s.sS = s.sT = 0.0;
        s.cnt = 0;
        s.has = 0;
        return s;
    }
    // one valid row, left to right like the reference closures; the integral's joint term with the previous point is
    // formed unconditionally and selected away for the first point of a run (adding +0.0 never changes these sums)
    static __device__ __forceinline__ void accumulate(State &s, int64_t t, uint64_t raw) {
        const double v = val(raw);
        if (SUMS) s.sum += v;
        if (MINMAX) {
            if (v < s.mn) s.mn = v;  // minmax.go:20 `v < m`
            if (v > s.mx) s.mx = v;
        }
        if (INTEG) {
            const double T = (double)t;
            const double dt = T - s.lT;
            if (STEP) s.sS += s.has ? s.lV * dt : 0.0;            // integral.go:57
            if (TRAP) s.sT += s.has ? (s.lV + v) / 2 * dt : 0.0;  // integral.go:28
            s.lT = T;
            s.lV = v;
            s.has = 1;
        }
    }
    // count / first / last of the rows of one phase that joined the run cost nothing per row: popcount / ffs / fls of
    // their validity bits (bit j of mask = row j of the thread's phase)
    static __device__ __forceinline__ void note(State &s, uint32_t mask, const int64_t *trow, const uint64_t *vrow, const int swz) {
        if (mask) {
            if ((NEED_FIRST || INTEG) && s.cnt == 0) {
                const int j = (__ffs(mask) - 1) ^ swz;
                if (NEED_FIRST) s.first = vrow[j];
                if (INTEG) s.fT = (double)trow[j];
            }
            if (NEED_LAST) s.last = vrow[(31 - __clz(mask)) ^ swz];
            s.cnt += __popc(mask);
        }
    }
    // one synthetic row of the interpolated frame joins a run as its FIRST row (fused Interpolate -> Aggregate)
    static __device__ __forceinline__ void inject(State &s, int64_t t, uint64_t raw, bool valid) {
        if (!valid) return;
        accumulate(s, t, raw);
        if (s.cnt == 0) {
            s.first = raw;
            s.fT = (double)t;
        }
        s.last = raw;
        s.cnt += 1;
    }
    // L then R, adjacent runs of the same window (both noted: has == (cnt != 0))
    static __device__ __forceinline__ State combine(const State &L, const State &R) {
        State o = identity();
        const bool l = L.cnt != 0, r = R.cnt != 0;
        if (SUMS) o.sum = L.sum + R.sum;
        if (MINMAX) {
            o.mn = (R.mn < L.mn) ? R.mn : L.mn;
            o.mx = (R.mx > L.mx) ? R.mx : L.mx;
        }
        if (NEED_FIRST) o.first = l ? L.first : R.first;
        if (NEED_LAST) o.last = r ? R.last : L.last;
        if (INTEG) {
            const double dt = R.fT - L.lT;
            if (STEP) {
                const double j = (L.sS + L.lV * dt) + R.sS;
                o.sS = l ? (r ? j : L.sS) : R.sS;
            }
            if (TRAP) {
                const double j = (L.sT + (L.lV + val(R.first)) / 2 * dt) + R.sT;
                o.sT = l ? (r ? j : L.sT) : R.sT;
            }
            o.fT = l ? L.fT : R.fT;
            o.lT = r ? R.lT : L.lT;
            o.lV = r ? R.lV : L.lV;
        }
        o.cnt = L.cnt + R.cnt;
        o.has = o.cnt != 0;
        return o;
    }
    static __device__ __forceinline__ State shfl_up(const State &s, int d) {
        State o = identity();
        if (SUMS) o.sum = __shfl_up_sync(0xffffffffu, s.sum, d);
        if (MINMAX) {
            o.mn = __shfl_up_sync(0xffffffffu, s.mn, d);
            o.mx = __shfl_up_sync(0xffffffffu, s.mx, d);
        }
        if (NEED_FIRST) o.first = __shfl_up_sync(0xffffffffu, (unsigned long long)s.first, d);
        if (NEED_LAST) o.last = __shfl_up_sync(0xffffffffu, (unsigned long long)s.last, d);
        if (INTEG) {
            o.fT = __shfl_up_sync(0xffffffffu, s.fT, d);
            o.lT = __shfl_up_sync(0xffffffffu, s.lT, d);
            o.lV = __shfl_up_sync(0xffffffffu, s.lV, d);
            if (STEP) o.sS = __shfl_up_sync(0xffffffffu, s.sS, d);
            if (TRAP) o.sT = __shfl_up_sync(0xffffffffu, s.sT, d);
        }
        o.cnt = __shfl_up_sync(0xffffffffu, s.cnt, d);
        o.has = o.cnt != 0;
        return o;
    }
    // FINAL values of window k in the layout bow.NewBuffer + SetOrDrop leave behind (bowbuffer.go:22-80): value 0 in null
    // slots; Count = 0 and Sum = 0.0 are valid for windows without valid rows (count.go:10, sum.go:11-13); the validity of
    // everything else follows the two bitmaps set here.  `width` = float64(w.LastValue - w.FirstValue) (weightedmean.go).
    static __device__ __forceinline__ void write_final(const McColOut &o, const WindowGeom &g, const double width, int64_t k,
                                                       const State &s, const Inc &inc) {
        if ((uint64_t)k >= (uint64_t)g.W) return;
        const uint32_t c = s.cnt;
        const bool ok = c != 0;
        if (SUMS) {
            if (o.cnt) o.cnt[k] = (int64_t)c;
            if (o.sum) o.sum[k] = ok ? s.sum : 0.0;
            if (o.mean) o.mean[k] = ok ? __ddiv_rn(s.sum, (double)c) : 0.0;  // arithmeticmean.go:28
        }
        if (MINMAX) {
            const double f = bits_as_f64(s.first);
            const bool sticky = !IS_INT && (f != f);  // the first valid value is NaN (minmax.go:14-24)
            if (o.mn) o.mn[k] = ok ? (sticky ? f : s.mn) : 0.0;
            if (o.mx) o.mx[k] = ok ? (sticky ? f : s.mx) : 0.0;
        }
        if (FIRSTLAST) {
            if (o.first) o.first[k] = ok ? s.first : 0;
            if (o.last) o.last[k] = ok ? s.last : 0;
        }
        if (STEP) {
            double r = 0.0;
            if (ok) r = s.sS + s.lV * ((double)window_last_value(g, k) - s.lT);  // integral.go:53-57
            if (o.step) o.step[k] = r;
            if (o.wstep) o.wstep[k] = ok ? __ddiv_rn(r, width) : 0.0;            // weightedmean.go:17
        }
        if (ok && o.vb_cnt) atomicOr(o.vb_cnt + (k >> 5), 1u << (k & 31));
        if (TRAP) {
            const bool okt = c + (inc.has ? 1u : 0u) >= 2u;  // integral.go:33-35: fewer than two points -> nil
            double r = 0.0;
            if (okt) {
                r = s.sT;
                if (inc.has) r += (s.lV + inc.v) / 2 * (inc.T - s.lT);
            }
            if (o.trap) o.trap[k] = r;
            if (o.wlin) o.wlin[k] = okt ? __ddiv_rn(r, width) : 0.0;             // weightedmean.go:31
            if (okt && o.vb_trap) atomicOr(o.vb_trap + (k >> 5), 1u << (k & 31));
        }
    }
    // ---- chunk edge records ----------------------------------------------------------------------------------------
    static __device__ __forceinline__ State state_of(const McCarry &c) {
        State s;
        s.sum = c.sum;
        s.mn = c.mn;
        s.mx = c.mx;
        s.first = c.first;
        s.last = c.last;
        s.fT = c.fT;
        s.lT = c.lT;
        s.lV = c.lV;
        s.sS = c.sS;
        s.sT = c.sT;
        s.cnt = (uint32_t)(c.cnt & ~MC_CLOSED);
        s.has = s.cnt != 0;
        return s;
    }
    static __device__ __forceinline__ void set_state(McCarry &c, const State &s, bool closed) {
        c.sum = s.sum;
        c.mn = s.mn;
        c.mx = s.mx;
        c.first = s.first;
        c.last = s.last;
        c.fT = s.fT;
        c.lT = s.lT;
        c.lV = s.lV;
        c.sS = s.sS;
        c.sT = s.sT;
        c.cnt = (int64_t)s.cnt | (closed ? MC_CLOSED : 0);
    }
    static __device__ __forceinline__ void set_inc(McCarry &c, const Inc &i) {
        c.incV = i.v;
        c.incT = i.T;
        c.inc_has = TRAP && i.has;
    }
    static __device__ __forceinline__ Inc inc_of(const McCarry &c) {
        Inc i;
        i.v = c.incV;
        i.T = c.incT;
        i.has = c.inc_has != 0;
        return i;
    }
    static __device__ __forceinline__ McCarry make_carry(const State &s, const Inc &inc, int64_t key, bool closed) {
        McCarry c;
        c.key = key;
        set_state(c, s, closed);
        set_inc(c, inc);
        c.edge_t = 0;
        c.edge_raw = 0;
        c.edge_valid = 0;
        c._pad = 0;
        return c;
    }
};

// ---- kernel arguments -----------------------------------------------------------------------------------------------
struct McArgs {
    const int64_t *time;
    WindowGeom g;
    double width;          // float64(w.LastValue - w.FirstValue): the interval, or the span of the whole-Bow window
    int64_t ntiles;
    int32_t chunk_tiles;   // tiles per CTA (contiguous)
    int32_t ncols;
    uint32_t *touched;
    int32_t *status;
    McColArgs col[MC_MAXC];
};
struct McMaps {
    CUtensorMap time;
    CUtensorMap val[MC_MAXC];
};

// what the time pass of a tile leaves in registers for the column passes
struct McTile {
    uint64_t bm;       // bit j (j >= 1): my row j starts a later window than my row j - 1
    uint64_t gm;       // ... and not the very next one (windows without rows in between)
    uint64_t kf;       // window of my first row
    uint64_t nk;       // window of the next thread's first row (when next_has)
    uint64_t rowmask;  // my rows that exist and count (rows before s0 that are dropped are cleared)
    int64_t nt;        // time of the next thread's first row (when next_has)
    int64_t r0;        // first row of the tile
    bool has_rows, next_has, closes_right;
};
// one thread, one column, one tile
template <class Pol>
struct McThread {
    typename Pol::State st;    // open window (tail)
    typename Pol::State head;  // first window that closed in this thread (valid when nclose > 0)
    typename Pol::Inc inc_head;
    uint64_t kcur;             // index of the open window
    int nclose;
};

constexpr int MC_MAXSLOTS = 12;
template <class Pol>
struct McShared {
    uint64_t full[MC_MAXSLOTS], empty[MC_MAXSLOTS];
    struct WTot {
        typename Pol::State st;
        uint32_t flag, _pad;
    } wtot[2][MC_NW];
    struct Cta {  // the window open at the right edge of the CTA's previous tile
        typename Pol::State st;
        int64_t key;
        uint32_t have, reaches_left;  // reaches_left: no window boundary since the chunk began
    } cta[2][MC_MAXC];
    struct ColEdge {  // first row of the tile in the column being processed
        uint64_t raw;
        uint32_t valid, _pad;
    } edge[2];
    struct ChunkEdge {  // first row of the chunk
        int64_t t;
        uint64_t raw;
        uint32_t valid, _pad;
    } chunk_edge[MC_MAXC];
    int64_t tile_kf[2], tile_t0[2];
};

// ---- fused Interpolate -> Aggregate ---------------------------------------------------------------------------------
// In the interpolated frame (rolling/interpolation.go:118-161) window k holds [a synthetic row at S_k iff missing[k]] ++
// its own rows, and every EMPTY window holds exactly its synthetic row (written by mc_fixup_kernel).
// The synthetic row of window k as a one-row run (identity when the window has a real start row):
template <class Pol>
__device__ __forceinline__ typename Pol::State mc_syn_state(const WindowGeom &g, const FusedSyn &S, const uint64_t k) {
    typename Pol::State p = Pol::identity();
    if (k < (uint64_t)S.len && S.missing[k] != 0)
        Pol::inject(p, (int64_t)((uint64_t)g.s0 + k * g.div.d), S.val[k], S.ok[k] != 0);
    return p;
}
// Where window `kcur` ends and the next REAL row (x, raw, valid) lies in window `knew`: the inclusive row of kcur — the
// first row of window kcur + 1 in that frame: its synthetic row if it has one, else the real row iff it sits on E_kcur.
template <class Pol>
__device__ __forceinline__ typename Pol::Inc mc_fused_inc(const WindowGeom &g, const FusedSyn &S, const uint64_t kcur,
                                                          const bool has_row, const uint64_t knew, const int64_t x,
                                                          const uint64_t raw, const bool valid) {
    // the column ends (no window after kcur), or kcur is a halo window of a shard (not owned: nothing to produce)
    if (!has_row || kcur >= (uint64_t)g.W) return Pol::no_inc();
    const uint64_t d = g.div.d, len = (uint64_t)S.len;
    const bool miss_new = knew < len && S.missing[knew] != 0;
    const uint64_t kn = kcur + 1;
    if (knew == kn && !miss_new) return Pol::make_inc(x == (int64_t)((uint64_t)g.s0 + kn * d), valid, raw, x);
    return kn < len ? Pol::make_inc(true, S.ok[kn] != 0, S.val[kn], (int64_t)((uint64_t)g.s0 + kn * d)) : Pol::no_inc();
}

// one copy of the final write per call site family (it is a dozen predicated stores and two atomics)
template <class Pol>
__device__ __noinline__ void mc_write(const McArgs &A, const McColArgs &col, const int64_t k, const typename Pol::State s,
                                      const typename Pol::Inc inc) {
    Pol::write_final(col.out, A.g, A.width, k, s, inc);
}

// ---- time pass --------------------------------------------------------------------------------------------------------
// a warp stages the rows of its own 32 threads (tiles at the edges of the column, staged by plain loads)
__device__ __forceinline__ void mc_stage_rows(uint64_t *box, const uint64_t *src, const int64_t n, const int64_t r0, const int p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int e = lane; e < 32 * MC_P; e += 32) {
        const int tt = warp * 32 + e / MC_P, cc = e % MC_P;
        const int64_t row = r0 + (int64_t)tt * MC_RE + p * MC_P + cc;
        box[tt * MC_P + (cc ^ mc_swz(tt))] = row < n ? src[row] : 0;
    }
    __syncwarp();
}

// bits [k0, k1] of the bitmap of windows that hold a row
__device__ __noinline__ static void mc_touch_range(uint32_t *touched, uint64_t k0, uint64_t k1, const int64_t W) {
    if (W <= 0 || k0 >= (uint64_t)W) return;
    if (k1 >= (uint64_t)W) k1 = (uint64_t)W - 1;
    for (uint64_t w = k0 >> 5; w <= (k1 >> 5) && k0 <= k1; ++w) {
        const uint64_t lo = w << 5;
        const uint32_t a = k0 > lo ? (uint32_t)(k0 - lo) : 0u, b = k1 < lo + 31 ? (uint32_t)(k1 - lo) : 31u;
        atomicOr(touched + w, (0xFFFFFFFFu >> (31 - b)) & (0xFFFFFFFFu << a));
    }
}
struct McGap {
    uint64_t kcur;
    int64_t eabs;
};
// row with time x lies beyond the window AFTER the open one: windows without rows in between (out of line, rare)
__device__ __noinline__ static McGap mc_time_gap(const McArgs &A, const int64_t x, const uint64_t kset, const uint64_t kold) {
    const WindowGeom &g = A.g;
    mc_touch_range(A.touched, kset, kold, g.W);
    McGap r;
    r.kcur = div_u64((uint64_t)x - (uint64_t)g.s0, g.div);
    r.eabs = (int64_t)((uint64_t)g.s0 + (r.kcur + 1) * g.div.d);
    return r;
}

// state of the time scan of one thread over one tile (registers)
struct McScan {
    uint64_t bm, gm;
    uint64_t kcur;   // window of the last row seen
    uint64_t kset;   // first window whose `touched` bit is still to be set
    int64_t eabs;    // absolute end of window kcur
    int64_t xlast;
};

// one phase of the time pass: order check and window boundary test of every row
template <bool FULL>
__device__ __forceinline__ void mc_time_phase(McScan &sc, bool &bad, const McArgs &A, const int64_t *trow, const int swz,
                                              const int p, const int nm) {
    const uint64_t d = A.g.div.d;
    const longlong2 *t2 = reinterpret_cast<const longlong2 *>(trow);
    uint32_t bmp = 0, gmp = 0;
    auto row = [&](const int j, const int64_t x) {
        if (FULL || j < nm) {
            bad |= x < sc.xlast;
            sc.xlast = x;
            if (x >= sc.eabs) {  // row j starts a later window
                bmp |= 1u << j;
                if ((uint64_t)x - (uint64_t)sc.eabs < d) {
                    ++sc.kcur;
                    sc.eabs = (int64_t)((uint64_t)sc.eabs + d);
                } else {
                    gmp |= 1u << j;
                    const McGap r = mc_time_gap(A, x, sc.kset, sc.kcur);
                    sc.kcur = sc.kset = r.kcur;
                    sc.eabs = r.eabs;
                }
            }
        }
    };
#pragma unroll
    for (int q = 0; q < MC_P / 2; ++q) {
        const longlong2 tq = t2[q ^ (swz >> 1)];
        row(2 * q, tq.x);
        row(2 * q + 1, tq.y);
    }
    sc.bm |= (uint64_t)bmp << (p * MC_P);
    sc.gm |= (uint64_t)gmp << (p * MC_P);
}

// The same for a complete tile, without a branch per row: a boundary advances the window end by one interval under a
// predicate; a row that lies beyond even that end (windows without rows in between: rare) sends the whole warp through
// the generic version above.
__device__ __forceinline__ void mc_time_phase_fast(McScan &sc, bool &bad, const McArgs &A, const int64_t *trow, const int swz,
                                                   const int p) {
    const uint64_t d = A.g.div.d;
    const longlong2 *t2 = reinterpret_cast<const longlong2 *>(trow);
    uint32_t bmp = 0;  // (bit of row j enters at bit 15 and is shifted down 15 - j times)
    int64_t eabs = sc.eabs, xl = sc.xlast;
    bool gap = false, unsorted = false;
    auto row = [&](const int64_t x) {
        unsorted |= x < xl;
        xl = x;
        const bool ge = x >= eabs;
        bmp = (bmp >> 1) | (ge ? 0x8000u : 0u);
        if (ge) eabs = (int64_t)((uint64_t)eabs + d);
        gap |= x >= eabs;
    };
#pragma unroll(MC_UNROLL)
    for (int q = 0; q < MC_P / 2; ++q) {
        const longlong2 tq = t2[q ^ (swz >> 1)];
        row(tq.x);
        row(tq.y);
    }
    if (__any_sync(0xffffffffu, gap)) {
        mc_time_phase<true>(sc, bad, A, trow, swz, p, MC_P);
        return;
    }
    bad |= unsorted;
    sc.eabs = eabs;
    sc.xlast = xl;
    sc.kcur += __popc(bmp);
    sc.bm |= (uint64_t)bmp << (p * MC_P);
}

// ---- value pass ---------------------------------------------------------------------------------------------------------
// A window that closed inside the rows of a thread is complete (its rows of the current phase already noted): its
// index, its inclusive row, and either the final values or — for the first window a thread closes, which began at or
// before its first row — the head that the stitch combines with the neighbours.  Row j of the phase started the next
// window.  trow == nullptr: the time boxes are not resident (one global read of that row's time where it matters).
template <class Pol, bool FUSED>
__device__ __forceinline__ void mc_flush(McThread<Pol> &c, const typename Pol::State &q, const int j, const McArgs &A,
                                         const McColArgs &col, const int64_t *trow, const uint64_t *vrow, const int swz,
                                         const uint32_t vb16, const uint32_t gm16, const int64_t grow0) {
    using State = typename Pol::State;
    using Inc = typename Pol::Inc;
    const WindowGeom &g = A.g;
    const uint64_t d = g.div.d;
    const bool gap = (gm16 >> j) & 1u;
    const uint64_t kold = c.kcur;
    int64_t xj = 0;
    if (trow)
        xj = trow[j ^ swz];
    else if (gap || FUSED)
        xj = A.time[grow0 + j];
    c.kcur = gap ? div_u64((uint64_t)xj - (uint64_t)g.s0, g.div) : kold + 1;
    Inc inc = Pol::no_inc();
    if (FUSED || Pol::NEXT_VALUE) {
        const uint64_t rj = vrow[j ^ swz];
        const bool vj = (vb16 >> j) & 1u;
        if (FUSED)
            inc = mc_fused_inc<Pol>(g, col.syn, kold, true, c.kcur, xj, rj, vj);
        else
            inc = Pol::make_inc(xj == (int64_t)((uint64_t)g.s0 + (kold + 1) * d), vj, rj, xj);
    }
    if (c.nclose == 0) {  // (began at or before my first row: its synthetic row, if any, comes with the neighbour's tail)
        c.head = q;
        c.inc_head = inc;
    } else {
        State w = q;
        if (FUSED) w = Pol::combine(mc_syn_state<Pol>(g, col.syn, kold), q);  // began inside my rows: synthetic row first
        mc_write<Pol>(A, col, (int64_t)kold, w, inc);
    }
    ++c.nclose;
}

// One phase of one column, generic: any number of window boundaries per thread, rows that do not exist (tiles at the
// edges of the column).  The loop is rolled: this is the slow path.
template <class Pol, bool HAS_NULLS, bool FULL, bool FUSED>
__device__ __forceinline__ void mc_value_phase(McThread<Pol> &c, const uint64_t bm, const uint64_t gm, const McArgs &A,
                                            const McColArgs &col, const int p, const int64_t *trow, const uint64_t *vrow,
                                            const uint64_t vbits, const int64_t grow0) {
    const int swz = mc_swz(threadIdx.x);
    const uint32_t bm16 = (uint32_t)(bm >> (p * MC_P)) & 0xFFFFu;
    const uint32_t gm16 = (uint32_t)(gm >> (p * MC_P)) & 0xFFFFu;
    const uint32_t vb16 = (uint32_t)(vbits >> (p * MC_P)) & 0xFFFFu;
    int segstart = 0;  // first row of this phase that belongs to the open window
#pragma unroll 1
    for (int j = 0; j < MC_P; ++j) {
        if ((bm16 >> j) & 1u) {  // row j starts a later window: the open one is complete
            Pol::note(c.st, vb16 & mc_below(j) & ~mc_below(segstart), trow, vrow, swz);
            mc_flush<Pol, FUSED>(c, c.st, j, A, col, trow, vrow, swz, vb16, gm16, grow0);
            c.st = Pol::identity();
            segstart = j;
        }
        if ((vb16 >> j) & 1u) Pol::accumulate(c.st, Pol::NEED_TIME ? trow[j ^ swz] : 0, vrow[j ^ swz]);
    }
    Pol::note(c.st, vb16 & ~mc_below(segstart), trow, vrow, swz);
}

// One phase of one column of a complete tile, without a branch per row.  A thread meets its window boundaries at other
// rows than its neighbours, so a branch taken once per window by a thread is taken several times per phase by the warp.
// Instead: the rows before the thread's boundary accumulate into `a` (the open window), the rows from it on into `b`
// (the next one), every update under a per-row predicate; the boundary itself is dealt with once, after the loop.
// A phase in which some thread of the warp has two boundaries (windows of a few rows) takes the generic path.
template <class Pol, bool HAS_NULLS, bool FUSED>
__device__ __forceinline__ void mc_value_phase_fast(McThread<Pol> &c, const uint64_t bm, const uint64_t gm, const McArgs &A,
                                                    const McColArgs &col, const int p, const int64_t *trow,
                                                    const uint64_t *vrow, const uint64_t vbits, const int64_t grow0) {
    using State = typename Pol::State;
    const int swz = mc_swz(threadIdx.x);
    const uint32_t bm16 = (uint32_t)(bm >> (p * MC_P)) & 0xFFFFu;
    const int nb = __popc(bm16);
    if (__any_sync(0xffffffffu, nb >= 2)) {
        mc_value_phase<Pol, HAS_NULLS, true, FUSED>(c, bm, gm, A, col, p, trow, vrow, vbits, grow0);
        return;
    }
    const uint32_t gm16 = (uint32_t)(gm >> (p * MC_P)) & 0xFFFFu;
    const uint32_t vb16 = HAS_NULLS ? (uint32_t)(vbits >> (p * MC_P)) & 0xFFFFu : 0xFFFFu;
    const longlong2 *t2 = reinterpret_cast<const longlong2 *>(trow);
    const ulonglong2 *v2 = reinterpret_cast<const ulonglong2 *>(vrow);
    if (!__any_sync(0xffffffffu, nb != 0)) {  // no boundary in the rows of the whole warp: one run, nothing to select
        State a = c.st;
        uint32_t vbq = vb16;
#pragma unroll(MC_UNROLL)
        for (int q = 0; q < MC_P / 2; ++q) {
            const ulonglong2 vq = v2[q ^ (swz >> 1)];
            longlong2 tq = make_longlong2(0, 0);
            if (Pol::NEED_TIME) tq = t2[q ^ (swz >> 1)];
            if (!HAS_NULLS || (vbq & 1u)) Pol::accumulate(a, tq.x, vq.x);
            if (!HAS_NULLS || (vbq & 2u)) Pol::accumulate(a, tq.y, vq.y);
            vbq >>= 2;
        }
        Pol::note(a, vb16, trow, vrow, swz);
        c.st = a;
        return;
    }
    const int jb = nb ? __ffs(bm16) - 1 : MC_P;
    const uint32_t mA = vb16 & mc_below(jb), mB = vb16 & ~mc_below(jb);
    State a = c.st, b = Pol::identity();
    // integrals: ONE chain of (time, value) of the last valid row serves both runs - the first row of `b` has no
    // predecessor in its run (hasB is false), every other row's predecessor is the last valid row before it
    double cT = a.lT, cV = a.lV;
    bool hasA = a.has != 0, hasB = false;
    uint32_t mAq = mA, mBq = mB;  // bits 0 / 1 = the two rows of the trip
    auto row = [&](const uint32_t bit, const int64_t t, const uint64_t raw) {
        const bool pA = (mAq & bit) != 0;
        const bool pB = HAS_NULLS ? (mBq & bit) != 0 : !pA;
        const double v = Pol::val(raw);
        if (Pol::SUMS) {
            mc_padd(a.sum, v, mAq, bit);
            mc_padd(b.sum, v, mBq, bit);
        }
        if (Pol::MINMAX) {
            if (pA && v < a.mn) a.mn = v;  // minmax.go:20 `v < m`
            if (pB && v < b.mn) b.mn = v;
            if (pA && v > a.mx) a.mx = v;
            if (pB && v > b.mx) b.mx = v;
        }
        if (Pol::INTEG) {
            const double T = (double)t;
            const double dt = T - cT;
            if (Pol::STEP) {
                const double s = cV * dt;  // integral.go:57
                mc_padd(a.sS, s, hasA ? mAq : 0u, bit);
                mc_padd(b.sS, s, hasB ? mBq : 0u, bit);
            }
            if (Pol::TRAP) {
                const double s = (cV + v) / 2 * dt;  // integral.go:28
                mc_padd(a.sT, s, hasA ? mAq : 0u, bit);
                mc_padd(b.sT, s, hasB ? mBq : 0u, bit);
            }
            if (!HAS_NULLS || pA || pB) {
                cT = T;
                cV = v;
            }
            hasA |= pA;
            hasB |= pB;
        }
    };
#pragma unroll(MC_UNROLL)
    for (int q = 0; q < MC_P / 2; ++q) {
        const ulonglong2 vq = v2[q ^ (swz >> 1)];
        longlong2 tq = make_longlong2(0, 0);
        if (Pol::NEED_TIME) tq = t2[q ^ (swz >> 1)];
        row(1u, tq.x, vq.x);
        row(2u, tq.y, vq.y);
        mAq >>= 2;
        mBq >>= 2;
    }
    if (Pol::INTEG) {
        if (nb == 0) {  // every row joined the open window
            a.lT = cT;
            a.lV = cV;
        } else {
            if (mA) {  // the open window's last point is its last valid row of this phase
                const int jl = (31 - __clz(mA)) ^ swz;
                a.lT = (double)trow[jl];
                a.lV = Pol::val(vrow[jl]);
            }
            b.lT = cT;
            b.lV = cV;
        }
        a.has = hasA;
        b.has = hasB;
    }
    Pol::note(a, mA, trow, vrow, swz);
    if (nb) {
        Pol::note(b, mB, trow, vrow, swz);
        mc_flush<Pol, FUSED>(c, a, jb, A, col, trow, vrow, swz, vb16, gm16, grow0);
        c.st = b;
    } else {
        c.st = a;
    }
}

// ---- end of a tile (per column): stitch the per-thread pieces -----------------------------------------------------------
template <class Pol, bool FUSED>
__device__ __forceinline__ void mc_stitch(McThread<Pol> &c, const uint64_t nraw, const uint32_t nvalid, const McTile &ti,
                                          const McArgs &A, const McColArgs &col, McShared<Pol> &sh, const int ci, const int cp,
                                          const int tp, const bool last_tile, const int64_t xlast_tile) {
    using State = typename Pol::State;
    using Inc = typename Pol::Inc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const WindowGeom &g = A.g;
    const uint64_t d = g.div.d;

    // fused: the open window began inside my rows when I closed one before it: its synthetic row comes first
    if (FUSED && c.nclose) c.st = Pol::combine(mc_syn_state<Pol>(g, col.syn, c.kcur), c.st);

    // ---- the boundary at my right edge -------------------------------------------------------------------------------
    bool tail_open = ti.has_rows;  // (meaningful for the last thread of the tile only)
    if (ti.closes_right) {
        State fresh = Pol::identity();
        Inc inc = Pol::no_inc();
        if (ti.next_has) {
            if (FUSED) {  // (the window the next thread starts in begins with its synthetic row)
                inc = mc_fused_inc<Pol>(g, col.syn, c.kcur, true, ti.nk, ti.nt, nraw, nvalid != 0);
                if (c.kcur < (uint64_t)g.W) fresh = mc_syn_state<Pol>(g, col.syn, ti.nk);
            } else {
                inc = Pol::make_inc(ti.nt == (int64_t)((uint64_t)g.s0 + (c.kcur + 1) * d), nvalid != 0, nraw, ti.nt);
            }
        }
        if (c.nclose == 0) {
            c.head = c.st;
            c.inc_head = inc;
        } else {
            mc_write<Pol>(A, col, (int64_t)c.kcur, c.st, inc);
        }
        ++c.nclose;
        c.st = fresh;  // the tail now belongs to the window the next thread starts in
        tail_open = false;
    }

    // ---- windows spanning threads: segmented inclusive scan of the tails (flag = a window closed in the thread) -----------
    const uint32_t ball = __ballot_sync(0xffffffffu, c.nclose != 0);
    State sc = c.st;
    {
        // step dd changes some lane only if dd consecutive lanes (the lowest >= 1) carry no flag: skip the others
        uint32_t run = ~ball & ~1u;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            if (run == 0) break;  // (warp uniform)
            const State o = Pol::shfl_up(sc, dd);
            // no flag in lanes (lane-dd, lane]  <=>  the run ending at lane-dd belongs to my open segment
            const uint32_t span = lane >= dd ? (0xFFFFFFFFu >> (31 - lane)) & ~((2u << (lane - dd)) - 1u) : 0u;
            if (lane >= dd && (ball & span) == 0) sc = Pol::combine(o, sc);
            run &= run << dd;
        }
    }
    if (lane == 31) {
        sh.wtot[cp][warp].st = sc;
        sh.wtot[cp][warp].flag = ball != 0;
    }
    State ex = Pol::shfl_up(sc, 1);
    if (lane == 0) ex = Pol::identity();
    mc_bar_consumers();

    // ---- the window open at the left edge of the tile (carried from the CTA's previous tile) ---------------------------
    const typename McShared<Pol>::Cta &cta = sh.cta[tp][ci];
    const int64_t tile_kf = sh.tile_kf[tp];
    const bool cta_have = cta.have != 0;
    const bool cta_same = cta_have && cta.key == tile_kf;     // it continues into this tile
    const bool cta_ends = cta_have && !cta_same;              // it ended exactly at the tile boundary
    const bool left_open = !cta_have || (cta_same && cta.reaches_left != 0);  // tile's first window reaches the chunk's left edge
    State acc = Pol::identity();
    if (cta_same) acc = cta.st;
    if (cta_ends) {
        Inc inc;
        const int64_t t0 = sh.tile_t0[tp];
        if (FUSED) {
            inc = mc_fused_inc<Pol>(g, col.syn, (uint64_t)cta.key, true, (uint64_t)tile_kf, t0, sh.edge[cp].raw, sh.edge[cp].valid != 0);
            if ((uint64_t)cta.key < (uint64_t)g.W) acc = mc_syn_state<Pol>(g, col.syn, (uint64_t)tile_kf);
        } else {
            inc = Pol::make_inc(t0 == (int64_t)((uint64_t)g.s0 + ((uint64_t)cta.key + 1) * d), sh.edge[cp].valid != 0,
                                sh.edge[cp].raw, t0);
        }
        if (tid == 0) {
            if (cta.reaches_left) {  // ... and began at or before the chunk's first row: the fix-up decides
                McCarry r = Pol::make_carry(cta.st, inc, cta.key, true);
                r.edge_t = sh.chunk_edge[ci].t;
                r.edge_raw = sh.chunk_edge[ci].raw;
                r.edge_valid = sh.chunk_edge[ci].valid;
                col.rec[2 * blockIdx.x] = r;
            } else {
                mc_write<Pol>(A, col, cta.key, cta.st, inc);
            }
        }
    }
    bool any_prev = cta_ends;  // a window closed before my warp
    for (int u = 0; u < warp; ++u) {
        const State ws = sh.wtot[cp][u].st;
        if (sh.wtot[cp][u].flag) {
            acc = ws;
            any_prev = true;
        } else {
            acc = Pol::combine(acc, ws);
        }
    }
    const bool flag_before = (ball & ((1u << lane) - 1u)) != 0;
    const State excl = flag_before ? ex : Pol::combine(acc, ex);
    const bool any_excl = any_prev || flag_before;

    if (c.nclose) {  // this thread closes the window that was open at its left edge
        const State hd = Pol::combine(excl, c.head);
        if (!any_excl && left_open) {  // ... which reaches the left edge of the chunk: the fix-up decides whether it began earlier
            McCarry r = Pol::make_carry(hd, c.inc_head, (int64_t)ti.kf, true);
            if (cta_have) {
                r.edge_t = sh.chunk_edge[ci].t;
                r.edge_raw = sh.chunk_edge[ci].raw;
                r.edge_valid = sh.chunk_edge[ci].valid;
            } else {  // first tile of the chunk
                r.edge_t = sh.tile_t0[tp];
                r.edge_raw = sh.edge[cp].raw;
                r.edge_valid = sh.edge[cp].valid;
            }
            col.rec[2 * blockIdx.x] = r;
        } else {
            mc_write<Pol>(A, col, (int64_t)ti.kf, hd, c.inc_head);
        }
    }
    if (tid == MC_NT - 1) {  // the window open at the right edge of the tile
        const bool flag_incl = flag_before || c.nclose != 0;
        const State incl = flag_incl ? sc : Pol::combine(acc, sc);
        const bool any_incl = any_prev || flag_incl;
        const bool reaches = left_open && !any_incl;
        if (!cta_have) {  // first tile of the chunk: remember its first row for the head record
            sh.chunk_edge[ci].t = sh.tile_t0[tp];
            sh.chunk_edge[ci].raw = sh.edge[cp].raw;
            sh.chunk_edge[ci].valid = sh.edge[cp].valid;
        }
        typename McShared<Pol>::Cta &nx = sh.cta[tp ^ 1][ci];
        nx.st = incl;
        nx.key = (int64_t)c.kcur;
        nx.have = tail_open;
        nx.reaches_left = reaches;
        if (last_tile) {  // chunk records
            const Inc noinc = Pol::no_inc();
            McCarry tl = Pol::make_carry(incl, noinc, -1, false);
            tl.edge_t = xlast_tile;  // last row of the chunk (cross-chunk order check)
            if (reaches) {  // no boundary anywhere in the chunk: it lies inside one window
                McCarry hr = Pol::make_carry(incl, noinc, tile_kf, false);
                if (cta_have) {
                    hr.edge_t = sh.chunk_edge[ci].t;
                    hr.edge_raw = sh.chunk_edge[ci].raw;
                    hr.edge_valid = sh.chunk_edge[ci].valid;
                } else {
                    hr.edge_t = sh.tile_t0[tp];
                    hr.edge_raw = sh.edge[cp].raw;
                    hr.edge_valid = sh.edge[cp].valid;
                }
                col.rec[2 * blockIdx.x] = hr;
            } else if (tail_open) {
                tl.key = (int64_t)c.kcur;
            }
            col.rec[2 * blockIdx.x + 1] = tl;
        }
    }
}

// Boxes in shared memory.  Policies whose rows need their time (integrals) keep the NP time boxes of the tile RESIDENT in
// slots [0, NP) while the value boxes of all columns go through a ring of the other slots; otherwise every slot is part
// of ONE ring that carries time and value boxes in the order they are consumed (a time box is released as soon as its
// phase of the time pass is over).  Order of a tile's boxes: T(p), V(0, p) for p = 0..NP-1 — the time pass runs
// interleaved with the first column — then V(c, p) for the other columns.
#ifdef MC_DEBUG_CLOCKS
__device__ unsigned long long mc_dbg[8];
#define MC_CLK(...) __VA_ARGS__
#else
#define MC_CLK(...)
#endif
template <class Pol, bool HAS_NULLS, bool FUSED>
__global__ void __launch_bounds__(MC_THREADS, MC_CFG_CTAS)
    segmc_kernel(const __grid_constant__ McArgs A, const __grid_constant__ McMaps M) {
    constexpr bool RESIDENT = Pol::NEED_TIME;
    constexpr int RING0 = RESIDENT ? MC_NP : 0;  // first slot of the ring
    constexpr int S = MC_R - RING0 > 0 ? MC_R - RING0 : 1;  // its length
    static_assert(S >= 1, "the ring needs a box");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    McShared<Pol> &sh = *reinterpret_cast<McShared<Pol> *>(smem_raw);
    static_assert(sizeof(McShared<Pol>) <= MC_HEADER, "header layout");
    uint8_t *boxes = smem_raw + MC_HEADER;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const WindowGeom &g = A.g;
    const int64_t tile_lo = (int64_t)blockIdx.x * A.chunk_tiles;
    int64_t tile_hi = tile_lo + A.chunk_tiles;
    if (tile_hi > A.ntiles) tile_hi = A.ntiles;
    auto tile_full = [&](int64_t tile) {  // staged by TMA: whole tile + one more row exist, no row before s0
        const int64_t r0 = tile * MC_T;
        return g.n - r0 > MC_T && r0 >= g.early_rows;
    };
    if (tid == 0) {
        for (int s = 0; s < MC_R; ++s) {
            mbar_init(&sh.full[s], 1);
            mbar_init(&sh.empty[s], MC_NW);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    if (tid < MC_MAXC) {
        sh.cta[0][tid].have = 0;
        sh.cta[0][tid].reaches_left = 0;
        sh.cta[0][tid].key = -1;
    }
    __syncthreads();

    if (warp == MC_NW) {  // ---- producer: one thread issues every TMA box of the chunk ---------------------------------
        if (lane != 0) return;
        uint32_t it = 0;    // ring items issued
        MC_CLK(long long pw = 0; const long long pt0 = clock64();)
        uint32_t tpar = 1;  // resident time slots: a fresh mbarrier passes a wait on parity 1
        auto issue = [&](const int slot, const uint32_t par, const CUtensorMap *tm, const int p, const int32_t trow, const bool fullt) {
            MC_CLK(const long long w0 = clock64();)
            mbar_wait_backoff(&sh.empty[slot], par);
            MC_CLK(pw += clock64() - w0;)
            if (fullt) {
                mbar_arrive_expect_tx(&sh.full[slot], MC_BOX);
                tma_box_2d(boxes + slot * MC_BOX, tm, p * MC_P, trow, &sh.full[slot]);
            } else {
                mbar_arrive(&sh.full[slot]);  // (staged by the consumers themselves)
            }
        };
        for (int64_t tile = tile_lo; tile < tile_hi; ++tile, tpar ^= 1u) {
            const bool fullt = tile_full(tile);
            const int32_t trow = (int32_t)(tile * MC_NT);
            for (int ci = 0; ci < A.ncols; ++ci)
                for (int p = 0; p < MC_NP; ++p) {
                    if (ci == 0) {
                        if (RESIDENT) {
                            issue(p, tpar, &M.time, p, trow, fullt);
                        } else {
                            issue((int)(it % S), ((it / S) & 1u) ^ 1u, &M.time, p, trow, fullt);
                            ++it;
                        }
                    }
                    issue(RING0 + (int)(it % S), ((it / S) & 1u) ^ 1u, &M.val[ci], p, trow, fullt);
                    ++it;
                }
        }
        MC_CLK(atomicAdd(&mc_dbg[0], (unsigned long long)pw); atomicAdd(&mc_dbg[1], (unsigned long long)(clock64() - pt0));)
        return;
    }

    // ---- consumers ------------------------------------------------------------------------------------------------------
    const int swz = mc_swz(tid);
    const uint64_t d = g.div.d;
    uint32_t it = 0, tpar = 0;
    int tp = 0, cp = 0;  // parities of the double-buffered exchange areas: per tile, per stitch
    bool bad = false;
    MC_CLK(long long cwt = 0, cwv = 0, cst = 0, ctp = 0, cvp = 0; const long long ct0 = clock64();)
    auto release = [&](const int slot, const bool staged) {  // (after __syncwarp)
        if (lane == 0) {
            if (staged) fence_proxy_async();  // plain stores into the box precede the next TMA write
            mbar_arrive(&sh.empty[slot]);
        }
    };
    for (int64_t tile = tile_lo; tile < tile_hi; ++tile, tpar ^= 1u, tp ^= 1) {
        const bool fullt = tile_full(tile);
        const bool last_tile = tile + 1 == tile_hi;
        const int64_t r0 = tile * MC_T, my0 = r0 + (int64_t)tid * MC_RE;
        int nrows = MC_RE, early = 0;
        if (!fullt) {
            const int64_t left = g.n - my0;
            nrows = left < 0 ? 0 : (left > MC_RE ? MC_RE : (int)left);
            const int64_t e = g.early_rows - my0;  // rows before s0 (negative timestamps / left halo of a shard)
            early = e < 0 ? 0 : (e > nrows ? nrows : (int)e);
        }
        McTile ti;
        ti.r0 = r0;
        ti.has_rows = nrows > 0;
        ti.rowmask = nrows == 64 ? ~0ull : ((1ull << nrows) - 1ull);
        if (early > 0 && !g.early_keep) ti.rowmask &= early == 64 ? 0ull : ~((1ull << early) - 1ull);
        ti.bm = ti.gm = ti.kf = ti.nk = 0;
        ti.nt = 0;
        ti.next_has = ti.closes_right = false;
        McScan scan;
        scan.bm = scan.gm = scan.kcur = scan.kset = 0;
        scan.eabs = scan.xlast = 0;
        int64_t x0 = 0, tprev = 0;
        // time of the tile's last row (chunk tail record: cross-chunk order check)
        int64_t xlast_tile = 0;
        if (last_tile && tid == MC_NT - 1) {
            int64_t lr = r0 + MC_T - 1;
            if (lr >= g.n) lr = g.n - 1;
            xlast_tile = A.time[lr];
        }
#pragma unroll 1
        for (int ci = 0; ci < A.ncols; ++ci) {
            const McColArgs &col = A.col[ci];
            McThread<Pol> c;
            c.st = Pol::identity();
            c.head = Pol::identity();
            c.inc_head = Pol::no_inc();
            c.kcur = ti.kf;
            c.nclose = 0;
            uint64_t nraw = 0;
            uint32_t nvalid = 0;
            uint64_t vbits = ti.rowmask;
            uint32_t vnext = 1;  // validity of the next thread's first row
            if (HAS_NULLS) {
                const uint64_t *bw = reinterpret_cast<const uint64_t *>(col.validity);
                const int64_t nwords = (((g.n + 7) / 8 + 15) & ~(int64_t)15) / 8;  // device bitmaps are padded to 16 bytes
                const int64_t w0 = my0 >> 6;
                const int sh0 = (int)(my0 & 63);                                    // (0 when RE = 64)
                uint64_t w = w0 < nwords ? bw[w0] : 0ull;
                if (MC_RE < 64 && sh0) w = (w >> sh0) | ((w0 + 1 < nwords ? bw[w0 + 1] : 0ull) << (64 - sh0));
                vbits &= w;
                if (Pol::NEXT_VALUE || FUSED) {
                    const int64_t nb = my0 + MC_RE;
                    vnext = (nb >> 6) < nwords ? (uint32_t)((bw[nb >> 6] >> (nb & 63)) & 1ull) : 0u;
                }
            }
#pragma unroll 1
            for (int p = 0; p < MC_NP; ++p) {
                int nm = MC_P;
                if (!fullt) {
                    nm = nrows - p * MC_P;
                    nm = nm < 0 ? 0 : (nm > MC_P ? MC_P : nm);
                }
                if (ci == 0) {  // ---- the time pass runs interleaved with the first column -------------------------------
                    int tslot = p;  // (RESIDENT)
                    uint32_t par = tpar;
                    if (!RESIDENT) {
                        tslot = (int)(it % S);
                        par = (it / S) & 1u;
                        ++it;
                    }
                    MC_CLK(long long w0 = clock64();)
                    mbar_wait(&sh.full[tslot], par);
                    MC_CLK(cwt += clock64() - w0; w0 = clock64();)
                    int64_t *tbox = reinterpret_cast<int64_t *>(boxes + tslot * MC_BOX);
                    if (!fullt) mc_stage_rows(reinterpret_cast<uint64_t *>(tbox), reinterpret_cast<const uint64_t *>(A.time), g.n, r0, p);
                    const int64_t *trow = tbox + tid * MC_P;
                    if (p == 0 && nrows > 0) {  // window of my first row and its absolute end (rows before s0 collapse onto window 0)
                        x0 = trow[0 ^ swz];
                        scan.kcur = early > 0 ? 0 : div_u64((uint64_t)x0 - (uint64_t)g.s0, g.div);
                        scan.kset = scan.kcur + 1;
                        ti.kf = scan.kcur;
                        c.kcur = ti.kf;
                        scan.eabs = (int64_t)((uint64_t)g.s0 + (scan.kcur + 1) * d);
                        scan.xlast = x0;
                        // the row after mine (the next thread's first row)
                        if (fullt) {
                            ti.next_has = tid < MC_NT - 1;
                            if (ti.next_has) ti.nt = tbox[(tid + 1) * MC_P + (0 ^ mc_swz(tid + 1))];
                        } else {
                            const int64_t nxt = my0 + MC_RE;
                            ti.next_has = nrows == MC_RE && tid < MC_NT - 1 && nxt < g.n;
                            if (ti.next_has) ti.nt = A.time[nxt];
                        }
                    }
                    if (fullt)
                        mc_time_phase_fast(scan, bad, A, trow, swz, p);
                    else if (nm > 0)
                        mc_time_phase<false>(scan, bad, A, trow, swz, p, nm);
                    if (p == MC_NP - 1 && nrows > 0) {  // the row before mine: order across threads / tiles
                        bool has_prev;
                        if (fullt) {
                            has_prev = tid > 0 || r0 > 0;
                            if (tid > 0)
                                tprev = tbox[(tid - 1) * MC_P + ((MC_P - 1) ^ mc_swz(tid - 1))];
                            else if (r0 > 0)
                                tprev = A.time[r0 - 1];
                        } else {
                            has_prev = my0 > 0;
                            if (has_prev) tprev = A.time[my0 - 1];
                        }
                        bool starts_new = true;  // does my first row begin its window?
                        if (has_prev) {
                            bad |= x0 < tprev;
                            starts_new = tprev < (int64_t)((uint64_t)g.s0 + ti.kf * d);
                        }
                        // windows that begin in my rows hold a row
                        const uint64_t k0 = (starts_new && scan.kset == ti.kf + 1) ? ti.kf : scan.kset;
                        if (starts_new && scan.kset != ti.kf + 1 && ti.kf < (uint64_t)g.W)
                            atomicOr(A.touched + (ti.kf >> 5), 1u << (ti.kf & 31));
                        if (k0 <= scan.kcur) mc_touch_range(A.touched, k0, scan.kcur, g.W);
                    }
                    ti.bm = scan.bm;
                    ti.gm = scan.gm;
                    MC_CLK(ctp += clock64() - w0;)
                    if (!RESIDENT) {
                        __syncwarp();
                        release(tslot, !fullt);
                    }
                }
                // ---- value box of (column, phase) ------------------------------------------------------------------------
                const int vslot = RING0 + (int)(it % S);
                MC_CLK(long long w1 = clock64();)
                mbar_wait(&sh.full[vslot], (it / S) & 1u);
                MC_CLK(cwv += clock64() - w1; w1 = clock64();)
                ++it;
                uint64_t *vbox = reinterpret_cast<uint64_t *>(boxes + vslot * MC_BOX);
                if (!fullt) mc_stage_rows(vbox, col.values, g.n, r0, p);
                if (p == 0) {
                    if (tid == 0) {  // the tile's first row in this column (tile / chunk edge records)
                        sh.edge[cp].raw = vbox[0 ^ swz];
                        sh.edge[cp].valid = (uint32_t)(vbits & 1ull);
                    }
                    if ((Pol::NEXT_VALUE || FUSED) && ti.next_has) {
                        nvalid = vnext;
                        nraw = fullt ? vbox[(tid + 1) * MC_P + (0 ^ mc_swz(tid + 1))] : col.values[my0 + MC_RE];
                    }
                }
                {
                    const int64_t *trow = RESIDENT ? reinterpret_cast<const int64_t *>(boxes + p * MC_BOX) + tid * MC_P : nullptr;
                    const int64_t grow0 = my0 + p * MC_P;
                    if (fullt)
                        mc_value_phase_fast<Pol, HAS_NULLS, FUSED>(c, ti.bm, ti.gm, A, col, p, trow, vbox + tid * MC_P, vbits, grow0);
                    else if (nm > 0)
                        mc_value_phase<Pol, HAS_NULLS, false, FUSED>(c, ti.bm, ti.gm, A, col, p, trow, vbox + tid * MC_P, vbits, grow0);
                }
                __syncwarp();
                MC_CLK(cvp += clock64() - w1;)
                release(vslot, !fullt);
                if (RESIDENT && ci == A.ncols - 1) release(p, !fullt);
            }
            if (ci == 0) {  // ---- end of the time pass: does my open window end at my right edge? --------------------------------
                if (ti.has_rows) {
                    if (ti.next_has) {
                        ti.closes_right = ti.nt >= scan.eabs;
                        ti.nk = ((uint64_t)ti.nt - (uint64_t)scan.eabs < d) ? scan.kcur + 1
                                                                              : mc_div_slow((uint64_t)ti.nt - (uint64_t)g.s0, d, g.div.inv_rd);
                    } else if (!fullt) {
                        ti.closes_right = my0 + nrows == g.n;  // I own the last row of the column
                    }
                }
                if (tid == 0) {
                    sh.tile_kf[tp] = (int64_t)ti.kf;
                    sh.tile_t0[tp] = x0;
                }
            }
            MC_CLK(const long long s0 = clock64();)
            mc_stitch<Pol, FUSED>(c, nraw, nvalid, ti, A, col, sh, ci, cp, tp, last_tile, xlast_tile);
            MC_CLK(cst += clock64() - s0;)
            cp ^= 1;
        }
    }
    MC_CLK(if (lane == 0 && warp == 1) { atomicAdd(&mc_dbg[2], (unsigned long long)cwt); atomicAdd(&mc_dbg[3], (unsigned long long)cwv); atomicAdd(&mc_dbg[4], (unsigned long long)ctp); atomicAdd(&mc_dbg[5], (unsigned long long)cvp); atomicAdd(&mc_dbg[6], (unsigned long long)cst); atomicAdd(&mc_dbg[7], (unsigned long long)(clock64() - ct0)); })
    if (bad) atomicOr(A.status, ST_UNSORTED);
}

// Joins the chunk records of one column, strictly left to right.  Thread j owns the windows whose first row lies in
// chunk j and that are not complete inside it: the one at the left edge of the chunk (head record, unless it continues
// a window of an earlier chunk) and the one open at its right edge (tail record).  On the fused path the remaining
// threads of the grid give every window WITHOUT rows its synthetic start row (interpolation_test.go:83-100).
template <class Pol, bool FUSED>
__global__ void mc_fixup_kernel(const __grid_constant__ McArgs A, const int nchunks) {
    using Inc = typename Pol::Inc;
    using State = typename Pol::State;
    const McColArgs &col = A.col[blockIdx.y];
    const WindowGeom &g = A.g;
    const uint64_t d = g.div.d;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (FUSED) {
        const FusedSyn &S = col.syn;
        for (int64_t m = gid; m < g.W; m += (int64_t)gridDim.x * blockDim.x) {
            if ((A.touched[m >> 5] >> (m & 31)) & 1u) continue;
            State sm = Pol::identity();
            if (m < S.len && S.missing[m]) Pol::inject(sm, (int64_t)((uint64_t)g.s0 + (uint64_t)m * d), S.val[m], S.ok[m] != 0);
            // its inclusive row = the first row of window m + 1 in the interpolated frame
            Inc inc = Pol::no_inc();
            const int64_t k1 = m + 1;
            if (k1 < S.len) {
                const int64_t s1 = (int64_t)((uint64_t)g.s0 + (uint64_t)k1 * d);
                if (S.missing[k1]) {
                    inc = Pol::make_inc(true, S.ok[k1] != 0, S.val[k1], s1);
                } else {  // window m + 1 holds a row exactly at its start
                    const int64_t r = S.first[k1];
                    if (r < g.n) {
                        const bool v = col.validity ? (col.validity[r >> 3] >> (r & 7)) & 1 : true;
                        inc = Pol::make_inc(A.time[r] == s1, v, col.values[r], s1);
                    }
                }
            }
            Pol::write_final(col.out, g, A.width, m, sm, inc);
        }
    }
    const int64_t j = gid;
    if (j >= nchunks) return;
    const McCarry *rec = col.rec;
    auto finish = [&](const McCarry &a) { Pol::write_final(col.out, g, A.width, a.key, Pol::state_of(a), Pol::inc_of(a)); };
    auto walk = [&](McCarry a, int64_t i) {
        const int64_t key = a.key;
        for (; i < nchunks; ++i) {
            const McCarry h = rec[2 * i];
            if (h.key != key) {  // the window ended exactly at the chunk boundary
                if (FUSED) {
                    Pol::set_inc(a, mc_fused_inc<Pol>(g, col.syn, (uint64_t)key, true, (uint64_t)h.key, h.edge_t, h.edge_raw,
                                                      h.edge_valid != 0));
                } else {
                    const int64_t E = (int64_t)((uint64_t)g.s0 + ((uint64_t)key + 1) * d);
                    Pol::set_inc(a, Pol::make_inc(h.edge_t == E, h.edge_valid != 0, h.edge_raw, h.edge_t));
                }
                finish(a);
                return;
            }
            Pol::set_state(a, Pol::combine(Pol::state_of(a), Pol::state_of(h)), false);
            Pol::set_inc(a, Pol::inc_of(h));
            if (h.cnt & MC_CLOSED) {
                finish(a);
                return;
            }
        }
        Pol::set_inc(a, Pol::no_inc());
        finish(a);  // the column ends inside the window
    };
    McCarry hd = rec[2 * j];
    const McCarry tl = rec[2 * j + 1];
    bool starts = true;
    if (j > 0) {
        const McCarry pt = rec[2 * (j - 1) + 1];
        int64_t open_key = pt.key;
        if (open_key < 0) {
            const McCarry ph = rec[2 * (j - 1)];
            if (!(ph.cnt & MC_CLOSED)) open_key = ph.key;
        }
        starts = open_key != hd.key;
        if (blockIdx.y == 0 && hd.edge_t < pt.edge_t) atomicOr(A.status, ST_UNSORTED);
    }
    if (starts) {
        if (FUSED) {  // the window begins at the chunk's first row: its synthetic start row comes first
            const int64_t k = hd.key;
            if (k < col.syn.len && col.syn.missing[k]) {
                State p = Pol::identity();
                Pol::inject(p, (int64_t)((uint64_t)g.s0 + (uint64_t)k * d), col.syn.val[k], col.syn.ok[k] != 0);
                Pol::set_state(hd, Pol::combine(p, Pol::state_of(hd)), (hd.cnt & MC_CLOSED) != 0);
            }
        }
        if (hd.cnt & MC_CLOSED)
            finish(hd);
        else
            walk(hd, j + 1);
    }
    if (tl.key >= 0) walk(tl, j + 1);
}

// ---- host side ------------------------------------------------------------------------------------------------------------
struct McKnobs {
    int ctas;
    int l2p;
    PFN_cuTensorMapEncodeTiled_v12000 encode;
};
inline const McKnobs &mc_knobs() {
    static McKnobs k;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *a = getenv("BOWGPU_MC_CTAS"), *b = getenv("BOWGPU_TMAP_L2");
        k.ctas = a ? atoi(a) : MC_CFG_CTAS;
        if (k.ctas < 1) k.ctas = 1;
        k.l2p = b ? atoi(b) : 1;  // L2 promotion of the box fetches (0 none, 1 64B, 2 128B, 3 256B)
        if (k.l2p < 0 || k.l2p > 3) k.l2p = 3;
        k.encode = nullptr;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            k.encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    });
    return k;
}

// Tensor map viewing a column of n 8-byte elements as [n / RE][RE]; box = [NT][P], 128-byte swizzle.  Only whole rows of
// the view are addressable, which is all the kernel asks for (tiles that are not complete are staged by plain loads).
inline int mc_make_tmap(CUtensorMap *m, const void *col, int64_t n) {
    memset(m, 0, sizeof *m);
    const int64_t outer = n / MC_RE;
    if (outer == 0) return 0;
    const McKnobs &k = mc_knobs();
    if (!k.encode) return (int)cudaErrorNotSupported;
    const cuuint64_t dims[2] = {(cuuint64_t)MC_RE, (cuuint64_t)outer};
    const cuuint64_t strides[1] = {(cuuint64_t)MC_RE * 8};
    const cuuint32_t box[2] = {(cuuint32_t)MC_P, (cuuint32_t)MC_NT};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = k.encode(m, CU_TENSOR_MAP_DATA_TYPE_INT64, 2, const_cast<void *>(col), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, (CUtensorMapL2promotion)k.l2p,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

inline int mc_chunks_for(int sm_count, int64_t ntiles, int *chunk_tiles) {
    int64_t grid = (int64_t)sm_count * mc_knobs().ctas;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    const int64_t ct = (ntiles + grid - 1) / grid;
    *chunk_tiles = (int)ct;
    return (int)((ntiles + ct - 1) / ct);
}

template <class Pol, bool HAS_NULLS, bool FUSED>
int mc_launch_impl(const McLaunch &L, int sm_count, cudaStream_t stream, cudaEvent_t e0, cudaEvent_t e1) {
    const int64_t ntiles = (L.g.n + MC_T - 1) / MC_T;
    if (ntiles == 0 || L.ncols <= 0) return 0;
    auto kern = segmc_kernel<Pol, HAS_NULLS, FUSED>;
    static std::atomic<bool> configured[64];  // function attributes are per device (one ctx per GPU may live in one process)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MC_SMEM);
        if (e != cudaSuccess) return (int)e;
        configured[dev & 63].store(true, std::memory_order_release);
    }
    McArgs A;
    memset(&A, 0, sizeof A);
    McMaps M;
    A.time = L.time;
    A.g = L.g;
    A.width = L.g.whole ? (double)(L.g.whole_last - L.g.whole_first) : (double)(int64_t)L.g.div.d;
    A.ntiles = ntiles;
    A.ncols = L.ncols;
    A.touched = L.touched;
    A.status = L.status;
    int chunk_tiles = 1;
    const int nchunks = mc_chunks_for(sm_count, ntiles, &chunk_tiles);
    A.chunk_tiles = chunk_tiles;
    int rc = mc_make_tmap(&M.time, L.time, L.g.n);
    for (int c = 0; c < L.ncols && !rc; ++c) {
        A.col[c] = L.col[c];
        rc = mc_make_tmap(&M.val[c], L.col[c].values, L.g.n);
    }
    for (int c = L.ncols; c < MC_MAXC; ++c) memset(&M.val[c], 0, sizeof(CUtensorMap));
    if (rc) return rc;
    if (e0) cudaEventRecord(e0, stream);
#ifdef MC_DEBUG_CLOCKS
    { unsigned long long z[8] = {}; cudaMemcpyToSymbol(mc_dbg, z, sizeof z); }
#endif
    kern<<<(unsigned)nchunks, MC_THREADS, MC_SMEM, stream>>>(A, M);
#ifdef MC_DEBUG_CLOCKS
    { unsigned long long z[8]; cudaStreamSynchronize(stream); cudaMemcpyFromSymbol(z, mc_dbg, sizeof z); const double n = nchunks;
      fprintf(stderr, "[mc clocks / CTA] producer: wait_empty %.0f of %.0f | consumer(warp1): wait_T %.0f wait_V %.0f time_pass %.0f value_pass %.0f stitch %.0f total %.0f\n", z[0]/n, z[1]/n, z[2]/n, z[3]/n, z[4]/n, z[5]/n, z[6]/n, z[7]/n); }
#endif
    if (e1) cudaEventRecord(e1, stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    const int fb = 128;
    int64_t fthreads = nchunks;
    if (FUSED) {  // the windows without rows are shared out over a few CTAs per SM
        const int64_t cap = (int64_t)sm_count * 8 * fb;
        fthreads = L.g.W < cap ? L.g.W : cap;
        if (fthreads < nchunks) fthreads = nchunks;
    }
    const dim3 fgrid((unsigned)((fthreads + fb - 1) / fb), (unsigned)L.ncols);
    mc_fixup_kernel<Pol, FUSED><<<fgrid, fb, 0, stream>>>(A, nchunks);
    return (int)cudaGetLastError();
}

template <class Pol, bool HAS_NULLS>
int mc_launch(const McLaunch &L, int sm_count, cudaStream_t stream, cudaEvent_t e0, cudaEvent_t e1) {
    return L.col[0].syn.missing ? mc_launch_impl<Pol, HAS_NULLS, true>(L, sm_count, stream, e0, e1)
                                : mc_launch_impl<Pol, HAS_NULLS, false>(L, sm_count, stream, e0, e1);
}

}  // namespace bowgpu
