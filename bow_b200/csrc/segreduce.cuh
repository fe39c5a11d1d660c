// segreduce.cuh — the generic streaming segmented reduction behind Rolling.Aggregate.
//
// One pass over the time column and one value column computes, for every window, a window STATE
// defined by a policy (a monoid: identity / accumulate one row / combine two adjacent runs).
// seg_basic.cu instantiates it for Count/Sum/Mean/Min/Max/First/Last, seg_integral.cu for
// IntegralStep/IntegralTrapezoid/WeightedAverageStep/WeightedAverageLinear.
//
// Work decomposition (load balance independent of window sizes, SURVEY 7 "hard parts"):
//   * fixed-size ROW tiles (T = NT*RE rows, contiguous in memory) are assigned round-robin to persistent
//     CTAs.  Thread t of the CTA owns the RE = NP*P CONSECUTIVE rows [t*RE, (t+1)*RE) of the tile and
//     streams through them in NP phases of P = 16 rows, keeping its reduction state in registers.
//   * staging: per phase ONE 2-D TMA box per column (cp.async.bulk.tensor.2d, SASS UTMALDG) of a tensor
//     map that views the column as [n/RE][RE] elements: box = [NT rows][P+2 columns] lands in shared
//     memory as [thread][18] (pitch 144 bytes: 16-byte reads of a quarter warp hit all 32 banks once).
//     The boxes of the NP phases of a tile interleave inside the same contiguous T*8-byte region of
//     global memory, so DRAM sees one sequential region per tile; mbarrier-tracked slots give
//     double buffering at phase granularity.
//   * inside its rows a thread reduces sequentially, left to right, exactly like the reference closure.
//     A row whose time reaches the running absolute window end closes the open window (one exact division
//     per thread and tile, none per row): the first such window of a thread is kept aside (head), later
//     ones began and ended inside the thread and are written out directly; the rows after the last
//     boundary are the thread's tail.
//   * per TILE (not per phase) the tails are stitched: the right-edge boundary of a thread is found by
//     comparing window indices with the next thread (shuffle), partial states of windows spanning several
//     threads by a warp-shuffle segmented scan (flag = "a window closed in this thread") plus a short
//     cross-warp pass;
//   * every tile leaves a head record (its first window) and a tail record (the window open at its end);
//     a tiny fix-up kernel joins records of equal window index left to right (deterministic, no atomics).
//
// Policy interface (all static, device):
//   State, Carry, Out, Inc                      types (Inc = the inclusive row of a window, rolling.go:201-209)
//   identity(), accumulate(State&, t, raw), note(State&, mask, trow, vrow, swz), combine(L, R), shfl_up(s, d)
//   make_inc(at_end, valid_next, raw_next, t_next)
//   write(out, g, k, state, inc)                final values of a window
//   make_carry(state, inc, key, closed) -> Carry; carry_* accessors; write_carry
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "kernels.h"

namespace bowgpu {

// tile shape and residency (compile-time; -DSEG_CFG_* only for tuning sweeps, scripts/tune_seg.sh)
#ifndef SEG_CFG_NT
#define SEG_CFG_NT 128
#endif
#ifndef SEG_CFG_NP
#define SEG_CFG_NP 4
#endif
#ifndef SEG_CFG_CTAS
#define SEG_CFG_CTAS 3
#endif
#ifndef SEG_CFG_UNROLL
#define SEG_CFG_UNROLL 2
#endif
#ifndef SEG_CFG_SWZ
#define SEG_CFG_SWZ 1
#endif
// SEG_SWZ: boxes are [NT][P] with TMA's 128-byte swizzle (16-byte chunk c of row r lands in chunk c ^ (r & 7)) instead of
// [NT][P + 2] with a padded pitch: the same conflict-free LDS.128, but no padding columns through the TMA unit (the
// staging pattern alone moves 7.1 TB/s swizzled against 6.6 TB/s padded, profiles/r1_box_stream_microbench.txt).
constexpr bool SEG_SWZ = SEG_CFG_SWZ != 0;
#ifndef SEG_CFG_FUSED_COLD
#define SEG_CFG_FUSED_COLD 2
#endif
#ifndef SEG_CFG_FUSED_EXPECT
#define SEG_CFG_FUSED_EXPECT 1
#endif
#ifndef SEG_CFG_FUSED_PREFETCH
#define SEG_CFG_FUSED_PREFETCH 1
#endif
#ifndef SEG_CFG_LINEAR
#define SEG_CFG_LINEAR 1
#endif
#ifndef SEG_CFG_WARPSYNC
#define SEG_CFG_WARPSYNC 2  // 0 never, 1 every instantiation, 2 the fused integral kernels only (see the kernel)
#endif
constexpr int SEG_UNROLL = SEG_CFG_UNROLL;  // row pairs per trip of the streaming loop
constexpr int SEG_NT = SEG_CFG_NT;
constexpr int SEG_P = 16;                     // rows per phase (box start columns must be 16-byte aligned in global memory)
constexpr int SEG_NP = SEG_CFG_NP;            // phases per tile
constexpr int SEG_RE = SEG_P * SEG_NP;        // consecutive rows owned by one thread in a tile
constexpr int SEG_T = SEG_NT * SEG_RE;        // rows per tile
constexpr int SEG_COLS = SEG_SWZ ? SEG_P : SEG_P + 2;  // box columns (unswizzled: the last two only pad the pitch to 144 bytes)
constexpr int SEG_BOX_BYTES = SEG_NT * SEG_COLS * 8;
constexpr int SEG_SLOT_BYTES = 2 * SEG_BOX_BYTES;  // time box + value box
constexpr int SEG_BITS_BYTES = SEG_T / 8;
constexpr int SEG_BITS_COPY = SEG_BITS_BYTES + 16;
constexpr int SEG_BITS_STRIDE = (SEG_BITS_COPY + 127) / 128 * 128;
constexpr int SEG_MAX_STAGES = 4;
constexpr int SEG_NW = SEG_NT / 32;
constexpr int SEG_SLOT_ALIGN = SEG_SWZ ? 1024 : 128;  // the swizzle pattern is a function of the shared-memory address
constexpr int SEG_HEADER_MIN = 640;
constexpr int SEG_HEADER_BYTES = (SEG_HEADER_MIN + 2 * SEG_BITS_STRIDE + SEG_SLOT_ALIGN - 1) / SEG_SLOT_ALIGN * SEG_SLOT_ALIGN - 2 * SEG_BITS_STRIDE;
static_assert((SEG_P * 8) % 16 == 0 && (SEG_COLS * 8) % 16 == 0, "box geometry: 16-byte aligned starts");
static_assert(SEG_SWZ ? SEG_COLS * 8 == 128 : (SEG_COLS * 2) % 8 == 4, "pitch: one swizzle row, or an odd multiple of 16 bytes");
// element j of a thread's row sits at index j ^ seg_swz(tid)
__device__ __forceinline__ int seg_swz(int tid) { return SEG_SWZ ? (tid & 7) << 1 : 0; }
static_assert(SEG_T % 128 == 0, "tile validity bytes must be a multiple of 16");
static_assert(SEG_BOX_BYTES % SEG_SLOT_ALIGN == 0, "boxes keep every slot aligned");

template <class Pol>
struct SegArgs {
    const int64_t *time;
    const uint64_t *values;
    const uint8_t *validity;  // null = all valid
    WindowGeom g;
    typename Pol::Out out;
    typename Pol::Carry *carry_head;  // [ntiles]
    typename Pol::Carry *carry_tail;  // [ntiles]
    typename Pol::Carry *skip;        // [seg_skip_records(ntiles)] or null: combined records of runs of 32 / 1024 / 32768 tiles
    int32_t *status;
    FusedSyn syn;                     // fused Interpolate -> Aggregate (FUSED instantiations only)
    int32_t *gate;                    // start gate of side-by-side lanes (null = none), see SegLaunch
    int32_t gate_lanes;
};

template <class Pol>
struct WarpTotal {
    typename Pol::State st;
    uint32_t flag;
    uint32_t _pad;
};

// first row of a warp's first thread, published for the last lane of the previous warp (and, for warp 0, the
// first row of the tile for the tile's head record)
struct WarpEdge {
    int64_t kf;      // window index of the row
    int64_t t;
    uint64_t raw;
    uint32_t flags;  // bit 0: the thread has rows, bit 1: the row's value is valid
    uint32_t _pad;
};
constexpr uint32_t EDGE_HAS = 1u, EDGE_VALID = 2u;

// exact division on the rare paths
__device__ __noinline__ static uint64_t div_slow(uint64_t x, uint64_t d, double inv_rd) {
    DivU64 dv{d, inv_rd};
    return div_u64(x, dv);
}

// ---- fused Interpolate -> Aggregate -------------------------------------------------------------------------------
// In the interpolated frame (rolling/interpolation.go:118-161) window k holds [a synthetic row at S_k iff missing[k]]
// ++ its own rows, and every EMPTY window holds exactly its synthetic row.  The inclusive row of window k is the first
// row of window k+1 in that frame: the synthetic row at S_{k+1} when there is one, else the real first row if it sits
// exactly on E_k.  fused_boundary() is called where window `kcur` ends and the next REAL row (x, raw, valid) - if any -
// lies in window `knew`: it returns the inclusive row of kcur, writes the empty windows in between, and leaves in
// `fresh` the state window knew starts from (its synthetic row, or the identity).
template <class Pol>
__device__ __forceinline__ typename Pol::Inc fused_boundary(const SegArgs<Pol> &A, const uint64_t kcur, const bool has_row,
                                                            const uint64_t knew, const int64_t x, const uint64_t raw,
                                                            const bool valid, typename Pol::State &fresh) {
    using Inc = typename Pol::Inc;
    const WindowGeom &g = A.g;
    const FusedSyn &S = A.syn;
    fresh = Pol::identity();
    // the column ends (no window after kcur), or kcur is a halo window of a shard (not owned: nothing to produce)
    if (!has_row || kcur >= (uint64_t)g.W) return Pol::make_inc(false, false, 0, 0);
    const uint64_t d = g.div.d;
    const uint64_t len = (uint64_t)S.len;
    auto start_of = [&](uint64_t k) { return (int64_t)((uint64_t)g.s0 + k * d); };
    auto real_inc = [&](uint64_t k) {  // the real row as inclusive row of window k (only if it sits exactly on E_k)
        return Pol::make_inc(x == start_of(k + 1), valid, raw, x);
    };
    auto syn_inc = [&](uint64_t k) {
        return k < len ? Pol::make_inc(true, S.ok[k] != 0, S.val[k], start_of(k)) : Pol::make_inc(false, false, 0, 0);
    };
    const bool miss_new = knew < len && S.missing[knew] != 0;
    const uint64_t kn = kcur + 1;
    const Inc inc = (knew == kn && !miss_new) ? real_inc(kcur) : syn_inc(kn);
    const uint64_t kend = knew < (uint64_t)g.W ? knew : (uint64_t)g.W;  // only owned windows are written
    for (uint64_t m = kn; m < kend; ++m) {  // empty windows: one synthetic row each (interpolation.go:83-100 golden)
        typename Pol::State sm = Pol::identity();
        Pol::inject(sm, start_of(m), S.val[m], S.ok[m] != 0);
        const Inc im = (m + 1 < knew || miss_new) ? syn_inc(m + 1) : real_inc(m);
        Pol::write(A.out, g, (int64_t)m, sm, im);
    }
    if (miss_new) Pol::inject(fresh, start_of(knew), S.val[knew], S.ok[knew] != 0);
    return inc;
}

// Out-of-line copy for the streaming loop: a window boundary is a rare event there (once per window), and the inlined
// body in each of the loop's four row slots pushed the fused kernel out of the instruction cache (configs[2]: 1.025 ms
// per column pass at 2.5e8 rows against 0.787 ms for the unfused kernel).
template <class Pol>
__device__ __noinline__ typename Pol::Inc fused_boundary_cold(const SegArgs<Pol> &A, const uint64_t kcur, const uint64_t knew,
                                                              const int64_t x, const uint64_t raw, const bool valid,
                                                              typename Pol::State &fresh) {
    return fused_boundary<Pol>(A, kcur, true, knew, x, raw, valid, fresh);
}

// SEG_CFG_FUSED_COLD == 2: the whole close of a window in the streaming loop out of line — inclusive row, empty windows,
// and the head capture / window write.  The head window and its inclusive row are handed over by address, which parks
// them in local memory while the rows stream: they are written once per thread and tile and read once before the stitch,
// but as registers they cost the fused streaming loop 14 register moves PER ROW (ptxas ping-pongs the merged copies
// around the branch; ncu: 27 % of the kernel's instructions on the line of the boundary test).
template <class Pol>
__device__ __noinline__ typename Pol::State fused_close_cold(const SegArgs<Pol> &A, const uint64_t kold, const uint64_t knew,
                                                             const int64_t x, const uint64_t raw, const bool valid,
                                                             const typename Pol::State st, typename Pol::State *head,
                                                             typename Pol::Inc *inc_head, const int nclose) {
    typename Pol::State fresh;
    const typename Pol::Inc inc = fused_boundary<Pol>(A, kold, true, knew, x, raw, valid, fresh);
    if (nclose == 0) {
        *head = st;
        *inc_head = inc;
    } else {
        Pol::write(A.out, A.g, (int64_t)kold, st, inc);
    }
    return fresh;
}

// In the fused kernel the state a new window starts from comes out of the boundary path (its synthetic row); without a
// hint ptxas keeps the merge cheap on THAT side and pays ~17 register moves per row on the side where nothing happens.
template <bool FUSED>
__device__ __forceinline__ bool seg_rare(bool b) {
#if SEG_CFG_FUSED_EXPECT
    if (FUSED) return __builtin_expect(b, 0);
#endif
    return b;
}

// per-thread state that lives in registers across the phases of one tile
template <class Pol>
struct SegThread {
    typename Pol::State st;    // open window (tail)
    typename Pol::State head;  // first window that closed in this thread (valid when nclose > 0)
    typename Pol::Inc inc_head;
    int64_t eabs;              // absolute end of the open window
    uint64_t kcur;             // its index
    uint64_t kf;               // window of my first row
    int64_t xlast;             // my last row's time
    int64_t first_t;
    uint64_t first_raw;
    uint32_t first_flags;
    int nclose;
    // fused kernel: the first window boundary of the thread in this tile, met by the straight-line phase and resolved later
    // (seg_resolve_deferred): the window that closed, the row that closed it
    uint64_t dk;
    int64_t dx;
    uint64_t draw;
    uint32_t dvalid;
    int deferred;
};

// Fused Interpolate -> Aggregate: what a window boundary needs — the inclusive row of the window that closed and the
// synthetic start row of the one that began (global loads, ~300 dependent instructions executed by the one or two lanes of
// a warp that met a boundary) — used to run inside the phase, where the other three warps of the CTA wait for it at the
// phase barrier (ncu: 26 % of the kernel's stall samples on that barrier, 1.86 barrier stalls per issued instruction
// against 1.0 in the unfused kernel).  The straight-line phase now only PARKS the first boundary of a thread (state of the
// closed window in local memory, the boundary row in registers) and goes on with the next window from the identity; the
// rest happens once per tile, before the stitch, in all threads at the same time.  States are monoids, so joining the
// synthetic start row afterwards (combine(fresh, rows)) gives the window the same value up to float64 association.
template <class Pol>
__device__ __forceinline__ void seg_resolve_deferred(SegThread<Pol> &c, const SegArgs<Pol> &A, typename Pol::Inc *finc) {
    if (!c.deferred) return;
    typename Pol::State fresh;
    *finc = fused_boundary<Pol>(A, c.dk, true, c.dk + 1, c.dx, c.draw, c.dvalid != 0, fresh);
    c.st = Pol::combine(fresh, c.st);
    c.deferred = 0;
}

// ---- straight-line phase (policies with Pol::LINEAR_PHASE) ---------------------------------------------------------------
// The generic row loop below tests every row for a window boundary and for validity with branches; between two branches
// ptxas has one row to schedule, so the float64 chain of a row (convert, subtract, multiply, add: ~50 cycles) is exposed
// once per row and the integral kernels ran latency bound at a third of the issue rate.  When the 16 rows of a thread's
// phase hold AT MOST ONE window boundary and no empty window follows it (decided from the phase's last row: rows are
// sorted), the phase is executed as ONE basic block instead: every row forms its joint term with the previous valid row
// unconditionally and adds it, under predicates, to the sums of the open window (rows before the boundary) or of the
// window that begins at the boundary row; the boundary itself is dealt with once, after the rows.  The rows of different
// positions are independent up to the predicated selects, so their float64 chains overlap.  Anything else — a second
// boundary, windows without rows, tiles at the edges of the column — takes the generic loop.
template <class Pol, bool HAS_NULLS, bool FUSED>
__device__ __forceinline__ bool seg_phase_linear(SegThread<Pol> &c, const SegArgs<Pol> &A, const int64_t *trow,
                                                 const uint64_t *vrow, const int swz, const uint32_t vbits, bool &bad,
                                                 typename Pol::State *fhead, typename Pol::Inc *finc) {
    using State = typename Pol::State;
    using Inc = typename Pol::Inc;
    constexpr int P = SEG_P;
    const WindowGeom &g = A.g;
    const uint64_t d = g.div.d;
    const longlong2 *t2 = reinterpret_cast<const longlong2 *>(trow);
    const ulonglong2 *v2 = reinterpret_cast<const ulonglong2 *>(vrow);
    const int sh = swz >> 1;
    const int64_t xend = trow[(P - 1) ^ swz];
    int b = P;  // rows [0, b) belong to the open window, row b (if any) starts the next one
    if (xend >= c.eabs) {
        if ((uint64_t)xend - (uint64_t)c.eabs >= d) return false;  // a second boundary or windows without rows
        b = 0;
#pragma unroll
        for (int q = 0; q < P / 2; ++q) {
            const longlong2 tq = t2[q ^ sh];
            b += (tq.x < c.eabs) + (tq.y < c.eabs);
        }
    }
    if (FUSED && b < P) seg_resolve_deferred<Pol>(c, A, finc);  // (a second boundary in my rows of this tile: rare)
    typename Pol::Lin L = Pol::lin_begin(c.st);
    int64_t xprev = c.xlast;
    bool unsorted = false;
#pragma unroll
    for (int q = 0; q < P / 2; ++q) {
        const longlong2 tq = t2[q ^ sh];
        const ulonglong2 vq = v2[q ^ sh];
        unsorted |= tq.x < xprev;
        unsorted |= tq.y < tq.x;
        xprev = tq.y;
        Pol::lin_row(L, 2 * q, b, tq.x, vq.x, !HAS_NULLS || ((vbits >> (2 * q)) & 1u));
        Pol::lin_row(L, 2 * q + 1, b, tq.y, vq.y, !HAS_NULLS || ((vbits >> (2 * q + 1)) & 1u));
    }
    bad |= unsorted;
    c.xlast = xend;
    if (b == P) {  // no boundary: the open window takes all my rows
        Pol::lin_end_open(c.st, L);
        Pol::note(c.st, vbits, trow, vrow, swz);
        return true;
    }
    // ---- one boundary, at row b ------------------------------------------------------------------------------------------
    const uint32_t below = (1u << b) - 1u;
    State sa = c.st, sb = Pol::identity();
    Pol::lin_end_split(sa, sb, L, vbits & below, vbits & ~below, trow, vrow, swz);
    Pol::note(sa, vbits & below, trow, vrow, swz);
    Pol::note(sb, vbits & ~below, trow, vrow, swz);
    const int64_t xb = trow[b ^ swz];
    const uint64_t rb = vrow[b ^ swz];
    const bool vb = (vbits >> b) & 1u;
    const uint64_t kold = c.kcur;
    ++c.kcur;
    if (!FUSED) {
        const Inc inc = Pol::make_inc(xb == c.eabs, vb, rb, xb);
        if (c.nclose == 0) {
            c.head = sa;
            c.inc_head = inc;
        } else {
            Pol::write(A.out, g, (int64_t)kold, sa, inc);
        }
        c.st = sb;
    } else {
        // (the window that begins at row b starts from its synthetic row, if it has one: joined like two adjacent runs)
        if (c.nclose == 0) {  // park it (seg_resolve_deferred)
            *fhead = sa;
            c.dk = kold;
            c.dx = xb;
            c.draw = rb;
            c.dvalid = vb;
            c.deferred = 1;
            c.st = sb;
        } else {
            const State fresh = fused_close_cold<Pol>(A, kold, c.kcur, xb, rb, vb, sa, fhead, finc, c.nclose);
            c.st = Pol::combine(fresh, sb);
        }
    }
    c.eabs = (int64_t)((uint64_t)c.eabs + d);
    ++c.nclose;
    return true;
}

// One phase: P rows of every thread.  FULL: all rows of the tile exist and none lies before s0 (the common case,
// free of per-row existence predicates).
template <class Pol, bool HAS_NULLS, bool FULL, bool FUSED>
__device__ __forceinline__ void seg_phase(SegThread<Pol> &c, const SegArgs<Pol> &A, const int64_t r0, const int phase,
                                          const uint8_t *slot, const uint32_t *bsm, bool &bad,
                                          typename Pol::State *fhead, typename Pol::Inc *finc) {
    using Inc = typename Pol::Inc;
    constexpr int P = SEG_P;
    const int tid = threadIdx.x;
    const WindowGeom &g = A.g;
    const uint64_t d = g.div.d;
    const int64_t *trow = reinterpret_cast<const int64_t *>(slot) + tid * SEG_COLS;
    const uint64_t *vrow = reinterpret_cast<const uint64_t *>(slot + SEG_BOX_BYTES) + tid * SEG_COLS;
    const int swz = seg_swz(tid);
    const int ti0 = tid * SEG_RE + phase * P;  // tile-relative index of my first row of this phase
    int nmine = P;
    if (!FULL) {
        const int64_t left = g.n - r0 - ti0;
        nmine = left < 0 ? 0 : (left > P ? P : (int)left);
    }
    if (!FULL && nmine == 0) return;

    uint32_t vbits = (1u << P) - 1u;
    if (HAS_NULLS) {
        const uint32_t lo = bsm[ti0 >> 5], hi = bsm[(ti0 >> 5) + 1];
        vbits &= __funnelshift_r(lo, hi, ti0 & 31);
    }
    bool early_any = false;
    if (!FULL) {
        vbits &= (1u << nmine) - 1u;
        const int64_t e = g.early_rows - (r0 + ti0);  // rows before s0 (negative timestamps / left halo of a shard)
        if (e > 0) {
            early_any = true;
            if (!g.early_keep) vbits &= e >= P ? 0u : ~((1u << (int)e) - 1u);
        }
    }

    const longlong2 *t2 = reinterpret_cast<const longlong2 *>(trow);
    const ulonglong2 *v2 = reinterpret_cast<const ulonglong2 *>(vrow);
    longlong2 ta = t2[0 ^ (swz >> 1)];
    ulonglong2 va = v2[0 ^ (swz >> 1)];

    if (phase == 0) {  // window of my first row and its absolute end (rows before s0 collapse onto window 0)
        const bool early0 = !FULL && early_any;
        c.kf = early0 ? 0 : div_u64((uint64_t)ta.x - (uint64_t)g.s0, g.div);
        c.kcur = c.kf;
        c.eabs = (int64_t)((uint64_t)g.s0 + (c.kf + 1) * d);
        c.first_t = ta.x;
        c.first_raw = va.x;
        c.first_flags = EDGE_HAS | ((vbits & 1u) ? EDGE_VALID : 0u);
        c.xlast = ta.x;
#if SEG_CFG_FUSED_PREFETCH
        if (FUSED) {  // the synthetic row of the next window: what the boundary path will ask for, pulled towards the SM now
            const uint64_t kn = c.kf + 1;
            if (kn < (uint64_t)A.syn.len) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(A.syn.missing + kn));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(A.syn.val + kn));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(A.syn.ok + kn));
            }
        }
#endif
    }

    if constexpr (Pol::LINEAR_PHASE && SEG_CFG_LINEAR) {
        if (FULL && !(FUSED && SEG_CFG_FUSED_COLD != 2))
            if (seg_phase_linear<Pol, HAS_NULLS, FUSED>(c, A, trow, vrow, swz, vbits, bad, fhead, finc)) return;
    }

    if (FUSED) seg_resolve_deferred<Pol>(c, A, finc);  // (the generic loop closes windows on the spot)

    int segstart = 0;  // first row of this phase that belongs to the open window
    // One row: order check against the previous row, window boundary test against the running absolute end,
    // accumulate.  The loop over row pairs is only partially unrolled (SEG_CFG_UNROLL pairs per trip) so that the
    // streaming path stays resident in the instruction cache; the next pair is fetched one trip ahead.
    auto row = [&](const int j, const int64_t xj, const uint64_t rj) {
        if (FULL || j < nmine) {
            bad |= xj < c.xlast;
            c.xlast = xj;
            if (seg_rare<FUSED>(xj >= c.eabs)) {  // row j starts a later window: the open one is complete
                Pol::note(c.st, vbits & ((1u << j) - 1u) & ~((1u << segstart) - 1u), trow, vrow, swz);
                if (!FUSED) {
                    const Inc inc = Pol::make_inc(xj == c.eabs, (vbits >> j) & 1u, rj, xj);
                    if (c.nclose == 0) {
                        c.head = c.st;
                        c.inc_head = inc;
                    } else {
                        Pol::write(A.out, g, (int64_t)c.kcur, c.st, inc);
                    }
                    ++c.nclose;
                    c.st = Pol::identity();
                    segstart = j;
                    if ((uint64_t)xj - (uint64_t)c.eabs < d) {
                        ++c.kcur;
                        c.eabs = (int64_t)((uint64_t)c.eabs + d);
                    } else {
                        c.kcur = div_slow((uint64_t)xj - (uint64_t)g.s0, d, g.div.inv_rd);
                        c.eabs = (int64_t)((uint64_t)g.s0 + (c.kcur + 1) * d);
                    }
                } else {
                    const uint64_t kold = c.kcur;
                    if ((uint64_t)xj - (uint64_t)c.eabs < d) {
                        ++c.kcur;
                        c.eabs = (int64_t)((uint64_t)c.eabs + d);
                    } else {
                        c.kcur = div_slow((uint64_t)xj - (uint64_t)g.s0, d, g.div.inv_rd);
                        c.eabs = (int64_t)((uint64_t)g.s0 + (c.kcur + 1) * d);
                    }
#if SEG_CFG_FUSED_COLD == 2
                    c.st = fused_close_cold<Pol>(A, kold, c.kcur, xj, rj, (vbits >> j) & 1u, c.st, fhead, finc, c.nclose);
                    ++c.nclose;
                    segstart = j;
#else
                    typename Pol::State fresh;
#if SEG_CFG_FUSED_COLD
                    const Inc inc = fused_boundary_cold<Pol>(A, kold, c.kcur, xj, rj, (vbits >> j) & 1u, fresh);
#else
                    const Inc inc = fused_boundary<Pol>(A, kold, true, c.kcur, xj, rj, (vbits >> j) & 1u, fresh);
#endif
                    if (c.nclose == 0) {
                        c.head = c.st;
                        c.inc_head = inc;
                    } else {
                        Pol::write(A.out, g, (int64_t)kold, c.st, inc);
                    }
                    ++c.nclose;
                    c.st = fresh;
                    segstart = j;
#endif
                }
            }
            if ((vbits >> j) & 1u) Pol::accumulate(c.st, xj, rj);
        }
    };
#pragma unroll(SEG_UNROLL)
    for (int q = 0; q < P / 2; ++q) {
        const longlong2 tb = t2[(q + 1 < P / 2 ? q + 1 : q) ^ (swz >> 1)];  // (stay inside the row)
        const ulonglong2 vb = v2[(q + 1 < P / 2 ? q + 1 : q) ^ (swz >> 1)];
        row(2 * q, ta.x, va.x);
        row(2 * q + 1, ta.y, va.y);
        ta = tb;
        va = vb;
    }
    Pol::note(c.st, vbits & ~((1u << segstart) - 1u), trow, vrow, swz);
}

// End of a tile: stitch the per-thread pieces.
template <class Pol, bool FULL, bool FUSED>
__device__ __forceinline__ void seg_stitch(SegThread<Pol> &c, const SegArgs<Pol> &A, const int64_t tile,
                                           WarpTotal<Pol> *wtot, const WarpEdge *wedge, bool &bad) {
    using State = typename Pol::State;
    using Inc = typename Pol::Inc;
    using Carry = typename Pol::Carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const WindowGeom &g = A.g;
    const int64_t r0 = tile * SEG_T;
    const bool has_rows = FULL || (c.first_flags & EDGE_HAS);

    // ---- the boundary at my right edge: compare with the first row of the next thread ----------------------
    uint64_t nk = __shfl_down_sync(0xffffffffu, (unsigned long long)c.kf, 1);
    int64_t nt = __shfl_down_sync(0xffffffffu, (long long)c.first_t, 1);
    uint64_t nraw = (Pol::NEXT_VALUE || FUSED) ? __shfl_down_sync(0xffffffffu, (unsigned long long)c.first_raw, 1) : 0;
    uint32_t nflags = __shfl_down_sync(0xffffffffu, c.first_flags, 1);
    if (lane == 31) {
        if (warp + 1 < SEG_NW) {
            const WarpEdge e = wedge[warp + 1];
            nk = (uint64_t)e.kf;
            nt = e.t;
            nraw = e.raw;
            nflags = e.flags;
        } else {
            nflags = 0;  // last thread of the tile: nobody to the right
        }
    }
    const bool next_has = (nflags & EDGE_HAS) != 0;
    bool closes_right = false;
    if (has_rows) {
        if (next_has) {
            closes_right = nk != c.kcur;
            bad |= nt < c.xlast;
        } else if (!FULL) {
            int64_t mine = g.n - r0 - (int64_t)tid * SEG_RE;
            if (mine > SEG_RE) mine = SEG_RE;
            closes_right = r0 + (int64_t)tid * SEG_RE + mine == g.n;  // I own the last row of the column
        }
    }
    bool tail_open = has_rows;  // (meaningful for the last thread of the tile only)
    if (closes_right) {
        State fresh = Pol::identity();
        Inc inc;
        if (FUSED)  // (also writes the empty windows between mine and the next thread's, and seeds the next window)
            inc = fused_boundary<Pol>(A, c.kcur, next_has, nk, nt, nraw, (nflags & EDGE_VALID) != 0, fresh);
        else
            inc = next_has ? Pol::make_inc(nt == c.eabs, (nflags & EDGE_VALID) != 0, nraw, nt)
                           : Pol::make_inc(false, false, 0, 0);
        if (c.nclose == 0) {
            c.head = c.st;
            c.inc_head = inc;
        } else {
            Pol::write(A.out, g, (int64_t)c.kcur, c.st, inc);
        }
        ++c.nclose;
        c.st = fresh;  // the tail now belongs to the window the next thread starts in
        tail_open = false;
    }

    // ---- stitch windows spanning threads: segmented inclusive scan of the tails ------------------------------
    const uint32_t ball = __ballot_sync(0xffffffffu, c.nclose != 0);
    State sc = c.st;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
        const State o = Pol::shfl_up(sc, dd);
        // no flag in lanes (lane-dd, lane]  <=>  the run ending at lane-dd belongs to my open segment
        const uint32_t span = lane >= dd ? (0xFFFFFFFFu >> (31 - lane)) & ~((2u << (lane - dd)) - 1u) : 0u;
        if (lane >= dd && (ball & span) == 0) sc = Pol::combine(o, sc);
    }
    if (lane == 31) {
        wtot[warp].st = sc;
        wtot[warp].flag = ball != 0;
    }
    State ex = Pol::shfl_up(sc, 1);
    if (lane == 0) ex = Pol::identity();
    __syncthreads();
    State acc = Pol::identity();
    bool any_prev = false;
    for (int u = 0; u < warp; ++u) {
        const State ws = wtot[u].st;
        if (wtot[u].flag) {
            acc = ws;
            any_prev = true;
        } else {
            acc = Pol::combine(acc, ws);
        }
    }
    const bool flag_before = (ball & ((1u << lane) - 1u)) != 0;
    const State excl = flag_before ? ex : Pol::combine(acc, ex);
    const bool any_excl = any_prev || flag_before;

    const WarpEdge tile_first = wedge[0];
    if (c.nclose) {  // this thread closes the window that was open at its left edge
        const State h = Pol::combine(excl, c.head);
        if (!any_excl) {  // ... which reaches the left edge of the tile: the fix-up decides whether it began earlier
            Carry r = Pol::make_carry(h, c.inc_head, (int64_t)c.kf, true);
            Pol::carry_set_edge(r, tile_first.t, tile_first.raw, (tile_first.flags & EDGE_VALID) != 0);
            A.carry_head[tile] = r;
        } else {
            Pol::write(A.out, g, (int64_t)c.kf, h, c.inc_head);
        }
    }
    if (tid == SEG_NT - 1) {  // tile-level records
        const bool flag_incl = flag_before || c.nclose != 0;
        const State incl = flag_incl ? sc : Pol::combine(acc, sc);
        const bool any_incl = any_prev || flag_incl;
        const Inc noinc = Pol::make_inc(false, false, 0, 0);
        Carry tl = Pol::make_carry(incl, noinc, -1, false);
        Pol::carry_set_edge(tl, c.xlast, 0, false);  // last row of the tile (cross-tile order check)
        if (!any_incl) {  // no boundary anywhere in the tile: it lies inside one window
            Carry hd = Pol::make_carry(incl, noinc, tile_first.kf, false);
            Pol::carry_set_edge(hd, tile_first.t, tile_first.raw, (tile_first.flags & EDGE_VALID) != 0);
            A.carry_head[tile] = hd;
        } else if (tail_open) {
            Pol::carry_set_key(tl, (int64_t)c.kcur);
        }
        A.carry_tail[tile] = tl;
    }
}

template <class Pol, bool HAS_NULLS, int MIN_CTAS, bool FUSED>
__global__ void __launch_bounds__(SEG_NT, MIN_CTAS)
    segreduce_kernel(const __grid_constant__ SegArgs<Pol> A, const __grid_constant__ CUtensorMap tm_time,
                     const __grid_constant__ CUtensorMap tm_val, const int64_t ntiles, const int nstages) {
    extern __shared__ __align__(SEG_SLOT_ALIGN) uint8_t smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);                    // [SEG_MAX_STAGES]
    int *done = reinterpret_cast<int *>(smem_raw + 32);                         // [SEG_MAX_STAGES] warps through with a slot
    WarpEdge *wedge = reinterpret_cast<WarpEdge *>(smem_raw + 64);              // [SEG_NW]
    WarpTotal<Pol> *wtot = reinterpret_cast<WarpTotal<Pol> *>(smem_raw + 64 + SEG_NW * sizeof(WarpEdge));
    static_assert(64 + SEG_NW * (sizeof(WarpEdge) + sizeof(WarpTotal<Pol>)) <= SEG_HEADER_MIN, "header layout");
    uint8_t *bits = smem_raw + SEG_HEADER_BYTES;                                // [2][SEG_BITS_STRIDE]
    uint8_t *slots = bits + 2 * SEG_BITS_STRIDE;
    static_assert((SEG_HEADER_BYTES + 2 * SEG_BITS_STRIDE) % SEG_SLOT_ALIGN == 0, "slots must be aligned");

    const int tid = threadIdx.x;
    const WindowGeom &g = A.g;
    constexpr bool WARP_SYNC = SEG_CFG_WARPSYNC == 1 || (SEG_CFG_WARPSYNC == 2 && FUSED && Pol::LINEAR_PHASE);
    const int64_t bitmap_total = ((g.n + 7) / 8 + 15) & ~(int64_t)15;  // device bitmaps are padded to 16 bytes

    auto tile_full = [&](int64_t tile) {  // staged by TMA: whole tile + one more row exist, no row before s0
        const int64_t r0 = tile * SEG_T;
        return g.n - r0 > SEG_T && r0 >= g.early_rows;
    };
    // item q of this CTA: tile = blockIdx.x + (q / NP) * gridDim.x, phase = q % NP
    auto issue_item = [&](int64_t q) {  // one thread
        const int64_t tile = blockIdx.x + (q / SEG_NP) * (int64_t)gridDim.x;
        if (tile >= ntiles || !tile_full(tile)) return;
        const int phase = (int)(q % SEG_NP);
        const int s = (int)(q % nstages);
        uint8_t *slot = slots + (size_t)s * SEG_SLOT_BYTES;
        uint32_t bbytes = 0;
        if (HAS_NULLS && phase == 0) {
            const int64_t b0 = tile * SEG_BITS_BYTES;
            int64_t b1 = b0 + SEG_BITS_COPY;
            if (b1 > bitmap_total) b1 = bitmap_total;
            bbytes = (uint32_t)(b1 - b0);
        }
        mbar_arrive_expect_tx(&full[s], 2 * SEG_BOX_BYTES + bbytes);
        tma_box_2d(slot, &tm_time, phase * SEG_P, (int32_t)(tile * SEG_NT), &full[s]);
        tma_box_2d(slot + SEG_BOX_BYTES, &tm_val, phase * SEG_P, (int32_t)(tile * SEG_NT), &full[s]);
        if (bbytes)
            bulk_g2s(bits + ((q / SEG_NP) & 1) * SEG_BITS_STRIDE, A.validity + tile * SEG_BITS_BYTES, bbytes, &full[s]);
    };

    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) mbar_init(&full[s], 1);
        for (int s = 0; s < SEG_MAX_STAGES; ++s) done[s] = 0;
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();
    if (A.gate) {  // side-by-side lanes start together (bounded wait: a lane that never shows up must not hang the others)
        if (tid == 0) {
            if (blockIdx.x == 0) atomicAdd(A.gate, 1);
            const long long t0 = clock64();
            while (*reinterpret_cast<volatile int32_t *>(A.gate) < A.gate_lanes && clock64() - t0 < 400000) __nanosleep(100);
        }
        __syncthreads();
    }
    if (tid == 0)
        for (int q = 0; q < nstages; ++q) issue_item(q);

    uint32_t parity = 0;  // bit s: phase parity of slot s
    bool bad = false;
    int64_t q = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r0 = tile * SEG_T;
        const bool fullt = tile_full(tile);
        const uint32_t *bsm = reinterpret_cast<const uint32_t *>(bits + ((q / SEG_NP) & 1) * SEG_BITS_STRIDE);
        SegThread<Pol> c;
        c.st = Pol::identity();
        c.head = Pol::identity();
        c.inc_head = Pol::make_inc(false, false, 0, 0);
        c.nclose = 0;
        c.deferred = 0;
        c.dk = 0;
        c.dx = 0;
        c.draw = 0;
        c.dvalid = 0;
        c.first_flags = 0;
        c.first_t = 0;
        c.first_raw = 0;
        c.kf = c.kcur = 0;
        c.eabs = 0;
        c.xlast = 0;
        // fused kernel: the head window of the thread is parked in local memory while the rows stream (fused_close_cold)
        typename Pol::State fhead = Pol::identity();
        typename Pol::Inc finc = Pol::make_inc(false, false, 0, 0);
#pragma unroll 1
        for (int phase = 0; phase < SEG_NP; ++phase, ++q) {
            const int s = (int)(q % nstages);
            uint8_t *slot = slots + (size_t)s * SEG_SLOT_BYTES;
            if (fullt) {
                mbar_wait(&full[s], (parity >> s) & 1u);
                parity ^= 1u << s;
                seg_phase<Pol, HAS_NULLS, true, FUSED>(c, A, r0, phase, slot, bsm, bad, &fhead, &finc);
            } else {
                // tile at an edge of the column (or holding rows before s0): staged by plain loads with bounds checks
                int64_t *ts = reinterpret_cast<int64_t *>(slot);
                uint64_t *vs = reinterpret_cast<uint64_t *>(slot + SEG_BOX_BYTES);
                for (int e = tid; e < SEG_NT * SEG_P; e += SEG_NT) {
                    const int tt = e / SEG_P, cc = e - tt * SEG_P;
                    const int64_t row = r0 + (int64_t)tt * SEG_RE + phase * SEG_P + cc;
                    const bool in = row < g.n;
                    ts[tt * SEG_COLS + (cc ^ seg_swz(tt))] = in ? A.time[row] : 0;
                    vs[tt * SEG_COLS + (cc ^ seg_swz(tt))] = in ? A.values[row] : 0;
                }
                if (HAS_NULLS && phase == 0) {
                    uint32_t *bw = const_cast<uint32_t *>(bsm);
                    const uint32_t *src = reinterpret_cast<const uint32_t *>(A.validity) + tile * (SEG_BITS_BYTES / 4);
                    const int64_t words_left = bitmap_total / 4 - tile * (SEG_BITS_BYTES / 4);
                    for (int w = tid; w < SEG_BITS_COPY / 4; w += SEG_NT) bw[w] = w < words_left ? src[w] : 0u;
                }
                __syncthreads();
                seg_phase<Pol, HAS_NULLS, false, FUSED>(c, A, r0, phase, slot, bsm, bad, &fhead, &finc);
            }
            if (phase == 0 && (tid & 31) == 0) {  // publish my first row for the previous warp's last lane
                WarpEdge e;
                e.kf = (int64_t)c.kf;
                e.t = c.first_t;
                e.raw = c.first_raw;
                e.flags = c.first_flags;
                e._pad = 0;
                wedge[tid >> 5] = e;
            }
            // WARP_SYNC: a slot is refilled by whichever warp lets go of it LAST (a counter per slot), not behind a CTA-wide
            // barrier per phase: a warp that met a window boundary in this phase no longer holds up the other three (they
            // run ahead by up to nstages - 1 phases).  Tiles staged by plain loads keep the barrier.  Measured on B200
            // (gpurun_out/s19_*): the fused integral kernel gains 12 % (configs[2] fused 15.1 -> 13.3 ms), the DRAM-bound
            // instantiations lose 3 % (their lock step keeps the DRAM pages of a tile together) - so it is on for the
            // former only.
            if (WARP_SYNC && fullt) {
                __syncwarp();
                if ((tid & 31) == 0) {
                    __threadfence_block();
                    if (atomicAdd(&done[s], 1) == SEG_NW - 1) {
                        done[s] = 0;
                        issue_item(q + nstages);
                    }
                }
            } else {
                __syncthreads();  // every read of this slot is done
                if (tid == 0) issue_item(q + nstages);
            }
        }
        if (WARP_SYNC) __syncthreads();  // every warp is through the tile's phases and has published its first row (wedge)
        if (FUSED) seg_resolve_deferred<Pol>(c, A, &finc);
        if (FUSED && SEG_CFG_FUSED_COLD == 2) {
            c.head = fhead;
            c.inc_head = finc;
        }
        if (fullt)
            seg_stitch<Pol, true, FUSED>(c, A, tile, wtot, wedge, bad);
        else
            seg_stitch<Pol, false, FUSED>(c, A, tile, wtot, wedge, bad);
        __syncthreads();  // wtot / wedge are reused by the next tile
    }
    if (bad) atomicOr(A.status, ST_UNSORTED);
}

// ---- skip records: windows that span very many tiles -------------------------------------------------------------------
// The fix-up walk below visits one tile record per step, which is fine for windows of a few tiles and hopeless for
// aggregation.Aggregate over a whole Bow (ONE window: 122 070 tiles at 1e9 rows).  When windows are that long the tile
// records are first combined bottom-up: level l holds, for every aligned group of 32^l tiles that lies INSIDE one window
// (no tile of it closes a window), the left-to-right combination of its records; the walk then jumps over whole groups.
constexpr int SKIP_FAN = 32, SKIP_LEVELS = 3;
inline int64_t seg_skip_level_count(int64_t ntiles, int level) {  // level 1..SKIP_LEVELS
    int64_t n = ntiles;
    for (int l = 0; l < level; ++l) n = (n + SKIP_FAN - 1) / SKIP_FAN;
    return n;
}
inline int64_t seg_skip_records(int64_t ntiles) {
    int64_t t = 0;
    for (int l = 1; l <= SKIP_LEVELS; ++l) t += seg_skip_level_count(ntiles, l);
    return t;
}
__host__ __device__ inline int64_t seg_skip_offset(int64_t ntiles, int level) {  // first record of `level` inside the skip array
    int64_t off = 0, n = ntiles;
    for (int l = 1; l < level; ++l) {
        n = (n + SKIP_FAN - 1) / SKIP_FAN;
        off += n;
    }
    return off;
}
// record of group i of `level` = children [32 i, 32 i + 32) of the level below (level 0 = the tiles' head records); a group
// that is incomplete, holds a closing tile or mixes windows gets key -2
template <class Pol>
__global__ void seg_skip_build_kernel(const __grid_constant__ SegArgs<Pol> A, const int64_t ntiles, const int level) {
    using Carry = typename Pol::Carry;
    int64_t nc = ntiles;  // children at the level below
    for (int l = 1; l < level; ++l) nc = (nc + SKIP_FAN - 1) / SKIP_FAN;
    const int64_t ngroups = (nc + SKIP_FAN - 1) / SKIP_FAN;
    const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= ngroups) return;
    const Carry *child = level == 1 ? A.carry_head : A.skip + seg_skip_offset(ntiles, level - 1);
    Carry *out = A.skip + seg_skip_offset(ntiles, level);
    const int64_t c0 = gi * SKIP_FAN;
    Carry acc;
    bool ok = c0 + SKIP_FAN <= nc;
    if (ok) {
        acc = child[c0];
        ok = Pol::carry_key(acc) >= 0 && !Pol::carry_closed(acc);
        const int64_t key = Pol::carry_key(acc);
        for (int j = 1; ok && j < SKIP_FAN; ++j) {
            const Carry h = child[c0 + j];
            ok = Pol::carry_key(h) == key && !Pol::carry_closed(h);
            if (ok) Pol::carry_combine(acc, h);
        }
        if (ok) Pol::carry_set_key(acc, key);
    }
    if (!ok) {
        acc = child[c0 < nc ? c0 : 0];
        Pol::carry_set_key(acc, -2);
    }
    out[gi] = acc;
}

// Joins the per-tile records, strictly left to right.  Thread j owns the windows whose first row lies in tile j
// and that are not complete inside it: the one at the left edge of the tile (head record, unless it continues a
// window of an earlier tile) and the one open at its right edge (tail record).
template <class Pol, bool FUSED>
__global__ void seg_fixup_kernel(const __grid_constant__ SegArgs<Pol> A, const int64_t ntiles) {
    using Carry = typename Pol::Carry;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ntiles) return;
    const WindowGeom &g = A.g;
    auto walk = [&](Carry a, int64_t i) {
        const int64_t key = Pol::carry_key(a);
        for (; i < ntiles; ++i) {
            if (A.skip && (i % SKIP_FAN) == 0) {  // jump over whole groups of tiles that lie inside this window
                bool jumped = false;
                int64_t span = SKIP_FAN * SKIP_FAN * SKIP_FAN;
                for (int l = SKIP_LEVELS; l >= 1 && !jumped; --l, span /= SKIP_FAN)
                    if (i % span == 0 && i + span <= ntiles) {
                        const Carry g_ = A.skip[seg_skip_offset(ntiles, l) + i / span];
                        if (Pol::carry_key(g_) == key) {
                            Pol::carry_combine(a, g_);
                            i += span - 1;  // (the loop adds one)
                            jumped = true;
                        }
                    }
                if (jumped) continue;
            }
            const Carry h = A.carry_head[i];
            if (Pol::carry_key(h) != key) {  // the window ended exactly at the tile boundary
                if (FUSED) {  // (also writes the empty windows up to the next tile's first window)
                    typename Pol::State fresh;
                    Pol::carry_set_inc(a, fused_boundary<Pol>(A, (uint64_t)key, true, (uint64_t)Pol::carry_key(h),
                                                              Pol::carry_edge_t(h), Pol::carry_edge_raw(h),
                                                              Pol::carry_edge_valid(h), fresh));
                } else {
                    const int64_t E = (int64_t)((uint64_t)g.s0 + ((uint64_t)key + 1) * g.div.d);
                    Pol::carry_inc_from_edge(a, h, E);
                }
                Pol::write_carry(A.out, g, key, a);
                return;
            }
            Pol::carry_combine(a, h);
            if (Pol::carry_closed(h)) {
                Pol::write_carry(A.out, g, key, a);
                return;
            }
        }
        Pol::carry_clear_inc(a);
        Pol::write_carry(A.out, g, key, a);  // the column ends inside the window
    };
    Carry hd = A.carry_head[j];
    const Carry tl = A.carry_tail[j];
    bool starts = true;
    if (j > 0) {
        const Carry pt = A.carry_tail[j - 1];
        int64_t open_key = Pol::carry_key(pt);
        if (open_key < 0) {
            const Carry ph = A.carry_head[j - 1];
            if (!Pol::carry_closed(ph)) open_key = Pol::carry_key(ph);
        }
        starts = open_key != Pol::carry_key(hd);
        if (Pol::carry_edge_t(hd) < Pol::carry_edge_t(pt)) atomicOr(A.status, ST_UNSORTED);
    }
    if (starts) {
        if (FUSED) {  // the window begins at the tile's first row: its synthetic start row comes first
            const int64_t k = Pol::carry_key(hd);
            if (k < A.syn.len && A.syn.missing[k])
                Pol::carry_prepend_point(hd, (int64_t)((uint64_t)g.s0 + (uint64_t)k * g.div.d), A.syn.val[k], A.syn.ok[k] != 0);
        }
        if (Pol::carry_closed(hd))
            Pol::write_carry(A.out, g, Pol::carry_key(hd), hd);
        else
            walk(hd, j + 1);
    }
    if (Pol::carry_key(tl) >= 0) walk(tl, j + 1);
}

// launch knobs (defaults chosen on B200): pipeline depth and resident CTAs per SM
inline int seg_smem_bytes(int nstages) { return SEG_HEADER_BYTES + 2 * SEG_BITS_STRIDE + nstages * SEG_SLOT_BYTES; }
inline void seg_knobs(int &nstages, int &ctas, int def_stages, int def_ctas) {
    const char *a = getenv("BOWGPU_SEG_STAGES"), *b = getenv("BOWGPU_SEG_CTAS");
    nstages = a ? atoi(a) : def_stages;
    ctas = b ? atoi(b) : def_ctas;
    if (ctas < 1) ctas = 1;
    if (nstages < 1) nstages = 1;
    if (nstages > SEG_MAX_STAGES) nstages = SEG_MAX_STAGES;
    while (nstages > 1 && (seg_smem_bytes(nstages) + 1024) * ctas > 233472) --nstages;
}

// Tensor map viewing a column of n 8-byte elements as [n / RE][RE]; box = [NT][P+2].  Only whole rows of the view
// are addressable, which is all the kernel asks for (tiles that are not complete are staged by plain loads).
// process-wide launch knobs, resolved once (bowgpu_aggregate_host drives the kernels from several host threads)
struct SegKnobs {
    PFN_cuTensorMapEncodeTiled_v12000 encode;
    int l2p;  // L2 promotion of the box fetches (tuning knob; 0 none, 1 64B, 2 128B, 3 256B)
};
inline const SegKnobs &seg_global_knobs() {
    static SegKnobs k;
    static std::once_flag once;
    std::call_once(once, [] {
        k.encode = nullptr;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            k.encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
        const char *e = getenv("BOWGPU_TMAP_L2");
        k.l2p = e ? atoi(e) : 1;
        if (k.l2p < 0 || k.l2p > 3) k.l2p = 3;
    });
    return k;
}
inline int seg_make_tmap(CUtensorMap *m, const void *col, int64_t n) {
    memset(m, 0, sizeof *m);
    const int64_t outer = n / SEG_RE;
    if (outer == 0) return 0;
    const SegKnobs &K = seg_global_knobs();
    if (!K.encode) return (int)cudaErrorNotSupported;
    const cuuint64_t dims[2] = {(cuuint64_t)SEG_RE, (cuuint64_t)outer};
    const cuuint64_t strides[1] = {(cuuint64_t)SEG_RE * 8};
    const cuuint32_t box[2] = {(cuuint32_t)SEG_COLS, (cuuint32_t)SEG_NT};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = K.encode(m, CU_TENSOR_MAP_DATA_TYPE_INT64, 2, const_cast<void *>(col), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, SEG_SWZ ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                (CUtensorMapL2promotion)K.l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

template <class Pol, bool HAS_NULLS, int MIN_CTAS, bool FUSED>
int seg_launch_impl(const SegArgs<Pol> &A, int sm_count, cudaStream_t stream, cudaEvent_t e0, cudaEvent_t e1) {
    const int64_t ntiles = (A.g.n + SEG_T - 1) / SEG_T;
    if (ntiles == 0) return 0;
    auto kern = segreduce_kernel<Pol, HAS_NULLS, MIN_CTAS, FUSED>;
    // resolved once per instantiation, under std::call_once: several host threads launch concurrently
    static int nstages = 0, ctas = 0;
    static std::once_flag once;
    std::call_once(once, [] { seg_knobs(nstages, ctas, 2, MIN_CTAS); });
    const int smem = seg_smem_bytes(nstages);
    static std::atomic<bool> configured[64];  // function attributes are per device (one ctx per GPU may live in one process)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured[dev & 63].store(true, std::memory_order_release);
    }
    CUtensorMap tm_time, tm_val;
    int rc = seg_make_tmap(&tm_time, A.time, A.g.n);
    if (!rc) rc = seg_make_tmap(&tm_val, A.values, A.g.n);
    if (rc) return rc;
    int64_t grid = (int64_t)sm_count * ctas;
    if (grid > ntiles) grid = ntiles;
    if (e0) cudaEventRecord(e0, stream);
    kern<<<(unsigned)grid, SEG_NT, smem, stream>>>(A, tm_time, tm_val, ntiles, nstages);
    if (e1) cudaEventRecord(e1, stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    const int fb = 128;
    SegArgs<Pol> F = A;
    // windows of (on average) dozens of tiles: combine the tile records bottom-up first (see seg_skip_build_kernel)
    const bool long_windows = A.skip && ntiles > 4 * SKIP_FAN && A.g.n / (A.g.W > 0 ? A.g.W : 1) > (int64_t)16 * SEG_T;
    if (long_windows) {
        for (int l = 1; l <= SKIP_LEVELS; ++l) {
            const int64_t groups = seg_skip_level_count(ntiles, l);
            seg_skip_build_kernel<Pol><<<(unsigned)((groups + fb - 1) / fb), fb, 0, stream>>>(A, ntiles, l);
        }
    } else {
        F.skip = nullptr;
    }
    seg_fixup_kernel<Pol, FUSED><<<(unsigned)((ntiles + fb - 1) / fb), fb, 0, stream>>>(F, ntiles);
    return (int)cudaGetLastError();
}

template <class Pol, bool HAS_NULLS, int MIN_CTAS>
int seg_launch(const SegArgs<Pol> &A, int sm_count, cudaStream_t stream, cudaEvent_t e0, cudaEvent_t e1) {
    return A.syn.missing ? seg_launch_impl<Pol, HAS_NULLS, MIN_CTAS, true>(A, sm_count, stream, e0, e1)
                         : seg_launch_impl<Pol, HAS_NULLS, MIN_CTAS, false>(A, sm_count, stream, e0, e1);
}

}  // namespace bowgpu
