// segreduce.cuh — the generic streaming segmented reduction behind Rolling.Aggregate.
//
// One pass over the time column and one value column computes, for every window, a window STATE
// defined by a policy (a monoid: identity / accumulate one row / combine two adjacent runs).
// seg_basic.cu instantiates it for Count/Sum/Mean/Min/Max/First/Last, seg_integral.cu for
// IntegralStep/IntegralTrapezoid/WeightedAverageStep/WeightedAverageLinear.
//
// Work decomposition (load balance independent of window sizes, SURVEY 7 "hard parts"):
//   * fixed-size ROW tiles (T = NT*R rows) are assigned round-robin to persistent CTAs and staged
//     into shared memory by TMA bulk copies (tile_pipe.cuh);
//   * thread t reduces its R consecutive rows sequentially, left to right, exactly like the
//     reference closure does inside a window.  A window "closes" after row i when row i+1 lies in a
//     later window; the thread owning row i detects that from its R+1 timestamps by comparing
//     against a running absolute window end (one exact division per thread and tile, none per row);
//   * the thread keeps the state of the rows before its first closing (head) and after its last
//     closing (tail); windows lying strictly inside one thread (windows shorter than R rows) are
//     handled by an out-of-line rolled loop of the policy;
//   * partial states of windows spanning several threads are stitched by a warp-shuffle segmented
//     scan (flags = "a window closed inside this thread") plus a short cross-warp pass;
//   * partial states of windows spanning several tiles go to per-tile head / tail carry records
//     and are stitched left to right by a tiny fix-up kernel (deterministic, no atomics).
//
// Policy interface (all static, device):
//   State, Carry, Out, Inc                      types (Inc = what a closing row knows about the row after it)
//   identity(), accumulate(State&, t, raw), set_meta(State&, mask, ti0), combine(L, R), shfl_up(s, d)
//   make_inc(at_end, valid_next, raw_next, t_next)   the inclusive row of a window (rolling.go:201-209)
//   write(out, g, k, state, inc, vsm)           final values of a window that lies inside one tile
//   make_carry(state, inc, vsm, key, closed)    -> Carry;  carry_key / carry_closed / carry_combine / write_carry
//   middle(sh_out, W, s0, d, inv_rd, trow, vrow, vraw, jfirst, jlast, nexist)   windows strictly inside one thread's rows
#pragma once
#include <cstdlib>

#include "kernels.h"
#include "tile_pipe.cuh"

namespace bowgpu {

// tile shape and residency (compile-time; -DSEG_CFG_* only for tuning sweeps, scripts/tune_seg.sh)
#ifndef SEG_CFG_NT
#define SEG_CFG_NT 128
#endif
#ifndef SEG_CFG_R
#define SEG_CFG_R 17
#endif
#ifndef SEG_CFG_CTAS
#define SEG_CFG_CTAS 3
#endif
constexpr int SEG_NT = SEG_CFG_NT;
constexpr int SEG_R = SEG_CFG_R;
constexpr int SEG_MAX_STAGES = 8;
using SegG = TileGeom<SEG_NT, SEG_R>;
constexpr int SEG_NW = SEG_NT / 32;
constexpr int SEG_STAGE_BYTES = SegG::STAGE_BYTES;
constexpr int SEG_HEADER_BYTES = 1024;

template <class Pol>
struct SegArgs {
    const int64_t *time;
    const uint64_t *values;
    const uint8_t *validity;  // null = all valid
    WindowGeom g;
    typename Pol::Out out;
    typename Pol::Carry *carry_head;  // [ntiles]
    typename Pol::Carry *carry_tail;  // [ntiles]
    int32_t *status;
};

template <class Pol>
struct WarpTotal {
    typename Pol::State st;
    uint32_t flag;
    uint32_t _pad;
};

// exact division on the rare paths
__device__ __noinline__ static uint64_t div_slow(uint64_t x, uint64_t d, double inv_rd) {
    DivU64 dv{d, inv_rd};
    return div_u64(x, dv);
}

// One tile.  FULL: every row of the tile, the row before it and the two rows after it exist and no
// row lies before s0 — the common case, free of per-row existence predicates.
template <class Pol, bool HAS_NULLS, bool FULL>
__device__ __forceinline__ void seg_tile(const SegArgs<Pol> &P, const typename Pol::Out *sh_out, const int64_t tile,
                                         const uint8_t *sb, WarpTotal<Pol> *wtot, volatile int *sh_flags, bool &bad) {
    using G = SegG;
    using State = typename Pol::State;
    using Inc = typename Pol::Inc;
    using Carry = typename Pol::Carry;
    constexpr int R = G::R;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const WindowGeom &g = P.g;
    const uint64_t d = g.div.d;
    const int64_t *tsm = reinterpret_cast<const int64_t *>(sb);
    const uint64_t *vsm = reinterpret_cast<const uint64_t *>(sb + G::TIME_BYTES);
    const uint32_t *bsm = reinterpret_cast<const uint32_t *>(sb + G::TIME_BYTES + G::VAL_BYTES);
    const int64_t r0 = tile * G::T;
    const int64_t nrem = g.n - r0;  // rows from the tile start to the end of the column (> 0)
    const int nrem_i = FULL ? G::T + 2 : (nrem > G::T + 2 ? G::T + 2 : (int)nrem);
    const int ti0 = tid * R;
    const int nmine = FULL ? R : (nrem_i - ti0 < 0 ? 0 : (nrem_i - ti0 > R ? R : nrem_i - ti0));  // rows I own
    const bool early_tile = !FULL && r0 < g.early_rows;

    // validity of my R rows and of the row after them (bit R)
    uint32_t vraw = (2u << R) - 1u;
    if (HAS_NULLS) {
        const uint32_t lo = bsm[ti0 >> 5], hi = bsm[(ti0 >> 5) + 1];
        vraw &= __funnelshift_r(lo, hi, ti0 & 31);
    }
    uint32_t vbits = vraw & ((1u << R) - 1u);
    if (!FULL) {
        vbits &= (1u << nmine) - 1u;
        if (early_tile && !g.early_keep) {  // rows before s0 are dropped (see WindowGeom)
            const int64_t e = g.early_rows - (r0 + ti0);
            if (e > 0) vbits &= e >= R ? 0u : ~((1u << (int)e) - 1u);
        }
    }

    int64_t x[R + 1];
    uint64_t raw[R + 1];
#pragma unroll
    for (int j = 0; j <= R; ++j) x[j] = tsm[ti0 + 2 + j];
#pragma unroll
    for (int j = 0; j < R; ++j) raw[j] = vsm[ti0 + j];
    raw[R] = Pol::NEXT_VALUE ? vsm[ti0 + R] : 0;

    // precondition check: time sorted ascending (every adjacent pair is checked exactly once)
    if (FULL || (nmine > 0 && r0 + ti0 > 0)) bad |= x[0] < tsm[ti0 + 1];
#pragma unroll
    for (int j = 1; j < R; ++j)
        if (FULL || j < nmine) bad |= x[j] < x[j - 1];

    // window of my first row (rows before s0 collapse onto window 0) and its absolute end
    uint64_t kf = 0;
    if (nmine > 0) {
        const bool early0 = early_tile && r0 + ti0 < g.early_rows;
        kf = early0 ? 0 : div_u64((uint64_t)x[0] - (uint64_t)g.s0, g.div);
    }
    int64_t eabs = (int64_t)((uint64_t)g.s0 + (kf + 1) * d);
    if (tid == 0) {  // does the window of my first row continue from the previous tile?
        int lo_open = 0;
        if (r0 > 0) {
            const bool earlyp = early_tile && r0 - 1 < g.early_rows;
            lo_open = earlyp ? kf == 0 : (uint64_t)tsm[1] - (uint64_t)g.s0 >= kf * d;
        }
        sh_flags[0] = lo_open;
    }

    State st = Pol::identity();
    State head = Pol::identity();
    Inc inc_head = Pol::make_inc(false, false, 0, 0);
    uint32_t cmask = 0;  // bit j: the window of row j closes after row j
#pragma unroll
    for (int j = 0; j < R; ++j) {
        if (FULL || j < nmine) {
            if ((vbits >> j) & 1u) Pol::accumulate(st, x[j], raw[j]);
            const bool next_exists = FULL || ti0 + j + 1 < nrem_i;
            if (!next_exists || x[j + 1] >= eabs) {
                if (cmask == 0) {
                    head = st;
                    if (Pol::NEXT_VALUE)
                        inc_head = Pol::make_inc(next_exists && x[j + 1] == eabs, (vraw >> (j + 1)) & 1u, raw[j + 1],
                                                 x[j + 1]);
                }
                cmask |= 1u << j;
                st = Pol::identity();
                if (next_exists) {
                    if ((uint64_t)x[j + 1] - (uint64_t)eabs < d) {
                        eabs = (int64_t)((uint64_t)eabs + d);
                    } else {
                        const uint64_t k = div_slow((uint64_t)x[j + 1] - (uint64_t)g.s0, d, g.div.inv_rd);
                        eabs = (int64_t)((uint64_t)g.s0 + (k + 1) * d);
                    }
                }
            }
        }
    }
    const int jfirst = __ffs(cmask) - 1, jlast = 31 - __clz(cmask);  // valid when cmask != 0
    if (cmask) {
        Pol::set_meta(head, vbits & ((2u << jfirst) - 1u), ti0);
        Pol::set_meta(st, vbits & ~((2u << jlast) - 1u), ti0);
    } else {
        Pol::set_meta(st, vbits, ti0);
    }
    // windows lying strictly inside this thread's rows (short windows only)
    if (jlast > jfirst) {
        const int nexist = FULL ? R + 1 : (nrem_i - ti0 > R + 1 ? R + 1 : nrem_i - ti0);  // of my R+1 entries
        Pol::middle(sh_out, g.W, g.s0, d, g.div.inv_rd, tsm + ti0 + 2, vsm + ti0, vraw, jfirst, jlast, nexist);
    }

    // ---- stitch windows spanning threads: segmented inclusive scan of the tails -------------
    uint32_t f = cmask != 0;
    State sc = st;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
        const State o = Pol::shfl_up(sc, dd);
        const uint32_t of = __shfl_up_sync(0xffffffffu, f, dd);
        if (lane >= dd) {
            if (!f) sc = Pol::combine(o, sc);
            f |= of;
        }
    }
    const uint32_t ball = __ballot_sync(0xffffffffu, cmask != 0);
    if (lane == 31) {
        wtot[warp].st = sc;
        wtot[warp].flag = ball != 0;
    }
    State ex = Pol::shfl_up(sc, 1);
    if (lane == 0) ex = Pol::identity();
    __syncthreads();
    const bool left_open = sh_flags[0] != 0;
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

    if (cmask) {  // this thread closes the window that was open at its left edge
        const State h = Pol::combine(excl, head);
        if (!any_excl && left_open)
            P.carry_head[tile] = Pol::make_carry(h, inc_head, vsm, (int64_t)kf, true);
        else
            Pol::write(P.out, g, (int64_t)kf, h, inc_head, vsm);
    }
    if (tid == SEG_NT - 1) {  // tile-level records: the window open at the right edge
        const bool flag_incl = flag_before || cmask != 0;
        const State incl = flag_incl ? sc : Pol::combine(acc, sc);
        const bool any_incl = any_prev || flag_incl;
        // the last row of the tile closes its window: end of data, or my last row closed
        const bool closes = nrem <= G::T || ((cmask >> (R - 1)) & 1u);
        const Inc noinc = Pol::make_inc(false, false, 0, 0);
        // tail records take their window index from the next tile's head record (key 0 = present)
        Carry c = Pol::make_carry(incl, noinc, vsm, 0, false);
        Carry none = c;
        Pol::carry_set_key(none, -1);
        if (!any_incl && left_open) {  // the whole tile lies inside one window that began earlier
            Pol::carry_set_key(c, (int64_t)kf);  // (every thread of the tile has the same kf)
            P.carry_head[tile] = c;              // not closed
            P.carry_tail[tile] = none;
        } else {
            if (!left_open) P.carry_head[tile] = none;
            P.carry_tail[tile] = closes ? none : c;
        }
    }
}

template <class Pol, bool HAS_NULLS, int MIN_CTAS>
__global__ void __launch_bounds__(SEG_NT, MIN_CTAS)
    segreduce_kernel(const SegArgs<Pol> P, const int64_t ntiles, const int nstages) {
    using G = SegG;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);                            // [SEG_MAX_STAGES]
    WarpTotal<Pol> *wtot = reinterpret_cast<WarpTotal<Pol> *>(smem_raw + 64);           // [SEG_NW]
    static_assert(64 + SEG_NW * sizeof(WarpTotal<Pol>) + 16 <= 640, "header layout");
    volatile int *sh_flags = reinterpret_cast<volatile int *>(smem_raw + 640);
    typename Pol::Out *sh_out = reinterpret_cast<typename Pol::Out *>(smem_raw + 704);
    static_assert(704 + sizeof(typename Pol::Out) <= SEG_HEADER_BYTES, "header layout");
    uint8_t *stages = smem_raw + SEG_HEADER_BYTES;

    const int tid = threadIdx.x;
    const WindowGeom &g = P.g;
    TileSrc src{P.time, P.values, HAS_NULLS ? P.validity : nullptr, g.n};

    if (tid == 0) {
        *sh_out = P.out;
        for (int s = 0; s < nstages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();
    if (tid == 0) {
        int64_t tl = blockIdx.x;
        for (int s = 0; s < nstages && tl < ntiles; ++s, tl += gridDim.x)
            issue_tile<G, true>(src, tl, stages + (size_t)s * SEG_STAGE_BYTES, &full[s]);
    }

    int stage = 0;
    uint32_t phase = 0;
    bool bad = false;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(&full[stage], phase);
        const uint8_t *sb = stages + (size_t)stage * SEG_STAGE_BYTES;
        const int64_t r0 = tile * G::T;
        const bool full_tile = r0 > 0 && g.n - r0 >= G::T + 2 && r0 >= g.early_rows;
        if (full_tile)
            seg_tile<Pol, HAS_NULLS, true>(P, sh_out, tile, sb, wtot, sh_flags, bad);
        else
            seg_tile<Pol, HAS_NULLS, false>(P, sh_out, tile, sb, wtot, sh_flags, bad);
        __syncthreads();  // every read of this stage (and of wtot) is done
        if (tid == 0) {
            const int64_t nxt = tile + (int64_t)nstages * gridDim.x;
            if (nxt < ntiles) issue_tile<G, true>(src, nxt, stages + (size_t)stage * SEG_STAGE_BYTES, &full[stage]);
        }
        if (++stage == nstages) {
            stage = 0;
            phase ^= 1u;
        }
    }
    if (bad) atomicOr(P.status, ST_UNSORTED);
}

// Stitches windows spanning tiles, strictly left to right: tail of tile j, then the head records
// of the following tiles until the one where the window closes.
template <class Pol>
__global__ void seg_fixup_kernel(const SegArgs<Pol> P, const int64_t ntiles) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ntiles) return;
    typename Pol::Carry a = P.carry_tail[j];
    if (Pol::carry_key(a) < 0) return;
    int64_t key = -1;
    for (int64_t i = j + 1; i < ntiles; ++i) {
        const typename Pol::Carry h = P.carry_head[i];
        const int64_t hk = Pol::carry_key(h);
        if (hk < 0 || (key >= 0 && hk != key)) break;
        key = hk;
        Pol::carry_combine(a, h);
        if (Pol::carry_closed(h)) break;
    }
    if (key < 0) return;  // cannot happen: an open tail is always continued by the next tile's head
    Pol::write_carry(P.out, P.g, key, a);
}

// launch knobs (defaults chosen on B200): pipeline depth and resident CTAs per SM
inline void seg_knobs(int &nstages, int &ctas, int def_stages, int def_ctas) {
    const char *a = getenv("BOWGPU_SEG_STAGES"), *b = getenv("BOWGPU_SEG_CTAS");
    nstages = a ? atoi(a) : def_stages;
    ctas = b ? atoi(b) : def_ctas;
    if (ctas < 1) ctas = 1;
    if (nstages < 1) nstages = 1;
    if (nstages > SEG_MAX_STAGES) nstages = SEG_MAX_STAGES;
    while (nstages > 1 && (SEG_HEADER_BYTES + nstages * SEG_STAGE_BYTES + 1024) * ctas > 232448) --nstages;
}

template <class Pol, bool HAS_NULLS, int MIN_CTAS>
int seg_launch(const SegArgs<Pol> &A, int sm_count, cudaStream_t stream, cudaEvent_t e0, cudaEvent_t e1) {
    const int64_t ntiles = (A.g.n + SegG::T - 1) / SegG::T;
    if (ntiles == 0) return 0;
    auto kern = segreduce_kernel<Pol, HAS_NULLS, MIN_CTAS>;
    static int nstages = 0, ctas = 0;
    if (!nstages) seg_knobs(nstages, ctas, 2, MIN_CTAS);
    const int smem = SEG_HEADER_BYTES + nstages * SEG_STAGE_BYTES;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    int64_t grid = (int64_t)sm_count * ctas;
    if (grid > ntiles) grid = ntiles;
    if (e0) cudaEventRecord(e0, stream);
    kern<<<(unsigned)grid, SEG_NT, smem, stream>>>(A, ntiles, nstages);
    if (e1) cudaEventRecord(e1, stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    const int fb = 128;
    seg_fixup_kernel<Pol><<<(unsigned)((ntiles + fb - 1) / fb), fb, 0, stream>>>(A, ntiles);
    return (int)cudaGetLastError();
}

}  // namespace bowgpu
