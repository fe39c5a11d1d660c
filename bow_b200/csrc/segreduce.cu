// segreduce — the segmented-reduction kernel family behind Rolling.Aggregate.
//
// Replaces, for one input column, the reference's per-aggregation re-iteration of every window
// (rolling/aggregation.go:190-238) and the per-window closures Count / Sum / ArithmeticMean / Min /
// Max / First / Last (rolling/aggregation/{count,sum,arithmeticmean,minmax,firstlast}.go) by ONE
// streaming pass over the time column and the value column.
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
//     handled by an out-of-line rolled loop;
//   * partial states of windows spanning several threads are stitched by a warp-shuffle segmented
//     scan (flags = "a window closed inside this thread") plus a short cross-warp pass;
//   * partial states of windows spanning several tiles go to per-tile head / tail carry records
//     and are stitched left to right by a tiny fix-up kernel (deterministic, no atomics).
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
#include "kernels.h"
#include "tile_pipe.cuh"

#include <math_constants.h>

#include <cstdlib>

namespace bowgpu {

namespace {

constexpr int SEG_NT = 128;
constexpr int SEG_R = 17;
constexpr int SEG_MAX_STAGES = 8;
constexpr int SEG_MIN_CTAS = 3;
using SegG = TileGeom<SEG_NT, SEG_R>;
constexpr int SEG_NW = SEG_NT / 32;
constexpr int SEG_STAGE_BYTES = SegG::TIME_BYTES + SegG::VAL_BYTES + SegG::BITS_STRIDE;
constexpr int SEG_HEADER_BYTES = 1024;
constexpr int64_t CLOSED_BIT = (int64_t)1 << 62;
static_assert(SegG::T < 4096, "cnt / fi are packed in 12 bits each");

// state of a run of rows; meta = cnt | fi << 12 (tile-relative first valid row), li = last valid row
struct BState {
    double sum, mn, mx;
    uint32_t meta;
    uint32_t li;
};
__device__ __forceinline__ uint32_t st_cnt(const BState &s) { return s.meta & 0xFFFu; }
__device__ __forceinline__ uint32_t st_fi(const BState &s) { return s.meta >> 12; }

__device__ __forceinline__ BState st_identity() {
    BState s;
    s.sum = 0.0;
    s.mn = CUDART_INF;
    s.mx = -CUDART_INF;
    s.meta = 0;
    s.li = 0;
    return s;
}

template <uint32_t OPS>
__device__ __forceinline__ BState st_combine(const BState &L, const BState &R) {
    BState o;
    o.sum = L.sum + R.sum;
    if (OPS & OPS_MINMAX) {
        o.mn = (R.mn < L.mn) ? R.mn : L.mn;
        o.mx = (R.mx > L.mx) ? R.mx : L.mx;
    }
    const uint32_t lc = L.meta & 0xFFFu, rc = R.meta & 0xFFFu;
    o.meta = (lc + rc) | ((lc ? L.meta : R.meta) & 0xFFF000u);
    if (OPS & OPS_FIRSTLAST) o.li = rc ? R.li : L.li;
    return o;
}

template <uint32_t OPS>
__device__ __forceinline__ BState st_shfl_up(const BState &s, int d) {
    BState o;
    o.sum = __shfl_up_sync(0xffffffffu, s.sum, d);
    if (OPS & OPS_MINMAX) {
        o.mn = __shfl_up_sync(0xffffffffu, s.mn, d);
        o.mx = __shfl_up_sync(0xffffffffu, s.mx, d);
    }
    o.meta = __shfl_up_sync(0xffffffffu, s.meta, d);
    if (OPS & OPS_FIRSTLAST) o.li = __shfl_up_sync(0xffffffffu, s.li, d);
    return o;
}

template <uint32_t OPS, bool IS_INT>
__device__ __forceinline__ void st_accumulate(BState &s, uint64_t raw) {
    const double v = IS_INT ? (double)(int64_t)raw : bits_as_f64(raw);  // GetFloat64, bowgetters.go:218-229
    s.sum += v;
    if (OPS & OPS_MINMAX) {
        if (v < s.mn) s.mn = v;
        if (v > s.mx) s.mx = v;
    }
}

// cnt / fi / li of the rows selected by `mask` (bit j = row ti0 + j of the tile)
__device__ __forceinline__ void st_set_meta(BState &s, uint32_t mask, int ti0) {
    const uint32_t cnt = __popc(mask);
    const uint32_t fi = mask ? (uint32_t)(ti0 + __ffs(mask) - 1) : 0u;
    s.meta = cnt | (fi << 12);
    s.li = mask ? (uint32_t)(ti0 + 31 - __clz(mask)) : 0u;
}

struct WarpTotal {
    BState st;
    uint32_t flag;
    uint32_t _pad;
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

// exact division on the rare paths
__device__ __noinline__ uint64_t div_slow(uint64_t x, uint64_t d, double inv_rd) {
    DivU64 dv{d, inv_rd};
    return div_u64(x, dv);
}

// Windows that begin AND end strictly inside one thread's rows (only when windows are shorter than
// R rows): rows (jfirst, jlast] of the thread are re-reduced window by window from shared memory in a
// rolled loop and written out directly.  Kept out of line so the unrolled fast path stays small.
template <uint32_t OPS, bool IS_INT>
__device__ __noinline__ void middle_windows(const BasicOut *outp, int64_t W, uint64_t d, double inv_rd, int64_t s0,
                                            const int64_t *trow, const uint64_t *vrow, uint32_t vbits, int jfirst,
                                            int jlast) {
    const BasicOut o = *outp;
    BState st = st_identity();
    uint32_t seg = 0;  // valid rows of the current window
    uint64_t kcur = div_slow((uint64_t)trow[jfirst + 1] - (uint64_t)s0, d, inv_rd);
    uint64_t erel = (kcur + 1) * d;
    for (int j = jfirst + 1; j <= jlast; ++j) {
        if ((vbits >> j) & 1u) {
            st_accumulate<OPS, IS_INT>(st, vrow[j]);
            seg |= 1u << j;
        }
        const uint64_t xn = (uint64_t)trow[j + 1] - (uint64_t)s0;
        if (j == jlast || xn >= erel) {
            const uint64_t fb = seg ? vrow[__ffs(seg) - 1] : 0, lb = seg ? vrow[31 - __clz(seg)] : 0;
            write_window<IS_INT>(o, W, (int64_t)kcur, __popc(seg), st.sum, st.mn, st.mx, fb, lb);
            st = st_identity();
            seg = 0;
            if (xn - erel < d) {
                ++kcur;
                erel += d;
            } else {
                kcur = div_slow(xn, d, inv_rd);
                erel = (kcur + 1) * d;
            }
        }
    }
}

// One tile.  FULL: every row of the tile, the row before it and the row after it exist and no row
// lies before s0 — the common case, free of per-row existence predicates.
template <uint32_t OPS, bool IS_INT, bool HAS_NULLS, bool FULL>
__device__ __forceinline__ void seg_tile(const SegLaunch &P, const BasicOut *sh_out, const int64_t tile,
                                         const uint8_t *sb, WarpTotal *wtot, volatile int *sh_flags, bool &bad) {
    using G = SegG;
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

    uint32_t vbits = (1u << R) - 1u;
    if (HAS_NULLS) {
        const uint32_t lo = bsm[ti0 >> 5], hi = bsm[(ti0 >> 5) + 1];
        vbits &= __funnelshift_r(lo, hi, ti0 & 31);
    }
    if (!FULL) {
        vbits &= (1u << nmine) - 1u;
        if (early_tile && !g.early_keep) {  // rows before s0 are dropped (see WindowGeom)
            const int64_t e = g.early_rows - (r0 + ti0);
            if (e > 0) vbits &= e >= R ? 0u : ~((1u << (int)e) - 1u);
        }
    }

    int64_t x[R + 1];
    uint64_t raw[R];
#pragma unroll
    for (int j = 0; j <= R; ++j) x[j] = tsm[ti0 + 2 + j];
#pragma unroll
    for (int j = 0; j < R; ++j) raw[j] = vsm[ti0 + j];

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

    BState st = st_identity();
    BState head = st_identity();
    uint32_t cmask = 0;  // bit j: the window of row j closes after row j
#pragma unroll
    for (int j = 0; j < R; ++j) {
        if (FULL || j < nmine) {
            if ((vbits >> j) & 1u) st_accumulate<OPS, IS_INT>(st, raw[j]);
            const bool next_exists = FULL || ti0 + j + 1 < nrem_i;
            if (!next_exists || x[j + 1] >= eabs) {
                if (cmask == 0) head = st;
                cmask |= 1u << j;
                st = st_identity();
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
        st_set_meta(head, vbits & ((2u << jfirst) - 1u), ti0);
        st_set_meta(st, vbits & ~((2u << jlast) - 1u), ti0);
    } else {
        st_set_meta(st, vbits, ti0);
    }
    // windows lying strictly inside this thread's rows (short windows only)
    if (jlast > jfirst)
        middle_windows<OPS, IS_INT>(sh_out, g.W, d, g.div.inv_rd, g.s0, tsm + ti0 + 2, vsm + ti0, vbits, jfirst, jlast);

    // ---- stitch windows spanning threads: segmented inclusive scan of the tails -------------
    uint32_t f = cmask != 0;
    BState sc = st;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
        const BState o = st_shfl_up<OPS>(sc, dd);
        const uint32_t of = __shfl_up_sync(0xffffffffu, f, dd);
        if (lane >= dd) {
            if (!f) sc = st_combine<OPS>(o, sc);
            f |= of;
        }
    }
    const uint32_t ball = __ballot_sync(0xffffffffu, cmask != 0);
    if (lane == 31) {
        wtot[warp].st = sc;
        wtot[warp].flag = ball != 0;
    }
    BState ex = st_shfl_up<OPS>(sc, 1);
    if (lane == 0) ex = st_identity();
    __syncthreads();
    const bool left_open = sh_flags[0] != 0;
    BState acc = st_identity();
    bool any_prev = false;
    for (int u = 0; u < warp; ++u) {
        const BState ws = wtot[u].st;
        if (wtot[u].flag) {
            acc = ws;
            any_prev = true;
        } else {
            acc = st_combine<OPS>(acc, ws);
        }
    }
    const bool flag_before = (ball & ((1u << lane) - 1u)) != 0;
    const BState excl = flag_before ? ex : st_combine<OPS>(acc, ex);
    const bool any_excl = any_prev || flag_before;

    if (cmask) {  // this thread closes the window that was open at its left edge
        const BState h = st_combine<OPS>(excl, head);
        const uint32_t hc = st_cnt(h);
        const uint64_t fb = hc ? vsm[st_fi(h)] : 0, lb = hc ? vsm[h.li] : 0;
        if (!any_excl && left_open) {
            BasicCarry c;
            c.key = (int64_t)kf;
            c.cnt = (int64_t)hc | CLOSED_BIT;
            c.sum = h.sum;
            c.mn = h.mn;
            c.mx = h.mx;
            c.first = fb;
            c.last = lb;
            c._pad = 0;
            P.carry_head[tile] = c;
        } else {
            write_window<IS_INT>(P.out, g.W, (int64_t)kf, hc, h.sum, h.mn, h.mx, fb, lb);
        }
    }
    if (tid == SEG_NT - 1) {  // tile-level records: the window open at the right edge
        const bool flag_incl = flag_before || cmask != 0;
        const BState incl = flag_incl ? sc : st_combine<OPS>(acc, sc);
        const bool any_incl = any_prev || flag_incl;
        // the last row of the tile closes its window: end of data, or my last row closed
        const bool closes = nrem <= G::T || ((cmask >> (R - 1)) & 1u);
        const uint32_t ic = st_cnt(incl);
        BasicCarry c;
        c.key = 0;  // tail records take their window index from the next tile's head record
        c.cnt = (int64_t)ic;
        c.sum = incl.sum;
        c.mn = incl.mn;
        c.mx = incl.mx;
        c.first = ic ? vsm[st_fi(incl)] : 0;
        c.last = ic ? vsm[incl.li] : 0;
        c._pad = 0;
        BasicCarry none = c;
        none.key = -1;
        if (!any_incl && left_open) {  // the whole tile lies inside one window that began earlier
            c.key = (int64_t)kf;       // (every thread of the tile has the same kf)
            P.carry_head[tile] = c;    // not closed
            P.carry_tail[tile] = none;
        } else {
            if (!left_open) P.carry_head[tile] = none;
            P.carry_tail[tile] = closes ? none : c;
        }
    }
}

template <uint32_t OPS, bool IS_INT, bool HAS_NULLS>
__global__ void __launch_bounds__(SEG_NT, SEG_MIN_CTAS)
    segreduce_basic_kernel(const SegLaunch P, const int64_t ntiles, const int nstages) {
    using G = SegG;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);                       // [SEG_MAX_STAGES]
    WarpTotal *wtot = reinterpret_cast<WarpTotal *>(smem_raw + 64);                // [SEG_NW]
    volatile int *sh_flags = reinterpret_cast<volatile int *>(smem_raw + 64 + SEG_NW * sizeof(WarpTotal));
    BasicOut *sh_out = reinterpret_cast<BasicOut *>(smem_raw + 512);
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
        const bool full_tile = r0 > 0 && g.n - r0 > G::T && r0 >= g.early_rows;
        if (full_tile)
            seg_tile<OPS, IS_INT, HAS_NULLS, true>(P, sh_out, tile, sb, wtot, sh_flags, bad);
        else
            seg_tile<OPS, IS_INT, HAS_NULLS, false>(P, sh_out, tile, sb, wtot, sh_flags, bad);
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
__global__ void seg_fixup_kernel(const SegLaunch P, const int64_t ntiles) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ntiles) return;
    BasicCarry a = P.carry_tail[j];
    if (a.key < 0) return;
    int64_t key = -1;
    for (int64_t i = j + 1; i < ntiles; ++i) {
        const BasicCarry h = P.carry_head[i];
        if (h.key < 0 || (key >= 0 && h.key != key)) break;
        key = h.key;
        const int64_t hc = h.cnt & ~CLOSED_BIT;
        a.sum = a.sum + h.sum;
        a.mn = (h.mn < a.mn) ? h.mn : a.mn;
        a.mx = (h.mx > a.mx) ? h.mx : a.mx;
        a.first = a.cnt ? a.first : h.first;
        a.last = hc ? h.last : a.last;
        a.cnt += hc;
        if (h.cnt & CLOSED_BIT) break;
    }
    if (key < 0) return;  // cannot happen: an open tail is always continued by the next tile's head
    if (P.is_int)
        write_window<true>(P.out, P.g.W, key, a.cnt, a.sum, a.mn, a.mx, a.first, a.last);
    else
        write_window<false>(P.out, P.g.W, key, a.cnt, a.sum, a.mn, a.mx, a.first, a.last);
}

template <uint32_t OPS, bool IS_INT, bool HAS_NULLS>
int launch_inst(const SegLaunch &L, int64_t ntiles, int sm_count, cudaStream_t stream, cudaEvent_t e0,
                cudaEvent_t e1) {
    auto kern = segreduce_basic_kernel<OPS, IS_INT, HAS_NULLS>;
    // tuning knobs (defaults chosen on B200): pipeline depth and resident CTAs per SM
    static int nstages = 0, ctas = 0;
    if (!nstages) {
        const char *a = getenv("BOWGPU_SEG_STAGES"), *b = getenv("BOWGPU_SEG_CTAS");
        nstages = a ? atoi(a) : 2;
        ctas = b ? atoi(b) : 3;
        if (nstages < 1) nstages = 1;
        if (nstages > SEG_MAX_STAGES) nstages = SEG_MAX_STAGES;
        while (nstages > 1 && (SEG_HEADER_BYTES + nstages * SEG_STAGE_BYTES + 1024) * ctas > 232448) --nstages;
    }
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
    kern<<<(unsigned)grid, SEG_NT, smem, stream>>>(L, ntiles, nstages);
    if (e1) cudaEventRecord(e1, stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    const int fb = 128;
    seg_fixup_kernel<<<(unsigned)((ntiles + fb - 1) / fb), fb, 0, stream>>>(L, ntiles);
    return (int)cudaGetLastError();
}

template <uint32_t OPS>
int launch_ops(const SegLaunch &L, int64_t ntiles, int sm, cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1) {
    const bool nulls = L.validity != nullptr;
    if (L.is_int)
        return nulls ? launch_inst<OPS, true, true>(L, ntiles, sm, s, e0, e1)
                     : launch_inst<OPS, true, false>(L, ntiles, sm, s, e0, e1);
    return nulls ? launch_inst<OPS, false, true>(L, ntiles, sm, s, e0, e1)
                 : launch_inst<OPS, false, false>(L, ntiles, sm, s, e0, e1);
}

}  // namespace

int64_t seg_num_tiles(int64_t n) { return (n + SegG::T - 1) / SegG::T; }
size_t seg_carry_bytes(int64_t n) { return (size_t)seg_num_tiles(n) * 2 * sizeof(BasicCarry); }

int launch_segreduce_basic(const SegLaunch &L, int sm_count, cudaStream_t stream, cudaEvent_t e0, cudaEvent_t e1) {
    const int64_t ntiles = seg_num_tiles(L.g.n);
    if (ntiles == 0) return 0;
    if (L.ops & OPS_FIRSTLAST) return launch_ops<OPS_SUMCNT | OPS_MINMAX | OPS_FIRSTLAST>(L, ntiles, sm_count, stream, e0, e1);
    if (L.ops & OPS_MINMAX) return launch_ops<OPS_SUMCNT | OPS_MINMAX>(L, ntiles, sm_count, stream, e0, e1);
    return launch_ops<OPS_SUMCNT>(L, ntiles, sm_count, stream, e0, e1);
}

}  // namespace bowgpu
