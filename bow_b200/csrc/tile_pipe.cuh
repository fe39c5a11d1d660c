// Row-tile staging pipeline shared by the streaming kernels (bounds, segreduce, interpolate).
//
// A persistent CTA walks row tiles  tile = blockIdx.x + it * gridDim.x  of T = NT * R rows.  For
// every tile one elected thread issues 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) of the
// time rows (with a halo on both sides), the value rows and the validity bytes into one of
// STAGES shared-memory stages; completion is tracked by one mbarrier per stage.  All threads wait
// on the barrier, work out of shared memory, __syncthreads(), and the elected thread refills the
// stage with the tile STAGES iterations ahead.  No registers are tied up by loads in flight, and
// every global access is a full-line bulk transfer.
//
// Thread t of the CTA owns the R consecutive rows [t*R, t*R+R) of the tile.  R is odd, so the
// 8-byte shared-memory reads of a half-warp hit 16 distinct bank pairs (conflict-free).
#pragma once
#include "common.cuh"

namespace bowgpu {

template <int NT_, int R_>
struct TileGeom {
    static constexpr int NT = NT_;
    static constexpr int R = R_;
    static constexpr int T = NT * R;                 // rows per tile
#ifndef TILE_HALO_LO
#define TILE_HALO_LO 16
#endif
    static constexpr int HALO = TILE_HALO_LO;        // time rows staged before the tile: 16 rows = 128 bytes keep every
                                                     // bulk copy of the time column 128-byte aligned in global memory
    static constexpr int TIME_ENTRIES = T + HALO + 2;  // rows r0-HALO .. r0+T+1
    static constexpr int TIME_BYTES = TIME_ENTRIES * 8;
    static constexpr int VAL_ENTRIES = T + 2;        // rows r0 .. r0+T+1 (the row after the tile is the
    static constexpr int VAL_BYTES = VAL_ENTRIES * 8;  // inclusive row of a window closing at the tile end)
    static constexpr int BITS_BYTES = T / 8;         // validity bytes of one tile
    static constexpr int BITS_COPY = BITS_BYTES + 16;  // + the bits of the rows after the tile
    static constexpr int BITS_STRIDE = BITS_COPY + 16;
    static constexpr int STAGE_BYTES = TIME_BYTES + VAL_BYTES + BITS_STRIDE;
    static_assert(R % 2 == 1, "R must be odd (bank-conflict-free thread-consecutive reads)");
    static_assert(T % 128 == 0, "tile validity bytes must be a multiple of 16");
    static_assert(R <= 31, "per-thread validity bits are extracted from two 32-bit words");
};

// Stage `bytes` from global to shared: the 16-byte multiple part as one bulk copy (returned as tx
// bytes), a trailing 8-byte element (only at the very end of a caller-owned buffer) by a plain
// load/store of the issuing thread.
__device__ __forceinline__ uint32_t stage_bytes(void *dst, const void *src, uint32_t bytes, uint64_t *bar,
                                                bool issue) {
    uint32_t bulk = bytes & ~15u;
    if (!issue) {
        if (bytes & 8u) *(uint64_t *)((char *)dst + bulk) = *(const uint64_t *)((const char *)src + bulk);
        return bulk;
    }
    if (bulk) bulk_g2s(dst, src, bulk, bar);
    return bulk;
}

struct TileSrc {
    const int64_t *time;       // n rows, 16B aligned
    const uint64_t *values;    // n rows, 16B aligned (may be null: bounds kernel)
    const uint8_t *validity;   // device bitmap at bit offset 0, padded to a 16B multiple (may be null)
    int64_t n;
};

// Called by ONE thread.  Stage layout: [time TIME_BYTES][values VAL_BYTES][bits BITS_STRIDE].
// (TIME_BYTES, VAL_BYTES and BITS_COPY are multiples of 16, so every destination is 16-byte aligned.)
template <class G, bool WITH_VALUES>
__device__ __forceinline__ void issue_tile(const TileSrc &src, int64_t tile, uint8_t *stage, uint64_t *bar) {
    const int64_t r0 = tile * G::T;
    const int64_t lo = r0 == 0 ? 0 : r0 - G::HALO;
    int64_t hi = r0 + G::T + 2;
    if (hi > src.n) hi = src.n;
    uint8_t *tdst = stage + (r0 == 0 ? G::HALO * 8 : 0);
    const uint32_t tbytes = (uint32_t)(hi - lo) * 8u;
    int64_t vhi = r0 + G::VAL_ENTRIES;
    if (vhi > src.n) vhi = src.n;
    const uint32_t vbytes = (uint32_t)(vhi - r0) * 8u;
    uint32_t bbytes = 0;
    if (WITH_VALUES && src.validity) {
        int64_t total = ((src.n + 7) / 8 + 15) & ~(int64_t)15;  // device bitmaps are padded to 16B
        int64_t b0 = r0 / 8;
        int64_t b1 = b0 + G::BITS_COPY;
        if (b1 > total) b1 = total;
        bbytes = (uint32_t)(b1 - b0);
    }
    // pass 1: plain tails + tx accounting, pass 2: bulk copies (after expect_tx)
    uint32_t tx = stage_bytes(tdst, src.time + lo, tbytes, bar, false);
    if (WITH_VALUES) tx += stage_bytes(stage + G::TIME_BYTES, src.values + r0, vbytes, bar, false);
    tx += bbytes;
    mbar_arrive_expect_tx(bar, tx);
    stage_bytes(tdst, src.time + lo, tbytes, bar, true);
    if (WITH_VALUES) stage_bytes(stage + G::TIME_BYTES, src.values + r0, vbytes, bar, true);
    if (bbytes) bulk_g2s(stage + G::TIME_BYTES + G::VAL_BYTES, src.validity + r0 / 8, bbytes, bar);
}

}  // namespace bowgpu
