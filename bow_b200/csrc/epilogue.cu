// epilogue — per-window finishing pass of Rolling.Aggregate (W threads, trivially balanced).
//
// The streaming kernels only write VALUES of windows that hold at least one valid row.  This pass
// produces what bow.NewBuffer + Buffer.SetOrDrop leave behind in the reference
// (rolling/aggregation.go:198,223-227, bowbuffer.go:22-80): validity bitmaps (LSB first, unused
// trailing bits 0), value 0 in null slots, Count = 0 / Sum = 0.0 for empty windows (count.go:10,
// sum.go:11-13), the WindowStart column (windowstart.go:8-13) and transformation.Factor
// (factor.go:7-20: float64 x*n, int64 int64(float64(x)*n), nil passes through).
#include "../../include/bowgpu.h"
#include "kernels.h"

namespace bowgpu {

namespace {

constexpr int EPI_MAX = 16;
struct EpiBatch {
    EpilogueSpec s[EPI_MAX];
};

// Each warp owns 128 consecutive windows; lane l handles windows base + l + 32*i (i = 0..3): coalesced 8-byte
// accesses with four independent loads in flight per array, and one ballot per i yields a whole 32-bit word of
// the validity bitmap.  Specs are walked in a rolled loop; the per-window count array shared by the
// aggregations of one input column is fetched once (consecutive specs with the same pointer reuse it).
#ifndef EPI_CFG_NT
#define EPI_CFG_NT 256
#endif
#ifndef EPI_CFG_WPT
#define EPI_CFG_WPT 4
#endif
constexpr int EPI_NT = EPI_CFG_NT, EPI_WPT = EPI_CFG_WPT;  // threads per block, windows per thread
#ifndef EPI_CFG_GROUP_WPT
#define EPI_CFG_GROUP_WPT 1
#endif
// the grouped pass is a handful of loads and stores per window: one window per thread keeps the most loads in flight
// (configs[1]: 1.67 M windows, step 0.3132 -> 0.3105 ms against four windows per thread)
constexpr int EPIG_WPT = EPI_CFG_GROUP_WPT;

__global__ void __launch_bounds__(EPI_NT) epilogue_kernel(const __grid_constant__ EpiBatch B, const int nspecs,
                                                          const WindowGeom g) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = ((int64_t)blockIdx.x * (EPI_NT / 32) + warp) * (32 * EPI_WPT);
    if (base >= g.W) return;
    int64_t k[EPI_WPT];
    bool in[EPI_WPT];
#pragma unroll
    for (int i = 0; i < EPI_WPT; ++i) {
        k[i] = base + lane + 32 * i;
        in[i] = k[i] < g.W;
    }
    const int64_t *last_cnt = nullptr;
    int64_t c[EPI_WPT] = {};
    for (int si = 0; si < nspecs; ++si) {
        const EpilogueSpec &sp = B.s[si];
        uint64_t *vals = reinterpret_cast<uint64_t *>(sp.values);
        if (sp.cnt != last_cnt) {
#pragma unroll
            for (int i = 0; i < EPI_WPT; ++i) c[i] = (sp.cnt && in[i]) ? sp.cnt[k[i]] : 0;
            last_cnt = sp.cnt;
        }
        double sv[EPI_WPT];
        uint8_t okv[EPI_WPT];
#pragma unroll
        for (int i = 0; i < EPI_WPT; ++i) {
            sv[i] = (sp.sum_src && in[i]) ? sp.sum_src[k[i]] : 0.0;
            okv[i] = (sp.ok && in[i]) ? sp.ok[k[i]] : 0;
        }
        const bool always = sp.op == BOWGPU_AGG_WINDOW_START || sp.op == BOWGPU_AGG_COUNT || sp.op == BOWGPU_AGG_SUM;
        const bool count_in_place = sp.op == BOWGPU_AGG_COUNT && sp.nfactors == 0 && sp.values == (const void *)sp.cnt;
#pragma unroll
        for (int i = 0; i < EPI_WPT; ++i) {
            bool valid = false;
            if (in[i]) {
                const bool okk = sp.ok ? okv[i] != 0 : c[i] > 0;
                valid = always || okk;
                bool have = false;  // value already computed in a register
                uint64_t bits = 0;
                if (sp.op == BOWGPU_AGG_WINDOW_START) {
                    bits = (uint64_t)window_first_value(g, k[i]);
                    have = true;
                } else if (sp.op == BOWGPU_AGG_COUNT) {
                    bits = (uint64_t)c[i];
                    have = !count_in_place;  // the segreduce kernel wrote the counts straight into this output
                } else if (!okk) {
                    bits = 0;  // null slot, or Sum of an empty / all-null window = 0.0
                    have = true;
                } else if (sp.op == BOWGPU_AGG_MEAN) {
                    bits = f64_as_bits(__ddiv_rn(sv[i], (double)c[i]));  // arithmeticmean.go:28
                    have = true;
                } else if (sp.op == BOWGPU_AGG_WAVG_STEP || sp.op == BOWGPU_AGG_WAVG_LINEAR) {
                    // integral / float64(w.LastValue - w.FirstValue), weightedmean.go:17,31
                    bits = f64_as_bits(__ddiv_rn(sv[i], g.whole ? (double)(g.whole_last - g.whole_first) : (double)(int64_t)g.div.d));
                    have = true;
                }
                if (valid && sp.nfactors > 0) {
                    if (!have) bits = vals[k[i]];
                    for (int f = 0; f < sp.nfactors; ++f) {
                        if (sp.out_is_int)
                            bits = (uint64_t)f64_to_i64_go(__dmul_rn((double)(int64_t)bits, sp.factors[f]));
                        else
                            bits = f64_as_bits(__dmul_rn(bits_as_f64(bits), sp.factors[f]));
                    }
                    have = true;
                }
                if (have) vals[k[i]] = bits;
            }
            // one ballot = the 32 validity bits of windows base + 32 i .. + 31; the last word is stored bytewise so that
            // nothing beyond ceil(W/8) bytes is touched (the caller's buffer ends there)
            const uint32_t ball = __ballot_sync(0xffffffffu, valid);
            const int64_t w0 = base + 32 * i;
            if (lane == 0 && w0 < g.W) {
                if (w0 + 32 <= g.W) {
                    *reinterpret_cast<uint32_t *>(sp.validity + (w0 >> 3)) = ball;
                } else {
                    for (int64_t b = 0; w0 + 8 * b < g.W; ++b) sp.validity[(w0 >> 3) + b] = (uint8_t)(ball >> (8 * b));
                }
            }
        }
    }
}

__global__ void __launch_bounds__(EPI_NT) epilogue_group_kernel(const __grid_constant__ EpiGroup G, const WindowGeom g) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = ((int64_t)blockIdx.x * (EPI_NT / 32) + warp) * (32 * EPIG_WPT);
    if (base >= g.W) return;
    int64_t c[EPIG_WPT];
#pragma unroll
    for (int i = 0; i < EPIG_WPT; ++i) {
        const int64_t k = base + lane + 32 * i;
        c[i] = (G.cnt && k < g.W) ? G.cnt[k] : 0;
    }
    double sv[EPIG_DIV][EPIG_WPT];
#pragma unroll
    for (int j = 0; j < EPIG_DIV; ++j)
#pragma unroll
        for (int i = 0; i < EPIG_WPT; ++i) {
            const int64_t k = base + lane + 32 * i;
            sv[j][i] = (j < G.n_div && k < g.W) ? G.div_src[j][k] : 0.0;  // (independent of the count load: both in flight;
                                                                         // slots of windows without valid rows were never written - initcheck
                                                                         // reports the read - and are discarded below: valid == false)
        }
    // float64(w.LastValue - w.FirstValue), weightedmean.go:17,31 (the interval, except for the whole-Bow window)
    bool by_width = false;
#pragma unroll
    for (int j = 0; j < EPIG_DIV; ++j) by_width |= j < G.n_div && !G.div_by_cnt[j];
    double width = 1.0;  // (64-bit integer -> double conversions are not cheap: only when an output divides by the interval)
    if (by_width) width = g.whole ? (double)(g.whole_last - g.whole_first) : (double)(int64_t)g.div.d;
#pragma unroll
    for (int i = 0; i < EPIG_WPT; ++i) {
        const int64_t k = base + lane + 32 * i;
        const bool in = k < g.W;
        const bool valid = in && c[i] > 0;
        if (in) {
            if (!valid)
                for (int j = 0; j < G.n_null; ++j) G.null_vals[j][k] = 0;  // null slot; Sum of an empty window = 0.0
#pragma unroll
            for (int j = 0; j < EPIG_DIV; ++j)
                if (j < G.n_div)  // arithmeticmean.go:28 / weightedmean.go:17,31; null slots hold 0
                    G.div_dst[j][k] = valid ? __ddiv_rn(sv[j][i], G.div_by_cnt[j] ? (double)c[i] : width) : 0.0;
            for (int j = 0; j < G.n_ws; ++j) G.ws[j][k] = window_first_value(g, k);
        }
        // one ballot = the 32 validity bits of windows w0 .. w0 + 31; a partial last word is stored bytewise so that
        // nothing beyond ceil(W/8) bytes is touched
        const uint32_t ball = __ballot_sync(0xffffffffu, valid);
        const uint32_t ball_in = __ballot_sync(0xffffffffu, in);
        const int64_t w0 = base + 32 * i;
        if (lane < G.n_bm_cnt + G.n_bm_all && w0 < g.W) {  // lane j stores the word of bitmap j
            const bool whole = w0 + 32 <= g.W;
            uint8_t *bm = lane < G.n_bm_cnt ? G.bm_cnt[lane] : G.bm_all[lane - G.n_bm_cnt];
            const uint32_t word = lane < G.n_bm_cnt ? ball : ball_in;
            if (whole)
                *reinterpret_cast<uint32_t *>(bm + (w0 >> 3)) = word;
            else
                for (int64_t b = 0; w0 + 8 * b < g.W; ++b) bm[(w0 >> 3) + b] = (uint8_t)(word >> (8 * b));
        }
    }
}

}  // namespace

int launch_epilogue_group(const EpiGroup &G, WindowGeom g, cudaStream_t stream) {
    if (g.W <= 0) return 0;
    const int64_t per_block = (int64_t)EPI_NT * EPIG_WPT;
    epilogue_group_kernel<<<(unsigned)((g.W + per_block - 1) / per_block), EPI_NT, 0, stream>>>(G, g);
    return (int)cudaGetLastError();
}

int launch_epilogue(const EpilogueSpec *specs, int nspecs, WindowGeom g, cudaStream_t stream) {
    if (g.W <= 0) return 0;
    const int64_t per_block = (int64_t)EPI_NT * EPI_WPT;
    const unsigned grid = (unsigned)((g.W + per_block - 1) / per_block);
    for (int b = 0; b < nspecs; b += EPI_MAX) {
        EpiBatch B;
        const int m = nspecs - b < EPI_MAX ? nspecs - b : EPI_MAX;
        for (int i = 0; i < m; ++i) B.s[i] = specs[b + i];
        epilogue_kernel<<<grid, EPI_NT, 0, stream>>>(B, m, g);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

}  // namespace bowgpu
