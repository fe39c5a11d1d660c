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

// One thread per window, looping over the specs of the batch: the per-window counts shared by the
// aggregations of one input column are fetched from DRAM once (repeats hit L1).
template <int NS>
__global__ void __launch_bounds__(256) epilogue_kernel(const __grid_constant__ EpiBatch B, const int nspecs,
                                                       const WindowGeom g) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = k < g.W;
    const int lane = threadIdx.x & 31;
    // phase 1: every global read of this window, all in flight together (the pass is latency bound otherwise)
    int64_t cv[NS];
    double sv[NS];
    uint8_t okv[NS];
#pragma unroll
    for (int si = 0; si < NS; ++si) {
        cv[si] = 0;
        sv[si] = 0.0;
        okv[si] = 0;
        if (si < nspecs && in) {
            const EpilogueSpec &sp = B.s[si];
            if (sp.cnt) cv[si] = sp.cnt[k];
            if (sp.sum_src) sv[si] = sp.sum_src[k];
            if (sp.ok) okv[si] = sp.ok[k];
        }
    }
#pragma unroll
    for (int si = 0; si < NS; ++si) {
        if (si >= nspecs) break;
        const EpilogueSpec &sp = B.s[si];
        bool valid = false;
        if (in) {
            const int64_t c = cv[si];
            const bool okk = sp.ok ? okv[si] != 0 : c > 0;
            uint64_t *vals = reinterpret_cast<uint64_t *>(sp.values);
            const bool always = sp.op == BOWGPU_AGG_WINDOW_START || sp.op == BOWGPU_AGG_COUNT || sp.op == BOWGPU_AGG_SUM;
            valid = always || okk;
            bool have = false;  // value already computed in a register
            uint64_t bits = 0;
            if (sp.op == BOWGPU_AGG_WINDOW_START) {
                bits = (uint64_t)g.s0 + (uint64_t)k * g.div.d;
                have = true;
            } else if (sp.op == BOWGPU_AGG_COUNT) {
                bits = (uint64_t)c;
                have = true;
            } else if (!okk) {
                bits = 0;  // null slot, or Sum of an empty / all-null window = 0.0
                have = true;
            } else if (sp.op == BOWGPU_AGG_MEAN) {
                bits = f64_as_bits(__ddiv_rn(sv[si], (double)c));  // arithmeticmean.go:28
                have = true;
            } else if (sp.op == BOWGPU_AGG_WAVG_STEP || sp.op == BOWGPU_AGG_WAVG_LINEAR) {
                // integral / float64(w.LastValue - w.FirstValue), weightedmean.go:17,31
                bits = f64_as_bits(__ddiv_rn(sv[si], (double)(int64_t)g.div.d));
                have = true;
            }
            if (valid && sp.nfactors > 0) {
                if (!have) bits = vals[k];
                for (int i = 0; i < sp.nfactors; ++i) {
                    if (sp.out_is_int)
                        bits = (uint64_t)f64_to_i64_go(__dmul_rn((double)(int64_t)bits, sp.factors[i]));
                    else
                        bits = f64_as_bits(__dmul_rn(bits_as_f64(bits), sp.factors[i]));
                }
                have = true;
            }
            if (have) vals[k] = bits;
        }
        const uint32_t ball = __ballot_sync(0xffffffffu, valid);
        if ((lane & 7) == 0 && in) sp.validity[k >> 3] = (uint8_t)(ball >> lane);
    }
}

}  // namespace

int launch_epilogue(const EpilogueSpec *specs, int nspecs, WindowGeom g, cudaStream_t stream) {
    if (g.W <= 0) return 0;
    const int nt = 256;
    for (int b = 0; b < nspecs; b += EPI_MAX) {
        EpiBatch B;
        const int m = nspecs - b < EPI_MAX ? nspecs - b : EPI_MAX;
        for (int i = 0; i < m; ++i) B.s[i] = specs[b + i];
        const unsigned grid = (unsigned)((g.W + nt - 1) / nt);
        if (m <= 4)
            epilogue_kernel<4><<<grid, nt, 0, stream>>>(B, m, g);
        else if (m <= 8)
            epilogue_kernel<8><<<grid, nt, 0, stream>>>(B, m, g);
        else
            epilogue_kernel<EPI_MAX><<<grid, nt, 0, stream>>>(B, m, g);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

}  // namespace bowgpu
