// finish — what is left of Rolling.Aggregate after the streaming kernels (segmc.cuh) wrote the final values of every
// window that holds a row: the dispatcher of the segmc policies, the per-output validity bitmaps, the WindowStart
// columns (windowstart.go:8-13), value 0 in the slots of windows WITHOUT rows (bow.NewBuffer zero-fills,
// bowbuffer.go:22-40; Count = 0 and Sum = 0.0 are valid there, count.go:10, sum.go:11-13) and transformation.Factor
// (factor.go:7-20).  One thread per window, one ballot per 32 windows = one word of every bitmap.
#include "../../include/bowgpu.h"
#include "segmc.cuh"

namespace bowgpu {

int launch_segmc_b3(const McLaunch &, int, cudaStream_t, cudaEvent_t, cudaEvent_t);
int launch_segmc_b7(const McLaunch &, int, cudaStream_t, cudaEvent_t, cudaEvent_t);
int launch_segmc_i2(const McLaunch &, int, cudaStream_t, cudaEvent_t, cudaEvent_t);
int launch_segmc_i3(const McLaunch &, int, cudaStream_t, cudaEvent_t, cudaEvent_t);
int launch_segmc_all(const McLaunch &, int, cudaStream_t, cudaEvent_t, cudaEvent_t);

// The instantiated policies are supersets: a launch takes the smallest one that covers what its columns need (an
// output nobody asked for is a null pointer, the state behind it costs a few instructions per row).
int launch_segmc(const McLaunch &L, int sm_count, cudaStream_t stream, cudaEvent_t e0, cudaEvent_t e1) {
    if (L.iops == 0)
        return (L.bops & MC_FIRSTLAST) ? launch_segmc_b7(L, sm_count, stream, e0, e1) : launch_segmc_b3(L, sm_count, stream, e0, e1);
    if (L.bops == 0)
        return (L.iops & MC_STEP) ? launch_segmc_i3(L, sm_count, stream, e0, e1) : launch_segmc_i2(L, sm_count, stream, e0, e1);
    return launch_segmc_all(L, sm_count, stream, e0, e1);
}

int mc_max_chunks(int sm_count) { return sm_count * mc_knobs().ctas; }

namespace {

constexpr int FIN_MAX = 96;  // destinations per launch (kernel parameter space)
struct FinishBatch {
    FinishDst d[FIN_MAX];
};

__global__ void __launch_bounds__(256) finish_kernel(const __grid_constant__ FinishBatch B, const int ndst,
                                                     const uint32_t *__restrict__ touched, const WindowGeom g) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // word of 32 windows
    const int64_t k0 = w * 32, k = k0 + lane;
    if (k0 >= g.W) return;
    const bool in = k < g.W;
    const uint32_t in_mask = __ballot_sync(0xffffffffu, in);
    const uint32_t tw = touched[w];
    const bool empty = in && !((tw >> lane) & 1u);
    const bool whole_word = k0 + 32 <= g.W;
    for (int j = 0; j < ndst; ++j) {
        const FinishDst &D = B.d[j];
        uint32_t word = in_mask;
        if (D.kind == FIN_SRC) word = D.src[w] & in_mask;
        if (D.kind == FIN_WINDOW_START) {
            if (in) D.values[k] = (uint64_t)window_first_value(g, k);
        } else if (empty && D.zero_empty) {
            D.values[k] = 0;
        }
        if (lane == (j & 31)) {  // a partial last word is stored bytewise: nothing beyond ceil(W/8) bytes is touched
            if (whole_word)
                *reinterpret_cast<uint32_t *>(D.validity + (k0 >> 3)) = word;
            else
                for (int64_t b = 0; k0 + 8 * b < g.W; ++b) D.validity[(k0 >> 3) + b] = (uint8_t)(word >> (8 * b));
        }
    }
}

__global__ void __launch_bounds__(256) factor_kernel(uint64_t *values, const uint8_t *validity, const int64_t W,
                                                     const int out_is_int, const int nf, const double f0, const double f1,
                                                     const double f2, const double f3) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= W || !((validity[k >> 3] >> (k & 7)) & 1)) return;  // nil passes through (factor.go:9-11)
    const double f[4] = {f0, f1, f2, f3};
    uint64_t bits = values[k];
    for (int i = 0; i < nf; ++i) {
        if (out_is_int)
            bits = (uint64_t)f64_to_i64_go(__dmul_rn((double)(int64_t)bits, f[i]));  // int64(float64(x) * n), factor.go:15-16
        else
            bits = f64_as_bits(__dmul_rn(bits_as_f64(bits), f[i]));
    }
    values[k] = bits;
}

}  // namespace

int launch_finish(const FinishDst *dst, int ndst, const uint32_t *touched, WindowGeom g, cudaStream_t stream, int *launches) {
    if (g.W <= 0) return 0;
    const int64_t threads = (g.W + 31) / 32 * 32;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    for (int b = 0; b < ndst; b += FIN_MAX) {
        FinishBatch B;
        const int m = ndst - b < FIN_MAX ? ndst - b : FIN_MAX;
        for (int i = 0; i < m; ++i) B.d[i] = dst[b + i];
        finish_kernel<<<grid, 256, 0, stream>>>(B, m, touched, g);
        if (launches) ++*launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

int launch_factor(uint64_t *values, const uint8_t *validity, int64_t W, int out_is_int, int nfactors, const double *factors,
                  cudaStream_t stream) {
    if (W <= 0 || nfactors <= 0) return 0;
    double f[4] = {1, 1, 1, 1};
    for (int i = 0; i < nfactors && i < 4; ++i) f[i] = factors[i];
    factor_kernel<<<(unsigned)((W + 255) / 256), 256, 0, stream>>>(values, validity, W, out_is_int, nfactors, f[0], f[1], f[2], f[3]);
    return (int)cudaGetLastError();
}

}  // namespace bowgpu
