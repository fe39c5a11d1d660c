// Microbenchmark: the data path of the segmc kernel alone — a producer warp issues 2-D TMA boxes [128][16] (128-byte
// swizzle) of two columns, alternating, into a ring of R boxes with full / empty mbarriers; four consumer warps wait,
// touch the box and release it.  NO compute.  Tiles are dealt to the CTAs round robin or in contiguous chunks.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a ring_bw.cu -o ring_bw
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include "../../bow_b200/csrc/common.cuh"
using namespace bowgpu;

__device__ __forceinline__ void arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int R, int NP, bool CHUNKED, int WORK>
__global__ void __launch_bounds__(160) ring_stream(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                                                   int64_t ntiles, int chunk, unsigned long long *sink) {
    constexpr int NT = 128, P = 16, BOX = NT * P * 8;
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t *full = reinterpret_cast<uint64_t *>(sm), *empty = full + 16;
    uint8_t *boxes = sm + 1024;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < R; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 4);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();
    int64_t t0, t1, ts;
    if (CHUNKED) {
        t0 = (int64_t)blockIdx.x * chunk, t1 = t0 + chunk, ts = 1;
        if (t1 > ntiles) t1 = ntiles;
    } else {
        t0 = blockIdx.x, t1 = ntiles, ts = gridDim.x;
    }
    if (warp == 4) {
        if (lane) return;
        uint32_t it = 0;
        for (int64_t tile = t0; tile < t1; tile += ts)
            for (int p = 0; p < NP; ++p)
                for (int c = 0; c < 2; ++c, ++it) {
                    const int s = it % R;
                    mbar_wait(&empty[s], ((it / R) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&full[s], BOX);
                    tma_box_2d(boxes + s * BOX, c ? &tm_b : &tm_a, p * P, (int32_t)(tile * NT), &full[s]);
                }
        return;
    }
    uint64_t acc = 0;
    uint32_t it = 0;
    for (int64_t tile = t0; tile < t1; tile += ts)
        for (int p = 0; p < NP; ++p)
            for (int c = 0; c < 2; ++c, ++it) {
                const int s = it % R;
                mbar_wait(&full[s], (it / R) & 1u);
                const ulonglong2 *v = reinterpret_cast<const ulonglong2 *>(boxes + s * BOX) + tid * 8;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const ulonglong2 x = v[q ^ (tid & 7)];
                    acc += x.x ^ x.y;
#pragma unroll
                    for (int w = 0; w < WORK; ++w) acc = acc * 6364136223846793005ull + x.x;  // dependent ALU work per pair
                }
                __syncwarp();
                if (lane == 0) arrive(&empty[s]);
            }
    if (acc == 0x1234567) *sink = acc;
}

static PFN_cuTensorMapEncodeTiled_v12000 enc;
static CUtensorMap make(const void *p, int64_t n, int RE) {
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)RE, (cuuint64_t)(n / RE)};
    const cuuint64_t strides[1] = {(cuuint64_t)RE * 8};
    const cuuint32_t box[2] = {16, 128};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_INT64, 2, const_cast<void *>(p), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_64B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
    return m;
}

template <int R, int NP, bool CHUNKED, int WORK>
void run(const int64_t *a, const int64_t *b, int64_t n, unsigned long long *sink, int ctas) {
    constexpr int RE = 16 * NP;
    const int smem = 1024 + R * 16384;
    auto k = ring_stream<R, NP, CHUNKED, WORK>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    CUtensorMap ta = make(a, n, RE), tb = make(b, n, RE);
    const int64_t ntiles = n / (128 * RE);
    const int grid = 148 * ctas;
    const int chunk = (int)((ntiles + grid - 1) / grid);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0);
        k<<<grid, 160, smem>>>(ta, tb, ntiles, chunk, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    printf("R=%d NP=%d %s work=%d ctas/SM=%d smem=%d: %.3f ms  %.0f GB/s (%s)\n", R, NP, CHUNKED ? "chunked" : "round-robin", WORK, ctas, smem,
           best, 16.0 * ntiles * 128 * RE / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    const int64_t n = 100000000;
    int64_t *a, *b;
    unsigned long long *sink;
    cudaMalloc(&a, n * 8 + 4096);
    cudaMalloc(&b, n * 8 + 4096);
    cudaMalloc(&sink, 8);
    cudaMemset(a, 1, n * 8);
    cudaMemset(b, 2, n * 8);
    run<6, 4, false, 0>(a, b, n, sink, 2);
    run<6, 4, true, 0>(a, b, n, sink, 2);
    run<4, 4, true, 0>(a, b, n, sink, 2);
    run<4, 4, true, 0>(a, b, n, sink, 3);
    run<2, 4, true, 0>(a, b, n, sink, 2);
    run<6, 2, true, 0>(a, b, n, sink, 2);
    run<6, 4, true, 4>(a, b, n, sink, 2);
    run<6, 4, true, 8>(a, b, n, sink, 2);
    run<6, 4, true, 16>(a, b, n, sink, 2);
    run<6, 4, false, 16>(a, b, n, sink, 2);
    run<4, 4, true, 16>(a, b, n, sink, 3);
    run<12, 4, true, 16>(a, b, n, sink, 1);
    return 0;
}
