// Microbenchmark: pure staging bandwidth of the segreduce access pattern — per tile NP 2-D TMA boxes [NT][P+2] of a
// [n/RE][RE] int64 view, double buffered, persistent CTAs, NO compute.  Tells whether the kernel's 5.95 TB/s plateau is the
// memory system's ceiling for this pattern.   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a box_bw.cu -o box_bw
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include "../../bow_b200/csrc/common.cuh"
using namespace bowgpu;

template <int NT, int P, int NP, int PAD>
__global__ void __launch_bounds__(NT) box_stream(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                                                 int64_t ntiles, unsigned long long *sink) {
    constexpr int COLS = P + PAD, BOX = NT * COLS * 8, SLOT = 2 * BOX, NS = 2;
    extern __shared__ __align__(128) uint8_t sm[];
    uint64_t *full = reinterpret_cast<uint64_t *>(sm);
    uint8_t *slots = sm + 1024;
    const int tid = threadIdx.x;
    auto issue = [&](int64_t it, int phase, int s) {
        const int64_t tile = blockIdx.x + it * (int64_t)gridDim.x;
        if (tile >= ntiles) return;
        mbar_arrive_expect_tx(&full[s], SLOT);
        tma_box_2d(slots + (size_t)s * SLOT, &tm_a, phase * P, (int32_t)(tile * NT), &full[s]);
        tma_box_2d(slots + (size_t)s * SLOT + BOX, &tm_b, phase * P, (int32_t)(tile * NT), &full[s]);
    };
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();
    if (tid == 0)
        for (int q = 0; q < NS; ++q) issue(q / NP, q % NP, q);
    uint32_t parity = 0;
    uint64_t acc = 0;
    int s = 0;
    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it)
        for (int phase = 0; phase < NP; ++phase) {
            mbar_wait(&full[s], (parity >> s) & 1u);
            parity ^= 1u << s;
            acc += reinterpret_cast<const uint64_t *>(slots + (size_t)s * SLOT)[tid * COLS];  // touch the data
            __syncthreads();
            if (tid == 0) {
                const int ahead = phase + NS;
                issue(it + ahead / NP, ahead % NP, s);
            }
            s ^= 1;
        }
    if (acc == 0x1234567) *sink = acc;
}

static PFN_cuTensorMapEncodeTiled_v12000 enc;
static CUtensorMap make(const void *p, int64_t n, int RE, int COLS, int NT, bool swz) {
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)RE, (cuuint64_t)(n / RE)};
    const cuuint64_t strides[1] = {(cuuint64_t)RE * 8};
    const cuuint32_t box[2] = {(cuuint32_t)COLS, (cuuint32_t)NT};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_INT64, 2, const_cast<void *>(p), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_64B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
    return m;
}

template <int NT, int P, int NP, int PAD = 2>
void run(const int64_t *a, const int64_t *b, int64_t n, unsigned long long *sink, int ctas) {
    constexpr int RE = P * NP, COLS = P + PAD, SLOT = 2 * NT * COLS * 8;
    const int smem = 1024 + 2 * SLOT;
    auto k = box_stream<NT, P, NP, PAD>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    CUtensorMap ta = make(a, n, RE, COLS, NT, PAD == 0), tb = make(b, n, RE, COLS, NT, PAD == 0);
    const int64_t ntiles = n / ((int64_t)NT * RE);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0);
        k<<<148 * ctas, NT, smem>>>(ta, tb, ntiles, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    printf("PAD=%d NT=%d P=%d NP=%d ctas/SM=%d smem=%d: %.3f ms  %.0f GB/s useful (%s)\n", PAD, NT, P, NP, ctas, smem, best,
           16.0 * ntiles * NT * RE / best / 1e6, cudaGetErrorString(cudaGetLastError()));
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
    run<128, 16, 4>(a, b, n, sink, 3);
    run<128, 16, 4, 0>(a, b, n, sink, 3);
    run<128, 16, 4, 0>(a, b, n, sink, 2);
    run<128, 16, 2, 0>(a, b, n, sink, 3);
    run<256, 16, 4, 0>(a, b, n, sink, 1);
    run<128, 16, 2>(a, b, n, sink, 3);
    run<128, 16, 8>(a, b, n, sink, 3);
    run<128, 32, 2>(a, b, n, sink, 1);
    run<64, 32, 2>(a, b, n, sink, 3);
    run<64, 32, 4>(a, b, n, sink, 3);
    run<128, 16, 4>(a, b, n, sink, 2);
    run<256, 16, 4>(a, b, n, sink, 1);
    return 0;
}
