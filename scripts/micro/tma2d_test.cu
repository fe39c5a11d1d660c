// standalone check of the 2-D TMA box load used by segreduce (tensor map [n/RE][RE] int64, box [NT][18])
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include "../../bow_b200/csrc/common.cuh"
using namespace bowgpu;
constexpr int RE = 64, NT = 128, COLS = 18;

__global__ void k(const __grid_constant__ CUtensorMap tm, int64_t *out, int phase, int tile) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm);
    int64_t *box = reinterpret_cast<int64_t *>(sm + 128);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, NT * COLS * 8);
        tma_box_2d(box, &tm, phase * 16, tile * NT, bar);
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < NT * COLS; i += blockDim.x) out[i] = box[i];
}

int main() {
    const int64_t n = 100000;
    int64_t *d, *o;
    cudaMalloc(&d, n * 8);
    cudaMalloc(&o, NT * COLS * 8);
    int64_t *h = new int64_t[n];
    for (int64_t i = 0; i < n; ++i) h[i] = i;
    cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    printf("entry point: %s qr=%d fn=%p\n", cudaGetErrorString(e), (int)qr, fn);
    auto encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    const cuuint64_t dims[2] = {RE, (cuuint64_t)(n / RE)};
    const cuuint64_t strides[1] = {RE * 8};
    const cuuint32_t box[2] = {COLS, NT};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_INT64, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    for (int phase = 0; phase < 4; ++phase) {
        k<<<1, 128, 128 + NT * COLS * 8>>>(tm, o, phase, 1);
        e = cudaDeviceSynchronize();
        printf("phase %d kernel: %s\n", phase, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        int64_t ho[NT * COLS];
        cudaMemcpy(ho, o, sizeof ho, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int t = 0; t < NT; ++t)
            for (int c = 0; c < COLS; ++c) {
                const int64_t col = phase * 16 + c;
                const int64_t want = col < RE ? (int64_t)(NT + t) * RE + col : 0;
                if (ho[t * COLS + c] != want && bad++ < 5) printf("  t=%d c=%d got %lld want %lld\n", t, c, (long long)ho[t * COLS + c], (long long)want);
            }
        printf("phase %d mismatches: %d\n", phase, bad);
    }
    return 0;
}
