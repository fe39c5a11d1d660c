// Microbenchmark: DRAM bandwidth when every THREAD streams its own contiguous chunk (128-byte granules, many
// concurrent streams) versus the classic CTA-contiguous pattern.  Decides whether thread-persistent row streams
// are viable for segreduce.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a stream_bw.cu -o stream_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// each thread owns chunk_bytes contiguous bytes; per step it reads `gran` bytes (as 16B vectors) and advances
template <int GRAN>
__global__ void per_thread_streams(const uint4 *__restrict__ src, size_t chunk_bytes, size_t nchunks, unsigned long long *sink) {
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    uint64_t acc = 0;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += nthreads) {
        const uint4 *p = src + c * (chunk_bytes / 16);
        for (size_t off = 0; off < chunk_bytes / 16; off += GRAN / 16) {
            uint4 v[GRAN / 16];
#pragma unroll
            for (int i = 0; i < GRAN / 16; ++i) v[i] = __ldcs(p + off + i);
#pragma unroll
            for (int i = 0; i < GRAN / 16; ++i) acc += v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
        }
    }
    if (acc == 0x1234567) *sink = acc;
}

__global__ void coalesced(const uint4 *__restrict__ src, size_t n16, unsigned long long *sink) {
    uint64_t acc = 0;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i + 3 * nthreads < n16; i += 4 * nthreads) {
        uint4 a = __ldcs(src + i), b = __ldcs(src + i + nthreads), c = __ldcs(src + i + 2 * nthreads), d = __ldcs(src + i + 3 * nthreads);
        acc += a.x ^ b.y ^ c.z ^ d.w;
    }
    if (acc == 0x1234567) *sink = acc;
}

int main() {
    const size_t bytes = (size_t)3200 << 20;  // 3.2 GB
    uint4 *d;
    unsigned long long *sink;
    cudaMalloc(&d, bytes);
    cudaMalloc(&sink, 8);
    cudaMemset(d, 1, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto time_it = [&](auto launch, const char *name) {
        launch();
        cudaDeviceSynchronize();
        float best = 1e9;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf("%-46s %8.3f ms  %7.1f GB/s  (%s)\n", name, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    time_it([&] { coalesced<<<148 * 8, 256>>>(d, bytes / 16, sink); }, "coalesced 16B loads, 4 in flight");
    for (size_t chunk : {2048, 7168, 16384, 65536}) {
        for (int ctas : {3, 6, 12}) {
            char name[128];
            const size_t nch = bytes / chunk;
            snprintf(name, sizeof name, "thread streams chunk=%zuB gran=128 ctas/SM=%d", chunk, ctas);
            time_it([&] { per_thread_streams<128><<<148 * ctas, 128>>>(d, chunk, nch, sink); }, name);
            snprintf(name, sizeof name, "thread streams chunk=%zuB gran=256 ctas/SM=%d", chunk, ctas);
            time_it([&] { per_thread_streams<256><<<148 * ctas, 128>>>(d, chunk, nch, sink); }, name);
        }
    }
    return 0;
}
