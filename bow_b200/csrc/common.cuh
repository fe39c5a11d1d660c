// Shared device helpers for libbowgpu (sm_100a only): mbarrier + 1-D TMA bulk staging, exact
// 64-bit division by the window interval, small bit utilities.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ < 1000
#error "libbowgpu is written for sm_100a (B200) only"
#endif

namespace bowgpu {

// ---- mbarrier / cp.async.bulk (TMA 1-D bulk copy, no tensor map) -----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t tx_bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(tx_bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, completion signalled on `bar` (bytes: multiple of 16, both addresses 16B aligned)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// global -> shared 2-D tiled TMA copy (cp.async.bulk.tensor, SASS UTMALDG): box of the tensor map at element
// coordinates (c0 = innermost, c1); the whole box (out-of-bounds parts zero filled) counts towards complete_tx
__device__ __forceinline__ void tma_box_2d(void *dst_smem, const void *tmap, int32_t c0, int32_t c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst_smem)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

// ---- exact floor(x / d) for uint64 x and a launch-constant divisor d -------------------------
// Two rounds of a round-toward-zero double estimate (never above the true quotient) followed by a
// short correction loop.  inv_rd must be the double just below 1.0/d (host: nextafter(1.0/d, 0)).
struct DivU64 {
    uint64_t d;
    double inv_rd;
};
__host__ __device__ __forceinline__ uint64_t div_u64(uint64_t x, const DivU64 &dv) {
#ifdef __CUDA_ARCH__
    uint64_t q = __double2ull_rz(__dmul_rz(__ull2double_rz(x), dv.inv_rd));
    uint64_t r = x - q * dv.d;
    uint64_t q2 = __double2ull_rz(__dmul_rz(__ull2double_rz(r), dv.inv_rd));
    q += q2;
    r -= q2 * dv.d;
    while (r >= dv.d) {
        r -= dv.d;
        ++q;
    }
    return q;
#else
    return x / dv.d;
#endif
}

// Go int64(float64) on amd64 (CVTTSD2SI): NaN / out of range -> INT64_MIN (reference bowconvert.go:28-29)
__host__ __device__ __forceinline__ int64_t f64_to_i64_go(double x) {
    if (!(x >= -9223372036854775808.0 && x < 9223372036854775808.0)) return INT64_MIN;
    return (int64_t)x;
}

__device__ __forceinline__ double bits_as_f64(uint64_t b) { return __longlong_as_double((long long)b); }
__device__ __forceinline__ uint64_t f64_as_bits(double d) { return (uint64_t)__double_as_longlong(d); }

// 64-bit shuffles
__device__ __forceinline__ double shfl_up_f64(double v, int delta) { return __shfl_up_sync(0xffffffffu, v, delta); }
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int delta) {
    return (uint64_t)__shfl_up_sync(0xffffffffu, (unsigned long long)v, delta);
}

}  // namespace bowgpu
