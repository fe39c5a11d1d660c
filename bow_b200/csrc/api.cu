// C ABI of libbowgpu.so (include/bowgpu.h): contexts, device-resident frames, the rolling object and
// the Aggregate / bounds drivers.  Host logic only — every row of data is touched by CUDA kernels.
#include "../../include/bowgpu.h"

#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "kernels.h"
#include "parquet.h"

using namespace bowgpu;

// ================================================================================================
// internal objects
// ================================================================================================
struct bowgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    std::string err;
    // error flags written by kernels (ST_*), sticky until read
    int32_t *d_status = nullptr;
    int64_t *d_scalars = nullptr;  // 8 device int64 for tiny results
    // scratch arena (grow only, bump allocated per call)
    uint8_t *arena = nullptr;
    size_t arena_cap = 0, arena_top = 0;
    // stream-ordered pool for column buffers: freed blocks stay cached (release threshold = max), so the
    // Interpolate -> Aggregate chain does not pay cudaMalloc / cudaFree of multi-GB columns on every call
    cudaMemPool_t pool = nullptr;
    // pinned staging for pageable host memory
    uint8_t *pinned[2] = {nullptr, nullptr};
    cudaEvent_t pinned_ev[2] = {nullptr, nullptr};
    size_t pinned_bytes = 0;
    // worker contexts of bowgpu_aggregate_host (own stream, arena, pool each), created on first use
    std::vector<bowgpu_ctx *> workers;
    // side streams of Rolling.Aggregate: the streaming launches of several value columns run side by side, each on its
    // share of the SMs, so that the time tiles they all read are fetched from DRAM once and served from L2 after that
    // staging of file reads (bowgpu_parquet_read): reader threads pread() straight into pinned chunks
    std::vector<uint8_t *> file_stage;
    std::vector<cudaEvent_t> file_stage_ev;
    std::vector<cudaStream_t> side;
    std::vector<cudaEvent_t> side_done;
    cudaEvent_t side_fork = nullptr;
    // timing
    int timing = 0;  // 0 off, 1 per call, 2 accumulate over calls
    cudaEvent_t ev_total[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_main;  // pairs
    int ev_main_used = 0;
    unsigned main_seq = 0;  // mode 2: every 4th main launch is bracketed by events
    int launches = 0, main_launches = 0;
};

struct DevCol {
    uint64_t *values = nullptr;
    uint8_t *validity = nullptr;  // bit offset 0, padded to 16 bytes, padding zero; null = all valid
    int32_t dtype = 0;
    int64_t null_count = 0;
    bool own_values = false, own_validity = false;
};

struct bowgpu_frame {
    bowgpu_ctx *ctx = nullptr;
    int64_t n = 0;
    std::vector<DevCol> cols;
};

struct PrevCell {
    uint64_t bits = 0;
    int32_t valid = 0;
    int32_t dtype = 0;
};

struct bowgpu_rolling {
    bowgpu_frame *frame = nullptr;
    int32_t time_col = 0;
    int64_t interval = 0, offset = 0;
    int32_t inclusive = 0;
    int64_t s0 = 0, W = 0;
    int64_t t_first = 0, t_last = 0;
    int64_t early_rows = 0;  // rows with t < s0
    int64_t t_after_early = 0;
    bool has_after_early = false;
    bool has_prev = false;
    bool shard = false;  // range-partitioned shard: rows before s0 are the left halo (never part of a window)
    bool whole = false;  // aggregation.Aggregate over the whole Bow: one window, see WindowGeom
    int64_t whole_first = 0, whole_last = 0;
    std::vector<PrevCell> prev;
};

namespace {

struct Guard {  // selects the ctx device for the duration of a call (cgo calls land on any OS thread)
    int prev = -1;
    explicit Guard(const bowgpu_ctx *c) {
        cudaGetDevice(&prev);
        if (prev != c->device) cudaSetDevice(c->device);
    }
    ~Guard() {}
};

int32_t fail(bowgpu_ctx *c, int32_t code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CK(call)                                                                                                  \
    do {                                                                                                          \
        cudaError_t e__ = (cudaError_t)(call);                                                                    \
        if (e__ != cudaSuccess)                                                                                   \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? BOWGPU_ENOMEM : BOWGPU_ECUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                                             \
    } while (0)

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

cudaError_t pool_alloc(bowgpu_ctx *ctx, void **p, size_t bytes) {
    return cudaMallocFromPoolAsync(p, bytes, ctx->pool, ctx->stream);
}
void pool_free(bowgpu_ctx *ctx, void *p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}
int64_t bitmap_bytes_padded(int64_t n) { return (int64_t)align_up((size_t)((n + 7) / 8), 16) + 16; }

// ---- scratch arena ---------------------------------------------------------------------------------
int32_t arena_reserve(bowgpu_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->arena_cap) return BOWGPU_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->arena) cudaFree(ctx->arena);
    ctx->arena = nullptr;
    ctx->arena_cap = 0;
    size_t cap = align_up(bytes + (bytes >> 3), (size_t)1 << 20);
    CK(cudaMalloc(&ctx->arena, cap));
    ctx->arena_cap = cap;
    return BOWGPU_OK;
}
void arena_reset(bowgpu_ctx *ctx) { ctx->arena_top = 0; }
void *arena_take(bowgpu_ctx *ctx, size_t bytes) {
    size_t off = align_up(ctx->arena_top, 256);
    if (off + bytes > ctx->arena_cap) return nullptr;
    ctx->arena_top = off + bytes;
    return ctx->arena + off;
}

// ---- host <-> device copies (complete before return) -------------------------------------------------
bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int32_t ensure_pinned(bowgpu_ctx *ctx) {
    if (ctx->pinned[0]) return BOWGPU_OK;
    ctx->pinned_bytes = (size_t)32 << 20;
    for (int i = 0; i < 2; ++i) {
        CK(cudaHostAlloc((void **)&ctx->pinned[i], ctx->pinned_bytes, cudaHostAllocDefault));
        CK(cudaEventCreateWithFlags(&ctx->pinned_ev[i], cudaEventDisableTiming));
    }
    return BOWGPU_OK;
}

// Pageable sources go through two pinned chunks (CPU memcpy of chunk i+1 overlaps the DMA of chunk i);
// pinned sources are DMA'd directly.  The caller synchronizes the stream before returning to Go.
int32_t copy_h2d(bowgpu_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return BOWGPU_OK;
    if (is_pinned(src)) {
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return BOWGPU_OK;
    }
    int32_t rc = ensure_pinned(ctx);
    if (rc) return rc;
    size_t done = 0;
    int b = 0;
    while (done < bytes) {
        size_t m = std::min(ctx->pinned_bytes, bytes - done);
        CK(cudaEventSynchronize(ctx->pinned_ev[b]));
        memcpy(ctx->pinned[b], (const char *)src + done, m);
        CK(cudaMemcpyAsync((char *)dst + done, ctx->pinned[b], m, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->pinned_ev[b], ctx->stream));
        done += m;
        b ^= 1;
    }
    return BOWGPU_OK;
}

int32_t copy_d2h(bowgpu_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return BOWGPU_OK;
    if (is_pinned(dst)) {
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        return BOWGPU_OK;
    }
    int32_t rc = ensure_pinned(ctx);
    if (rc) return rc;
    // ping-pong: DMA chunk i+1 while the CPU copies chunk i out of the pinned buffer
    size_t off = 0, prev_off = 0, prev_len = 0;
    int b = 0, prev_b = -1;
    while (off < bytes) {
        const size_t m = std::min(ctx->pinned_bytes, bytes - off);
        CK(cudaMemcpyAsync(ctx->pinned[b], (const char *)src + off, m, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaEventRecord(ctx->pinned_ev[b], ctx->stream));
        if (prev_b >= 0) {
            CK(cudaEventSynchronize(ctx->pinned_ev[prev_b]));
            memcpy((char *)dst + prev_off, ctx->pinned[prev_b], prev_len);
        }
        prev_b = b;
        prev_off = off;
        prev_len = m;
        off += m;
        b ^= 1;
    }
    if (prev_b >= 0) {
        CK(cudaEventSynchronize(ctx->pinned_ev[prev_b]));
        memcpy((char *)dst + prev_off, ctx->pinned[prev_b], prev_len);
    }
    return BOWGPU_OK;
}

void count_launch(bowgpu_ctx *ctx, int n = 1, bool main = false) {
    ctx->launches += n;
    if (main) ctx->main_launches += n;
}

int32_t check_status(bowgpu_ctx *ctx) {  // stream must be synchronized by the caller's copy
    int32_t st = 0;
    CK(cudaMemcpyAsync(&st, ctx->d_status, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (st) {
        CK(cudaMemsetAsync(ctx->d_status, 0, 4, ctx->stream));
        if (st & ST_UNSORTED) return fail(ctx, BOWGPU_EUNSORTED, "time column is not sorted ascending");
    }
    return BOWGPU_OK;
}

void timing_begin(bowgpu_ctx *ctx) {
    if (ctx->timing != 2) {  // mode 2 accumulates over calls until bowgpu_ctx_last_timing
        ctx->launches = 0;
        ctx->main_launches = 0;
        ctx->ev_main_used = 0;
    }
    if (ctx->timing == 1) cudaEventRecord(ctx->ev_total[0], ctx->stream);
}
void timing_end(bowgpu_ctx *ctx) {
    if (ctx->timing == 1) cudaEventRecord(ctx->ev_total[1], ctx->stream);
}
// returns a pair of events for one main-kernel launch (null when timing is off)
void timing_main_pair(bowgpu_ctx *ctx, cudaEvent_t *e0, cudaEvent_t *e1) {
    *e0 = *e1 = nullptr;
    if (!ctx->timing) return;
    // mode 2 runs inside timed loops: four event records per call cost 11 us of a 0.31 ms step (measured), so only
    // every 4th main launch is bracketed and the call itself is not
    if (ctx->timing == 2 && (ctx->main_seq++ & 3u)) return;
    if ((size_t)ctx->ev_main_used + 2 > ctx->ev_main.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        ctx->ev_main.push_back(a);
        ctx->ev_main.push_back(b);
    }
    *e0 = ctx->ev_main[ctx->ev_main_used++];
    *e1 = ctx->ev_main[ctx->ev_main_used++];
}

WindowGeom make_geom(const bowgpu_rolling *r, bool inclusive_eff) {
    WindowGeom g;
    g.n = r->frame->n;
    g.s0 = r->s0;
    g.W = r->W;
    g.div.d = (uint64_t)r->interval;
    g.div.inv_rd = std::nextafter(1.0 / (double)r->interval, 0.0);
    g.early_rows = r->early_rows;
    // rows before s0 stay in window 0 iff that window holds a row in [S0, E0) — or exactly at E0 when the
    // iteration is inclusive (rolling.go:194-211: they are only part of the slice if lastRowIndex is set)
    g.early_keep = 0;
    g.shard = r->shard;
    g.whole = r->whole;
    g.whole_first = r->whole_first;
    g.whole_last = r->whole_last;
    g._pad = 0;
    if (!r->shard && r->early_rows > 0 && r->has_after_early) {
        const uint64_t rel = (uint64_t)r->t_after_early - (uint64_t)r->s0;
        g.early_keep = rel < (uint64_t)r->interval || (inclusive_eff && rel == (uint64_t)r->interval);
    }
    return g;
}

bool agg_is_basic(int op) { return op >= BOWGPU_AGG_COUNT && op <= BOWGPU_AGG_LAST; }
bool agg_is_integral(int op) { return op >= BOWGPU_AGG_INTEGRAL_STEP && op <= BOWGPU_AGG_WAVG_LINEAR; }

}  // namespace

// ================================================================================================
// context
// ================================================================================================
extern "C" int32_t bowgpu_abi_version(void) { return BOWGPU_ABI_VERSION; }

extern "C" const char *bowgpu_status_string(int32_t s) {
    switch (s) {
    case BOWGPU_OK: return "ok";
    case BOWGPU_EINVAL: return "invalid argument";
    case BOWGPU_ETYPE: return "unsupported column type";
    case BOWGPU_EFIRSTNULL: return "first value of the interval column is null";
    case BOWGPU_EPREVROW: return "prevRow must have only one row";
    case BOWGPU_ENOINTERVALCOL: return "must keep interval column";
    case BOWGPU_ECAPACITY: return "output capacity too small";
    case BOWGPU_EUNSORTED: return "interval column is not sorted";
    case BOWGPU_ENULLTIME: return "interval column holds nulls";
    case BOWGPU_ECUDA: return "CUDA error";
    case BOWGPU_ENOMEM: return "out of memory";
    case BOWGPU_EUNSUPPORTED: return "not supported by the GPU backend";
    case BOWGPU_EIO: return "file cannot be read or is malformed";
    }
    return "unknown status";
}

extern "C" int32_t bowgpu_ctx_create(int32_t device, void *stream, bowgpu_ctx **out) {
    if (!out) return BOWGPU_EINVAL;
    *out = nullptr;
    bowgpu_ctx *ctx = new (std::nothrow) bowgpu_ctx();
    if (!ctx) return BOWGPU_ENOMEM;
    ctx->device = device;
    auto bail = [&](int32_t code) {
        delete ctx;
        return code;
    };
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return bail(BOWGPU_ECUDA);
    }
    if (cudaSetDevice(device) != cudaSuccess) return bail(BOWGPU_ECUDA);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(BOWGPU_ECUDA);
    if (prop.major != 10) return bail(BOWGPU_EUNSUPPORTED);  // kernels are built for sm_100a only
    ctx->sm_count = prop.multiProcessorCount;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(BOWGPU_ECUDA);
        ctx->own_stream = true;
    }
    {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof props);
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        if (cudaMemPoolCreate(&ctx->pool, &props) != cudaSuccess) return bail(BOWGPU_ECUDA);
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    if (cudaMalloc(&ctx->d_status, 256) != cudaSuccess) return bail(BOWGPU_ENOMEM);
    ctx->d_scalars = (int64_t *)((char *)ctx->d_status + 64);
    cudaMemsetAsync(ctx->d_status, 0, 256, ctx->stream);
    cudaEventCreate(&ctx->ev_total[0]);
    cudaEventCreate(&ctx->ev_total[1]);
    cudaStreamSynchronize(ctx->stream);
    *out = ctx;
    return BOWGPU_OK;
}

extern "C" void bowgpu_ctx_destroy(bowgpu_ctx *ctx) {
    if (!ctx) return;
    for (bowgpu_ctx *w : ctx->workers) bowgpu_ctx_destroy(w);
    ctx->workers.clear();
    Guard gd(ctx);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->d_status) cudaFree(ctx->d_status);
    for (int i = 0; i < 2; ++i) {
        if (ctx->pinned[i]) cudaFreeHost(ctx->pinned[i]);
        if (ctx->pinned_ev[i]) cudaEventDestroy(ctx->pinned_ev[i]);
        if (ctx->ev_total[i]) cudaEventDestroy(ctx->ev_total[i]);
    }
    for (auto e : ctx->ev_main) cudaEventDestroy(e);
    for (auto b : ctx->file_stage) cudaFreeHost(b);
    for (auto e : ctx->file_stage_ev) cudaEventDestroy(e);
    for (auto st : ctx->side) cudaStreamDestroy(st);
    for (auto e : ctx->side_done) cudaEventDestroy(e);
    if (ctx->side_fork) cudaEventDestroy(ctx->side_fork);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *bowgpu_last_error(const bowgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int32_t bowgpu_ctx_synchronize(bowgpu_ctx *ctx) {
    if (!ctx) return BOWGPU_EINVAL;
    Guard gd(ctx);
    return check_status(ctx);
}

extern "C" int32_t bowgpu_ctx_enable_timing(bowgpu_ctx *ctx, int32_t enable) {
    if (!ctx) return BOWGPU_EINVAL;
    ctx->timing = enable;
    ctx->launches = ctx->main_launches = ctx->ev_main_used = 0;
    ctx->main_seq = 0;
    return BOWGPU_OK;
}

extern "C" int32_t bowgpu_ctx_last_timing(bowgpu_ctx *ctx, bowgpu_timing *out) {
    if (!ctx || !out) return BOWGPU_EINVAL;
    Guard gd(ctx);
    memset(out, 0, sizeof *out);
    out->launches = ctx->launches;
    out->main_launches = ctx->main_launches;
    if (!ctx->timing) return BOWGPU_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->timing == 1) CK(cudaEventElapsedTime(&out->total_ms, ctx->ev_total[0], ctx->ev_total[1]));
    float acc = 0.f;
    for (int i = 0; i + 1 < ctx->ev_main_used; i += 2) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev_main[i], ctx->ev_main[i + 1]));
        acc += ms;
    }
    out->main_ms = acc;
    if (ctx->timing == 2) {
        out->main_launches = ctx->ev_main_used / 2;  // the launches main_ms was measured on
        ctx->launches = ctx->main_launches = ctx->ev_main_used = 0;
    }
    return BOWGPU_OK;
}

extern "C" int32_t bowgpu_ctx_sm_count(const bowgpu_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

extern "C" int32_t bowgpu_ctx_trim(bowgpu_ctx *ctx) {
    if (!ctx) return BOWGPU_EINVAL;
    Guard gd(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemPoolTrimTo(ctx->pool, 0));
    return BOWGPU_OK;
}

// ================================================================================================
// frames
// ================================================================================================
namespace {

void free_col(bowgpu_ctx *ctx, DevCol &c) {
    if (c.own_values && c.values) pool_free(ctx, c.values);
    if (c.own_validity && c.validity) pool_free(ctx, c.validity);
    c = DevCol();
}

// allocates an owned column of n rows (values padded to a 16-byte multiple + one spare vector)
int32_t alloc_col(bowgpu_ctx *ctx, DevCol &c, int64_t n, int32_t dtype, bool with_validity) {
    c.dtype = dtype;
    size_t vb = align_up((size_t)n * 8, 16) + 32;
    CK(pool_alloc(ctx, (void **)&c.values, vb));
    c.own_values = true;
    if (with_validity) {
        size_t bb = (size_t)bitmap_bytes_padded(n);
        CK(pool_alloc(ctx, (void **)&c.validity, bb));
        c.own_validity = true;
        CK(cudaMemsetAsync(c.validity, 0, bb, ctx->stream));
    }
    return BOWGPU_OK;
}

int32_t count_nulls(bowgpu_ctx *ctx, DevCol &c, int64_t n) {
    unsigned long long *d = (unsigned long long *)ctx->d_scalars;
    CK(cudaMemsetAsync(d, 0, 8, ctx->stream));
    CK(launch_bitmap_popcount(c.validity, n, d, ctx->stream));
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    c.null_count = n - (int64_t)h;
    return BOWGPU_OK;
}

}  // namespace

extern "C" int32_t bowgpu_frame_create(bowgpu_ctx *ctx, const bowgpu_col *cols, int32_t ncols, int32_t mem,
                                       bowgpu_frame **out) {
    if (!ctx || !out || ncols < 0 || (ncols > 0 && !cols)) return BOWGPU_EINVAL;
    *out = nullptr;
    Guard gd(ctx);
    const int64_t n = ncols ? cols[0].length : 0;
    for (int j = 0; j < ncols; ++j) {
        if (cols[j].length != n || cols[j].offset < 0 || n < 0)
            return fail(ctx, BOWGPU_EINVAL, "column %d: length %lld != %lld", j, (long long)cols[j].length, (long long)n);
        if (cols[j].dtype != BOWGPU_FLOAT64 && cols[j].dtype != BOWGPU_INT64)
            return fail(ctx, BOWGPU_ETYPE, "column %d: only Int64 / Float64 columns run on the GPU path", j);
        if (n > 0 && !cols[j].values) return fail(ctx, BOWGPU_EINVAL, "column %d: null values buffer", j);
    }
    bowgpu_frame *f = new (std::nothrow) bowgpu_frame();
    if (!f) return BOWGPU_ENOMEM;
    f->ctx = ctx;
    f->n = n;
    f->cols.resize(ncols);
    int32_t rc = BOWGPU_OK;
    arena_reset(ctx);
    for (int j = 0; j < ncols && rc == BOWGPU_OK; ++j) {
        const bowgpu_col &s = cols[j];
        DevCol &c = f->cols[j];
        const bool want_validity = s.validity != nullptr && s.null_count != 0 && n > 0;
        const char *vsrc = (const char *)s.values + s.offset * 8;
        if (mem == BOWGPU_MEM_DEVICE && n > 0 && ((uintptr_t)vsrc & 15) == 0) {
            c.values = (uint64_t *)vsrc;  // zero copy
            c.dtype = s.dtype;
            if (want_validity) {
                size_t bb = (size_t)bitmap_bytes_padded(n);
                cudaError_t e = pool_alloc(ctx, (void **)&c.validity, bb);
                if (e != cudaSuccess) {
                    rc = fail(ctx, BOWGPU_ENOMEM, "cudaMalloc(%zu): %s", bb, cudaGetErrorString(e));
                    break;
                }
                c.own_validity = true;
            }
        } else {
            rc = alloc_col(ctx, c, n, s.dtype, want_validity);
            if (rc) break;
            if (n > 0) {
                if (mem == BOWGPU_MEM_DEVICE) {
                    cudaError_t e = cudaMemcpyAsync(c.values, vsrc, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream);
                    if (e != cudaSuccess) rc = fail(ctx, BOWGPU_ECUDA, "D2D copy: %s", cudaGetErrorString(e));
                } else {
                    rc = copy_h2d(ctx, c.values, vsrc, (size_t)n * 8);
                }
            }
        }
        if (rc == BOWGPU_OK && want_validity) {
            const int64_t byte0 = s.offset >> 3;
            const int64_t bit_off = s.offset & 7;
            const int64_t nbytes = ((s.offset + n + 7) >> 3) - byte0;
            const uint8_t *bsrc = s.validity + byte0;
            const int64_t dst_bytes = bitmap_bytes_padded(n);
            if (mem == BOWGPU_MEM_HOST) {
                rc = arena_reserve(ctx, (size_t)nbytes + 256);
                if (rc) break;
                arena_reset(ctx);
                uint8_t *tmp = (uint8_t *)arena_take(ctx, (size_t)nbytes);
                rc = copy_h2d(ctx, tmp, bsrc, (size_t)nbytes);
                bsrc = tmp;
            }
            if (rc == BOWGPU_OK) {
                int e = launch_bitmap_realign(bsrc, bit_off, n, c.validity, dst_bytes, ctx->stream);
                if (e) rc = fail(ctx, BOWGPU_ECUDA, "bitmap_realign: %s", cudaGetErrorString((cudaError_t)e));
            }
            if (rc == BOWGPU_OK) {
                c.null_count = s.null_count;
                if (s.null_count < 0) rc = count_nulls(ctx, c, n);
                if (rc == BOWGPU_OK && c.null_count == 0) {  // bitmap without nulls: drop it (faster kernels)
                    cudaStreamSynchronize(ctx->stream);
                    if (c.own_validity) pool_free(ctx, c.validity);
                    c.validity = nullptr;
                    c.own_validity = false;
                }
            }
        }
    }
    if (rc == BOWGPU_OK) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);  // Go may release its buffers after return
        if (e != cudaSuccess) rc = fail(ctx, BOWGPU_ECUDA, "frame upload: %s", cudaGetErrorString(e));
    }
    if (rc != BOWGPU_OK) {
        cudaStreamSynchronize(ctx->stream);
        for (auto &c : f->cols) free_col(ctx, c);
        delete f;
        return rc;
    }
    *out = f;
    return BOWGPU_OK;
}

extern "C" void bowgpu_frame_destroy(bowgpu_frame *f) {
    if (!f) return;
    bowgpu_ctx *ctx = f->ctx;
    Guard gd(ctx);
    bool borrowed = false;  // zero-copy columns belong to the caller: it may release them right after this call
    for (auto &c : f->cols) borrowed |= c.values && !c.own_values;
    if (borrowed) cudaStreamSynchronize(ctx->stream);
    for (auto &c : f->cols) free_col(ctx, c);  // stream ordered: kernels still using the columns finish first
    delete f;
}

extern "C" int64_t bowgpu_frame_num_rows(const bowgpu_frame *f) { return f ? f->n : 0; }
extern "C" int32_t bowgpu_frame_num_cols(const bowgpu_frame *f) { return f ? (int32_t)f->cols.size() : 0; }
extern "C" int32_t bowgpu_frame_col_dtype(const bowgpu_frame *f, int32_t col) {
    return (f && col >= 0 && col < (int32_t)f->cols.size()) ? f->cols[col].dtype : 0;
}
extern "C" int32_t bowgpu_frame_col_has_validity(const bowgpu_frame *f, int32_t col) {
    return (f && col >= 0 && col < (int32_t)f->cols.size()) ? f->cols[col].validity != nullptr : 0;
}
extern "C" int32_t bowgpu_frame_col_device_ptrs(const bowgpu_frame *f, int32_t col, void **values, uint8_t **validity) {
    if (!f || col < 0 || col >= (int32_t)f->cols.size()) return BOWGPU_EINVAL;
    if (values) *values = f->cols[col].values;
    if (validity) *validity = f->cols[col].validity;
    return BOWGPU_OK;
}

extern "C" int32_t bowgpu_frame_download_range(const bowgpu_frame *f, int64_t row0, int64_t nrows, bowgpu_out_col *outs,
                                               int32_t ncols) {
    if (!f || !outs || ncols != (int32_t)f->cols.size() || row0 < 0 || nrows < 0 || row0 + nrows > f->n)
        return BOWGPU_EINVAL;
    bowgpu_ctx *ctx = f->ctx;
    Guard gd(ctx);
    for (int j = 0; j < ncols; ++j) {
        const DevCol &c = f->cols[j];
        outs[j].dtype = c.dtype;
        if (nrows == 0) continue;
        int32_t rc = copy_d2h(ctx, outs[j].values, c.values + row0, (size_t)nrows * 8);
        if (rc) return rc;
        if (!outs[j].validity) continue;
        const size_t ob = (size_t)((nrows + 7) / 8);
        if (!c.validity) {
            memset(outs[j].validity, 0xff, ob);
            if (nrows & 7) outs[j].validity[ob - 1] = (uint8_t)((1u << (nrows & 7)) - 1u);
            continue;
        }
        if (row0 == 0 && nrows == f->n) {
            rc = copy_d2h(ctx, outs[j].validity, c.validity, ob);
            if (rc) return rc;
        } else {  // realign the sub-range on the device, then copy
            const int64_t dst_bytes = bitmap_bytes_padded(nrows);
            rc = arena_reserve(ctx, (size_t)dst_bytes + 256);
            if (rc) return rc;
            arena_reset(ctx);
            uint8_t *tmp = (uint8_t *)arena_take(ctx, (size_t)dst_bytes);
            CK(launch_bitmap_realign(c.validity + (row0 >> 3), row0 & 7, nrows, tmp, dst_bytes, ctx->stream));
            rc = copy_d2h(ctx, outs[j].validity, tmp, ob);
            if (rc) return rc;
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return BOWGPU_OK;
}

extern "C" int32_t bowgpu_frame_download(const bowgpu_frame *f, bowgpu_out_col *outs, int32_t ncols) {
    return bowgpu_frame_download_range(f, 0, f ? f->n : 0, outs, ncols);
}

// ================================================================================================
// Parquet ingest (bowparquet.go:44-155): metadata on the host, data on the device (parquet.cu)
// ================================================================================================
// File bytes -> device: FILE_READERS threads pread() slices of the ranges into their own pair of pinned chunks and queue the
// DMA on the ctx stream (one memcpy page cache -> pinned per byte; a single thread copying out of an mmap tops out near
// 8 GB/s and pays a page fault per 4 KB).  The reference reads with 4 goroutines as well (pr.NP = 4, bowparquet.go:52).
constexpr int FILE_READERS = 4;
constexpr size_t FILE_SLICE = (size_t)16 << 20;
static int32_t upload_file_ranges(bowgpu_ctx *ctx, int fd, const std::vector<PqRange> &ranges, uint8_t *image) {
    struct Slice {
        int64_t file_off, len, image_off;
    };
    std::vector<Slice> slices;
    for (const PqRange &rg : ranges)
        for (int64_t o = 0; o < rg.len; o += (int64_t)FILE_SLICE)
            slices.push_back({rg.file_off + o, std::min<int64_t>((int64_t)FILE_SLICE, rg.len - o), rg.image_off + o});
    if (slices.empty()) return BOWGPU_OK;
    while (ctx->file_stage.size() < (size_t)2 * FILE_READERS) {
        uint8_t *b = nullptr;
        cudaEvent_t e;
        CK(cudaHostAlloc((void **)&b, FILE_SLICE, cudaHostAllocDefault));
        ctx->file_stage.push_back(b);
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->file_stage_ev.push_back(e);
    }
    const int nthreads = (int)std::min<size_t>(FILE_READERS, slices.size());
    std::atomic<size_t> next{0};
    std::atomic<int> err{0};
    auto work = [&](int t) {
        cudaSetDevice(ctx->device);
        int b = 0;
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= slices.size() || err.load()) break;
            const Slice &sl = slices[i];
            uint8_t *buf = ctx->file_stage[2 * t + b];
            cudaEvent_t ev = ctx->file_stage_ev[2 * t + b];
            if (cudaEventSynchronize(ev) != cudaSuccess) { err = 1; break; }
            int64_t done = 0;
            while (done < sl.len) {
                const ssize_t r = pread(fd, buf + done, (size_t)(sl.len - done), (off_t)(sl.file_off + done));
                if (r <= 0) { err = 2; break; }
                done += r;
            }
            if (err.load()) break;
            if (cudaMemcpyAsync(image + sl.image_off, buf, (size_t)sl.len, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
                cudaEventRecord(ev, ctx->stream) != cudaSuccess) { err = 1; break; }
            b ^= 1;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    if (err.load() == 2) return fail(ctx, BOWGPU_EIO, "pread: %s", strerror(errno));
    if (err.load()) return fail(ctx, BOWGPU_ECUDA, "file upload: %s", cudaGetErrorString(cudaGetLastError()));
    return BOWGPU_OK;
}

struct bowgpu_parquet {
    PqFile *file = nullptr;
};

extern "C" int32_t bowgpu_parquet_open(const char *path, bowgpu_parquet **out, char *err, int32_t err_cap) {
    if (!path || !out) return BOWGPU_EINVAL;
    *out = nullptr;
    std::string msg;
    PqFile *f = nullptr;
    const int rc = pq_open(path, &f, msg);
    if (rc) {
        if (err && err_cap > 0) snprintf(err, (size_t)err_cap, "%s", msg.c_str());
        return rc;
    }
    bowgpu_parquet *p = new (std::nothrow) bowgpu_parquet();
    if (!p) {
        pq_close(f);
        return BOWGPU_ENOMEM;
    }
    p->file = f;
    *out = p;
    return BOWGPU_OK;
}
extern "C" void bowgpu_parquet_close(bowgpu_parquet *pq) {
    if (!pq) return;
    pq_close(pq->file);
    delete pq;
}
extern "C" int64_t bowgpu_parquet_num_rows(const bowgpu_parquet *pq) { return pq ? pq_num_rows(pq->file) : -1; }
extern "C" int32_t bowgpu_parquet_num_cols(const bowgpu_parquet *pq) { return pq ? (int32_t)pq_columns(pq->file).size() : -1; }
extern "C" const char *bowgpu_parquet_col_name(const bowgpu_parquet *pq, int32_t col) {
    if (!pq || col < 0 || col >= (int32_t)pq_columns(pq->file).size()) return nullptr;
    return pq_columns(pq->file)[col].name.c_str();
}
extern "C" int32_t bowgpu_parquet_col_dtype(const bowgpu_parquet *pq, int32_t col) {
    if (!pq || col < 0 || col >= (int32_t)pq_columns(pq->file).size()) return -1;
    return pq_columns(pq->file)[col].dtype;
}
extern "C" int32_t bowgpu_parquet_col_physical_type(const bowgpu_parquet *pq, int32_t col) {
    if (!pq || col < 0 || col >= (int32_t)pq_columns(pq->file).size()) return -1;
    return pq_columns(pq->file)[col].physical;
}
extern "C" int64_t bowgpu_frame_col_null_count(const bowgpu_frame *f, int32_t col) {
    if (!f || col < 0 || col >= (int32_t)f->cols.size()) return -1;
    return f->cols[col].validity ? f->cols[col].null_count : 0;
}

extern "C" int32_t bowgpu_parquet_plan(const bowgpu_parquet *pq, const int32_t *cols, int32_t ncols, int64_t *out4, char *err,
                                       int32_t err_cap) {
    if (!pq || !out4 || ncols < 0 || (ncols > 0 && !cols)) return BOWGPU_EINVAL;
    PqPlan plan;
    std::string msg;
    const int32_t rc = pq_plan(pq->file, cols, ncols, plan, msg);
    if (rc && err && err_cap > 0) snprintf(err, (size_t)err_cap, "%s", msg.c_str());
    out4[0] = (int64_t)plan.pages.size();
    out4[1] = plan.image_bytes;
    out4[2] = plan.scratch_bytes;
    out4[3] = plan.aux_entries;
    return rc;
}

extern "C" int32_t bowgpu_parquet_read(bowgpu_ctx *ctx, const bowgpu_parquet *pq, const int32_t *cols, int32_t ncols,
                                       bowgpu_frame **out) {
    if (!ctx || !pq || !out || ncols < 0 || (ncols > 0 && !cols)) return BOWGPU_EINVAL;
    *out = nullptr;
    Guard gd(ctx);
    PqPlan plan;
    std::string msg;
    const bool dbg = getenv("BOWGPU_PQ_DEBUG") != nullptr;  // stage times on stderr (synchronises between stages)
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(now() - t0).count();
    };
    auto t_stage = now();
    int32_t rc = pq_plan(pq->file, cols, ncols, plan, msg);
    if (rc) return fail(ctx, rc, "%s", msg.c_str());
    if (dbg) fprintf(stderr, "pq: plan %.2f ms (%zu pages)\n", ms_since(t_stage), plan.pages.size());
    const int64_t n = pq_num_rows(pq->file);
    const auto &pcols = pq_columns(pq->file);
    bowgpu_frame *f = new (std::nothrow) bowgpu_frame();
    if (!f) return BOWGPU_ENOMEM;
    f->ctx = ctx;
    f->n = n;
    f->cols.resize(ncols);
    uint8_t *image = nullptr, *scratch = nullptr;
    int32_t *aux = nullptr;
    auto cleanup = [&](int32_t code) {
        pool_free(ctx, image);
        pool_free(ctx, scratch);
        pool_free(ctx, aux);
        if (code) {
            cudaStreamSynchronize(ctx->stream);
            for (auto &c : f->cols) free_col(ctx, c);
            delete f;
        }
        return code;
    };
    for (int j = 0; j < ncols; ++j) {
        rc = alloc_col(ctx, f->cols[j], n, pcols[cols[j]].dtype, pcols[cols[j]].optional && n > 0);
        if (rc) return cleanup(rc);
    }
    const int npages = (int)plan.pages.size();
    if (n == 0 || npages == 0) {
        *out = f;
        return BOWGPU_OK;
    }
    // small tables (page descriptors, column outputs, valid-row counters) live in the arena
    const size_t pb = align_up((size_t)npages * sizeof(PqPage), 256), cb = align_up((size_t)ncols * sizeof(PqColOut), 256);
    const size_t nb = align_up((size_t)ncols * 8, 256);
    rc = arena_reserve(ctx, pb + cb + nb + 1024);
    if (rc) return cleanup(rc);
    arena_reset(ctx);
    PqPage *d_pages = (PqPage *)arena_take(ctx, pb);
    PqColOut *d_cols = (PqColOut *)arena_take(ctx, cb);
    unsigned long long *d_counts = (unsigned long long *)arena_take(ctx, nb);
    auto ck = [&](cudaError_t e, const char *what) -> int32_t {
        if (e == cudaSuccess) return BOWGPU_OK;
        return fail(ctx, e == cudaErrorMemoryAllocation ? BOWGPU_ENOMEM : BOWGPU_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    };
    if ((rc = ck(pool_alloc(ctx, (void **)&image, (size_t)plan.image_bytes), "parquet image"))) return cleanup(rc);
    if ((rc = ck(pool_alloc(ctx, (void **)&scratch, (size_t)plan.scratch_bytes), "parquet scratch"))) return cleanup(rc);
    if ((rc = ck(pool_alloc(ctx, (void **)&aux, (size_t)(plan.aux_entries + 16) * 4), "parquet index scratch"))) return cleanup(rc);
    // the file bytes of the chosen column chunks, as they are
    // (the padding behind every range / page is read by the word-granular loads of the decoder, never used: zero it)
    if ((rc = ck(cudaMemsetAsync(image, 0, (size_t)plan.image_bytes, ctx->stream), "memset"))) return cleanup(rc);
    if ((rc = ck(cudaMemsetAsync(scratch, 0, (size_t)plan.scratch_bytes, ctx->stream), "memset"))) return cleanup(rc);
    if (dbg) {
        cudaStreamSynchronize(ctx->stream);
        fprintf(stderr, "pq: alloc + memset %.2f ms\n", ms_since(t_stage));
        t_stage = now();
    }
    std::vector<PqColOut> hc(ncols);
    for (int j = 0; j < ncols; ++j) {
        hc[j].values = f->cols[j].values;
        hc[j].validity = (uint32_t *)f->cols[j].validity;
        hc[j].valid_count = d_counts + j;
    }
    if ((rc = ck(cudaMemsetAsync(d_counts, 0, nb, ctx->stream), "memset"))) return cleanup(rc);
    rc = copy_h2d(ctx, d_pages, plan.pages.data(), (size_t)npages * sizeof(PqPage));
    if (!rc) rc = copy_h2d(ctx, d_cols, hc.data(), (size_t)ncols * sizeof(PqColOut));
    if (rc) return cleanup(rc);
    rc = upload_file_ranges(ctx, pq_fd(pq->file), plan.ranges, image);
    if (rc) return cleanup(rc);
    if (dbg) {
        cudaStreamSynchronize(ctx->stream);
        fprintf(stderr, "pq: upload %.2f ms (%.1f MB)\n", ms_since(t_stage), plan.image_bytes / 1e6);
        t_stage = now();
    }
    // Decompress + decode after the last byte has arrived.  Launching the kernels of a column chunk as soon as its bytes were
    // queued (second stream, behind an event) was measured and is WORSE on B200: 44.5 instead of 38.5 ms uncompressed,
    // 1138 instead of 61 ms with Snappy and 1 MiB pages (the long-running one-warp-per-page kernel and the H2D stream get
    // in each other's way) — gpurun_out/s12_parquet.jsonl.
    if ((rc = ck((cudaError_t)launch_pq_decode(d_pages, 0, npages, image, scratch, aux, d_cols, ctx->d_status, ctx->stream), "parquet decode")))
        return cleanup(rc);
    std::vector<unsigned long long> counts(ncols);
    int32_t st = 0;
    rc = copy_d2h(ctx, counts.data(), d_counts, (size_t)ncols * 8);
    if (!rc) rc = ck(cudaMemcpyAsync(&st, ctx->d_status, 4, cudaMemcpyDeviceToHost, ctx->stream), "status");
    if (!rc) rc = ck(cudaStreamSynchronize(ctx->stream), "parquet decode");
    if (rc) return cleanup(rc);
    if (dbg) fprintf(stderr, "pq: decompress + decode %.2f ms\n", ms_since(t_stage));
    if (st & ST_PARQUET) {
        cudaMemsetAsync(ctx->d_status, 0, 4, ctx->stream);
        return cleanup(fail(ctx, BOWGPU_EIO, "malformed parquet page data (Snappy stream, levels or values out of bounds)"));
    }
    for (int j = 0; j < ncols; ++j) {
        DevCol &c = f->cols[j];
        c.null_count = n - (int64_t)counts[j];
        if (c.validity && c.null_count == 0) {  // no nulls: the kernels take the all-valid path
            pool_free(ctx, c.validity);
            c.validity = nullptr;
            c.own_validity = false;
        }
    }
    *out = f;
    return cleanup(BOWGPU_OK);
}

extern "C" int32_t bowgpu_frame_generate(bowgpu_ctx *ctx, const bowgpu_gen_spec *spec, bowgpu_frame **out) {
    if (!ctx || !spec || !out || spec->nrows < 0 || spec->ncols < 0 || spec->ncols > 31) return BOWGPU_EINVAL;
    *out = nullptr;
    Guard gd(ctx);
    if (spec->kind != BOWGPU_GEN_REGULAR && spec->kind != BOWGPU_GEN_BURSTY)
        return fail(ctx, BOWGPU_EUNSUPPORTED, "generator kind %d", spec->kind);
    if (spec->kind == BOWGPU_GEN_BURSTY && (spec->step <= 0 || spec->row0 < 0))
        return fail(ctx, BOWGPU_EINVAL, "bursty generator: step (window interval) must be positive");
    bowgpu_frame *f = new (std::nothrow) bowgpu_frame();
    if (!f) return BOWGPU_ENOMEM;
    f->ctx = ctx;
    f->n = spec->nrows;
    f->cols.resize(spec->ncols + 1);
    int32_t rc = alloc_col(ctx, f->cols[0], f->n, BOWGPU_INT64, false);
    if (rc == BOWGPU_OK && spec->kind == BOWGPU_GEN_REGULAR) {
        int e = launch_gen_regular((int64_t *)f->cols[0].values, f->n, spec->row0, spec->t0, spec->step, ctx->stream);
        if (e) rc = fail(ctx, BOWGPU_ECUDA, "gen_time: %s", cudaGetErrorString((cudaError_t)e));
    }
    if (rc == BOWGPU_OK && spec->kind == BOWGPU_GEN_BURSTY && f->n > 0) {
        // enough windows of the pattern to cover global rows [0, row0 + nrows): the mean window holds ~7.7e3 rows
        const int64_t nw = (spec->row0 + spec->nrows) / 2000 + 4096;
        rc = arena_reserve(ctx, (size_t)(nw + 1) * 8 + scan_scratch_bytes(nw) + 1024);
        if (rc == BOWGPU_OK) {
            arena_reset(ctx);
            int64_t *off = (int64_t *)arena_take(ctx, (size_t)(nw + 1) * 8);
            int64_t *tmp = (int64_t *)arena_take(ctx, scan_scratch_bytes(nw));
            int e = launch_gen_bursty_counts(off, nw, spec->seed, ctx->stream);
            if (!e) e = launch_exclusive_scan(off, nw, tmp, ctx->stream);
            int64_t total = 0;
            if (!e) e = (int)cudaMemcpyAsync(&total, off + nw, 8, cudaMemcpyDeviceToHost, ctx->stream);
            if (!e) e = (int)cudaStreamSynchronize(ctx->stream);
            if (!e && total < spec->row0 + spec->nrows)
                rc = fail(ctx, BOWGPU_EINVAL, "bursty generator: pattern covers %lld rows, %lld requested", (long long)total,
                          (long long)(spec->row0 + spec->nrows));
            if (!e && rc == BOWGPU_OK)
                e = launch_gen_bursty_time((int64_t *)f->cols[0].values, f->n, spec->row0, spec->t0, spec->step, spec->seed,
                                           off, nw, ctx->stream);
            if (e) rc = fail(ctx, BOWGPU_ECUDA, "gen_bursty: %s", cudaGetErrorString((cudaError_t)e));
        }
    }
    for (int c = 0; c < spec->ncols && rc == BOWGPU_OK; ++c) {
        DevCol &dc = f->cols[c + 1];
        const bool nulls = ((spec->null_mask >> c) & 1u) && spec->null_mod > 0;
        const bool is_int = (spec->int_mask >> c) & 1u;
        rc = alloc_col(ctx, dc, f->n, is_int ? BOWGPU_INT64 : BOWGPU_FLOAT64, nulls);
        if (rc) break;
        int e = launch_gen_values(dc.values, dc.validity, f->n, spec->row0, spec->seed, (uint64_t)(c + 1), is_int,
                                  nulls ? spec->null_mod : 0, ctx->stream);
        if (e) rc = fail(ctx, BOWGPU_ECUDA, "gen_values: %s", cudaGetErrorString((cudaError_t)e));
        if (rc == BOWGPU_OK && nulls) rc = count_nulls(ctx, dc, f->n);
    }
    if (rc == BOWGPU_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        rc = fail(ctx, BOWGPU_ECUDA, "generate: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc != BOWGPU_OK) {
        for (auto &c : f->cols) free_col(ctx, c);
        delete f;
        return rc;
    }
    *out = f;
    return BOWGPU_OK;
}


// ================================================================================================
// whole-column fills (bowfill.go)
// ================================================================================================
namespace {

// copies column `src` of n rows into a freshly allocated column of the new frame
int32_t clone_col(bowgpu_ctx *ctx, DevCol &dst, const DevCol &src, int64_t n) {
    int32_t rc = alloc_col(ctx, dst, n, src.dtype, src.validity != nullptr);
    if (rc) return rc;
    dst.null_count = src.null_count;
    if (n > 0) {
        CK(cudaMemcpyAsync(dst.values, src.values, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        if (src.validity)
            CK(cudaMemcpyAsync(dst.validity, src.validity, (size_t)bitmap_bytes_padded(n), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return BOWGPU_OK;
}

// fills column c of `f` into `dst` (allocated here); ref_col < 0 unless method == LINEAR
int32_t fill_col(bowgpu_ctx *ctx, const bowgpu_frame *f, int c, int method, int ref_col, DevCol &dst) {
    const DevCol &src = f->cols[c];
    const int64_t n = f->n;
    int32_t rc = alloc_col(ctx, dst, n, src.dtype, true);
    if (rc) return rc;
    rc = arena_reserve(ctx, fill_scratch_bytes(n) + 512);
    if (rc) return rc;
    arena_reset(ctx);
    int64_t *scratch = (int64_t *)arena_take(ctx, fill_scratch_bytes(n));
    FillLaunch L;
    memset(&L, 0, sizeof L);
    L.values = src.values;
    L.validity = src.validity;
    L.out_values = dst.values;
    L.out_validity = dst.validity;
    L.n = n;
    L.is_int = src.dtype == BOWGPU_INT64;
    if (ref_col >= 0) {
        L.ref_values = f->cols[ref_col].values;
        L.ref_validity = f->cols[ref_col].validity;
        L.ref_is_int = f->cols[ref_col].dtype == BOWGPU_INT64;
    }
    CK(launch_fill(method, L, scratch, ctx->stream));
    count_launch(ctx, 3, false);
    rc = count_nulls(ctx, dst, n);  // synchronizes: the arena scratch is free again afterwards
    if (rc) return rc;
    if (dst.null_count == 0) {
        pool_free(ctx, dst.validity);
        dst.validity = nullptr;
        dst.own_validity = false;
    }
    return BOWGPU_OK;
}

}  // namespace

extern "C" int32_t bowgpu_frame_fill(bowgpu_frame *frame, int32_t method, const int32_t *cols, int32_t ncols,
                                     bowgpu_frame **out) {
    if (!frame || !out || ncols < 0 || (ncols > 0 && !cols)) return BOWGPU_EINVAL;
    *out = nullptr;
    bowgpu_ctx *ctx = frame->ctx;
    Guard gd(ctx);
    if (method != BOWGPU_FILL_PREVIOUS && method != BOWGPU_FILL_NEXT && method != BOWGPU_FILL_MEAN)
        return fail(ctx, BOWGPU_EINVAL, "bow.Fill: method '%d' is not supported", method);
    const int nc = (int)frame->cols.size();
    std::vector<char> sel(nc, ncols == 0 ? 1 : 0);  // selectCols, bowfill.go:266-288
    for (int i = 0; i < ncols; ++i) {
        if (cols[i] < 0 || cols[i] > nc - 1) return fail(ctx, BOWGPU_EINVAL, "selectCols: colIndex '%d' out of range", cols[i]);
        sel[cols[i]] = 1;
    }
    bowgpu_frame *of = new (std::nothrow) bowgpu_frame();
    if (!of) return BOWGPU_ENOMEM;
    of->ctx = ctx;
    of->n = frame->n;
    of->cols.resize(nc);
    int32_t rc = BOWGPU_OK;
    timing_begin(ctx);
    for (int c = 0; c < nc && rc == BOWGPU_OK; ++c) {
        const DevCol &src = frame->cols[c];
        if (!sel[c] || !src.validity || src.null_count == 0 || frame->n == 0)
            rc = clone_col(ctx, of->cols[c], src, frame->n);  // NewSeriesFromCol, bowfill.go:129-132,176-179
        else
            rc = fill_col(ctx, frame, c, method, -1, of->cols[c]);
    }
    timing_end(ctx);
    if (rc == BOWGPU_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        rc = fail(ctx, BOWGPU_ECUDA, "fill: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc != BOWGPU_OK) {
        cudaStreamSynchronize(ctx->stream);
        for (auto &c : of->cols) free_col(ctx, c);
        delete of;
        return rc;
    }
    *out = of;
    return BOWGPU_OK;
}

extern "C" int32_t bowgpu_frame_fill_linear(bowgpu_frame *frame, int32_t ref_col, int32_t tofill_col, bowgpu_frame **out) {
    if (!frame || !out) return BOWGPU_EINVAL;
    *out = nullptr;
    bowgpu_ctx *ctx = frame->ctx;
    Guard gd(ctx);
    const int nc = (int)frame->cols.size();
    if (ref_col < 0 || ref_col > nc - 1) return fail(ctx, BOWGPU_EINVAL, "refColIndex is out of range");  // bowfill.go:18-28
    if (tofill_col < 0 || tofill_col > nc - 1) return fail(ctx, BOWGPU_EINVAL, "toFillColIndex is out of range");
    if (ref_col == tofill_col) return fail(ctx, BOWGPU_EINVAL, "refColIndex and toFillColIndex are equal");
    const DevCol &rcol = frame->cols[ref_col];
    const int64_t n = frame->n;
    const bool ref_empty = n == 0 || (rcol.validity && rcol.null_count == n);  // IsColEmpty, bowfill.go:37-39
    bool do_fill = !ref_empty && frame->cols[tofill_col].validity && frame->cols[tofill_col].null_count != 0;
    if (!ref_empty) {  // IsColSorted, bowfill.go:41-44
        int32_t rc = arena_reserve(ctx, fill_scratch_bytes(n) + 1024);
        if (rc) return rc;
        arena_reset(ctx);
        int64_t *scratch = (int64_t *)arena_take(ctx, fill_scratch_bytes(n));
        int32_t *d_flags = (int32_t *)arena_take(ctx, 16);
        CK(cudaMemsetAsync(d_flags, 0, 16, ctx->stream));
        CK(launch_sorted_flags(rcol.values, rcol.validity, rcol.dtype == BOWGPU_INT64, n, scratch, d_flags, ctx->stream));
        int32_t h = 0;
        CK(cudaMemcpyAsync(&h, d_flags, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (h == 3) return fail(ctx, BOWGPU_EUNSORTED, "refColIndex '%d' is empty or not sorted", ref_col);
    }
    bowgpu_frame *of = new (std::nothrow) bowgpu_frame();
    if (!of) return BOWGPU_ENOMEM;
    of->ctx = ctx;
    of->n = n;
    of->cols.resize(nc);
    int32_t rc = BOWGPU_OK;
    timing_begin(ctx);
    for (int c = 0; c < nc && rc == BOWGPU_OK; ++c) {
        if (c == tofill_col && do_fill)
            rc = fill_col(ctx, frame, c, BOWGPU_FILL_LINEAR, ref_col, of->cols[c]);
        else
            rc = clone_col(ctx, of->cols[c], frame->cols[c], n);
    }
    timing_end(ctx);
    if (rc == BOWGPU_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        rc = fail(ctx, BOWGPU_ECUDA, "fill: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc != BOWGPU_OK) {
        cudaStreamSynchronize(ctx->stream);
        for (auto &c : of->cols) free_col(ctx, c);
        delete of;
        return rc;
    }
    *out = of;
    return BOWGPU_OK;
}

// Bow.DropNils(colIndices...) (bow.go:188-224): drops every row holding a nil in one of the selected columns (all
// columns when none is given).  Stream compaction on the device; the result is a new frame.
extern "C" int32_t bowgpu_frame_drop_nils(bowgpu_frame *frame, const int32_t *cols, int32_t ncols, bowgpu_frame **out) {
    if (!frame || !out || ncols < 0 || (ncols > 0 && !cols)) return BOWGPU_EINVAL;
    *out = nullptr;
    bowgpu_ctx *ctx = frame->ctx;
    Guard gd(ctx);
    const int nc = (int)frame->cols.size();
    if (nc > 32) return fail(ctx, BOWGPU_EUNSUPPORTED, "DropNils supports at most 32 columns");
    std::vector<char> sel(nc, ncols == 0 ? 1 : 0);  // selectCols, bowfill.go:266-288
    for (int i = 0; i < ncols; ++i) {
        if (cols[i] < 0 || cols[i] > nc - 1) return fail(ctx, BOWGPU_EINVAL, "selectCols: colIndex '%d' out of range", cols[i]);
        sel[cols[i]] = 1;
    }
    const int64_t n = frame->n;
    bowgpu_frame *of = new (std::nothrow) bowgpu_frame();
    if (!of) return BOWGPU_ENOMEM;
    of->ctx = ctx;
    of->cols.resize(nc);
    auto bail = [&](int32_t code) {
        cudaStreamSynchronize(ctx->stream);
        for (auto &c : of->cols) free_col(ctx, c);
        delete of;
        return code;
    };
    bool any = false;
    for (int c = 0; c < nc; ++c) any |= sel[c] && frame->cols[c].validity && frame->cols[c].null_count != 0;
    int32_t rc = BOWGPU_OK;
    if (!any || n == 0) {  // bow.go:209-211: nothing to drop
        of->n = n;
        for (int c = 0; c < nc && rc == BOWGPU_OK; ++c) rc = clone_col(ctx, of->cols[c], frame->cols[c], n);
        if (rc == BOWGPU_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, BOWGPU_ECUDA, "DropNils copy");
        if (rc) return bail(rc);
        *out = of;
        return BOWGPU_OK;
    }
    DropLaunch L;
    memset(&L, 0, sizeof L);
    L.ncols = nc;
    L.n = n;
    for (int c = 0; c < nc; ++c) {
        L.values[c] = frame->cols[c].values;
        L.validity[c] = frame->cols[c].validity;
        L.selected[c] = sel[c];
    }
    void *scratch = nullptr;
    if (pool_alloc(ctx, &scratch, drop_scratch_bytes(n)) != cudaSuccess) return bail(fail(ctx, BOWGPU_ENOMEM, "DropNils scratch"));
    timing_begin(ctx);
    int64_t *d_total = nullptr, total = 0;
    int e = launch_drop_mark(L, scratch, ctx->stream, &d_total);
    if (e || cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        pool_free(ctx, scratch);
        return bail(fail(ctx, BOWGPU_ECUDA, "DropNils (mark): %s", cudaGetErrorString(cudaGetLastError())));
    }
    of->n = total;
    for (int c = 0; c < nc && rc == BOWGPU_OK; ++c) {
        const DevCol &src = frame->cols[c];
        const bool keeps_nulls = !sel[c] && src.validity && src.null_count != 0;
        rc = alloc_col(ctx, of->cols[c], total, src.dtype, keeps_nulls);  // (bitmaps come zero-initialised)
        L.out_values[c] = of->cols[c].values;
        L.out_validity[c] = of->cols[c].validity;
        if (!keeps_nulls) L.validity[c] = nullptr;
    }
    if (rc == BOWGPU_OK && total > 0) {
        e = launch_drop_compact(L, scratch, ctx->stream);
        if (e) rc = fail(ctx, BOWGPU_ECUDA, "DropNils (compact): %s", cudaGetErrorString((cudaError_t)e));
    }
    count_launch(ctx, 6);
    timing_end(ctx);
    pool_free(ctx, scratch);
    for (int c = 0; c < nc && rc == BOWGPU_OK; ++c)
        if (of->cols[c].validity) rc = count_nulls(ctx, of->cols[c], total);
    if (rc == BOWGPU_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, BOWGPU_ECUDA, "DropNils");
    if (rc) return bail(rc);
    *out = of;
    return BOWGPU_OK;
}

// Bow.IsColSorted(colIndex) (bowassertion.go:15-81): ascending or descending over the non-nil values; an empty column
// is not sorted.
extern "C" int32_t bowgpu_frame_is_col_sorted(bowgpu_frame *frame, int32_t col, int32_t *sorted) {
    if (!frame || !sorted) return BOWGPU_EINVAL;
    bowgpu_ctx *ctx = frame->ctx;
    Guard gd(ctx);
    if (col < 0 || col >= (int)frame->cols.size()) return fail(ctx, BOWGPU_EINVAL, "no column %d", col);
    const DevCol &c = frame->cols[col];
    const int64_t n = frame->n;
    *sorted = 0;
    if (n == 0 || (c.validity && c.null_count == n)) return BOWGPU_OK;  // IsColEmpty
    int32_t rc = arena_reserve(ctx, fill_scratch_bytes(n) + 1024);
    if (rc) return rc;
    arena_reset(ctx);
    int64_t *scratch = (int64_t *)arena_take(ctx, fill_scratch_bytes(n));
    int32_t *d_flags = (int32_t *)arena_take(ctx, 16);
    CK(cudaMemsetAsync(d_flags, 0, 16, ctx->stream));
    CK(launch_sorted_flags(c.values, c.validity, c.dtype == BOWGPU_INT64, n, scratch, d_flags, ctx->stream));
    int32_t h = 0;
    CK(cudaMemcpyAsync(&h, d_flags, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *sorted = h != 3;
    return BOWGPU_OK;
}

// scratch for the validity pyramids of the columns whose interpolation looks up previous / next valid rows
// (interp.cu); one stream-ordered block, released by the caller once the window kernel is enqueued
static void *attach_pyramids(bowgpu_ctx *ctx, InterpLaunch &L, int64_t n) {
    const size_t per = align_up(interp_pyramid_bytes(n), 256);
    int need = 0;
    for (int j = 0; j < L.ncols; ++j) {
        const InterpCol &c = L.cols[j];
        need += c.validity && (c.op == BOWGPU_INTERP_STEP_PREVIOUS || c.op == BOWGPU_INTERP_LINEAR || c.op == BOWGPU_INTERP_STEP_NEXT);
    }
    if (!need || n <= 4096) return nullptr;  // short columns: the word-by-word walk is bounded anyway
    uint8_t *blk = nullptr;
    if (pool_alloc(ctx, (void **)&blk, per * need) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;  // (the lookups fall back to walking the words)
    }
    int k = 0;
    for (int j = 0; j < L.ncols; ++j) {
        InterpCol &c = L.cols[j];
        if (c.validity && (c.op == BOWGPU_INTERP_STEP_PREVIOUS || c.op == BOWGPU_INTERP_LINEAR || c.op == BOWGPU_INTERP_STEP_NEXT))
            c.summary = (uint32_t *)(blk + per * k++);
    }
    return blk;
}

// Bow.SortByCol(colIndex) (bowsort.go:10-47): rows by ascending values of a nil-free column; *out = null when the column
// is already sorted (the reference returns b itself, bowsort.go:18-21).  Stable radix sort on the device (sort.cu).
extern "C" int32_t bowgpu_frame_sort_by_col(bowgpu_frame *frame, int32_t col, bowgpu_frame **out) {
    if (!frame || !out) return BOWGPU_EINVAL;
    *out = nullptr;
    bowgpu_ctx *ctx = frame->ctx;
    Guard gd(ctx);
    const int nc = (int)frame->cols.size();
    if (col < 0 || col >= nc) return fail(ctx, BOWGPU_EINVAL, "no column %d", col);
    if (nc > 32) return fail(ctx, BOWGPU_EUNSUPPORTED, "SortByCol supports at most 32 columns");
    const DevCol &kc = frame->cols[col];
    const int64_t n = frame->n;
    if (kc.validity && kc.null_count != 0)
        return fail(ctx, BOWGPU_EINVAL, "column to sort by has %lld nil values", (long long)kc.null_count);  // bowsort.go:11-15
    if (n < 2) return BOWGPU_OK;
    if (n > 0xffffffffll) return fail(ctx, BOWGPU_EUNSUPPORTED, "SortByCol: more than 2^32-1 rows");
    void *scratch = nullptr;
    if (pool_alloc(ctx, &scratch, sort_scratch_bytes(n)) != cudaSuccess) return fail(ctx, BOWGPU_ENOMEM, "SortByCol scratch");
    std::vector<unsigned long long> hist(8 * 256);
    int32_t flags = 0;
    timing_begin(ctx);
    int e = launch_sort_prepare(kc.values, kc.dtype == BOWGPU_INT64, n, scratch, ctx->stream, &flags, hist.data());
    if (e || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        pool_free(ctx, scratch);
        return fail(ctx, BOWGPU_ECUDA, "SortByCol (keys): %s", cudaGetErrorString(cudaGetLastError()));
    }
    count_launch(ctx, 1);
    if (!(flags & 1)) {  // sort.IsSorted: nothing to do
        timing_end(ctx);
        pool_free(ctx, scratch);
        return BOWGPU_OK;
    }
    if (flags & 2) {
        timing_end(ctx);
        pool_free(ctx, scratch);
        return fail(ctx, BOWGPU_EUNSUPPORTED, "SortByCol: NaN in an unsorted float64 column (`<` is no order there; undefined upstream)");
    }
    bowgpu_frame *of = new (std::nothrow) bowgpu_frame();
    if (!of) {
        pool_free(ctx, scratch);
        return BOWGPU_ENOMEM;
    }
    of->ctx = ctx;
    of->n = n;
    of->cols.resize(nc);
    auto bail = [&](int32_t code) {
        cudaStreamSynchronize(ctx->stream);
        pool_free(ctx, scratch);
        for (auto &c : of->cols) free_col(ctx, c);
        delete of;
        return code;
    };
    SortGather G;
    memset(&G, 0, sizeof G);
    int npasses = 0;
    e = launch_sort_passes(n, scratch, hist.data(), ctx->stream, &G.idx, &G.sorted_keys, &npasses);
    if (e) return bail(fail(ctx, BOWGPU_ECUDA, "SortByCol (passes): %s", cudaGetErrorString((cudaError_t)e)));
    G.ncols = nc;
    G.key_col = col;
    G.key_is_int = kc.dtype == BOWGPU_INT64;
    G.n = n;
    int32_t rc = BOWGPU_OK;
    for (int c = 0; c < nc && rc == BOWGPU_OK; ++c) {
        const DevCol &src = frame->cols[c];
        const bool nulls = src.validity && src.null_count != 0;
        rc = alloc_col(ctx, of->cols[c], n, src.dtype, nulls);
        of->cols[c].null_count = nulls ? src.null_count : 0;  // a permutation keeps the count
        G.values[c] = src.values;
        G.validity[c] = nulls ? src.validity : nullptr;
        G.out_values[c] = of->cols[c].values;
        G.out_validity[c] = of->cols[c].validity;
    }
    if (rc) return bail(rc);
    e = launch_sort_gather(G, ctx->stream);
    count_launch(ctx, 2 * npasses + 1);
    timing_end(ctx);
    if (e || cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return bail(fail(ctx, BOWGPU_ECUDA, "SortByCol (gather): %s", cudaGetErrorString(cudaGetLastError())));
    pool_free(ctx, scratch);
    *out = of;
    return BOWGPU_OK;
}

// ================================================================================================
// rolling
// ================================================================================================
// rows with t < r->s0 at the head of the frame (and the time of the first row after them)
static int32_t find_early_rows(bowgpu_ctx *ctx, bowgpu_rolling *r) {
    const DevCol &tc = r->frame->cols[r->time_col];
    const int64_t n = r->frame->n;
    CK(launch_lower_bound((const int64_t *)tc.values, n, r->s0, ctx->d_scalars, ctx->stream));
    int64_t p = 0;
    CK(cudaMemcpyAsync(&p, ctx->d_scalars, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    r->early_rows = p;
    r->has_after_early = false;
    if (p < n) {
        int64_t tp = 0;
        CK(cudaMemcpyAsync(&tp, tc.values + p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        r->t_after_early = tp;
        r->has_after_early = true;
    }
    return BOWGPU_OK;
}

extern "C" int32_t bowgpu_rolling_create(bowgpu_frame *frame, int32_t time_col, int64_t interval, int64_t offset,
                                         int32_t inclusive, const bowgpu_col *prev_row, bowgpu_rolling **out) {
    if (!frame || !out) return BOWGPU_EINVAL;
    *out = nullptr;
    bowgpu_ctx *ctx = frame->ctx;
    Guard gd(ctx);
    const int ncols = (int)frame->cols.size();
    if (time_col < 0 || time_col >= ncols) return fail(ctx, BOWGPU_EINVAL, "time column index %d out of range", time_col);
    if (frame->cols[time_col].dtype != BOWGPU_INT64)  // rolling.go:70-73
        return fail(ctx, BOWGPU_ETYPE, "impossible to create a new intervalRolling on column of type float64");
    if (interval <= 0) return fail(ctx, BOWGPU_EINVAL, "strictly positive interval required");  // rolling.go:115-117
    if (offset >= interval || offset <= -interval) offset %= interval;                           // rolling.go:119-126
    if (offset < 0) offset += interval;
    bowgpu_rolling *r = new (std::nothrow) bowgpu_rolling();
    if (!r) return BOWGPU_ENOMEM;
    r->frame = frame;
    r->time_col = time_col;
    r->interval = interval;
    r->offset = offset;
    r->inclusive = inclusive != 0;
    auto bail = [&](int32_t code) {
        delete r;
        return code;
    };
    if (prev_row) {  // enforcePrevRow, rolling.go:130-141
        if (prev_row[0].length == 0) {
            prev_row = nullptr;
        } else if (prev_row[0].length != 1) {
            return bail(fail(ctx, BOWGPU_EPREVROW, "prevRow must have only one row"));
        }
    }
    if (prev_row) {
        r->has_prev = true;
        r->prev.resize(ncols);
        for (int j = 0; j < ncols; ++j) {
            const bowgpu_col &p = prev_row[j];
            PrevCell &pc = r->prev[j];
            pc.dtype = p.dtype;
            pc.valid = 1;
            if (p.validity) pc.valid = (p.validity[p.offset >> 3] >> (p.offset & 7)) & 1;
            memcpy(&pc.bits, (const char *)p.values + p.offset * 8, 8);
        }
    }
    const DevCol &tc = frame->cols[time_col];
    const int64_t n = frame->n;
    if (tc.validity && tc.null_count != 0)
        return bail(fail(ctx, BOWGPU_ENULLTIME, "interval column holds %lld nulls: the GPU path requires a non-null, sorted interval column",
                         (long long)tc.null_count));
    if (n > 0) {
        int64_t ends[2];
        if (cudaMemcpyAsync(&ends[0], tc.values, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaMemcpyAsync(&ends[1], tc.values + (n - 1), 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            return bail(fail(ctx, BOWGPU_ECUDA, "reading interval column ends: %s", cudaGetErrorString(cudaGetLastError())));
        r->t_first = ends[0];
        r->t_last = ends[1];
        // first window start, rolling.go:96-99 (Go '/' truncates toward zero, like C)
        int64_t s0 = (int64_t)((uint64_t)((r->t_first / interval) * interval) + (uint64_t)offset);
        if (s0 > r->t_first) s0 = (int64_t)((uint64_t)s0 - (uint64_t)interval);
        r->s0 = s0;
        // countWindows, rolling.go:143-154
        r->W = s0 > r->t_last ? 0 : (int64_t)(((uint64_t)r->t_last - (uint64_t)s0) / (uint64_t)interval) + 1;
        if (r->t_first < s0) {  // negative timestamps: leading rows before the first window start
            int32_t rc = find_early_rows(ctx, r);
            if (rc) return bail(rc);
        }
    }
    *out = r;
    return BOWGPU_OK;
}

extern "C" int32_t bowgpu_rolling_create_shard(bowgpu_frame *frame, int32_t time_col, int64_t interval, int64_t s0,
                                               int64_t num_windows, int32_t inclusive, const bowgpu_col *prev_row,
                                               bowgpu_rolling **out) {
    if (!frame || !out || num_windows < 0) return BOWGPU_EINVAL;
    bowgpu_rolling *r = nullptr;
    int32_t rc = bowgpu_rolling_create(frame, time_col, interval, 0, inclusive, prev_row, &r);
    if (rc) return rc;
    r->s0 = s0;
    r->W = num_windows;
    r->offset = 0;
    r->shard = true;
    r->early_rows = 0;
    r->has_after_early = false;
    if (frame->n > 0 && r->t_first < s0) {  // left halo: rows before the shard's first window
        rc = find_early_rows(frame->ctx, r);
        if (rc) {
            delete r;
            return rc;
        }
    }
    *out = r;
    return BOWGPU_OK;
}

extern "C" void bowgpu_rolling_destroy(bowgpu_rolling *r) { delete r; }
extern "C" int64_t bowgpu_rolling_num_windows(const bowgpu_rolling *r) { return r ? r->W : 0; }
extern "C" int64_t bowgpu_rolling_first_window_start(const bowgpu_rolling *r) { return r ? r->s0 : 0; }
extern "C" int32_t bowgpu_rolling_inclusive(const bowgpu_rolling *r) { return r ? r->inclusive : 0; }
extern "C" int64_t bowgpu_rolling_early_rows(const bowgpu_rolling *r, int32_t *kept) {
    if (!r) return 0;
    if (kept) *kept = make_geom(r, r->inclusive).early_keep;
    return r->early_rows;
}

extern "C" int32_t bowgpu_rolling_bounds(bowgpu_rolling *r, int64_t *first, uint8_t *inclusive_bitmap) {
    if (!r || !first) return BOWGPU_EINVAL;
    bowgpu_ctx *ctx = r->frame->ctx;
    Guard gd(ctx);
    const WindowGeom g = make_geom(r, r->inclusive);
    const int64_t W = g.W;
    if (g.n == 0 || W == 0) {
        for (int64_t k = 0; k <= W; ++k) first[k] = g.n;  // first[W] = n; a row-less shard only has empty windows
        if (inclusive_bitmap) memset(inclusive_bitmap, 0, (size_t)((W + 7) / 8));
        return BOWGPU_OK;
    }
    const size_t fb = (size_t)(W + 1) * 8, ib = (size_t)((W + 7) / 8) + 64;
    int32_t rc = arena_reserve(ctx, fb + ib + 1024);
    if (rc) return rc;
    arena_reset(ctx);
    int64_t *d_first = (int64_t *)arena_take(ctx, fb);
    uint8_t *d_inc = (uint8_t *)arena_take(ctx, ib);
    timing_begin(ctx);
    BoundsLaunch L;
    L.time = (const int64_t *)r->frame->cols[r->time_col].values;
    L.g = g;
    L.first = d_first;
    L.status = ctx->d_status;
    cudaEvent_t e0, e1;
    timing_main_pair(ctx, &e0, &e1);
    CK(launch_bounds(L, ctx->sm_count, ctx->stream, e0, e1));
    count_launch(ctx, 1, true);
    if (inclusive_bitmap) {
        if (r->inclusive) {
            CK(launch_inclusive_bitmap(L.time, d_first, g, d_inc, ctx->stream));
            count_launch(ctx);
        } else {
            CK(cudaMemsetAsync(d_inc, 0, ib, ctx->stream));
        }
    }
    timing_end(ctx);
    rc = copy_d2h(ctx, first, d_first, fb);
    if (rc) return rc;
    if (inclusive_bitmap) {
        rc = copy_d2h(ctx, inclusive_bitmap, d_inc, (size_t)((W + 7) / 8));
        if (rc) return rc;
    }
    return check_status(ctx);
}

extern "C" int32_t bowgpu_agg_return_type(int32_t op, int32_t input_dtype) {
    switch (op) {
    case BOWGPU_AGG_WINDOW_START: return BOWGPU_INT64;  // IteratorDependent; the iterator column is Int64
    case BOWGPU_AGG_COUNT: return BOWGPU_INT64;
    case BOWGPU_AGG_FIRST:
    case BOWGPU_AGG_LAST: return input_dtype;  // InputDependent
    default: return BOWGPU_FLOAT64;
    }
}
extern "C" int32_t bowgpu_agg_needs_inclusive(int32_t op) {
    return op == BOWGPU_AGG_INTEGRAL_TRAPEZOID || op == BOWGPU_AGG_WAVG_LINEAR;
}

static int32_t whole_return_type(int32_t op, int32_t input_dtype) {  // whole.go:44-46: iterator type := input column type
    return op == BOWGPU_AGG_WINDOW_START ? input_dtype : bowgpu_agg_return_type(op, input_dtype);
}

// syn_by_col: null, or one FusedSyn per frame column (fused Interpolate -> Aggregate: the aggregation runs over the
// interpolated frame without materialising it)
// width of a wave of side-by-side streaming launches (BOWGPU_SEG_SIDE, 1 = one launch after the other on the ctx stream)
static int side_width() {
    static const int k = [] {
        const char *e = getenv("BOWGPU_SEG_SIDE");
        int v = e ? atoi(e) : 1;  // measured on B200: side by side is no faster
        return v < 1 ? 1 : (v > 16 ? 16 : v);
    }();
    return k;
}
// one launch for both aggregation families of a column (seg_full.cu); BOWGPU_SEG_MERGE=0 keeps them apart
static bool merge_families() {
    static const bool on = [] {
        const char *e = getenv("BOWGPU_SEG_MERGE");
        return !e || atoi(e) != 0;
    }();
    return on;
}
static int32_t ensure_side(bowgpu_ctx *ctx, int k) {
    if (k <= 1) return BOWGPU_OK;
    if (!ctx->side_fork) CK(cudaEventCreateWithFlags(&ctx->side_fork, cudaEventDisableTiming));
    while ((int)ctx->side.size() < k) {
        cudaStream_t st;
        cudaEvent_t ev;
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        ctx->side.push_back(st);
        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->side_done.push_back(ev);
    }
    return BOWGPU_OK;
}

static int32_t aggregate_core(bowgpu_rolling *r, const bowgpu_agg_spec *specs, int32_t nspecs, bowgpu_out_col *outs,
                              int32_t mem, const FusedSyn *syn_by_col) {
    if (!r || !specs || !outs || nspecs <= 0) return BOWGPU_EINVAL;
    bowgpu_frame *f = r->frame;
    bowgpu_ctx *ctx = f->ctx;
    Guard gd(ctx);
    const int ncols = (int)f->cols.size();
    bool keeps_interval = false, inclusive_eff = r->inclusive != 0;
    for (int j = 0; j < nspecs; ++j) {  // validateAggregation, aggregation.go:171-188
        if (specs[j].col < 0 || specs[j].col >= ncols) return fail(ctx, BOWGPU_EINVAL, "aggregation %d: no column %d", j, specs[j].col);
        if (specs[j].op < 0 || specs[j].op >= BOWGPU_AGG__COUNT) return fail(ctx, BOWGPU_EUNSUPPORTED, "aggregation %d: unknown opcode %d", j, specs[j].op);
        if (specs[j].nfactors < 0 || specs[j].nfactors > 4) return fail(ctx, BOWGPU_EINVAL, "aggregation %d: at most 4 factors", j);
        if (bowgpu_agg_needs_inclusive(specs[j].op)) inclusive_eff = true;
        if (specs[j].col == r->time_col) keeps_interval = true;
    }
    if (!keeps_interval && !r->whole)
        return fail(ctx, BOWGPU_ENOINTERVALCOL, "must keep interval column");  // aggregation.go:163-166
    for (int j = 0; j < nspecs; ++j)
        outs[j].dtype = r->whole ? whole_return_type(specs[j].op, f->cols[specs[j].col].dtype)
                                 : bowgpu_agg_return_type(specs[j].op, f->cols[specs[j].col].dtype);
    const WindowGeom g = make_geom(r, inclusive_eff);
    const int64_t W = g.W;
    if (W == 0) return BOWGPU_OK;

    // ---- plan: per distinct input column one streaming launch per kernel family ------------------------
    std::vector<int> cols_used;
    int n_basic_cols = 0, n_int_cols = 0;
    for (int j = 0; j < nspecs; ++j) {
        if (specs[j].op == BOWGPU_AGG_WINDOW_START) continue;
        if (std::find(cols_used.begin(), cols_used.end(), specs[j].col) == cols_used.end()) cols_used.push_back(specs[j].col);
    }
    for (int c : cols_used) {
        bool b = false, i = false;
        for (int j = 0; j < nspecs; ++j)
            if (specs[j].col == c) b |= agg_is_basic(specs[j].op), i |= agg_is_integral(specs[j].op);
        n_basic_cols += b;
        n_int_cols += i;
    }
    // ---- side-by-side launches: the streaming launches of one kernel family run in waves of up to K columns, every lane
    // of a wave on its own stream and on 1/K of the SMs' CTA slots.  Tiles are dealt round-robin inside a launch, so the
    // lanes of a wave move through the rows abreast and the time tiles all of them read come from DRAM once and from L2
    // after that (a lane that runs ahead takes the misses and falls back) — DESIGN.md 3.7.
    const int n_fam[2] = {n_basic_cols, n_int_cols};
    int K = std::min(side_width(), std::max(n_basic_cols, n_int_cols));
    if (K < 1) K = 1;
    int32_t rc = ensure_side(ctx, K);
    if (rc) return rc;
    const size_t wv = align_up((size_t)W * 8, 256), wb = align_up((size_t)((W + 7) / 8) + 16, 256);
    const size_t carry_bytes = align_up(std::max(std::max(seg_carry_bytes(g.n), integral_carry_bytes(g.n)), full_carry_bytes(g.n)), 256);
    const size_t skip_bytes = align_up(std::max(std::max(seg_skip_bytes(g.n), integral_skip_bytes(g.n)), full_skip_bytes(g.n)) + 256, 256);
    size_t need = 8192 + 256 * (size_t)(n_basic_cols + n_int_cols + 2) + (size_t)K * (skip_bytes + carry_bytes + 512) + (size_t)n_basic_cols * 2 * (wv + 256) +
                  (size_t)n_int_cols * 6 * (wv + 256);
    if (mem == BOWGPU_MEM_HOST) need += (size_t)nspecs * (wv + wb);
    rc = arena_reserve(ctx, need);
    if (rc) return rc;
    arena_reset(ctx);
    std::vector<void *> dvals(nspecs);
    std::vector<uint8_t *> dbits(nspecs);
    for (int j = 0; j < nspecs; ++j) {
        if (mem == BOWGPU_MEM_DEVICE) {
            dvals[j] = outs[j].values;
            dbits[j] = outs[j].validity;
        } else {
            dvals[j] = arena_take(ctx, wv);
            dbits[j] = (uint8_t *)arena_take(ctx, wb);
        }
        if (!dvals[j] || !dbits[j]) return fail(ctx, BOWGPU_EINVAL, "aggregation %d: null output buffer", j);
    }
    std::vector<uint8_t *> carry_of(K), skip_of(K);  // tile records: one set per lane
    for (int l = 0; l < K; ++l) {
        carry_of[l] = (uint8_t *)arena_take(ctx, carry_bytes);
        skip_of[l] = (uint8_t *)arena_take(ctx, skip_bytes);
    }

    timing_begin(ctx);
    // per spec: the per-window count that decides validity, and the base quantity a derived op divides
    std::vector<const int64_t *> spec_cnt(nspecs, nullptr);
    std::vector<const double *> spec_src(nspecs, nullptr);
    for (int fam = 0; fam < 2; ++fam) {
        int left = n_fam[fam], lane = 0, width = 0;  // launches of the family still to go; lane and width of the open wave
        int32_t *gate = nullptr;
        for (int c : cols_used) {
            const DevCol &dc = f->cols[c];
            // first spec of each op on this column receives the kernel output; duplicates are copied afterwards
            int primary[BOWGPU_AGG__COUNT], count[BOWGPU_AGG__COUNT];
            for (int o = 0; o < BOWGPU_AGG__COUNT; ++o) primary[o] = -1, count[o] = 0;
            for (int j = 0; j < nspecs; ++j)
                if (specs[j].col == c) {
                    if (primary[specs[j].op] < 0) primary[specs[j].op] = j;
                    ++count[specs[j].op];
                }
            bool any_basic = false, any_int = false;
            for (int o = 0; o < BOWGPU_AGG__COUNT; ++o)
                if (count[o]) any_basic |= agg_is_basic(o), any_int |= agg_is_integral(o);
            if (!(fam == 0 ? any_basic : any_int)) continue;
            // a column with aggregations of BOTH families: one merged launch (seg_full.cu) in the first pass
            const bool both = any_basic && any_int && merge_families();
            if (fam == 1 && both) continue;
            // ---- the lane this launch runs on
            if (lane == width) {  // open a wave (the main stream has already been told to wait for the previous one)
                width = std::min(K, left);
                lane = 0;
                gate = nullptr;
                if (width > 1) {
                    gate = (int32_t *)arena_take(ctx, 256);  // the lanes of a wave start their walk together
                    if (gate) CK(cudaMemsetAsync(gate, 0, 4, ctx->stream));
                    CK(cudaEventRecord(ctx->side_fork, ctx->stream));
                }
            }
            cudaStream_t st = ctx->stream;
            if (width > 1) {
                st = ctx->side[lane];
                CK(cudaStreamWaitEvent(st, ctx->side_fork, 0));
            }
            const int sm_share = std::max(1, ctx->sm_count / width);
            uint8_t *carry = carry_of[lane], *skip = skip_of[lane];

            auto prim = [&](int op) -> void * { return primary[op] >= 0 ? dvals[primary[op]] : nullptr; };
            // A base quantity (sum / integral) feeds its own op and a derived op that divides it in the epilogue
            // (mean, weighted averages).  It lands in the base op's output unless a Factor rescales that one in
            // place while a derived op still needs the raw values.
            auto pick_dst = [&](int base_op, int derived_op) -> double * {
                const int pb = primary[base_op], pd = primary[derived_op];
                if (pb >= 0 && (specs[pb].nfactors == 0 || count[derived_op] == 0)) return (double *)dvals[pb];
                if (count[derived_op] == 1 && count[base_op] == 0) return (double *)dvals[pd];  // divided in place
                if (count[derived_op] > 0) return (double *)arena_take(ctx, wv);
                return nullptr;
            };
            auto copy_dups = [&](int op, const void *src) -> int32_t {  // duplicates of an (op, column) pair
                for (int j = 0; j < nspecs; ++j) {
                    if (specs[j].col != c || specs[j].op != op || !src || src == dvals[j]) continue;
                    CK(cudaMemcpyAsync(dvals[j], src, (size_t)W * 8, cudaMemcpyDeviceToDevice, st));
                }
                return BOWGPU_OK;
            };
            if (fam == 0 && both) {
                FullLaunch L;
                memset(&L, 0, sizeof L);
                L.time = (const int64_t *)f->cols[r->time_col].values;
                L.values = dc.values;
                L.validity = dc.validity;
                L.is_int = dc.dtype == BOWGPU_INT64;
                L.g = g;
                L.carry_head = carry;
                L.carry_tail = carry + full_carry_bytes(g.n) / 2;
                L.skip = skip;
                L.status = ctx->d_status;
                if (syn_by_col) L.syn = syn_by_col[c];
                L.gate = gate;
                L.gate_lanes = width;
                int64_t *cnt = nullptr;
                if (primary[BOWGPU_AGG_COUNT] >= 0 && specs[primary[BOWGPU_AGG_COUNT]].nfactors == 0)
                    cnt = (int64_t *)dvals[primary[BOWGPU_AGG_COUNT]];
                else
                    cnt = (int64_t *)arena_take(ctx, wv);
                CK(cudaMemsetAsync(cnt, 0, (size_t)W * 8, st));
                L.out_basic.cnt = cnt;
                double *sum_dst = pick_dst(BOWGPU_AGG_SUM, BOWGPU_AGG_MEAN);
                L.out_basic.sum = sum_dst;
                L.out_basic.mn = (double *)prim(BOWGPU_AGG_MIN);
                L.out_basic.mx = (double *)prim(BOWGPU_AGG_MAX);
                L.out_basic.first = (uint64_t *)prim(BOWGPU_AGG_FIRST);
                L.out_basic.last = (uint64_t *)prim(BOWGPU_AGG_LAST);
                // the merged kernel always forms both integrals: the ones nobody asked for land in scratch
                L.out_integral.n_step = (int64_t *)arena_take(ctx, wv);
                L.out_integral.n_trap = (int64_t *)arena_take(ctx, wv);
                CK(cudaMemsetAsync(L.out_integral.n_step, 0, (size_t)W * 8, st));
                CK(cudaMemsetAsync(L.out_integral.n_trap, 0, (size_t)W * 8, st));
                double *step_dst = pick_dst(BOWGPU_AGG_INTEGRAL_STEP, BOWGPU_AGG_WAVG_STEP);
                double *trap_dst = pick_dst(BOWGPU_AGG_INTEGRAL_TRAPEZOID, BOWGPU_AGG_WAVG_LINEAR);
                L.out_integral.step = step_dst ? step_dst : (double *)arena_take(ctx, wv);
                L.out_integral.trap = trap_dst ? trap_dst : (double *)arena_take(ctx, wv);
                if (!L.out_integral.n_step || !L.out_integral.n_trap || !L.out_integral.step || !L.out_integral.trap)
                    return fail(ctx, BOWGPU_ENOMEM, "aggregate: scratch");
                cudaEvent_t e0, e1;
                timing_main_pair(ctx, &e0, &e1);
                CK(launch_segreduce_full(L, sm_share, st, e0, e1));
                count_launch(ctx, 1, true);
                count_launch(ctx, 1);
                if ((rc = copy_dups(BOWGPU_AGG_COUNT, cnt))) return rc;
                if ((rc = copy_dups(BOWGPU_AGG_SUM, sum_dst))) return rc;
                for (int op : {BOWGPU_AGG_MIN, BOWGPU_AGG_MAX, BOWGPU_AGG_FIRST, BOWGPU_AGG_LAST})
                    if ((rc = copy_dups(op, prim(op)))) return rc;
                if ((rc = copy_dups(BOWGPU_AGG_INTEGRAL_STEP, step_dst))) return rc;
                if ((rc = copy_dups(BOWGPU_AGG_INTEGRAL_TRAPEZOID, trap_dst))) return rc;
                for (int j = 0; j < nspecs; ++j) {
                    if (specs[j].col != c) continue;
                    if (agg_is_basic(specs[j].op)) {
                        spec_cnt[j] = cnt;
                        if (specs[j].op == BOWGPU_AGG_MEAN) spec_src[j] = sum_dst;
                    }
                    switch (specs[j].op) {
                    case BOWGPU_AGG_INTEGRAL_STEP: spec_cnt[j] = L.out_integral.n_step; break;
                    case BOWGPU_AGG_INTEGRAL_TRAPEZOID: spec_cnt[j] = L.out_integral.n_trap; break;
                    case BOWGPU_AGG_WAVG_STEP: spec_cnt[j] = L.out_integral.n_step, spec_src[j] = L.out_integral.step; break;
                    case BOWGPU_AGG_WAVG_LINEAR: spec_cnt[j] = L.out_integral.n_trap, spec_src[j] = L.out_integral.trap; break;
                    default: break;
                    }
                }
            } else if (fam == 0) {
                SegLaunch L;
                memset(&L, 0, sizeof L);
                L.time = (const int64_t *)f->cols[r->time_col].values;
                L.values = dc.values;
                L.validity = dc.validity;
                L.is_int = dc.dtype == BOWGPU_INT64;
                L.g = g;
                L.carry_head = (BasicCarry *)carry;
                L.carry_tail = (BasicCarry *)carry + seg_num_tiles(g.n);
                L.skip = (BasicCarry *)skip;
                L.status = ctx->d_status;
                if (syn_by_col) L.syn = syn_by_col[c];
                L.gate = gate;
                L.gate_lanes = width;
                // the valid-row count drives every validity bitmap: Count output if it can be used as is, else scratch
                int64_t *cnt = nullptr;
                if (primary[BOWGPU_AGG_COUNT] >= 0 && specs[primary[BOWGPU_AGG_COUNT]].nfactors == 0)
                    cnt = (int64_t *)dvals[primary[BOWGPU_AGG_COUNT]];
                else
                    cnt = (int64_t *)arena_take(ctx, wv);
                CK(cudaMemsetAsync(cnt, 0, (size_t)W * 8, st));
                L.out.cnt = cnt;
                double *sum_dst = pick_dst(BOWGPU_AGG_SUM, BOWGPU_AGG_MEAN);
                L.out.sum = sum_dst;
                L.out.mn = (double *)prim(BOWGPU_AGG_MIN);
                L.out.mx = (double *)prim(BOWGPU_AGG_MAX);
                L.out.first = (uint64_t *)prim(BOWGPU_AGG_FIRST);
                L.out.last = (uint64_t *)prim(BOWGPU_AGG_LAST);
                L.ops = OPS_SUMCNT;
                if (L.out.mn || L.out.mx) L.ops |= OPS_MINMAX;
                if (L.out.first || L.out.last) L.ops |= OPS_FIRSTLAST;
                cudaEvent_t e0, e1;
                timing_main_pair(ctx, &e0, &e1);
                CK(launch_segreduce_basic(L, sm_share, st, e0, e1));
                count_launch(ctx, 1, true);
                count_launch(ctx, 1);
                if ((rc = copy_dups(BOWGPU_AGG_COUNT, cnt))) return rc;
                if ((rc = copy_dups(BOWGPU_AGG_SUM, sum_dst))) return rc;
                for (int op : {BOWGPU_AGG_MIN, BOWGPU_AGG_MAX, BOWGPU_AGG_FIRST, BOWGPU_AGG_LAST})
                    if ((rc = copy_dups(op, prim(op)))) return rc;
                for (int j = 0; j < nspecs; ++j)
                    if (specs[j].col == c && agg_is_basic(specs[j].op)) {
                        spec_cnt[j] = cnt;
                        if (specs[j].op == BOWGPU_AGG_MEAN) spec_src[j] = sum_dst;  // every mean divides the shared sums
                    }
            } else {
                IntLaunch L;
                memset(&L, 0, sizeof L);
                L.time = (const int64_t *)f->cols[r->time_col].values;
                L.values = dc.values;
                L.validity = dc.validity;
                L.is_int = dc.dtype == BOWGPU_INT64;
                L.g = g;
                L.carry_head = carry;
                L.carry_tail = carry + integral_carry_bytes(g.n) / 2;
                L.skip = skip;
                L.status = ctx->d_status;
                if (syn_by_col) L.syn = syn_by_col[c];
                L.gate = gate;
                L.gate_lanes = width;
                const bool want_step = count[BOWGPU_AGG_INTEGRAL_STEP] || count[BOWGPU_AGG_WAVG_STEP];
                const bool want_trap = count[BOWGPU_AGG_INTEGRAL_TRAPEZOID] || count[BOWGPU_AGG_WAVG_LINEAR];
                if (want_step) {
                    L.out.n_step = (int64_t *)arena_take(ctx, wv);
                    CK(cudaMemsetAsync(L.out.n_step, 0, (size_t)W * 8, st));
                    L.out.step = pick_dst(BOWGPU_AGG_INTEGRAL_STEP, BOWGPU_AGG_WAVG_STEP);
                }
                if (want_trap) {
                    L.out.n_trap = (int64_t *)arena_take(ctx, wv);
                    CK(cudaMemsetAsync(L.out.n_trap, 0, (size_t)W * 8, st));
                    L.out.trap = pick_dst(BOWGPU_AGG_INTEGRAL_TRAPEZOID, BOWGPU_AGG_WAVG_LINEAR);
                }
                cudaEvent_t e0, e1;
                timing_main_pair(ctx, &e0, &e1);
                CK(launch_segreduce_integral(L, sm_share, st, e0, e1));
                count_launch(ctx, 1, true);
                count_launch(ctx, 1);
                if ((rc = copy_dups(BOWGPU_AGG_INTEGRAL_STEP, L.out.step))) return rc;
                if ((rc = copy_dups(BOWGPU_AGG_INTEGRAL_TRAPEZOID, L.out.trap))) return rc;
                for (int j = 0; j < nspecs; ++j) {
                    if (specs[j].col != c) continue;
                    switch (specs[j].op) {
                    case BOWGPU_AGG_INTEGRAL_STEP: spec_cnt[j] = L.out.n_step; break;
                    case BOWGPU_AGG_INTEGRAL_TRAPEZOID: spec_cnt[j] = L.out.n_trap; break;
                    case BOWGPU_AGG_WAVG_STEP: spec_cnt[j] = L.out.n_step, spec_src[j] = L.out.step; break;
                    case BOWGPU_AGG_WAVG_LINEAR: spec_cnt[j] = L.out.n_trap, spec_src[j] = L.out.trap; break;
                    default: break;
                    }
                }
            }
            if (width > 1) {  // the main stream goes on when every lane of the wave is through
                CK(cudaEventRecord(ctx->side_done[lane], st));
                CK(cudaStreamWaitEvent(ctx->stream, ctx->side_done[lane], 0));
            }
            ++lane;
            --left;
        }
    }
    // ---- epilogue: specs without Factor are grouped by the count array that decides their validity (fast path);
    // the others go through the generic per-spec pass
    {
        std::vector<EpiGroup> groups;
        std::vector<EpilogueSpec> generic;
        auto group_for = [&](const int64_t *cnt, auto full) -> EpiGroup & {
            for (auto &G : groups)
                if (G.cnt == cnt && !full(G)) return G;
            EpiGroup G;
            memset(&G, 0, sizeof G);
            G.cnt = cnt;
            groups.push_back(G);
            return groups.back();
        };
        for (int j = 0; j < nspecs; ++j) {
            const int op = specs[j].op;
            if (r->whole && op == BOWGPU_AGG_WINDOW_START && f->cols[specs[j].col].dtype != BOWGPU_INT64) {
                // whole.go:87 SetOrDropStrict: the int64 window start does not fit the Float64 buffer -> null
                CK(cudaMemsetAsync(dvals[j], 0, (size_t)W * 8, ctx->stream));
                CK(cudaMemsetAsync(dbits[j], 0, (size_t)((W + 7) / 8), ctx->stream));
                continue;
            }
            if (specs[j].nfactors > 0) {
                EpilogueSpec e;
                memset(&e, 0, sizeof e);
                e.op = op;
                e.out_is_int = outs[j].dtype == BOWGPU_INT64;
                e.cnt = spec_cnt[j];
                e.sum_src = spec_src[j];
                e.values = dvals[j];
                e.validity = dbits[j];
                e.nfactors = specs[j].nfactors;
                for (int i = 0; i < e.nfactors; ++i) e.factors[i] = specs[j].factors[i];
                generic.push_back(e);
                continue;
            }
            const bool always = op == BOWGPU_AGG_WINDOW_START || op == BOWGPU_AGG_COUNT || op == BOWGPU_AGG_SUM;
            const bool divides = op == BOWGPU_AGG_MEAN || op == BOWGPU_AGG_WAVG_STEP || op == BOWGPU_AGG_WAVG_LINEAR;
            const bool zeroes = !divides && op != BOWGPU_AGG_WINDOW_START && op != BOWGPU_AGG_COUNT;
            // WindowStart does not depend on any count: it rides along with the first group that has room
            const int64_t *gkey = spec_cnt[j];
            if (op == BOWGPU_AGG_WINDOW_START)
                for (int jj = 0; jj < nspecs && !gkey; ++jj)
                    if (specs[jj].nfactors == 0) gkey = spec_cnt[jj];
            EpiGroup &G = group_for(gkey, [&](const EpiGroup &g_) {
                return (always ? g_.n_bm_all >= EPIG_ALL : g_.n_bm_cnt >= EPIG_BM) || (divides && g_.n_div >= EPIG_DIV) ||
                       (zeroes && g_.n_null >= EPIG_NULL) || (op == BOWGPU_AGG_WINDOW_START && g_.n_ws >= EPIG_WS);
            });
            if (always)
                G.bm_all[G.n_bm_all++] = dbits[j];
            else
                G.bm_cnt[G.n_bm_cnt++] = dbits[j];
            if (divides) {
                G.div_dst[G.n_div] = (double *)dvals[j];
                G.div_src[G.n_div] = spec_src[j];
                G.div_by_cnt[G.n_div++] = op == BOWGPU_AGG_MEAN;
            }
            if (zeroes) G.null_vals[G.n_null++] = (uint64_t *)dvals[j];
            if (op == BOWGPU_AGG_WINDOW_START) G.ws[G.n_ws++] = (int64_t *)dvals[j];
        }
        for (const auto &G : groups) {
            CK(launch_epilogue_group(G, g, ctx->stream));
            count_launch(ctx);
        }
        if (!generic.empty()) {
            CK(launch_epilogue(generic.data(), (int)generic.size(), g, ctx->stream));
            count_launch(ctx, ((int)generic.size() + 15) / 16);
        }
    }
    timing_end(ctx);
    if (mem == BOWGPU_MEM_DEVICE) return BOWGPU_OK;  // asynchronous: errors surface in bowgpu_ctx_synchronize
    for (int j = 0; j < nspecs; ++j) {
        rc = copy_d2h(ctx, outs[j].values, dvals[j], (size_t)W * 8);
        if (rc) return rc;
        rc = copy_d2h(ctx, outs[j].validity, dbits[j], (size_t)((W + 7) / 8));
        if (rc) return rc;
    }
    return check_status(ctx);
}

// Rolling.Aggregate on the segmc kernels (segmc.cuh): the input columns are grouped by (dtype, has nulls), every group is
// ONE streaming launch that reads the time column once and each of its value columns once and writes the final values of
// all their aggregations; launch_finish then derives the validity bitmap of every output, the WindowStart columns and the
// zeros of windows without rows.
static int32_t aggregate_core_mc(bowgpu_rolling *r, const bowgpu_agg_spec *specs, int32_t nspecs, bowgpu_out_col *outs,
                                 int32_t mem, const FusedSyn *syn_by_col) {
    if (!r || !specs || !outs || nspecs <= 0) return BOWGPU_EINVAL;
    bowgpu_frame *f = r->frame;
    bowgpu_ctx *ctx = f->ctx;
    Guard gd(ctx);
    const int ncols = (int)f->cols.size();
    bool keeps_interval = false, inclusive_eff = r->inclusive != 0;
    for (int j = 0; j < nspecs; ++j) {  // validateAggregation, aggregation.go:171-188
        if (specs[j].col < 0 || specs[j].col >= ncols) return fail(ctx, BOWGPU_EINVAL, "aggregation %d: no column %d", j, specs[j].col);
        if (specs[j].op < 0 || specs[j].op >= BOWGPU_AGG__COUNT) return fail(ctx, BOWGPU_EUNSUPPORTED, "aggregation %d: unknown opcode %d", j, specs[j].op);
        if (specs[j].nfactors < 0 || specs[j].nfactors > 4) return fail(ctx, BOWGPU_EINVAL, "aggregation %d: at most 4 factors", j);
        if (bowgpu_agg_needs_inclusive(specs[j].op)) inclusive_eff = true;
        if (specs[j].col == r->time_col) keeps_interval = true;
    }
    if (!keeps_interval && !r->whole)
        return fail(ctx, BOWGPU_ENOINTERVALCOL, "must keep interval column");  // aggregation.go:163-166
    for (int j = 0; j < nspecs; ++j)
        outs[j].dtype = r->whole ? whole_return_type(specs[j].op, f->cols[specs[j].col].dtype)
                                 : bowgpu_agg_return_type(specs[j].op, f->cols[specs[j].col].dtype);
    const WindowGeom g = make_geom(r, inclusive_eff);
    const int64_t W = g.W;
    if (W == 0) return BOWGPU_OK;

    // ---- plan -------------------------------------------------------------------------------------------------------------
    std::vector<int> cols_used;
    for (int j = 0; j < nspecs; ++j) {
        if (specs[j].op == BOWGPU_AGG_WINDOW_START) continue;
        if (std::find(cols_used.begin(), cols_used.end(), specs[j].col) == cols_used.end()) cols_used.push_back(specs[j].col);
    }
    const size_t wv = align_up((size_t)W * 8, 256), wb = align_up((size_t)((W + 7) / 8) + 16, 256);
    const size_t ww = align_up((size_t)((W + 31) / 32) * 4 + 16, 256);  // one bitmap, in words
    const size_t rec_bytes = align_up((size_t)2 * mc_max_chunks(ctx->sm_count) * sizeof(McCarry), 256);
    size_t need = 8192 + (1 + 2 * cols_used.size()) * (ww + 256) + cols_used.size() * (rec_bytes + 256);
    if (mem == BOWGPU_MEM_HOST) need += (size_t)nspecs * (wv + wb + 512);
    int32_t rc = arena_reserve(ctx, need);
    if (rc) return rc;
    arena_reset(ctx);
    std::vector<void *> dvals(nspecs);
    std::vector<uint8_t *> dbits(nspecs);
    for (int j = 0; j < nspecs; ++j) {
        if (mem == BOWGPU_MEM_DEVICE) {
            dvals[j] = outs[j].values;
            dbits[j] = outs[j].validity;
        } else {
            dvals[j] = arena_take(ctx, wv);
            dbits[j] = (uint8_t *)arena_take(ctx, wb);
        }
        if (!dvals[j] || !dbits[j]) return fail(ctx, BOWGPU_EINVAL, "aggregation %d: null output buffer", j);
    }
    // bitmaps set by the streaming kernels: touched, then (valid rows, trapezoid defined) per used column - one memset
    uint8_t *bm_base = (uint8_t *)arena_take(ctx, (1 + 2 * cols_used.size()) * ww);
    if (!bm_base) return fail(ctx, BOWGPU_ENOMEM, "aggregate: scratch");
    uint32_t *touched = (uint32_t *)bm_base;
    timing_begin(ctx);
    CK(cudaMemsetAsync(bm_base, 0, (1 + 2 * cols_used.size()) * ww, ctx->stream));

    struct ColPlan {
        int col;
        McColArgs a;
        uint32_t bops, iops;
        int primary[BOWGPU_AGG__COUNT];
    };
    std::vector<ColPlan> plans(cols_used.size());
    for (size_t u = 0; u < cols_used.size(); ++u) {
        ColPlan &P = plans[u];
        const int c = cols_used[u];
        const DevCol &dc = f->cols[c];
        P.col = c;
        memset(&P.a, 0, sizeof P.a);
        P.bops = P.iops = 0;
        for (int o = 0; o < BOWGPU_AGG__COUNT; ++o) P.primary[o] = -1;
        for (int j = 0; j < nspecs; ++j)  // the first spec of each op on this column receives the kernel output
            if (specs[j].col == c && specs[j].op != BOWGPU_AGG_WINDOW_START && P.primary[specs[j].op] < 0) P.primary[specs[j].op] = j;
        auto prim = [&](int op) -> void * { return P.primary[op] >= 0 ? dvals[P.primary[op]] : nullptr; };
        P.a.values = dc.values;
        P.a.validity = dc.validity;
        McColOut &o = P.a.out;
        o.cnt = (int64_t *)prim(BOWGPU_AGG_COUNT);
        o.sum = (double *)prim(BOWGPU_AGG_SUM);
        o.mean = (double *)prim(BOWGPU_AGG_MEAN);
        o.mn = (double *)prim(BOWGPU_AGG_MIN);
        o.mx = (double *)prim(BOWGPU_AGG_MAX);
        o.first = (uint64_t *)prim(BOWGPU_AGG_FIRST);
        o.last = (uint64_t *)prim(BOWGPU_AGG_LAST);
        o.step = (double *)prim(BOWGPU_AGG_INTEGRAL_STEP);
        o.wstep = (double *)prim(BOWGPU_AGG_WAVG_STEP);
        o.trap = (double *)prim(BOWGPU_AGG_INTEGRAL_TRAPEZOID);
        o.wlin = (double *)prim(BOWGPU_AGG_WAVG_LINEAR);
        o.vb_cnt = (uint32_t *)(bm_base + (1 + 2 * u) * ww);
        o.vb_trap = (uint32_t *)(bm_base + (2 + 2 * u) * ww);
        if (o.cnt || o.sum || o.mean) P.bops |= MC_SUM;
        if (o.mn || o.mx) P.bops |= MC_SUM | MC_MINMAX;
        if (o.first || o.last) P.bops |= MC_SUM | MC_MINMAX | MC_FIRSTLAST;
        if (o.step || o.wstep) P.iops |= MC_STEP;
        if (o.trap || o.wlin) P.iops |= MC_TRAP;
        if (syn_by_col) P.a.syn = syn_by_col[c];
        P.a.rec = (McCarry *)arena_take(ctx, rec_bytes);
        if (!P.a.rec) return fail(ctx, BOWGPU_ENOMEM, "aggregate: scratch");
    }
    // ---- one streaming launch per (dtype, nulls) group of at most MC_MAXC columns ------------------------------------------
    {
        std::vector<bool> done(plans.size(), false);
        for (size_t u = 0; u < plans.size(); ++u) {
            if (done[u]) continue;
            const DevCol &d0 = f->cols[plans[u].col];
            McLaunch L;
            memset(&L, 0, sizeof L);
            L.time = (const int64_t *)f->cols[r->time_col].values;
            L.g = g;
            L.is_int = d0.dtype == BOWGPU_INT64;
            L.touched = touched;
            L.status = ctx->d_status;
            for (size_t v = u; v < plans.size() && L.ncols < MC_MAXC; ++v) {
                const DevCol &dv = f->cols[plans[v].col];
                if (done[v] || (dv.dtype == BOWGPU_INT64) != (L.is_int != 0) || (dv.validity != nullptr) != (d0.validity != nullptr)) continue;
                L.col[L.ncols++] = plans[v].a;
                L.bops |= plans[v].bops;
                L.iops |= plans[v].iops;
                done[v] = true;
            }
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            timing_main_pair(ctx, &e0, &e1);
            CK(launch_segmc(L, ctx->sm_count, ctx->stream, e0, e1));
            count_launch(ctx, 1, true);
            count_launch(ctx, 1);
        }
    }
    // ---- per-output finishing: validity bitmaps, WindowStart, zeros of windows without rows ---------------------------------
    {
        std::vector<FinishDst> dst;
        for (int j = 0; j < nspecs; ++j) {
            const int op = specs[j].op;
            if (r->whole && op == BOWGPU_AGG_WINDOW_START && f->cols[specs[j].col].dtype != BOWGPU_INT64) {
                // whole.go:87 SetOrDropStrict: the int64 window start does not fit the Float64 buffer -> null
                CK(cudaMemsetAsync(dvals[j], 0, (size_t)W * 8, ctx->stream));
                CK(cudaMemsetAsync(dbits[j], 0, (size_t)((W + 7) / 8), ctx->stream));
                continue;
            }
            FinishDst D;
            memset(&D, 0, sizeof D);
            D.values = (uint64_t *)dvals[j];
            D.validity = dbits[j];
            if (op == BOWGPU_AGG_WINDOW_START) {
                D.kind = FIN_WINDOW_START;
            } else {
                size_t u = std::find(cols_used.begin(), cols_used.end(), specs[j].col) - cols_used.begin();
                const ColPlan &P = plans[u];
                D.zero_empty = syn_by_col == nullptr;
                if (op == BOWGPU_AGG_COUNT || op == BOWGPU_AGG_SUM) {
                    D.kind = FIN_ALWAYS;
                } else {
                    D.kind = FIN_SRC;
                    D.src = (op == BOWGPU_AGG_INTEGRAL_TRAPEZOID || op == BOWGPU_AGG_WAVG_LINEAR) ? P.a.out.vb_trap : P.a.out.vb_cnt;
                }
                // duplicates of an (op, column) pair: copies of the first one
                if (P.primary[op] != j)
                    CK(cudaMemcpyAsync(dvals[j], dvals[P.primary[op]], (size_t)W * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            }
            dst.push_back(D);
        }
        int nl = 0;
        if (!dst.empty()) CK(launch_finish(dst.data(), (int)dst.size(), touched, g, ctx->stream, &nl));
        count_launch(ctx, nl);
        for (int j = 0; j < nspecs; ++j)
            if (specs[j].nfactors > 0) {  // transformation.Factor, in place on the valid slots
                CK(launch_factor((uint64_t *)dvals[j], dbits[j], W, outs[j].dtype == BOWGPU_INT64, specs[j].nfactors, specs[j].factors,
                                 ctx->stream));
                count_launch(ctx);
            }
    }
    timing_end(ctx);
    if (mem == BOWGPU_MEM_DEVICE) return BOWGPU_OK;  // asynchronous: errors surface in bowgpu_ctx_synchronize
    for (int j = 0; j < nspecs; ++j) {
        rc = copy_d2h(ctx, outs[j].values, dvals[j], (size_t)W * 8);
        if (rc) return rc;
        rc = copy_d2h(ctx, outs[j].validity, dbits[j], (size_t)((W + 7) / 8));
        if (rc) return rc;
    }
    return check_status(ctx);
}

// Two implementations of Rolling.Aggregate live side by side (same results, same ABI):
//   v1 (default)  one streaming launch per input column and kernel family (segreduce.cuh) + epilogue pass
//   mc            BOWGPU_SEG_IMPL=mc: the multi-column kernels of segmc.cuh (time read once per launch, final values
//                 written in place).  Parity-green, but measured slower than v1 on B200: opt-in only.
static bool use_mc() {
    static const bool mc = [] {
        const char *e = getenv("BOWGPU_SEG_IMPL");
        return e && strcmp(e, "mc") == 0;
    }();
    return mc;
}
static int32_t aggregate_dispatch(bowgpu_rolling *r, const bowgpu_agg_spec *specs, int32_t nspecs, bowgpu_out_col *outs,
                                  int32_t mem, const FusedSyn *syn_by_col) {
    return use_mc() ? aggregate_core_mc(r, specs, nspecs, outs, mem, syn_by_col) : aggregate_core(r, specs, nspecs, outs, mem, syn_by_col);
}

extern "C" int32_t bowgpu_rolling_aggregate(bowgpu_rolling *r, const bowgpu_agg_spec *specs, int32_t nspecs,
                                            bowgpu_out_col *outs, int32_t mem) {
    return aggregate_dispatch(r, specs, nspecs, outs, mem, nullptr);
}

// aggregation.Aggregate(b, intervalCol, aggrs...) (rolling/aggregation/whole.go:12-93): every aggregation over ONE window
// that holds the whole frame.  Runs the same kernels as Rolling.Aggregate with an interval that spans all rows.
extern "C" int32_t bowgpu_frame_aggregate_whole(bowgpu_frame *frame, int32_t time_col, const bowgpu_agg_spec *specs,
                                                int32_t nspecs, bowgpu_out_col *outs, int32_t mem) {
    if (!frame || !specs || !outs || nspecs <= 0) return BOWGPU_EINVAL;
    bowgpu_ctx *ctx = frame->ctx;
    Guard gd(ctx);
    const int ncols = (int)frame->cols.size();
    if (time_col < 0 || time_col >= ncols) return fail(ctx, BOWGPU_EINVAL, "time column index %d out of range", time_col);
    if (frame->cols[time_col].dtype != BOWGPU_INT64)
        return fail(ctx, BOWGPU_ETYPE, "the GPU path needs an Int64 interval column");
    for (int j = 0; j < nspecs; ++j) {
        if (specs[j].col < 0 || specs[j].col >= ncols) return fail(ctx, BOWGPU_EINVAL, "column aggregation %d: no column %d", j, specs[j].col);
        if (specs[j].op < 0 || specs[j].op >= BOWGPU_AGG__COUNT) return fail(ctx, BOWGPU_EUNSUPPORTED, "column aggregation %d: unknown opcode %d", j, specs[j].op);
        outs[j].dtype = whole_return_type(specs[j].op, frame->cols[specs[j].col].dtype);
    }
    const int64_t n = frame->n;
    if (n == 0) return BOWGPU_OK;  // whole.go:49-50: zero-length output columns
    const DevCol &tc = frame->cols[time_col];
    if (tc.validity && tc.null_count != 0)
        return fail(ctx, BOWGPU_ENULLTIME, "interval column holds %lld nulls: the GPU path requires a non-null, sorted interval column",
                    (long long)tc.null_count);
    int64_t ends[2];
    CK(cudaMemcpyAsync(&ends[0], tc.values, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&ends[1], tc.values + (n - 1), 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ends[1] < ends[0]) return fail(ctx, BOWGPU_EUNSORTED, "time column is not sorted ascending");
    if (ends[1] == INT64_MAX) return fail(ctx, BOWGPU_EUNSUPPORTED, "last time value is the largest int64");
    bowgpu_rolling r;
    r.frame = frame;
    r.time_col = time_col;
    r.interval = (int64_t)((uint64_t)ends[1] - (uint64_t)ends[0] + 1);  // all rows fall into window 0, none at its end
    if (r.interval <= 0) return fail(ctx, BOWGPU_EUNSUPPORTED, "time span does not fit 63 bits");
    r.s0 = ends[0];
    r.W = 1;
    r.t_first = ends[0];
    r.t_last = ends[1];
    r.whole = true;
    r.whole_first = f64_to_i64_go((double)ends[0]);  // int64(firstValue) through GetNextFloat64, whole.go:54-70
    r.whole_last = f64_to_i64_go((double)ends[1]);
    return bowgpu_rolling_aggregate(&r, specs, nspecs, outs, mem);
}

extern "C" int32_t bowgpu_rolling_interpolate(bowgpu_rolling *r, const int32_t *ops, int32_t nops, bowgpu_frame **out_frame,
                                              int64_t *n_out) {
    if (!r || !ops || !out_frame || !n_out) return BOWGPU_EINVAL;
    *out_frame = nullptr;
    *n_out = 0;
    bowgpu_frame *f = r->frame;
    bowgpu_ctx *ctx = f->ctx;
    Guard gd(ctx);
    const int ncols = (int)f->cols.size();
    // the reference appends [start row] ++ window by column POSITION (bowappend.go:28-47): the interpolations must
    // name every column of the Bow in schema order (which also keeps the interval column, interpolation.go:49-51)
    if (nops != ncols)
        return fail(ctx, BOWGPU_EINVAL, "interpolations must name every column in schema order (%d given, %d columns)", nops, ncols);
    if (ncols > INTERP_MAX_COLS) return fail(ctx, BOWGPU_EUNSUPPORTED, "interpolate supports at most %d columns", INTERP_MAX_COLS);
    for (int j = 0; j < nops; ++j) {  // validateInterpolation, interpolation.go:71-96
        if (ops[j] < 0 || ops[j] > BOWGPU_INTERP_STEP_NEXT) return fail(ctx, BOWGPU_EUNSUPPORTED, "interpolation %d: unknown opcode %d", j, ops[j]);
        if (ops[j] == BOWGPU_INTERP_WINDOW_START && f->cols[j].dtype != BOWGPU_INT64)
            return fail(ctx, BOWGPU_ETYPE, "interpolation %d: WindowStart accepts types [int64], got type float64", j);
    }
    const WindowGeom g = make_geom(r, r->inclusive);
    const int64_t W = g.W;
    bowgpu_frame *of = new (std::nothrow) bowgpu_frame();
    if (!of) return BOWGPU_ENOMEM;
    of->ctx = ctx;
    of->cols.resize(ncols);
    for (int j = 0; j < ncols; ++j) of->cols[j].dtype = f->cols[j].dtype;
    auto bail = [&](int32_t code) {
        cudaStreamSynchronize(ctx->stream);
        for (auto &c : of->cols) free_col(ctx, c);
        delete of;
        return code;
    };
    if (W == 0 || g.n == 0) {  // interpolation.go:59-61: empty result keeps the schema
        *out_frame = of;
        return BOWGPU_OK;
    }
    const size_t wv = align_up((size_t)(W + 1) * 8, 256), wb = align_up((size_t)W + 16, 256);
    const size_t need = 8192 + 3 * wv + (size_t)ncols * (wv + wb) + scan_scratch_bytes(W) + 1024;
    int32_t rc = arena_reserve(ctx, need);
    if (rc) return bail(rc);
    arena_reset(ctx);
    InterpLaunch L;
    memset(&L, 0, sizeof L);
    L.time = (const int64_t *)f->cols[r->time_col].values;
    int64_t *d_first = (int64_t *)arena_take(ctx, wv);
    L.first = d_first;
    L.off = (int64_t *)arena_take(ctx, wv);
    L.wsrc = (int64_t *)arena_take(ctx, wv);
    int64_t *scan_tmp = (int64_t *)arena_take(ctx, scan_scratch_bytes(W));
    L.g = g;
    L.inclusive = r->inclusive;
    L.ncols = ncols;
    if (r->has_prev) {
        L.prev_time = (int64_t)r->prev[r->time_col].bits;
        L.prev_time_valid = r->prev[r->time_col].valid;
    }
    for (int j = 0; j < ncols; ++j) {
        InterpCol &c = L.cols[j];
        c.values = f->cols[j].values;
        c.validity = (const uint32_t *)f->cols[j].validity;
        c.syn_val = (uint64_t *)arena_take(ctx, wv);
        c.syn_ok = (uint8_t *)arena_take(ctx, wb);
        c.op = ops[j];
        c.is_int = f->cols[j].dtype == BOWGPU_INT64;
        if (r->has_prev) {
            c.prev_bits = r->prev[j].bits;
            c.prev_valid = r->prev[j].valid;
        }
    }
    timing_begin(ctx);
    BoundsLaunch B;
    B.time = L.time;
    B.g = g;
    B.first = d_first;
    B.status = ctx->d_status;
    int e = launch_bounds(B, ctx->sm_count, ctx->stream, nullptr, nullptr);
    void *pyr = attach_pyramids(ctx, L, g.n);
    if (!e) e = launch_interp_windows(L, ctx->stream);
    pool_free(ctx, pyr);
    if (!e) e = launch_exclusive_scan(L.off, W, scan_tmp, ctx->stream);
    count_launch(ctx, 5);
    int64_t total = 0;
    if (e || cudaMemcpyAsync(&total, L.off + W, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return bail(fail(ctx, BOWGPU_ECUDA, "interpolate (windows/scan): %s", cudaGetErrorString(cudaGetLastError())));
    of->n = total;
    for (int j = 0; j < ncols; ++j) {
        const bool may_null = f->cols[j].validity != nullptr || ops[j] != BOWGPU_INTERP_WINDOW_START;
        rc = alloc_col(ctx, of->cols[j], total, f->cols[j].dtype, may_null);
        if (rc) return bail(rc);
        L.cols[j].out_values = of->cols[j].values;
        L.cols[j].out_validity = (uint32_t *)of->cols[j].validity;
    }
    cudaEvent_t e0, e1;
    timing_main_pair(ctx, &e0, &e1);
    // the tile -> window table lives in its own pool block: the arena may not grow while its buffers are in use
    int64_t *tile_k = nullptr;
    if (total > 0 && pool_alloc(ctx, (void **)&tile_k, (size_t)(interp_gather_tiles(total) + 1) * 8) != cudaSuccess)
        return bail(fail(ctx, BOWGPU_ENOMEM, "interpolate: tile table"));
    e = launch_interp_gather(L, total, tile_k, ctx->stream, e0, e1);
    pool_free(ctx, tile_k);
    count_launch(ctx, 2);
    count_launch(ctx, 1, true);
    timing_end(ctx);
    if (e) return bail(fail(ctx, BOWGPU_ECUDA, "interpolate (gather): %s", cudaGetErrorString((cudaError_t)e)));
    {  // bitmaps without nulls are dropped (the aggregation kernels run faster): one popcount per column, one copy
        unsigned long long *d_cnt = nullptr, h_cnt[INTERP_MAX_COLS];
        if (pool_alloc(ctx, (void **)&d_cnt, sizeof h_cnt) != cudaSuccess) return bail(fail(ctx, BOWGPU_ENOMEM, "interpolate: counters"));
        cudaMemsetAsync(d_cnt, 0, sizeof h_cnt, ctx->stream);
        for (int j = 0; j < ncols; ++j)
            if (of->cols[j].validity) launch_bitmap_popcount(of->cols[j].validity, total, d_cnt + j, ctx->stream);
        cudaError_t ce = cudaMemcpyAsync(h_cnt, d_cnt, sizeof h_cnt, cudaMemcpyDeviceToHost, ctx->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
        pool_free(ctx, d_cnt);
        if (ce != cudaSuccess) return bail(fail(ctx, BOWGPU_ECUDA, "interpolate (null counts): %s", cudaGetErrorString(ce)));
        for (int j = 0; j < ncols; ++j) {
            DevCol &c = of->cols[j];
            if (!c.validity) continue;
            c.null_count = total - (int64_t)h_cnt[j];
            if (c.null_count == 0) {
                pool_free(ctx, c.validity);
                c.validity = nullptr;
                c.own_validity = false;
            }
        }
    }
    rc = check_status(ctx);
    if (rc) return bail(rc);
    *out_frame = of;
    *n_out = total;
    return BOWGPU_OK;
}

// Rolling.Interpolate(ops...).Aggregate(specs...) WITHOUT materialising the interpolated frame (interpolation.go:30-161
// followed by aggregation.go:123-238).  bounds -> per-window synthetic start rows (interp_window_kernel) -> the ordinary
// segmented reduction with the synthetic rows injected at window boundaries (segreduce.cuh, FUSED).  Cases the fused
// kernels do not express take the materialising chain, with identical results: Options.Inclusive set by the user,
// rows before the first window start, range-partitioned shards, and "has start" rows that only match S_k through the
// float64 round trip (interpolation.go:121-128).
extern "C" int32_t bowgpu_rolling_interpolate_aggregate(bowgpu_rolling *r, const int32_t *ops, int32_t nops,
                                                        const bowgpu_agg_spec *specs, int32_t nspecs,
                                                        bowgpu_out_col *outs, int32_t mem) {
    if (!r || !ops || !specs || !outs || nspecs <= 0) return BOWGPU_EINVAL;
    bowgpu_frame *f = r->frame;
    bowgpu_ctx *ctx = f->ctx;
    Guard gd(ctx);
    const int ncols = (int)f->cols.size();
    auto unfused = [&]() -> int32_t {
        bowgpu_frame *fi = nullptr;
        int64_t n_out = 0;
        int32_t rc = bowgpu_rolling_interpolate(r, ops, nops, &fi, &n_out);
        if (rc) return rc;
        bowgpu_rolling *r2 = nullptr;
        rc = r->shard ? bowgpu_rolling_create_shard(fi, r->time_col, r->interval, r->s0, r->W, r->inclusive, nullptr, &r2)
                      : bowgpu_rolling_create(fi, r->time_col, r->interval, r->offset, r->inclusive, nullptr, &r2);
        if (rc == BOWGPU_OK && r2->W != r->W)  // the caller sized `outs` for the lattice of r
            rc = fail(ctx, BOWGPU_EUNSUPPORTED, "interpolation changes the window lattice (%lld -> %lld windows): use the two-step calls",
                      (long long)r->W, (long long)r2->W);
        if (rc == BOWGPU_OK) rc = aggregate_dispatch(r2, specs, nspecs, outs, mem, nullptr);
        if (rc == BOWGPU_OK && mem == BOWGPU_MEM_DEVICE) rc = check_status(ctx);  // the frame is about to go
        bowgpu_rolling_destroy(r2);
        bowgpu_frame_destroy(fi);
        return rc;
    };
    // (on a shard the rows before s0 are the left halo; elsewhere they are the negative-timestamp quirk of window 0)
    if (r->inclusive || (r->early_rows > 0 && !r->shard) || r->W == 0 || f->n == 0) return unfused();
    // validation of the interpolations: same rules as bowgpu_rolling_interpolate
    if (nops != ncols)
        return fail(ctx, BOWGPU_EINVAL, "interpolations must name every column in schema order (%d given, %d columns)", nops, ncols);
    if (ncols > INTERP_MAX_COLS) return fail(ctx, BOWGPU_EUNSUPPORTED, "interpolate supports at most %d columns", INTERP_MAX_COLS);
    for (int j = 0; j < nops; ++j) {
        if (ops[j] < 0 || ops[j] > BOWGPU_INTERP_STEP_NEXT) return fail(ctx, BOWGPU_EUNSUPPORTED, "interpolation %d: unknown opcode %d", j, ops[j]);
        if (ops[j] == BOWGPU_INTERP_WINDOW_START && f->cols[j].dtype != BOWGPU_INT64)
            return fail(ctx, BOWGPU_ETYPE, "interpolation %d: WindowStart accepts types [int64], got type float64", j);
    }
    if (ops[r->time_col] != BOWGPU_INTERP_WINDOW_START) return unfused();  // the frame's time column must carry S_k
    const WindowGeom g = make_geom(r, false);
    // a shard also interpolates the start row of the window after its last one (the inclusive row of that last window)
    WindowGeom gx = g;
    gx.W = g.W + (r->shard ? 1 : 0);
    const int64_t W = gx.W;
    // per-window arrays live in pool blocks: the aggregation core re-uses the arena
    const size_t wv = align_up((size_t)(W + 2) * 8, 256), wb = align_up((size_t)W + 16, 256);
    uint8_t *blk = nullptr;
    const size_t total = wv + wb + (size_t)ncols * (wv + wb);
    if (pool_alloc(ctx, (void **)&blk, total) != cudaSuccess) return fail(ctx, BOWGPU_ENOMEM, "fused interpolate: scratch");
    auto bail = [&](int32_t code) {
        pool_free(ctx, blk);
        return code;
    };
    InterpLaunch L;
    memset(&L, 0, sizeof L);
    L.time = (const int64_t *)f->cols[r->time_col].values;
    int64_t *d_first = (int64_t *)blk;
    L.first = d_first;
    L.missing = blk + wv;
    L.status = ctx->d_status;
    L.g = gx;
    L.inclusive = 0;
    L.ncols = ncols;
    if (r->has_prev) {
        L.prev_time = (int64_t)r->prev[r->time_col].bits;
        L.prev_time_valid = r->prev[r->time_col].valid;
    }
    std::vector<FusedSyn> syn(ncols);
    uint8_t *p = blk + wv + wb;
    for (int j = 0; j < ncols; ++j) {
        InterpCol &c = L.cols[j];
        c.values = f->cols[j].values;
        c.validity = (const uint32_t *)f->cols[j].validity;
        c.syn_val = (uint64_t *)p;
        c.syn_ok = p + wv;
        p += wv + wb;
        c.op = ops[j];
        c.is_int = f->cols[j].dtype == BOWGPU_INT64;
        if (r->has_prev) {
            c.prev_bits = r->prev[j].bits;
            c.prev_valid = r->prev[j].valid;
        }
        syn[j].missing = L.missing;
        syn[j].val = c.syn_val;
        syn[j].ok = c.syn_ok;
        syn[j].len = W;
        syn[j].first = d_first;
    }
    // synthetic rows of windows that have a start row are never read, but keep the arrays defined
    if (cudaMemsetAsync(blk + wv, 0, total - wv, ctx->stream) != cudaSuccess) return bail(fail(ctx, BOWGPU_ECUDA, "memset"));
    BoundsLaunch B;
    B.time = L.time;
    B.g = gx;
    B.first = d_first;
    B.status = ctx->d_status;
    // window boundaries: one streaming pass over the time column, or — when windows hold hundreds of rows — a binary
    // search per window (the fused reduction that follows checks the order of every row itself)
    static const int search_mode = [] {
        const char *e_ = getenv("BOWGPU_BOUNDS_SEARCH");  // 0 never, 1 always (tests), default: by rows per window
        return e_ ? atoi(e_) : -1;
    }();
    const bool by_search = search_mode == 1 || (search_mode != 0 && gx.n / (gx.W + 1) >= 256);
    int e = by_search ? launch_bounds_search(B, ctx->stream) : launch_bounds(B, ctx->sm_count, ctx->stream, nullptr, nullptr);
    void *pyr = attach_pyramids(ctx, L, gx.n);
    if (!e) e = launch_interp_windows(L, ctx->stream);
    pool_free(ctx, pyr);
    count_launch(ctx, 2);
    int32_t st = 0;
    if (e || cudaMemcpyAsync(&st, ctx->d_status, 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return bail(fail(ctx, BOWGPU_ECUDA, "fused interpolate (windows): %s", cudaGetErrorString(cudaGetLastError())));
    if (st & ST_INEXACT_START) {  // rare: ns timestamps whose window starts are hit only approximately
        cudaMemsetAsync(ctx->d_status, 0, 4, ctx->stream);
        pool_free(ctx, blk);
        if (st & ST_UNSORTED) return fail(ctx, BOWGPU_EUNSORTED, "time column is not sorted ascending");
        return unfused();
    }
    const int32_t rc = aggregate_dispatch(r, specs, nspecs, outs, mem, syn.data());
    pool_free(ctx, blk);  // stream ordered: released after the kernels that read it
    return rc;
}

// ================================================================================================
// one-shot, pipelined host -> host calls (one process, one or several GPUs)
// ================================================================================================
// rolling.IntervalRolling(b, col, interval, opts)[.Interpolate(...)].Aggregate(aggrs...) for a Bow that lives in HOST
// memory, results into host buffers, in ONE call (rolling.go:60, interpolation.go:30-69, aggregation.go:123-145).  The
// window range is cut into chunks (multiples of 64 windows, so validity bitmaps land byte aligned; every chunk carries
// its halo rows) exactly like the multi-GPU partitioning, and worker contexts - own stream, arena and pool
// each, a few per GPU, on every GPU the caller lists - run upload -> kernels -> download of different chunks
// concurrently: the device-to-host copies and the kernels of one chunk hide behind the host-to-device copy of the next,
// so the call takes the PCIe time of the inputs and little else.  Only the columns the operators read are uploaded.
// Worker threads belong to the library (never the caller's thread: a cgo call arrives on a Go runtime thread) and bind
// themselves to the CPUs next to their GPU, so that pinned staging buffers are allocated NUMA-local.
#include <sched.h>

#include <mutex>
namespace {

int64_t host_time_at(const bowgpu_col &c, int64_t i) { return ((const int64_t *)c.values)[c.offset + i]; }

int64_t host_lower_bound(const bowgpu_col &c, int64_t n, int64_t x) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (host_time_at(c, mid) < x)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}
bool host_valid(const bowgpu_col &c, int64_t i) {
    if (!c.validity || c.null_count == 0) return true;
    const int64_t b = c.offset + i;
    return (c.validity[b >> 3] >> (b & 7)) & 1;
}
// last row before `row` holding a valid value (0 if none) / one past the first row at or after `row` holding one (n if none)
int64_t host_prev_valid(const bowgpu_col &c, int64_t row) {
    for (int64_t i = row - 1; i >= 0; --i)
        if (host_valid(c, i)) return i;
    return 0;
}
int64_t host_next_valid_end(const bowgpu_col &c, int64_t n, int64_t row) {
    for (int64_t i = row; i < n; ++i)
        if (host_valid(c, i)) return i + 1;
    return n;
}

// CPUs next to a GPU (sysfs local_cpulist of its PCI function); the calling thread is bound to them
void bind_thread_near_device(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    for (char *p = bus; *p; ++p) *p = (char)tolower(*p);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE *f = fopen(path, "r");
    if (!f) return;
    char line[1024] = {0};
    const bool ok = fgets(line, sizeof line, f) != nullptr;
    fclose(f);
    if (!ok) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    int n = 0;
    for (char *p = line; *p && *p != '\n';) {  // "0-15,32-47"
        char *e;
        long a = strtol(p, &e, 10);
        if (e == p) break;
        long b = a;
        if (*e == '-') b = strtol(e + 1, &e, 10);
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) CPU_SET((int)c, &set), ++n;
        p = *e == ',' ? e + 1 : e;
    }
    if (n > 0) sched_setaffinity(0, sizeof set, &set);  // (tid 0 = the calling thread)
}

struct HostChunk {
    int64_t k_lo, k_hi;      // windows owned
    int64_t frame_lo;        // first row shipped (rows before row_lo: left halo of an interpolation)
    int64_t row_lo;          // first row of window k_lo
    int64_t halo_hi;         // one past the last row shipped
};

struct HostCall {
    const bowgpu_col *cols;
    int32_t ncols, time_col;
    int64_t interval, offset;
    int32_t inclusive;
    const int32_t *ops;  // interpolations (null: plain Aggregate)
    int32_t nops;
    const bowgpu_col *prev_row;
    const bowgpu_agg_spec *specs;
    int32_t nspecs;
    bowgpu_out_col *outs;
    int64_t out_capacity;
    int64_t *num_windows;
    const bowgpu_host_opts *opts;
};

// worker contexts of `ctx` on `device` (created on first use, kept for the next call)
int32_t workers_for(bowgpu_ctx *ctx, int device, int count, const std::vector<bowgpu_ctx *> &taken,
                    std::vector<bowgpu_ctx *> &out) {
    out.clear();
    for (bowgpu_ctx *w : ctx->workers)  // (a device listed twice gets two sets of workers)
        if (w->device == device && (int)out.size() < count && std::find(taken.begin(), taken.end(), w) == taken.end()) out.push_back(w);
    while ((int)out.size() < count) {
        bowgpu_ctx *w = nullptr;
        int32_t rc = bowgpu_ctx_create(device, nullptr, &w);
        if (rc) return fail(ctx, rc, "worker context on device %d", device);
        ctx->workers.push_back(w);
        out.push_back(w);
    }
    return BOWGPU_OK;
}

int32_t host_pipeline(bowgpu_ctx *ctx, const HostCall &H) {
    if (!ctx || !H.cols || !H.specs || !H.outs || H.ncols <= 0 || H.nspecs <= 0 || !H.num_windows) return BOWGPU_EINVAL;
    Guard gd(ctx);
    *H.num_windows = 0;
    const bowgpu_host_opts none = {};
    const bowgpu_host_opts &O = H.opts ? *H.opts : none;
    const bool interp = H.ops != nullptr;
    if (H.time_col < 0 || H.time_col >= H.ncols) return fail(ctx, BOWGPU_EINVAL, "time column index %d out of range", H.time_col);
    if (interp && H.nops != H.ncols)
        return fail(ctx, BOWGPU_EINVAL, "interpolations must name every column in schema order (%d given, %d columns)", H.nops, H.ncols);
    const bowgpu_col &tc = H.cols[H.time_col];
    const int64_t n = tc.length;
    auto plain = [&]() -> int32_t {  // upload everything, one pass on the ctx's own device (small inputs, corner cases)
        bowgpu_frame *f = nullptr;
        bowgpu_rolling *r = nullptr;
        int32_t rc = bowgpu_frame_create(ctx, H.cols, H.ncols, BOWGPU_MEM_HOST, &f);
        if (rc == BOWGPU_OK)
            rc = O.shard ? bowgpu_rolling_create_shard(f, H.time_col, H.interval, O.s0, O.num_windows, H.inclusive, H.prev_row, &r)
                         : bowgpu_rolling_create(f, H.time_col, H.interval, H.offset, H.inclusive, H.prev_row, &r);
        if (rc == BOWGPU_OK) {
            *H.num_windows = r->W;
            if (r->W > H.out_capacity) rc = fail(ctx, BOWGPU_ECAPACITY, "%lld windows, capacity %lld", (long long)r->W, (long long)H.out_capacity);
        }
        if (rc == BOWGPU_OK)
            rc = interp ? bowgpu_rolling_interpolate_aggregate(r, H.ops, H.nops, H.specs, H.nspecs, H.outs, BOWGPU_MEM_HOST)
                        : bowgpu_rolling_aggregate(r, H.specs, H.nspecs, H.outs, BOWGPU_MEM_HOST);
        bowgpu_rolling_destroy(r);
        bowgpu_frame_destroy(f);
        return rc;
    };
    const int64_t min_rows = O.chunk_rows > 0 ? 2 * O.chunk_rows : (int64_t)4 << 20;
    if (tc.dtype != BOWGPU_INT64 || H.interval <= 0 || n < min_rows || (tc.validity && tc.null_count != 0)) return plain();
    if (interp && (H.inclusive || H.ops[H.time_col] != BOWGPU_INTERP_WINDOW_START)) return plain();
    for (int j = 0; j < H.nspecs; ++j)
        if (H.specs[j].col < 0 || H.specs[j].col >= H.ncols) return fail(ctx, BOWGPU_EINVAL, "aggregation %d: no column %d", j, H.specs[j].col);
    // lattice, rolling.go:96-99,114-128,143-154 (a shard brings its own)
    const int64_t t_first = host_time_at(tc, 0), t_last = host_time_at(tc, n - 1);
    int64_t s0, W;
    if (O.shard) {
        s0 = O.s0;
        W = O.num_windows;
        if (W < 0) return BOWGPU_EINVAL;
    } else {
        int64_t off = H.offset;
        if (off >= H.interval || off <= -H.interval) off %= H.interval;
        if (off < 0) off += H.interval;
        s0 = (int64_t)((uint64_t)((t_first / H.interval) * H.interval) + (uint64_t)off);
        if (s0 > t_first) s0 = (int64_t)((uint64_t)s0 - (uint64_t)H.interval);
        if (t_last < t_first) return plain();  // obviously unsorted: the plain path reports it
        W = (int64_t)(((uint64_t)t_last - (uint64_t)s0) / (uint64_t)H.interval) + 1;
    }
    if (t_first < s0) return plain();  // rows before the first window start (or a left halo the caller shipped)
    *H.num_windows = W;
    if (W > H.out_capacity) return fail(ctx, BOWGPU_ECAPACITY, "%lld windows, capacity %lld", (long long)W, (long long)H.out_capacity);

    // columns actually read, remapped to a compact frame (an interpolation reads every column: positional append)
    std::vector<int> used, remap(H.ncols, -1);
    auto use = [&](int c) {
        if (remap[c] < 0) {
            remap[c] = (int)used.size();
            used.push_back(c);
        }
    };
    if (interp)
        for (int c = 0; c < H.ncols; ++c) use(c);
    use(H.time_col);
    for (int j = 0; j < H.nspecs; ++j) use(H.specs[j].col);
    std::vector<bowgpu_agg_spec> sp(H.specs, H.specs + H.nspecs);
    for (auto &x : sp) x.col = remap[x.col];
    for (int j = 0; j < H.nspecs; ++j) H.outs[j].dtype = bowgpu_agg_return_type(H.specs[j].op, H.cols[H.specs[j].col].dtype);

    // chunks of ~12M rows (a few ms of PCIe each), cut on multiples of 64 windows
    const int64_t target_rows = O.chunk_rows > 0 ? O.chunk_rows : (int64_t)12 << 20;
    int64_t nchunks = (n + target_rows - 1) / target_rows;
    if (nchunks > W / 64) nchunks = W / 64;
    if (nchunks < 2) return plain();
    auto start_of = [&](int64_t k) { return (int64_t)((uint64_t)s0 + (uint64_t)k * (uint64_t)H.interval); };
    std::vector<int> icols;  // columns whose interpolation looks at neighbouring valid rows
    if (interp)
        for (int c = 0; c < H.ncols; ++c)
            if (H.ops[c] == BOWGPU_INTERP_LINEAR || H.ops[c] == BOWGPU_INTERP_STEP_PREVIOUS || H.ops[c] == BOWGPU_INTERP_STEP_NEXT) icols.push_back(c);
    std::vector<HostChunk> chunks;
    int64_t k_prev = 0, row_prev = 0;
    for (int64_t c = 1; c <= nchunks; ++c) {
        const int64_t k = c == nchunks ? W : ((c * W) / nchunks) / 64 * 64;
        if (k <= k_prev) continue;
        const int64_t row = c == nchunks ? n : host_lower_bound(tc, n, start_of(k));
        HostChunk ch{k_prev, k, row_prev, row_prev, row};
        if (!interp) {
            if (c != nchunks && row < n) ch.halo_hi = row + 1;  // halo: the row an inclusive last window may borrow
        } else {
            // SURVEY 8e / partition.plan_interpolate: left halo back to the last valid row of every interpolated column (what
            // Linear / StepPrevious look up for the chunk's first windows), ONE EXTRA WINDOW on the right (its start row is the
            // inclusive row of the chunk's last window) and a right halo up to the next valid row of every column
            const int64_t extra = k < W ? 1 : 0;
            const int64_t k_last = k - 1 + extra;
            const int64_t end_rows = k_last + 1 < W ? host_lower_bound(tc, n, start_of(k_last + 1)) : n;
            const int64_t r_last = host_lower_bound(tc, n, start_of(k_last));
            int64_t hi = std::min<int64_t>(n, end_rows + 1);
            for (int ic : icols) hi = std::max(hi, host_next_valid_end(H.cols[ic], n, r_last));
            ch.halo_hi = std::min<int64_t>(n, hi);
            if (row_prev > 0) {
                int64_t lo = row_prev - 1;
                for (int ic : icols) lo = std::min(lo, host_prev_valid(H.cols[ic], row_prev));
                ch.frame_lo = std::max<int64_t>(0, lo);
            }
        }
        chunks.push_back(ch);
        k_prev = k;
        row_prev = row;
    }
    // workers: a few per device, on every device the caller lists
    std::vector<int> devs;
    if (O.devices && O.ndevices > 0)
        devs.assign(O.devices, O.devices + O.ndevices);
    else
        devs.push_back(ctx->device);
    // pageable inputs (the Go heap) pass through pinned staging chunks filled by the worker's own memcpy: more workers
    // = more copy threads (measured on the B200 box: 29 GB/s with 3)
    const int per_dev = O.workers_per_device > 0 ? O.workers_per_device : (is_pinned(tc.values) ? 3 : 6);
    std::vector<bowgpu_ctx *> workers;
    for (int d : devs) {
        std::vector<bowgpu_ctx *> w;
        const int want = (int)std::min<size_t>((size_t)per_dev, (chunks.size() + devs.size() - 1) / devs.size());
        int32_t rc = workers_for(ctx, d, std::max(want, 1), workers, w);
        if (rc) return rc;
        workers.insert(workers.end(), w.begin(), w.end());
    }
    const int nworkers = (int)workers.size();
    std::atomic<size_t> next{0};
    std::vector<int32_t> status(nworkers, BOWGPU_OK);
    std::vector<std::string> errs(nworkers);
    auto work = [&](int wi) {
        bowgpu_ctx *wc = workers[wi];
        cudaSetDevice(wc->device);
        bind_thread_near_device(wc->device);
        for (;;) {
            const size_t ci = next.fetch_add(1);
            if (ci >= chunks.size() || status[wi] != BOWGPU_OK) return;
            const HostChunk &ch = chunks[ci];
            std::vector<bowgpu_col> cc(used.size());
            for (size_t u = 0; u < used.size(); ++u) {
                cc[u] = H.cols[used[u]];
                cc[u].offset += ch.frame_lo;
                cc[u].length = ch.halo_hi - ch.frame_lo;
                if (cc[u].validity && cc[u].null_count != 0) cc[u].null_count = -1;  // counted on the device
            }
            std::vector<bowgpu_out_col> oc(H.nspecs);
            for (int j = 0; j < H.nspecs; ++j) {
                oc[j].values = (char *)H.outs[j].values + ch.k_lo * 8;
                oc[j].validity = H.outs[j].validity + ch.k_lo / 8;
            }
            bowgpu_frame *f = nullptr;
            bowgpu_rolling *r = nullptr;
            int32_t rc = bowgpu_frame_create(wc, cc.data(), (int32_t)cc.size(), BOWGPU_MEM_HOST, &f);
            if (rc == BOWGPU_OK)
                rc = bowgpu_rolling_create_shard(f, remap[H.time_col], H.interval, start_of(ch.k_lo), ch.k_hi - ch.k_lo, H.inclusive,
                                                 ci == 0 && ch.frame_lo == 0 ? H.prev_row : nullptr, &r);
            if (rc == BOWGPU_OK) {
                if (interp) {
                    std::vector<int32_t> ops(used.size());
                    for (size_t u = 0; u < used.size(); ++u) ops[u] = H.ops[used[u]];
                    rc = bowgpu_rolling_interpolate_aggregate(r, ops.data(), (int32_t)ops.size(), sp.data(), H.nspecs, oc.data(), BOWGPU_MEM_HOST);
                } else {
                    rc = aggregate_dispatch(r, sp.data(), H.nspecs, oc.data(), BOWGPU_MEM_HOST, nullptr);
                }
            }
            if (rc != BOWGPU_OK) {
                status[wi] = rc;
                errs[wi] = wc->err;
            }
            bowgpu_rolling_destroy(r);
            bowgpu_frame_destroy(f);
        }
    };
    std::vector<std::thread> threads;
    for (int wi = 0; wi < nworkers; ++wi) threads.emplace_back(work, wi);
    for (auto &t : threads) t.join();
    for (int wi = 0; wi < nworkers; ++wi)
        if (status[wi] != BOWGPU_OK) return fail(ctx, status[wi], "%s", errs[wi].c_str());
    return BOWGPU_OK;
}

}  // namespace

extern "C" int32_t bowgpu_aggregate_host_ex(bowgpu_ctx *ctx, const bowgpu_col *cols, int32_t ncols, int32_t time_col,
                                            int64_t interval, int64_t offset, int32_t inclusive,
                                            const bowgpu_agg_spec *specs, int32_t nspecs, bowgpu_out_col *outs,
                                            int64_t out_capacity, int64_t *num_windows, const bowgpu_host_opts *opts) {
    HostCall H{cols, ncols, time_col, interval, offset, inclusive, nullptr, 0, nullptr, specs, nspecs, outs, out_capacity, num_windows, opts};
    return host_pipeline(ctx, H);
}

extern "C" int32_t bowgpu_aggregate_host(bowgpu_ctx *ctx, const bowgpu_col *cols, int32_t ncols, int32_t time_col,
                                         int64_t interval, int64_t offset, int32_t inclusive,
                                         const bowgpu_agg_spec *specs, int32_t nspecs, bowgpu_out_col *outs,
                                         int64_t out_capacity, int64_t *num_windows) {
    return bowgpu_aggregate_host_ex(ctx, cols, ncols, time_col, interval, offset, inclusive, specs, nspecs, outs, out_capacity,
                                    num_windows, nullptr);
}

extern "C" int32_t bowgpu_interpolate_aggregate_host(bowgpu_ctx *ctx, const bowgpu_col *cols, int32_t ncols, int32_t time_col,
                                                     int64_t interval, int64_t offset, const bowgpu_col *prev_row,
                                                     const int32_t *ops, int32_t nops, const bowgpu_agg_spec *specs,
                                                     int32_t nspecs, bowgpu_out_col *outs, int64_t out_capacity,
                                                     int64_t *num_windows, const bowgpu_host_opts *opts) {
    if (!ops) return BOWGPU_EINVAL;
    HostCall H{cols, ncols, time_col, interval, offset, 0, ops, nops, prev_row, specs, nspecs, outs, out_capacity, num_windows, opts};
    return host_pipeline(ctx, H);
}
