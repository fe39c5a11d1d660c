// parquet.cu — bow.NewBowFromParquet (bowparquet.go:44-155) for the column types of the rolling path.
//
// The reference reads every leaf column with xitongsys/parquet-go v1.6.2 (go.mod:13; not vendored) into []interface{}
// and copies value by value into a bow.Buffer (bowparquet.go:97-107).  Here the HOST only walks metadata — the Thrift
// compact footer (FileMetaData) and the page headers of the chosen column chunks, a few dozen bytes per page — and the
// file bytes of those chunks go to the device as they are.  Everything that touches data runs on the GPU:
//   * pq_decompress_kernel: one warp per SNAPPY page (format: varint length, then literal / copy elements); the warp
//     parses an element in lock step (every lane reads the same tag bytes: one L1 transaction) and copies it with all 32
//     lanes, back-references included (overlapping copies repeat their period).
//   * pq_decode_kernel: one CTA per data page.  Definition levels (RLE / bit-packed hybrid, bit width 1) become the
//     Arrow validity bits of the page's rows (atomicOr into the zeroed column bitmap: pages start at any row), a block
//     scan of their popcounts gives every row the index of its value, and the PLAIN (or dictionary) values are
//     scattered to their rows — value 0 in null slots, as bow.NewBuffer leaves them (bowbuffer.go).
// Supported: flat schemas, INT64 / DOUBLE leaves (Boolean / String columns have no GPU type: the caller skips them or
// gets BOWGPU_ETYPE), required or optional, PLAIN and PLAIN_DICTIONARY / RLE_DICTIONARY, data pages v1 and v2,
// UNCOMPRESSED and SNAPPY (what parquet-go writes by default, bowparquet.go:181-183), any number of row groups.
#include "parquet.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/bowgpu.h"
#include "kernels.h"

namespace bowgpu {

// ================================================================================================
// host: Thrift compact protocol (just enough for parquet.thrift's FileMetaData and PageHeader)
// ================================================================================================
namespace {

struct TReader {
    const uint8_t *p, *end;
    bool ok = true;
    TReader(const uint8_t *b, const uint8_t *e) : p(b), end(e) {}
    uint8_t byte() {
        if (p >= end) {
            ok = false;
            return 0;
        }
        return *p++;
    }
    uint64_t varint() {
        uint64_t v = 0;
        for (int s = 0; s < 70; s += 7) {
            const uint8_t b = byte();
            v |= (uint64_t)(b & 0x7f) << s;
            if (!(b & 0x80) || !ok) return v;
        }
        ok = false;
        return v;
    }
    int64_t zigzag() {
        const uint64_t v = varint();
        return (int64_t)(v >> 1) ^ -(int64_t)(v & 1);
    }
    std::string binary() {
        const uint64_t n = varint();
        if (!ok || n > (uint64_t)(end - p)) {
            ok = false;
            return std::string();
        }
        std::string s((const char *)p, (size_t)n);
        p += n;
        return s;
    }
    // field header of a struct: returns false at the stop byte
    bool field(int &id, int &type) {
        const uint8_t b = byte();
        if (!ok || b == 0) return false;
        type = b & 0x0f;
        const int delta = b >> 4;
        id = delta ? id + delta : (int)zigzag();
        return ok;
    }
    void list(int &n, int &type) {
        const uint8_t b = byte();
        type = b & 0x0f;
        n = b >> 4;
        if (n == 15) n = (int)varint();
        if (n < 0) ok = false;
    }
    void skip(int type, int depth = 0) {
        if (depth > 32) ok = false;
        if (!ok) return;
        switch (type) {
        case 1: case 2: break;  // bool, value in the field header
        case 3: byte(); break;
        case 4: case 5: case 6: varint(); break;
        case 7: p += 8; if (p > end) ok = false; break;
        case 8: {  // (no string is built for what is skipped: the min / max statistics of every page header land here)
            const uint64_t n = varint();
            if (!ok || n > (uint64_t)(end - p)) ok = false;
            else p += n;
            break;
        }
        case 9: case 10: {
            int n, t;
            list(n, t);
            for (int i = 0; i < n && ok; ++i) {
                if (t == 1 || t == 2) byte();  // bools inside a list take a byte each
                else skip(t, depth + 1);
            }
            break;
        }
        case 11: {
            const int n = (int)varint();
            if (n > 0) {
                const uint8_t kv = byte();
                for (int i = 0; i < n && ok; ++i) {
                    skip(kv >> 4, depth + 1);
                    skip(kv & 0x0f, depth + 1);
                }
            }
            break;
        }
        case 12: {
            int id = 0, t;
            while (field(id, t)) skip(t, depth + 1);
            break;
        }
        default: ok = false;
        }
    }
};

struct ChunkMeta {
    int32_t type = -1, codec = 0;
    int64_t num_values = 0, total_compressed = 0, data_page_offset = -1, dict_page_offset = -1;
};
struct RowGroupMeta {
    int64_t num_rows = 0;
    std::vector<ChunkMeta> chunks;
};
struct SchemaElem {
    std::string name;
    int32_t type = -1, repetition = 0, num_children = 0;
    bool has_type = false;
};

void parse_column_meta(TReader &r, ChunkMeta &m) {
    int id = 0, t;
    while (r.field(id, t)) {
        switch (id) {
        case 1: m.type = (int32_t)r.zigzag(); break;
        case 4: m.codec = (int32_t)r.zigzag(); break;
        case 5: m.num_values = r.zigzag(); break;
        case 7: m.total_compressed = r.zigzag(); break;
        case 9: m.data_page_offset = r.zigzag(); break;
        case 11: m.dict_page_offset = r.zigzag(); break;
        default: r.skip(t);
        }
    }
}
void parse_column_chunk(TReader &r, ChunkMeta &m) {
    int id = 0, t;
    while (r.field(id, t)) {
        if (id == 3 && t == 12) parse_column_meta(r, m);
        else r.skip(t);
    }
}
void parse_row_group(TReader &r, RowGroupMeta &g) {
    int id = 0, t;
    while (r.field(id, t)) {
        if (id == 1 && t == 9) {
            int n, et;
            r.list(n, et);
            g.chunks.resize(n);
            for (int i = 0; i < n && r.ok; ++i) parse_column_chunk(r, g.chunks[i]);
        } else if (id == 3) {
            g.num_rows = r.zigzag();
        } else {
            r.skip(t);
        }
    }
}
void parse_schema_elem(TReader &r, SchemaElem &e) {
    int id = 0, t;
    while (r.field(id, t)) {
        switch (id) {
        case 1: e.type = (int32_t)r.zigzag(); e.has_type = true; break;
        case 3: e.repetition = (int32_t)r.zigzag(); break;
        case 4: e.name = r.binary(); break;
        case 5: e.num_children = (int32_t)r.zigzag(); break;
        default: r.skip(t);
        }
    }
}

struct PageHeader {
    int32_t type = -1, uncompressed = 0, compressed = 0;
    int32_t num_values = 0, encoding = 0, def_encoding = 3;
    int32_t def_bytes = 0, rep_bytes = 0;
    bool v2_compressed = true;
};
void parse_data_header(TReader &r, PageHeader &h, bool v2) {
    int id = 0, t;
    while (r.field(id, t)) {
        if (!v2) {
            switch (id) {
            case 1: h.num_values = (int32_t)r.zigzag(); break;
            case 2: h.encoding = (int32_t)r.zigzag(); break;
            case 3: h.def_encoding = (int32_t)r.zigzag(); break;
            default: r.skip(t);
            }
        } else {
            switch (id) {
            case 1: h.num_values = (int32_t)r.zigzag(); break;
            case 4: h.encoding = (int32_t)r.zigzag(); break;
            case 5: h.def_bytes = (int32_t)r.zigzag(); break;
            case 6: h.rep_bytes = (int32_t)r.zigzag(); break;
            case 7: h.v2_compressed = t == 1; break;  // bool field: the value is the type nibble
            default: r.skip(t);
            }
        }
    }
}
void parse_dict_header(TReader &r, PageHeader &h) {
    int id = 0, t;
    while (r.field(id, t)) {
        switch (id) {
        case 1: h.num_values = (int32_t)r.zigzag(); break;
        case 2: h.encoding = (int32_t)r.zigzag(); break;
        default: r.skip(t);
        }
    }
}
void parse_page_header(TReader &r, PageHeader &h) {
    int id = 0, t;
    while (r.field(id, t)) {
        switch (id) {
        case 1: h.type = (int32_t)r.zigzag(); break;
        case 2: h.uncompressed = (int32_t)r.zigzag(); break;
        case 3: h.compressed = (int32_t)r.zigzag(); break;
        case 5: if (t == 12) parse_data_header(r, h, false); else r.skip(t); break;
        case 7: if (t == 12) parse_dict_header(r, h); else r.skip(t); break;
        case 8: if (t == 12) parse_data_header(r, h, true); else r.skip(t); break;
        default: r.skip(t);
        }
    }
}

int64_t up16(int64_t x) { return (x + 15) & ~(int64_t)15; }

}  // namespace

struct PqFile {
    int fd = -1;
    const uint8_t *map = nullptr;
    int64_t size = 0;
    int64_t num_rows = 0;
    std::vector<PqColumn> cols;       // leaf columns in schema order
    std::vector<RowGroupMeta> groups;
};

int pq_open(const char *path, PqFile **out, std::string &err) {
    *out = nullptr;
    PqFile *f = new PqFile();
    auto bail = [&](int code, const std::string &m) {
        err = m;
        pq_close(f);
        return code;
    };
    f->fd = open(path, O_RDONLY);
    if (f->fd < 0) return bail(BOWGPU_EIO, std::string("open ") + path + ": " + strerror(errno));
    struct stat st;
    if (fstat(f->fd, &st) != 0) return bail(BOWGPU_EIO, std::string("fstat: ") + strerror(errno));
    f->size = st.st_size;
    if (f->size < 12) return bail(BOWGPU_EIO, "not a parquet file (too short)");
    void *m = mmap(nullptr, (size_t)f->size, PROT_READ, MAP_PRIVATE, f->fd, 0);
    if (m == MAP_FAILED) return bail(BOWGPU_EIO, std::string("mmap: ") + strerror(errno));
    f->map = (const uint8_t *)m;
    if (memcmp(f->map, "PAR1", 4) != 0 || memcmp(f->map + f->size - 4, "PAR1", 4) != 0)
        return bail(BOWGPU_EIO, "not a parquet file (magic bytes)");
    uint32_t flen;
    memcpy(&flen, f->map + f->size - 8, 4);
    if ((int64_t)flen + 12 > f->size) return bail(BOWGPU_EIO, "parquet footer length out of range");
    TReader r(f->map + f->size - 8 - flen, f->map + f->size - 8);
    std::vector<SchemaElem> schema;
    int id = 0, t;
    while (r.field(id, t)) {  // FileMetaData
        if (id == 2 && t == 9) {
            int n, et;
            r.list(n, et);
            schema.resize(n);
            for (int i = 0; i < n && r.ok; ++i) parse_schema_elem(r, schema[i]);
        } else if (id == 3) {
            f->num_rows = r.zigzag();
        } else if (id == 4 && t == 9) {
            int n, et;
            r.list(n, et);
            f->groups.resize(n);
            for (int i = 0; i < n && r.ok; ++i) parse_row_group(r, f->groups[i]);
        } else {
            r.skip(t);
        }
    }
    if (!r.ok || schema.empty()) return bail(BOWGPU_EIO, "malformed parquet footer");
    // flat schemas only: root + leaves (bowparquet.go:84-87 skips group nodes and reads every leaf)
    if (schema[0].num_children != (int)schema.size() - 1)
        return bail(BOWGPU_EUNSUPPORTED, "nested parquet schemas are not supported");
    for (size_t i = 1; i < schema.size(); ++i) {
        const SchemaElem &e = schema[i];
        if (e.num_children != 0 || !e.has_type) return bail(BOWGPU_EUNSUPPORTED, "nested parquet schemas are not supported");
        if (e.repetition == 2) return bail(BOWGPU_EUNSUPPORTED, "repeated parquet column " + e.name);
        PqColumn c;
        c.name = e.name;
        c.physical = e.type;
        c.dtype = e.type == 2 ? BOWGPU_INT64 : e.type == 5 ? BOWGPU_FLOAT64 : 0;  // mapParquetToBowTypes, bowparquet.go:20-25
        c.optional = e.repetition == 1;
        f->cols.push_back(c);
    }
    int64_t rows = 0;
    for (const auto &g : f->groups) {
        if (g.chunks.size() != f->cols.size()) return bail(BOWGPU_EIO, "row group with a different number of column chunks");
        rows += g.num_rows;
    }
    if (rows != f->num_rows) return bail(BOWGPU_EIO, "row group sizes do not add up to num_rows");
    *out = f;
    return 0;
}

void pq_close(PqFile *f) {
    if (!f) return;
    if (f->map) munmap((void *)f->map, (size_t)f->size);
    if (f->fd >= 0) close(f->fd);
    delete f;
}
int64_t pq_num_rows(const PqFile *f) { return f->num_rows; }
const std::vector<PqColumn> &pq_columns(const PqFile *f) { return f->cols; }
const uint8_t *pq_bytes(const PqFile *f) { return f->map; }
int pq_fd(const PqFile *f) { return f->fd; }

// The page headers of one column chunk (col j of the output, one row group).  Offsets that depend on the chunks before it
// (scratch, dictionary-index scratch, descriptor index of the dictionary page) are chunk-relative here; pq_plan rebases them.
namespace {
struct ChunkWalk {
    int j = 0;                 // output column
    const ChunkMeta *m = nullptr;
    const PqColumn *pc = nullptr;
    int64_t start = 0, image_off = 0, row0 = 0, num_rows = 0;
    std::vector<PqPage> pages;
    int64_t scratch = 0, aux = 0;
    int rc = 0;
    std::string err;
};

void walk_chunk(const PqFile *f, ChunkWalk &w) {
    const ChunkMeta &m = *w.m;
    const PqColumn &pc = *w.pc;
    auto bail = [&](int code, const std::string &what) {
        w.rc = code;
        w.err = "parquet column " + pc.name + ": " + what;
    };
    const int64_t start = w.start, end = start + m.total_compressed;
    int64_t pos = start, seen = 0, row = w.row0;
    int dict_idx = -1;
    const bool dbg = [] {
        const char *e = getenv("BOWGPU_PQ_DEBUG");
        return e && e[0] == '2';
    }();
    w.pages.reserve((size_t)(m.total_compressed / 6000 + 4));
    while (seen < m.num_values && pos < end) {
        TReader r(f->map + pos, f->map + end);
        PageHeader h;
        parse_page_header(r, h);
        // (a page is a few KB to a few MB; a header that claims more than 1 GiB uncompressed is damage, not data)
        if (!r.ok || h.compressed < 0 || h.uncompressed < 0 || h.uncompressed > (1 << 30) || h.num_values < 0 ||
            (r.p - f->map) + h.compressed > end)
            return bail(BOWGPU_EIO, "malformed page header");
        const int64_t body = r.p - f->map;
        pos = body + h.compressed;
        if (h.type != 0 && h.type != 2 && h.type != 3) continue;  // index pages: skipped
        PqPage p;
        memset(&p, 0, sizeof p);
        p.src = w.image_off + (body - start);
        p.comp_size = h.compressed;
        p.uncomp_size = h.uncompressed;
        p.num_values = h.num_values;
        p.col = w.j;
        p.codec = m.codec;
        p.optional = pc.optional;
        p.dict = -1;
        p.aux = -1;
        p.dst = -1;
        if (h.type == 2) {  // dictionary page
            if (h.encoding != 0 && h.encoding != 2) return bail(BOWGPU_EUNSUPPORTED, "dictionary page encoding " + std::to_string(h.encoding));
            p.kind = PQ_DICT;
            dict_idx = (int)w.pages.size();
        } else {
            p.kind = h.type == 0 ? PQ_DATA_V1 : PQ_DATA_V2;
            if (h.encoding == 2 || h.encoding == 8) {
                if (dict_idx < 0) return bail(BOWGPU_EIO, "dictionary-encoded page without a dictionary page");
                p.dict_enc = 1;
                p.dict = dict_idx;
                p.aux = w.aux;
                w.aux += h.num_values;
            } else if (h.encoding != 0) {
                return bail(BOWGPU_EUNSUPPORTED, "value encoding " + std::to_string(h.encoding) + " (only PLAIN and dictionary)");
            }
            if (p.kind == PQ_DATA_V1 && pc.optional && h.def_encoding != 3) return bail(BOWGPU_EUNSUPPORTED, "definition levels must be RLE encoded");
            if (p.kind == PQ_DATA_V2) {
                if (h.rep_bytes != 0) return bail(BOWGPU_EIO, "repetition levels in a flat column");
                p.lvl_bytes = h.def_bytes;
                if (!h.v2_compressed) p.codec = 0;
                if (p.lvl_bytes < 0 || p.lvl_bytes > h.compressed || p.lvl_bytes > h.uncompressed) return bail(BOWGPU_EIO, "malformed v2 page header");
            }
            p.row0 = row;
            row += h.num_values;
            seen += h.num_values;
        }
        if (p.codec == 1) {  // the part that is compressed (a v2 page keeps its levels in front, uncompressed)
            p.dst = w.scratch;
            w.scratch += up16(p.uncomp_size - p.lvl_bytes) + 16;
        }
        if (dbg)
            fprintf(stderr, "pq page col %d kind %d codec %d nv %d comp %d uncomp %d row0 %lld dict_enc %d lvl %d\n", p.col, p.kind,
                    p.codec, p.num_values, p.comp_size, p.uncomp_size, (long long)p.row0, p.dict_enc, p.lvl_bytes);
        w.pages.push_back(p);
    }
    if (seen != m.num_values || m.num_values != w.num_rows) return bail(BOWGPU_EIO, "pages do not add up to the row group");
}
}  // namespace

// Page headers sit thousands of bytes apart and each one says where the next begins: a walk is a chain of cache misses
// (0.5 us per page; 50 ms for the 97 659 pages of a 1 GB file written with the reference's 8 KB pages).  The chunks are
// independent, so they are walked by up to 8 threads.
int pq_plan(const PqFile *f, const int32_t *cols, int32_t ncols, PqPlan &plan, std::string &err) {
    plan = PqPlan();
    std::vector<ChunkWalk> walks;
    for (int j = 0; j < ncols; ++j) {
        const int ci = cols[j];
        if (ci < 0 || ci >= (int)f->cols.size()) {
            err = "no parquet column " + std::to_string(ci);
            return BOWGPU_EINVAL;
        }
        const PqColumn &pc = f->cols[ci];
        if (!pc.dtype) {
            err = "parquet column " + pc.name + ": only INT64 / DOUBLE columns have a GPU type";
            return BOWGPU_ETYPE;
        }
        int64_t row0 = 0;
        for (const auto &g : f->groups) {
            const ChunkMeta &m = g.chunks[ci];
            if (g.num_rows == 0 && m.num_values == 0) continue;  // (an empty table still has a row group)
            if (m.codec != 0 && m.codec != 1) {
                err = "parquet column " + pc.name + ": compression codec " + std::to_string(m.codec) + " (only UNCOMPRESSED and SNAPPY)";
                return BOWGPU_EUNSUPPORTED;
            }
            int64_t start = m.data_page_offset;
            if (m.dict_page_offset > 0 && m.dict_page_offset < start) start = m.dict_page_offset;
            if (start < 4 || m.total_compressed < 0 || start + m.total_compressed > f->size - 8) {
                err = "parquet column " + pc.name + ": chunk outside the file";
                return BOWGPU_EIO;
            }
            PqRange rg;
            rg.file_off = start;
            rg.len = m.total_compressed;
            rg.image_off = plan.image_bytes;
            plan.image_bytes += up16(rg.len) + 16;
            plan.ranges.push_back(rg);
            ChunkWalk w;
            w.j = j;
            w.m = &m;
            w.pc = &pc;
            w.start = start;
            w.image_off = rg.image_off;
            w.row0 = row0;
            w.num_rows = g.num_rows;
            walks.push_back(std::move(w));
            row0 += g.num_rows;
        }
    }
    {
        const int nt = (int)std::min<size_t>(8, walks.size());
        std::atomic<size_t> next{0};
        auto work = [&] {
            for (size_t i; (i = next.fetch_add(1)) < walks.size();) walk_chunk(f, walks[i]);
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work);
        work();
        for (auto &x : th) x.join();
    }
    size_t total = 0;
    for (const auto &w : walks) {
        if (w.rc) {
            err = w.err;
            return w.rc;
        }
        total += w.pages.size();
    }
    plan.pages.reserve(total);
    for (auto &w : walks) {  // rebase the chunk-relative offsets (one walk per entry of plan.ranges, same order)
        const int base = (int)plan.pages.size();
        plan.range_first_page.push_back(base);
        for (PqPage p : w.pages) {
            if (p.dst >= 0) p.dst += plan.scratch_bytes;
            if (p.aux >= 0) p.aux += plan.aux_entries;
            if (p.dict >= 0) p.dict += base;
            plan.pages.push_back(p);
        }
        plan.scratch_bytes += w.scratch;
        plan.aux_entries += w.aux;
    }
    plan.range_first_page.push_back((int32_t)plan.pages.size());
    plan.image_bytes += 64;
    plan.scratch_bytes += 64;
    return 0;
}

// ================================================================================================
// device
// ================================================================================================
namespace {

constexpr int PQ_THREADS = 256;
constexpr int PQ_BATCH = PQ_THREADS * 32;  // rows per trip of the decode loop
constexpr int PQ_RUNS = 512;               // runs of a hybrid stream parsed ahead

__device__ __forceinline__ void pq_fail(int32_t *status) { atomicOr(status, ST_PARQUET); }

// dst[0, len) = src[0, len) by one warp; src may be a global address written earlier by this warp (L2 reads)
__device__ __forceinline__ void warp_copy(uint8_t *dst, const uint8_t *src, const int64_t len, const int lane) {
    if (len < 64) {
        for (int64_t i = lane; i < len; i += 32) dst[i] = __ldcg(src + i);
        return;
    }
    const int head = (int)((8 - ((uintptr_t)dst & 7)) & 7);
    if (lane < head) dst[lane] = __ldcg(src + lane);
    uint64_t *d8 = reinterpret_cast<uint64_t *>(dst + head);
    const uint8_t *s = src + head;
    const int mis = (int)((uintptr_t)s & 7);
    const uint64_t *s8 = reinterpret_cast<const uint64_t *>(s - mis);
    const int64_t words = (len - head) >> 3;
    if (mis == 0) {
        for (int64_t w = lane; w < words; w += 32) d8[w] = __ldcg(s8 + w);
    } else {
        const int sh = mis * 8;
        for (int64_t w = lane; w < words; w += 32) d8[w] = (__ldcg(s8 + w) >> sh) | (__ldcg(s8 + w + 1) << (64 - sh));
    }
    const int64_t done = head + words * 8;
    for (int64_t i = done + lane; i < len; i += 32) dst[i] = __ldcg(src + i);
}

// One warp per page: Snappy block format.
__global__ void __launch_bounds__(128) pq_decompress_kernel(const PqPage *pages, const int first, const int npages,
                                                             const uint8_t *image, uint8_t *scratch, int32_t *status) {
    const int page = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (page >= npages) return;
    const PqPage pg = pages[first + page];
    if (pg.codec != 1 || pg.dst < 0) return;
    const uint8_t *in = image + pg.src + pg.lvl_bytes;
    const int64_t in_len = pg.comp_size - pg.lvl_bytes, out_len = pg.uncomp_size - pg.lvl_bytes;
    uint8_t *out = scratch + pg.dst;
    int64_t ip = 0, op = 0;
    uint64_t ulen = 0;  // preamble: uncompressed length
    for (int s = 0; s < 35 && ip < in_len; s += 7) {
        const uint8_t b = in[ip++];
        ulen |= (uint64_t)(b & 0x7f) << s;
        if (!(b & 0x80)) break;
    }
    if ((int64_t)ulen != out_len) {
        if (lane == 0) pq_fail(status);
        return;
    }
    while (ip < in_len) {
        const uint32_t tag = in[ip++];
        int64_t len, off = 0;
        if ((tag & 3) == 0) {  // literal
            len = (tag >> 2) + 1;
            if (len > 60) {
                const int nb = (int)len - 60;
                if (ip + nb > in_len) break;
                uint32_t v = 0;
                for (int b = 0; b < nb; ++b) v |= (uint32_t)in[ip + b] << (8 * b);
                ip += nb;
                len = (int64_t)v + 1;
            }
            if (ip + len > in_len || op + len > out_len) {
                ip = -1;
                break;
            }
            warp_copy(out + op, in + ip, len, lane);
            ip += len;
            op += len;
            __syncwarp();
            continue;
        }
        if ((tag & 3) == 1) {
            if (ip + 1 > in_len) break;
            len = 4 + ((tag >> 2) & 7);
            off = ((int64_t)(tag >> 5) << 8) | in[ip];
            ip += 1;
        } else if ((tag & 3) == 2) {
            if (ip + 2 > in_len) break;
            len = (tag >> 2) + 1;
            off = (int64_t)in[ip] | ((int64_t)in[ip + 1] << 8);
            ip += 2;
        } else {
            if (ip + 4 > in_len) break;
            len = (tag >> 2) + 1;
            off = (int64_t)in[ip] | ((int64_t)in[ip + 1] << 8) | ((int64_t)in[ip + 2] << 16) | ((int64_t)in[ip + 3] << 24);
            ip += 4;
        }
        if (off <= 0 || off > op || op + len > out_len) {
            ip = -1;
            break;
        }
        if (off >= len) {
            warp_copy(out + op, out + op - off, len, lane);
        } else {  // overlapping: the last `off` bytes repeat
            for (int64_t i = lane; i < len; i += 32) out[op + i] = __ldcg(out + op - off + (i % off));
        }
        op += len;
        __syncwarp();
    }
    if ((ip != in_len || op != out_len) && lane == 0) pq_fail(status);
}

// ---- RLE / bit-packed hybrid streams ------------------------------------------------------------------------------------
struct PqRun {
    int32_t start;   // first value of the run (index inside the page)
    int32_t count;
    int32_t packed;  // 1: bit-packed, `at` = byte offset of its first group; 0: RLE, `at` = the repeated value
    int32_t at;
};
struct PqStream {     // shared memory
    PqRun runs[PQ_RUNS];
    int32_t nruns;
    int32_t covered;  // values [0, covered) of the page are described by runs parsed so far (those before runs[0].start are done)
    int32_t pos;      // next byte of the stream to parse
    int32_t bad;
};
// thread 0: parse run headers until the table is full, `want` values are covered or the stream ends
// (`limit` = values of the page: what lies beyond is the padding of the last bit-packed group)
__device__ void pq_parse_runs(PqStream &S, const uint8_t *p, const int32_t len, const int bw, const int32_t want,
                              const int32_t limit) {
    while (S.nruns < PQ_RUNS && S.covered < want && S.pos < len) {
        uint32_t h = 0;
        int s = 0;
        for (;;) {
            if (S.pos >= len || s > 28) {
                S.bad = 1;
                return;
            }
            const uint8_t b = p[S.pos++];
            h |= (uint32_t)(b & 0x7f) << s;
            if (!(b & 0x80)) break;
            s += 7;
        }
        PqRun r;
        r.start = S.covered;
        if (h & 1) {  // bit-packed: (h >> 1) groups of 8 values
            const int64_t groups = h >> 1, bytes = groups * bw;
            r.count = (int32_t)min((int64_t)INT32_MAX / 2, groups * 8);
            r.packed = 1;
            r.at = S.pos;
            if (S.pos + bytes > len) {  // (writers may cut the padding of the last group short)
                const int64_t avail = len - S.pos;
                r.count = (int32_t)(avail * 8 / (bw ? bw : 1));
                S.pos = len;
            } else {
                S.pos += (int32_t)bytes;
            }
        } else {
            const int nb = (bw + 7) / 8;
            if (S.pos + nb > len) {
                S.bad = 1;
                return;
            }
            uint32_t v = 0;
            for (int b = 0; b < nb; ++b) v |= (uint32_t)p[S.pos + b] << (8 * b);
            S.pos += nb;
            r.count = (int32_t)(h >> 1);
            r.packed = 0;
            r.at = (int32_t)v;
        }
        if (r.count <= 0) continue;
        if (r.count > limit - r.start) r.count = limit - r.start;
        S.runs[S.nruns++] = r;
        S.covered += r.count;
    }
}
__device__ __forceinline__ int pq_find_run(const PqStream &S, const int32_t v) {  // last run with start <= v
    int lo = 0, hi = S.nruns - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (S.runs[mid].start <= v) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}
__device__ __forceinline__ uint32_t pq_bits(const uint8_t *p, const int64_t bit, const int bw) {  // bw <= 32
    const uint8_t *q = p + (bit >> 3);
    uint64_t w = 0;
    const int nb = (int)(((bit & 7) + bw + 7) >> 3);
    for (int b = 0; b < nb; ++b) w |= (uint64_t)q[b] << (8 * b);
    return (uint32_t)((w >> (bit & 7)) & ((1ull << bw) - 1ull));
}

__device__ __forceinline__ uint64_t pq_load_u64(const uint8_t *base, const int64_t idx) {  // base + 8 idx, any alignment
    const uint8_t *a = base + 8 * idx;
    const int mis = (int)((uintptr_t)a & 7);
    const uint64_t *w = reinterpret_cast<const uint64_t *>(a - mis);
    if (mis == 0) return w[0];
    return (w[0] >> (8 * mis)) | (w[1] << (64 - 8 * mis));
}

// One CTA per data page.
__global__ void __launch_bounds__(PQ_THREADS) pq_decode_kernel(const PqPage *pages, const int first, const uint8_t *image,
                                                                const uint8_t *scratch, int32_t *aux, const PqColOut *cols,
                                                                int32_t *status) {
    __shared__ PqStream S;
    __shared__ uint32_t words[PQ_THREADS];
    __shared__ int32_t prefix[PQ_THREADS];
    __shared__ int32_t warp_tot[PQ_THREADS / 32];
    __shared__ int32_t vbase_sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const PqPage pg = pages[first + blockIdx.x];
    if (pg.kind == PQ_DICT) return;
    const PqColOut out = cols[pg.col];
    const int32_t nv = pg.num_values;
    // the page body: levels (optional columns) then values
    const uint8_t *body = pg.dst >= 0 ? scratch + pg.dst : image + pg.src + pg.lvl_bytes;
    int64_t body_len = (pg.dst >= 0 || pg.kind == PQ_DATA_V2 ? pg.uncomp_size : pg.comp_size) - pg.lvl_bytes;
    const uint8_t *lv = nullptr;
    int32_t lv_len = 0;
    bool bad = false;
    if (pg.optional) {
        if (pg.kind == PQ_DATA_V2) {
            lv = image + pg.src;
            lv_len = pg.lvl_bytes;
        } else {
            if (body_len < 4) {
                bad = true;
            } else {
                lv_len = (int32_t)body[0] | ((int32_t)body[1] << 8) | ((int32_t)body[2] << 16) | ((int32_t)body[3] << 24);
                lv = body + 4;
                if (lv_len < 0 || 4 + (int64_t)lv_len > body_len) bad = true;
                body += 4 + (int64_t)lv_len;
                body_len -= 4 + (int64_t)lv_len;
            }
        }
    }
    if (bad) {
        if (tid == 0) pq_fail(status);
        return;
    }
    // ---- dictionary-encoded values: indices of the page's values into aux --------------------------------------------------
    const uint8_t *dict = nullptr;
    int32_t dict_n = 0;
    int32_t *ix = nullptr;
    if (pg.dict_enc) {
        const PqPage dp = pages[pg.dict];
        dict = dp.dst >= 0 ? scratch + dp.dst : image + dp.src;
        dict_n = dp.num_values;
        ix = aux + pg.aux;
        const int bw = body_len > 0 ? body[0] : 0;
        if (bw > 32 || body_len < 1) {
            if (tid == 0) pq_fail(status);
            return;
        }
        const uint8_t *sp = body + 1;
        const int32_t slen = (int32_t)(body_len - 1);
        if (tid == 0) {
            S.nruns = 0;
            S.covered = 0;
            S.pos = 0;
            S.bad = 0;
        }
        __syncthreads();
        int32_t done = 0;
        for (;;) {
            if (tid == 0) {
                S.nruns = 0;
                pq_parse_runs(S, sp, slen, bw, nv, nv);
            }
            __syncthreads();
            const int32_t cov = S.covered;
            if (S.bad || cov == done) break;  // (the indices of a page with nulls end before nv)
            for (int32_t v = done + tid; v < cov; v += PQ_THREADS) {
                const PqRun r = S.runs[pq_find_run(S, v)];
                const uint32_t x = r.packed ? pq_bits(sp + r.at, (int64_t)(v - r.start) * bw, bw) : (uint32_t)r.at;
                ix[v] = (int32_t)x;
            }
            done = cov;
            __syncthreads();
        }
        if (S.bad) {
            if (tid == 0) pq_fail(status);
            return;
        }
        __syncthreads();
        __threadfence_block();
    }
    // ---- rows -----------------------------------------------------------------------------------------------------------------
    if (tid == 0) {
        S.nruns = 0;
        S.covered = 0;
        S.pos = 0;
        S.bad = 0;
        vbase_sh = 0;
    }
    __syncthreads();
    int32_t r0 = 0;  // first row of the trip
    while (r0 < nv) {
        int32_t rows = min(PQ_BATCH, nv - r0);
        if (pg.optional) {
            if (tid == 0) {
                // drop the runs that lie before r0, then parse ahead
                int keep = 0;
                while (keep < S.nruns && S.runs[keep].start + S.runs[keep].count <= r0) ++keep;
                if (keep) {
                    for (int i = keep; i < S.nruns; ++i) S.runs[i - keep] = S.runs[i];
                    S.nruns -= keep;
                }
                pq_parse_runs(S, lv, lv_len, 1, min(nv, r0 + PQ_BATCH), nv);
            }
            __syncthreads();
            if (S.bad) break;
            int32_t cov = S.covered - r0;
            if (cov < rows) {  // the table filled up first (many short runs): a shorter trip, on a word boundary
                if (S.pos >= lv_len && S.nruns < PQ_RUNS) {  // the stream ended early
                    bad = true;
                    break;
                }
                rows = cov & ~31;
                if (rows <= 0) {
                    bad = true;
                    break;
                }
            }
            // validity word of rows [r0 + 32 tid, + 32)
            uint32_t w = 0;
            const int32_t wr0 = r0 + 32 * tid;
            if (32 * tid < rows) {
                const int32_t wend = min(r0 + rows, wr0 + 32);
                int ri = pq_find_run(S, wr0);
                int32_t v = wr0;
                while (v < wend) {
                    const PqRun r = S.runs[ri];
                    const int32_t e = min(wend, r.start + r.count);
                    const int n = e - v;
                    uint32_t bits;
                    if (r.packed)
                        bits = n > 0 ? pq_bits(lv + r.at, v - r.start, n) : 0u;  // (bit width 1: n values = n bits)
                    else
                        bits = r.at ? (n >= 32 ? 0xFFFFFFFFu : (1u << n) - 1u) : 0u;
                    w |= bits << (v - wr0);
                    v = e;
                    ++ri;
                }
            }
            words[tid] = w;
        } else {
            const int32_t left = rows - 32 * tid;
            words[tid] = left >= 32 ? 0xFFFFFFFFu : (left > 0 ? (1u << left) - 1u : 0u);
        }
        // exclusive scan of the popcounts (value index of the first row of every word)
        const int c = __popc(words[tid]);
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int before = vbase_sh;
        for (int u = 0; u < warp; ++u) before += warp_tot[u];
        prefix[tid] = before + incl - c;
        __syncthreads();
        int total = 0;
        for (int u = 0; u < PQ_THREADS / 32; ++u) total += warp_tot[u];
        // validity bits of the trip into the column bitmap (pages start at any row: two partial words per word)
        if (out.validity && words[tid]) {
            const int64_t gbit = pg.row0 + r0 + 32 * (int64_t)tid;
            const int sh = (int)(gbit & 31);
            atomicOr(out.validity + (gbit >> 5), words[tid] << sh);
            if (sh && (words[tid] >> (32 - sh))) atomicOr(out.validity + (gbit >> 5) + 1, words[tid] >> (32 - sh));
        }
        // values, row by row (coalesced stores; null slots hold 0)
        for (int32_t i = tid; i < rows; i += PQ_THREADS) {
            const uint32_t w = words[i >> 5];
            const int b = i & 31;
            uint64_t x = 0;
            if ((w >> b) & 1u) {
                const int32_t vi = prefix[i >> 5] + __popc(w & ((1u << b) - 1u));
                if (pg.dict_enc) {
                    const int32_t di = ix[vi];
                    if ((uint32_t)di < (uint32_t)dict_n) x = pq_load_u64(dict, di);
                    else bad = true;
                } else if (8 * ((int64_t)vi + 1) <= body_len) {
                    x = pq_load_u64(body, vi);
                } else {
                    bad = true;
                }
            }
            out.values[pg.row0 + r0 + i] = x;
        }
        __syncthreads();
        if (tid == 0) vbase_sh += total;
        r0 += rows;
        __syncthreads();
    }
    if (tid == 0 && out.valid_count) atomicAdd(out.valid_count, (unsigned long long)vbase_sh);
    if (bad || S.bad) pq_fail(status);
}

}  // namespace

int launch_pq_decode(const PqPage *d_pages, int first, int npages, const uint8_t *image, uint8_t *scratch, int32_t *aux,
                     const PqColOut *d_cols, int32_t *status, cudaStream_t stream) {
    if (npages == 0) return 0;
    const int wpb = 4;
    pq_decompress_kernel<<<(npages + wpb - 1) / wpb, wpb * 32, 0, stream>>>(d_pages, first, npages, image, scratch, status);
    pq_decode_kernel<<<npages, PQ_THREADS, 0, stream>>>(d_pages, first, image, scratch, aux, d_cols, status);
    return (int)cudaGetLastError();
}

}  // namespace bowgpu
