// Parquet ingest (bow.NewBowFromParquet, bowparquet.go:44-155) — internal interface between the host-side file
// walk (footer + page headers, parquet.cu) and the C ABI (api.cu).  Every byte of column DATA is decoded on the GPU:
// Snappy decompression, RLE / bit-packed definition levels -> Arrow validity bitmap, PLAIN and dictionary-encoded
// INT64 / DOUBLE values -> bow.NewBuffer layout (value 0 in null slots).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace bowgpu {

// one page of a column chunk as the decode kernels see it
struct PqPage {
    int64_t src;         // offset of the page body (right after its header) in the device image of the file bytes
    int64_t dst;         // offset of the uncompressed body in the scratch buffer (16-byte aligned); -1 = not compressed
    int64_t row0;        // data pages: first row of the page in its column
    int64_t aux;         // dictionary-encoded data pages: first entry of the page in the index scratch (int32 per value)
    int32_t comp_size;   // bytes of the body in the file
    int32_t uncomp_size; // bytes of the body once uncompressed
    int32_t num_values;  // data pages: rows (flat schema: values incl. nulls); dictionary pages: entries
    int32_t col;         // output column
    int32_t kind;        // PQ_DATA_V1 / PQ_DATA_V2 / PQ_DICT
    int32_t codec;       // 0 UNCOMPRESSED, 1 SNAPPY
    int32_t dict_enc;    // data pages: values are indices into the dictionary page `dict`
    int32_t optional;    // the column has definition levels (max level 1)
    int32_t lvl_bytes;   // v2: bytes of the levels stored in front of the values (never compressed)
    int32_t dict;        // descriptor index of the chunk's dictionary page, or -1
    int32_t _pad[2];
};
enum { PQ_DATA_V1 = 0, PQ_DATA_V2 = 1, PQ_DICT = 2 };

struct PqColOut {
    uint64_t *values;    // [n]
    uint32_t *validity;  // zeroed bitmap (bit offset 0) or null for a required column
    unsigned long long *valid_count;  // device counter (zeroed): valid rows written
};

struct PqColumn {
    std::string name;
    int32_t physical;   // parquet.Type: 0 BOOLEAN 1 INT32 2 INT64 3 INT96 4 FLOAT 5 DOUBLE 6 BYTE_ARRAY 7 FIXED_LEN_BYTE_ARRAY
    int32_t dtype;      // BOWGPU_INT64 / BOWGPU_FLOAT64, 0 = no GPU type (Boolean, String, anything else)
    bool optional;
};

struct PqRange {  // bytes of the file the selected columns need, and where they sit in the device image
    int64_t file_off, len, image_off;
};

struct PqPlan {
    std::vector<PqRange> ranges;
    std::vector<PqPage> pages;
    std::vector<int32_t> range_first_page;  // pages of ranges[r] = [range_first_page[r], range_first_page[r + 1])
    int64_t image_bytes = 0, scratch_bytes = 0, aux_entries = 0;
};

struct PqFile;  // mapped file + parsed footer
int pq_open(const char *path, PqFile **out, std::string &err);  // 0 or a BOWGPU_* status
void pq_close(PqFile *f);
int64_t pq_num_rows(const PqFile *f);
const std::vector<PqColumn> &pq_columns(const PqFile *f);
const uint8_t *pq_bytes(const PqFile *f);
int pq_fd(const PqFile *f);
// walks the page headers of the chosen leaf columns (indices into pq_columns, output column j = cols[j])
int pq_plan(const PqFile *f, const int32_t *cols, int32_t ncols, PqPlan &plan, std::string &err);

// decompress (one warp per compressed page) and decode (one CTA per data page); ST_PARQUET in *status on malformed data
// (pages [first, first + count) of the table; dictionary references are indices into the whole table)
int launch_pq_decode(const PqPage *d_pages, int first, int count, const uint8_t *image, uint8_t *scratch, int32_t *aux,
                     const PqColOut *d_cols, int32_t *status, cudaStream_t stream);

}  // namespace bowgpu
