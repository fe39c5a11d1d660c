/* A plain-C consumer of the drop-in boundary (include/bowgpu.h): no Python, no torch — what the cgo shim of
 * INTEGRATION.md does, written in C.  Replays two of the reference's golden tables
 * (rolling/aggregation/arithmeticmean_test.go:24-41, count_test.go:24-41 on sparseFloatBow, core_test.go:37-53)
 * and one interpolation (rolling/interpolation/linear_test.go) through host Arrow-layout buffers.
 *   gcc -O2 -I include tests/c/abi_demo.c -o tests/c/abi_demo -L bow_b200 -lbowgpu -Wl,-rpath,$PWD/bow_b200
 * exit status 0 = every value matches. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "bowgpu.h"

#define CHECK(call)                                                                                   \
    do {                                                                                              \
        int32_t st__ = (call);                                                                        \
        if (st__ != BOWGPU_OK) {                                                                      \
            fprintf(stderr, "%s -> %s: %s\n", #call, bowgpu_status_string(st__), bowgpu_last_error(ctx)); \
            return 2;                                                                                 \
        }                                                                                             \
    } while (0)

int main(int argc, char **argv) {
    bowgpu_ctx *ctx = NULL;
    if (bowgpu_abi_version() != BOWGPU_ABI_VERSION) return 3;
    int32_t st = bowgpu_ctx_create(0, NULL, &ctx);
    if (st != BOWGPU_OK) {
        fprintf(stderr, "bowgpu_ctx_create: %s (a B200 is required: there is no CPU fallback)\n", bowgpu_status_string(st));
        return 2;
    }
    /* sparseFloatBow: time int64, value float64 with nils (validity bitmap LSB first) */
    int64_t t[9] = {10, 11, 20, 40, 41, 50, 51, 61, 69};
    double v[9] = {10.0, 0, 0, 0, 10.0, 10.0, 20.0, 10.0, 20.0};
    uint8_t vbits[2] = {0xF1 /* rows 0,4,5,6,7 */, 0x01 /* row 8 */};
    bowgpu_col cols[2];
    memset(cols, 0, sizeof cols);
    cols[0].values = t, cols[0].length = 9, cols[0].dtype = BOWGPU_INT64;
    cols[1].values = v, cols[1].validity = vbits, cols[1].length = 9, cols[1].null_count = 3, cols[1].dtype = BOWGPU_FLOAT64;
    bowgpu_frame *frame = NULL;
    CHECK(bowgpu_frame_create(ctx, cols, 2, BOWGPU_MEM_HOST, &frame));
    bowgpu_rolling *r = NULL;
    CHECK(bowgpu_rolling_create(frame, 0, 10, 0, 0, NULL, &r)); /* IntervalRolling(b, "time", 10, Options{}) */
    const int64_t W = bowgpu_rolling_num_windows(r);
    if (W != 6) return 4;

    bowgpu_agg_spec specs[3];
    memset(specs, 0, sizeof specs);
    specs[0].op = BOWGPU_AGG_WINDOW_START, specs[0].col = 0;
    specs[1].op = BOWGPU_AGG_MEAN, specs[1].col = 1;
    specs[2].op = BOWGPU_AGG_COUNT, specs[2].col = 1;
    int64_t ws[6], cnt[6];
    double mean[6];
    uint8_t bm[3][1];
    bowgpu_out_col outs[3] = {{ws, bm[0], 0, 0}, {mean, bm[1], 0, 0}, {cnt, bm[2], 0, 0}};
    CHECK(bowgpu_rolling_aggregate(r, specs, 3, outs, BOWGPU_MEM_HOST));
    const int64_t want_ws[6] = {10, 20, 30, 40, 50, 60}, want_cnt[6] = {1, 0, 0, 1, 2, 2};
    const double want_mean[6] = {10.0, 0, 0, 10.0, 15.0, 15.0};
    const uint8_t want_mean_valid = 0x39; /* windows 0, 3, 4, 5 */
    int bad = 0;
    for (int k = 0; k < 6; ++k) bad |= ws[k] != want_ws[k] || cnt[k] != want_cnt[k] || mean[k] != want_mean[k];
    bad |= bm[0][0] != 0x3F || bm[2][0] != 0x3F || bm[1][0] != want_mean_valid;
    bad |= outs[0].dtype != BOWGPU_INT64 || outs[1].dtype != BOWGPU_FLOAT64 || outs[2].dtype != BOWGPU_INT64;
    printf("aggregate: %s\n", bad ? "MISMATCH" : "ok");

    /* Interpolate(WindowStart(time), Linear(value)) then download the interpolated Bow */
    const int32_t ops[2] = {BOWGPU_INTERP_WINDOW_START, BOWGPU_INTERP_LINEAR};
    bowgpu_frame *fi = NULL;
    int64_t n_out = 0;
    CHECK(bowgpu_rolling_interpolate(r, ops, 2, &fi, &n_out));
    int64_t it[16];
    double iv[16];
    uint8_t ib[2][2];
    bowgpu_out_col dl[2] = {{it, ib[0], 0, 0}, {iv, ib[1], 0, 0}};
    if (n_out > 16) return 5;
    CHECK(bowgpu_frame_download(fi, dl, 2));
    /* windows 20 and 40 start on a row; 30 (empty), 50?no: 50 is a row; 60 has no row at 60 -> synthetic rows at 30 and 60 */
    const int64_t want_t[11] = {10, 11, 20, 30, 40, 41, 50, 51, 60, 61, 69};
    int ibad = n_out != 11;
    for (int i = 0; i < 11 && !ibad; ++i) ibad |= it[i] != want_t[i];
    /* Linear at 30: between (10, 10.0) and (41, 10.0) -> 10.0 ; at 60: between (51, 20.0) and (61, 10.0) -> 11.0 */
    ibad |= !(((ib[1][0] >> 3) & 1) && fabs(iv[3] - 10.0) < 1e-12);
    ibad |= !(((ib[1][1] >> 0) & 1) && fabs(iv[8] - 11.0) < 1e-12);
    printf("interpolate: %s (%lld rows)\n", ibad ? "MISMATCH" : "ok", (long long)n_out);

    /* Bow.SortByCol(0) on the reference's duplicate-key table (bowsort_test.go:133-157); NULL = already sorted */
    int64_t st_[4] = {13, 12, 12, 10};
    double sa[4] = {3.9, 2.9, 2.8, 2.4};
    bowgpu_col scols[2];
    memset(scols, 0, sizeof scols);
    scols[0].values = st_, scols[0].length = 4, scols[0].dtype = BOWGPU_INT64;
    scols[1].values = sa, scols[1].length = 4, scols[1].dtype = BOWGPU_FLOAT64;
    bowgpu_frame *unsorted = NULL, *sorted = NULL, *again = NULL;
    CHECK(bowgpu_frame_create(ctx, scols, 2, BOWGPU_MEM_HOST, &unsorted));
    CHECK(bowgpu_frame_sort_by_col(unsorted, 0, &sorted));
    int sbad = sorted == NULL;
    if (!sbad) {
        int64_t ot[4];
        double oa[4];
        uint8_t ob[2][1];
        bowgpu_out_col sd[2] = {{ot, ob[0], 0, 0}, {oa, ob[1], 0, 0}};
        CHECK(bowgpu_frame_download(sorted, sd, 2));
        const int64_t want_st[4] = {10, 12, 12, 13};
        const double want_sa[4] = {2.4, 2.9, 2.8, 3.9};
        for (int i = 0; i < 4; ++i) sbad |= ot[i] != want_st[i] || oa[i] != want_sa[i];
        CHECK(bowgpu_frame_sort_by_col(sorted, 0, &again));
        sbad |= again != NULL;
    }
    printf("sort: %s\n", sbad ? "MISMATCH" : "ok");

    /* NewBowFromParquet (bowparquet.go:44-155): a file written by the reference's own writer (benchmarks/bow1-10-rows.parquet,
     * SNAPPY + PLAIN + 3-row pages), three of its numeric columns decoded on the device, then IntervalRolling(Int64_ref, 10)
     * + Count over the decoded frame without a host round trip */
    int pbad = 0;
    if (argc > 1) {
        bowgpu_parquet *pq = NULL;
        char msg[256];
        CHECK(bowgpu_parquet_open(argv[1], &pq, msg, (int32_t)sizeof msg));
        int32_t idx[3] = {-1, -1, -1};
        const char *want_names[3] = {"Int64_ref", "Int64_bow1", "Float64_bow1"};
        for (int32_t j = 0; j < bowgpu_parquet_num_cols(pq); ++j)
            for (int c = 0; c < 3; ++c)
                if (strcmp(bowgpu_parquet_col_name(pq, j), want_names[c]) == 0) idx[c] = j;
        pbad |= idx[0] < 0 || idx[1] < 0 || idx[2] < 0 || bowgpu_parquet_num_rows(pq) != 10;
        pbad |= bowgpu_parquet_col_dtype(pq, idx[0]) != BOWGPU_INT64 || bowgpu_parquet_col_dtype(pq, idx[2]) != BOWGPU_FLOAT64;
        bowgpu_frame *pf = NULL;
        CHECK(bowgpu_parquet_read(ctx, pq, idx, 3, &pf));
        int64_t pt[10], pi[10];
        double pv[10];
        uint8_t pb[3][2];
        bowgpu_out_col pd[3] = {{pt, pb[0], 0, 0}, {pi, pb[1], 0, 0}, {pv, pb[2], 0, 0}};
        CHECK(bowgpu_frame_download(pf, pd, 3));
        const int64_t want_t[10] = {3, 18, 28, 32, 42, 55, 63, 73, 89, 92};
        const int64_t want_i[10] = {8, 0, 0, 6, 5, 0, 0, 1, 0, 0};
        const double want_v[10] = {5.5, 9.5, 0.5, 8.5, 0, 5.5, 9.5, 0, 6.5, 0};
        for (int i = 0; i < 10; ++i) pbad |= pt[i] != want_t[i] || pi[i] != want_i[i] || pv[i] != want_v[i];
        pbad |= pb[1][0] != 0x99 || (pb[1][1] & 3) != 0x00 || pb[2][0] != 0x6F || (pb[2][1] & 3) != 0x01;
        pbad |= bowgpu_frame_col_null_count(pf, 0) != 0 || bowgpu_frame_col_null_count(pf, 1) != 6 || bowgpu_frame_col_null_count(pf, 2) != 3;
        bowgpu_rolling *pr = NULL;
        CHECK(bowgpu_rolling_create(pf, 0, 10, 0, 0, NULL, &pr));
        const int64_t PW = bowgpu_rolling_num_windows(pr); /* windows [0,10) .. [90,100) */
        pbad |= PW != 10;
        if (!pbad) {
            bowgpu_agg_spec ps[2];
            memset(ps, 0, sizeof ps);
            ps[0].op = BOWGPU_AGG_WINDOW_START, ps[0].col = 0;
            ps[1].op = BOWGPU_AGG_COUNT, ps[1].col = 2;
            int64_t pws[10], pc[10];
            uint8_t pm[2][2];
            bowgpu_out_col po[2] = {{pws, pm[0], 0, 0}, {pc, pm[1], 0, 0}};
            CHECK(bowgpu_rolling_aggregate(pr, ps, 2, po, BOWGPU_MEM_HOST));
            const int64_t want_c[10] = {1, 1, 1, 1, 0, 1, 1, 0, 1, 0};
            for (int k = 0; k < 10; ++k) pbad |= pws[k] != 10 * k || pc[k] != want_c[k];
        }
        bowgpu_rolling_destroy(pr);
        bowgpu_frame_destroy(pf);
        bowgpu_parquet_close(pq);
        printf("parquet: %s\n", pbad ? "MISMATCH" : "ok");
    }

    bowgpu_frame_destroy(sorted);
    bowgpu_frame_destroy(unsorted);
    bowgpu_frame_destroy(fi);
    bowgpu_rolling_destroy(r);
    bowgpu_frame_destroy(frame);
    bowgpu_ctx_destroy(ctx);
    return (bad || ibad || sbad || pbad) ? 1 : 0;
}
