// seg_basic.cu — launchers of the basic family (policy: seg_basic.cuh)
#include "seg_basic.cuh"

namespace bowgpu {

namespace {

template <uint32_t OPS>
int launch_ops(const SegLaunch &L, int sm, cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1) {
    const bool nulls = L.validity != nullptr;
    auto fill = [&](auto &A) {
        A.time = L.time;
        A.values = L.values;
        A.validity = L.validity;
        A.g = L.g;
        A.out = L.out;
        A.carry_head = L.carry_head;
        A.carry_tail = L.carry_tail;
        A.skip = L.skip;
        A.status = L.status;
        A.syn = L.syn;
        A.gate = L.gate;
        A.gate_lanes = L.gate_lanes;
    };
    if (L.is_int) {
        SegArgs<BasicPol<OPS, true>> A;
        fill(A);
        return nulls ? seg_launch<BasicPol<OPS, true>, true, SEG_CFG_CTAS>(A, sm, s, e0, e1)
                     : seg_launch<BasicPol<OPS, true>, false, SEG_CFG_CTAS>(A, sm, s, e0, e1);
    }
    SegArgs<BasicPol<OPS, false>> A;
    fill(A);
    return nulls ? seg_launch<BasicPol<OPS, false>, true, SEG_CFG_CTAS>(A, sm, s, e0, e1)
                 : seg_launch<BasicPol<OPS, false>, false, SEG_CFG_CTAS>(A, sm, s, e0, e1);
}

}  // namespace

int64_t seg_num_tiles(int64_t n) { return (n + SEG_T - 1) / SEG_T; }
size_t seg_carry_bytes(int64_t n) { return (size_t)seg_num_tiles(n) * 2 * sizeof(BasicCarry); }
size_t seg_skip_bytes(int64_t n) { return (size_t)seg_skip_records(seg_num_tiles(n)) * sizeof(BasicCarry); }

int launch_segreduce_basic(const SegLaunch &L, int sm_count, cudaStream_t stream, cudaEvent_t e0, cudaEvent_t e1) {
    if (L.ops & OPS_FIRSTLAST) return launch_ops<OPS_SUMCNT | OPS_MINMAX | OPS_FIRSTLAST>(L, sm_count, stream, e0, e1);
    if (L.ops & OPS_MINMAX) return launch_ops<OPS_SUMCNT | OPS_MINMAX>(L, sm_count, stream, e0, e1);
    return launch_ops<OPS_SUMCNT>(L, sm_count, stream, e0, e1);
}

}  // namespace bowgpu
