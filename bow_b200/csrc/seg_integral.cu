// seg_integral.cu — launchers of the integral family (policy: seg_integral.cuh)
#include "seg_integral.cuh"

namespace bowgpu {

namespace {

template <bool STEP, bool TRAP>
int launch_mode(const IntLaunch &L, int sm, cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1) {
    const bool nulls = L.validity != nullptr;
    auto fill = [&](auto &A) {
        A.time = L.time;
        A.values = L.values;
        A.validity = L.validity;
        A.g = L.g;
        A.out = L.out;
        A.carry_head = (ICarry *)L.carry_head;
        A.carry_tail = (ICarry *)L.carry_tail;
        A.skip = (ICarry *)L.skip;
        A.status = L.status;
        A.syn = L.syn;
        A.gate = L.gate;
        A.gate_lanes = L.gate_lanes;
    };
    if (L.is_int) {
        SegArgs<IntegralPol<STEP, TRAP, true>> A;
        fill(A);
        return nulls ? seg_launch<IntegralPol<STEP, TRAP, true>, true, SEG_INT_CTAS>(A, sm, s, e0, e1)
                     : seg_launch<IntegralPol<STEP, TRAP, true>, false, SEG_INT_CTAS>(A, sm, s, e0, e1);
    }
    SegArgs<IntegralPol<STEP, TRAP, false>> A;
    fill(A);
    return nulls ? seg_launch<IntegralPol<STEP, TRAP, false>, true, SEG_INT_CTAS>(A, sm, s, e0, e1)
                 : seg_launch<IntegralPol<STEP, TRAP, false>, false, SEG_INT_CTAS>(A, sm, s, e0, e1);
}

}  // namespace

size_t integral_carry_bytes(int64_t n) { return (size_t)((n + SEG_T - 1) / SEG_T) * 2 * sizeof(ICarry); }
size_t integral_skip_bytes(int64_t n) { return (size_t)seg_skip_records((n + SEG_T - 1) / SEG_T) * sizeof(ICarry); }

int launch_segreduce_integral(const IntLaunch &L, int sm_count, cudaStream_t stream, cudaEvent_t e0, cudaEvent_t e1) {
    const bool step = L.out.step != nullptr, trap = L.out.trap != nullptr;
    if (step && trap) return launch_mode<true, true>(L, sm_count, stream, e0, e1);
    if (trap) return launch_mode<false, true>(L, sm_count, stream, e0, e1);
    return launch_mode<true, false>(L, sm_count, stream, e0, e1);
}

}  // namespace bowgpu
