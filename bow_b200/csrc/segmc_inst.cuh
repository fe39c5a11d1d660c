// One policy of the segmc kernel family per translation unit (so that `make -j` compiles them side by side):
// the including .cu defines MC_INST_NAME, MC_INST_BOPS and MC_INST_IOPS.
#include "segmc.cuh"

namespace bowgpu {

int MC_INST_NAME(const McLaunch &L, int sm, cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1) {
    const bool nulls = L.col[0].validity != nullptr;
    if (L.is_int) {
        using Pol = McPol<MC_INST_BOPS, MC_INST_IOPS, true>;
        return nulls ? mc_launch<Pol, true>(L, sm, s, e0, e1) : mc_launch<Pol, false>(L, sm, s, e0, e1);
    }
    using Pol = McPol<MC_INST_BOPS, MC_INST_IOPS, false>;
    return nulls ? mc_launch<Pol, true>(L, sm, s, e0, e1) : mc_launch<Pol, false>(L, sm, s, e0, e1);
}

}  // namespace bowgpu
