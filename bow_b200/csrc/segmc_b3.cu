// segmc kernels, policy: basic ops mask 3 (MC_SUM | MC_MINMAX | MC_FIRSTLAST), integral ops mask 0 (MC_STEP | MC_TRAP)
#define MC_INST_NAME launch_segmc_b3
#define MC_INST_BOPS 3
#define MC_INST_IOPS 0
#include "segmc_inst.cuh"
