// segmc kernels, policy: basic ops mask 7 (MC_SUM | MC_MINMAX | MC_FIRSTLAST), integral ops mask 3 (MC_STEP | MC_TRAP)
#define MC_INST_NAME launch_segmc_all
#define MC_INST_BOPS 7
#define MC_INST_IOPS 3
#include "segmc_inst.cuh"
