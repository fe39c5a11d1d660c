// segmc kernels, policy: basic ops mask 0 (MC_SUM | MC_MINMAX | MC_FIRSTLAST), integral ops mask 2 (MC_STEP | MC_TRAP)
#define MC_INST_NAME launch_segmc_i2
#define MC_INST_BOPS 0
#define MC_INST_IOPS 2
#include "segmc_inst.cuh"
