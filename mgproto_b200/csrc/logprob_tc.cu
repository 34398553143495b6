// placeholder until the tcgen05 kernel lands (built only with -DMGP_WITH_TC)
#include "mgp_common.cuh"
