// placeholder: replaced by the tcgen05 fused attention kernels
#include "common.cuh"
#include "../../include/dle_b200.h"
extern "C" int dle_attn_fwd(const void*, const float*, void*, float*, int32_t, int32_t, int32_t, float, uint64_t, uint32_t, void*) { return DLE_ERR_NOSYS; }
extern "C" int dle_attn_bwd(const void*, const float*, const void*, const void*, const float*, void*, float*, int32_t, int32_t, int32_t, float, uint64_t, uint32_t, void*) { return DLE_ERR_NOSYS; }
