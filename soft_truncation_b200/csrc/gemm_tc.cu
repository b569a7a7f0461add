// placeholder until the tcgen05 backend lands
#include "common.cuh"
extern "C" __attribute__((visibility("default"))) int st_tc_available(void) { return 0; }
int st_gemm_tc_supported(const st_gemm_args*, const char** why) { *why = "not built"; return 0; }
int st_gemm_tc(const st_gemm_args*, cudaStream_t) { st_set_error("tcgen05 backend not built"); return ST_ERR_UNSUPPORTED; }
