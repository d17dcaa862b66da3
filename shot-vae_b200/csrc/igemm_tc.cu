// tcgen05 + TMA implicit-GEMM path (placeholder until the kernel lands: reports "unsupported").
#include "igemm.h"
bool igemm_fprop_tc_supported(const IgemmParams&) { return false; }
int igemm_fprop_tc(const IgemmParams&, cudaStream_t) { sv_set_error("tcgen05 path not built"); return SV_ERR_UNSUPPORTED; }
