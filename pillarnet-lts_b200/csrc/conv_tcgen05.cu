// placeholder until the tcgen05 implicit-GEMM kernel lands
#include "common.cuh"
namespace pn_detail {
int conv_tcgen05(const pn_conv_args*, cudaStream_t) { return PN_ERR_UNSUPPORTED; }
}
