// tcgen05 (3xTF32) implicit-GEMM convolution -- placeholder until the tensor-core path lands.
#include "common.cuh"
namespace dtb200 {
int launch_conv_tc(const dtb200_conv_params&, int, cudaStream_t) {
  return fail(DTB200_ERR_UNSUPPORTED, "conv: math=TC3X not built yet%s");
}
int launch_pack_tc(const float*, float*, int, int, int, cudaStream_t) {
  return fail(DTB200_ERR_UNSUPPORTED, "pack: math=TC3X not built yet%s");
}
uint64_t packed_floats_tc(int out_c, int in_c, int ksize) { return (uint64_t)out_c * in_c * ksize * ksize; }
int launch_cost_volume_tc(const dtb200_cost_volume_params&, cudaStream_t) {
  return fail(DTB200_ERR_UNSUPPORTED, "cost volume: math=TC3X not built yet%s");
}
}  // namespace dtb200
