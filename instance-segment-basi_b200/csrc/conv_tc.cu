// tcgen05 / TMEM / TMA implicit-GEMM convolution path (placeholder until the kernels land).
#include "common.cuh"
extern "C" {
int basi_tc_conv_supported(int, const basi_conv_desc*, const basi_tensor*, const basi_tensor*) { return 0; }
int basi_tc_pack_weights(const float*, void*, void*, int, int, int, void*) {
  basi::set_error("tc path not built");
  return BASI_E_INVALID;
}
int basi_tc_conv_create(int, const basi_conv_desc*, const basi_tensor*, const basi_tensor*, const void*, float*, int,
                        basi_tc_conv**) {
  basi::set_error("tc path not built");
  return BASI_E_INVALID;
}
int basi_tc_conv_run(basi_tc_conv*, void*) { return BASI_E_INVALID; }
void basi_tc_conv_destroy(basi_tc_conv*) {}
}
