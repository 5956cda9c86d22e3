// conv1d_f32<> instantiations, group 6 (see ovc_variants.h)
#include "ovc_variants.h"
namespace ovc {
#define X(name, K, D, WM, WN, CI, EPI, NG, XA)                                                         \
  cudaError_t launch_##name(const ConvArgs& a, int t_len, int row_tiles, int B, cudaStream_t st) {     \
    return ConvLaunch<ConvCfg<K, D, WM, WN, CI, EPI, NG, XA>>::launch(a, t_len, row_tiles, B, st);     \
  }                                                                                                    \
  cudaError_t prepare_##name() { return ConvLaunch<ConvCfg<K, D, WM, WN, CI, EPI, NG, XA>>::prepare(); }
OVC_VARIANTS_G6(X)
#undef X
}  // namespace ovc
