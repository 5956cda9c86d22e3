// ovc_variants.h -- the list of conv1d_f32<> instantiations, split into groups so that the
// translation units compile in parallel.  X(name, K, DIL, WM, WN, CI_CH, EPI, NG, XALIGN)
#pragma once
#include "ovc_conv.cuh"


// posterior encoder + flow (time axis = spectrogram frames): 64 rows x 128 frames, 128 threads
#define OVC_VARIANTS_G0(X)                                  \
  X(ENC_PRE, 1, 1, 2, 2, 8, EPI_LINEAR, 1, 4)               \
  X(FLOW_PRE, 1, 1, 2, 2, 8, EPI_LINEAR, 1, 16)             \
  X(WN_IN, 5, 1, 2, 2, 8, EPI_GATE, 1, 16)                  \
  X(WN_RS, 1, 1, 2, 2, 8, EPI_RESSKIP, 1, 16)               \
  X(ENC_PROJ, 1, 1, 2, 2, 8, EPI_PROJ, 1, 16)               \
  X(FLOW_POST, 1, 1, 1, 4, 8, EPI_COUPLE, 1, 16)            \
  X(UPS8_A, 3, 1, 4, 2, 8, EPI_UPS8, 1, 16)                 \
  X(UPS2_A, 3, 1, 4, 2, 8, EPI_UPS2, 1, 16)                 \
  X(UPS2_B, 3, 1, 2, 4, 8, EPI_UPS2, 1, 16)                 \
  X(TXT_K3D1, 3, 1, 2, 2, 8, EPI_LINEAR, 1, 16)   /* TTS text side: k = 3 convs on short sequences (64 x 128 tiles) */

// generator, class A: 128 rows x 128 samples (C = 512, 256, 128)
#define OVC_VARIANTS_G1(X)                                  \
  X(A_K3D1, 3, 1, 4, 2, 8, EPI_LINEAR, 1, 16)               \
  X(A_K3D3, 3, 3, 4, 2, 8, EPI_LINEAR, 1, 16)               \
  X(A_K3D5, 3, 5, 4, 2, 8, EPI_LINEAR, 1, 16)               \
  X(A_K7D1, 7, 1, 4, 2, 4, EPI_LINEAR, 1, 16)               \
  X(A_K7D3, 7, 3, 4, 2, 8, EPI_LINEAR, 2, 16)               \
  X(A_K7D5, 7, 5, 4, 2, 8, EPI_LINEAR, 2, 16)
#define OVC_VARIANTS_G2(X)                                  \
  X(A_K11D1, 11, 1, 4, 2, 4, EPI_LINEAR, 1, 16)             \
  X(A_K11D3, 11, 3, 4, 2, 8, EPI_LINEAR, 3, 16)             \
  X(A_K11D5, 11, 5, 4, 2, 8, EPI_LINEAR, 4, 16)

// generator, class B: 64 rows x 256 samples (C = 64)
#define OVC_VARIANTS_G3(X)                                  \
  X(B_K3D1, 3, 1, 2, 4, 8, EPI_LINEAR, 1, 16)               \
  X(B_K3D3, 3, 3, 2, 4, 8, EPI_LINEAR, 1, 16)               \
  X(B_K3D5, 3, 5, 2, 4, 8, EPI_LINEAR, 1, 16)               \
  X(B_K7D1, 7, 1, 2, 4, 8, EPI_LINEAR, 1, 16)               \
  X(B_K7D3, 7, 3, 2, 4, 8, EPI_LINEAR, 2, 16)               \
  X(B_K7D5, 7, 5, 2, 4, 8, EPI_LINEAR, 2, 16)
#define OVC_VARIANTS_G4(X)                                  \
  X(B_K11D1, 11, 1, 2, 4, 8, EPI_LINEAR, 1, 16)             \
  X(B_K11D3, 11, 3, 2, 4, 8, EPI_LINEAR, 3, 16)             \
  X(B_K11D5, 11, 5, 2, 4, 8, EPI_LINEAR, 4, 16)

// generator, class C: 32 rows x 512 samples (C = 32)
#define OVC_VARIANTS_G5(X)                                  \
  X(C_K3D1, 3, 1, 1, 8, 8, EPI_LINEAR, 1, 16)               \
  X(C_K3D3, 3, 3, 1, 8, 8, EPI_LINEAR, 1, 16)               \
  X(C_K3D5, 3, 5, 1, 8, 8, EPI_LINEAR, 1, 16)               \
  X(C_K7D1, 7, 1, 1, 8, 8, EPI_LINEAR, 1, 16)               \
  X(C_K7D3, 7, 3, 1, 8, 8, EPI_LINEAR, 2, 16)               \
  X(C_K7D5, 7, 5, 1, 8, 8, EPI_LINEAR, 2, 16)
#define OVC_VARIANTS_G6(X)                                  \
  X(C_K11D1, 11, 1, 1, 8, 8, EPI_LINEAR, 1, 16)             \
  X(C_K11D3, 11, 3, 1, 8, 8, EPI_LINEAR, 3, 16)             \
  X(C_K11D5, 11, 5, 1, 8, 8, EPI_LINEAR, 4, 16)

#define OVC_VARIANTS_ALL(X) \
  OVC_VARIANTS_G0(X) OVC_VARIANTS_G1(X) OVC_VARIANTS_G2(X) OVC_VARIANTS_G3(X) OVC_VARIANTS_G4(X) \
  OVC_VARIANTS_G5(X) OVC_VARIANTS_G6(X)

namespace ovc {

enum Variant : int {
#define X(name, K, D, WM, WN, CI, EPI, NG, XA) V_##name,
  OVC_VARIANTS_ALL(X)
#undef X
  V_COUNT
};

struct VariantInfo {
  const char* name;
  int K, DIL, CO_T, T_T, CI_CH, EPI, THREADS;
  size_t smem;
};

typedef cudaError_t (*LaunchFn)(const ConvArgs&, int t_len, int row_tiles, int B, cudaStream_t);
typedef cudaError_t (*PrepareFn)();

#define X(name, K, D, WM, WN, CI, EPI, NG, XA)                                        \
  cudaError_t launch_##name(const ConvArgs&, int, int, int, cudaStream_t);            \
  cudaError_t prepare_##name();
OVC_VARIANTS_ALL(X)
#undef X

}  // namespace ovc
