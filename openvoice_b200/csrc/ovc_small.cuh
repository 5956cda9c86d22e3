// ovc_small.cuh -- the non-GEMM kernels of the hot path: speaker-conditioning mat-vec,
// conv_post (+leaky_relu 0.01, tanh), latent copy-out.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ovc {

// ---------------------------------------------------------------------------------------------
// cond_kernel: every 1x1 conv the reference applies to the [B, gin, 1] speaker embedding
// (WN.cond_layer of enc_q and of the 4 couplings, modules.py:189-190; Generator.cond,
// models.py:274-275) is a mat-vec.  All of them run in ONE launch: row `i` of the stacked,
// pre-permuted matrix dotted with g_src / g_tgt / zeros (zero_g) of batch item b.  The bias of
// the conv that consumes the result (in_layer / conv_pre) is pre-added on the host, so the conv
// epilogues add a single per-(batch,row) vector.  One warp per output, warp-shuffle reduction.
// ---------------------------------------------------------------------------------------------
struct CondArgs {
  const float* w;        // [rows_w][gin]
  const float* bias;     // [rows_out]
  const int* w_row;      // [rows_out] matrix row feeding output row i
  const int* sel;        // [rows_out] 0 = zeros, 1 = g_src, 2 = g_tgt
  const float* g_src; const float* g_tgt;   // [B][gin]
  float* out;            // [B][rows_out]
  int rows_out; int gin;
};

__global__ void __launch_bounds__(256) cond_kernel(const CondArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  if (warp >= a.rows_out) return;
  const int sel = a.sel[warp];
  float s = 0.f;
  if (sel != 0) {
    const float* g = (sel == 1 ? a.g_src : a.g_tgt) + (size_t)b * a.gin;
    const float* w = a.w + (size_t)a.w_row[warp] * a.gin;
    for (int c = lane; c < a.gin; c += 32) s = fmaf(w[c], g[c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  if (lane == 0) a.out[(size_t)b * a.rows_out + warp] = s + a.bias[warp];
}

// ---------------------------------------------------------------------------------------------
// conv_post_kernel: y[b, t] = tanh( sum_{ci<C, k<7} w[ci,k] * lrelu_0.01(x[b, ci, t+k-3]) )
// (models.py:287-289; conv_post has no bias, models.py:266).  3.4 FLOP/B -> HBM-bound: every
// thread produces 4 consecutive samples from aligned 16-byte loads; neighbouring threads'
// halo vectors hit L1.
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) conv_post_kernel(const float* __restrict__ x, long long x_bs, int x_pitch,
                                                        const float* __restrict__ w, float* __restrict__ y,
                                                        long long y_bs, int y_len, const long long* lens, int tmax,
                                                        int mul) {
  __shared__ float ws[C * 7];
  for (int i = threadIdx.x; i < C * 7; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int lim = (lens ? (int)min((long long)tmax, lens[b]) : tmax) * mul;
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (t >= y_len) return;
  float* yp = y + (size_t)b * y_bs + t;
  if (t >= lim) {
    *reinterpret_cast<float4*>(yp) = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float* xb = x + (size_t)b * x_bs;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int ci = 0; ci < C; ++ci) {
    const float* xr = xb + (size_t)ci * x_pitch;
    float win[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int tt = t - 4 + 4 * i;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tt >= 0 && tt < lim) q = *reinterpret_cast<const float4*>(xr + tt);
      const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = (tt + j < lim) ? e[j] : 0.f;
        win[4 * i + j] = v > 0.f ? v : 0.01f * v;
      }
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const float wk = ws[ci * 7 + k];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(wk, win[1 + k + j], acc[j]);
    }
  }
  float o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = (t + j < lim) ? tanhf(acc[j]) : 0.f;
  *reinterpret_cast<float4*>(yp) = make_float4(o[0], o[1], o[2], o[3]);
}

// latent copy-out: internal [B][C][pitch] -> caller [B][C][tmax], zero past the length (the
// reference returns masked latents, models.py:220 and modules.py:449,454)
__global__ void __launch_bounds__(256) copy_latent_kernel(const float* __restrict__ src, int pitch, float* __restrict__ dst,
                                                          int tmax, int C, const long long* lens) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= tmax) return;
  const int len = (int)min((long long)tmax, lens[b]);
  const float v = t < len ? src[((size_t)b * C + c) * pitch + t] : 0.f;
  dst[((size_t)b * C + c) * tmax + t] = v;
}

}  // namespace ovc
