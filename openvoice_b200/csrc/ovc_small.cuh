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


// ---------------------------------------------------------------------------------------------
// stft_mag_kernel: the linear-spectrogram front end of convert (mel_processing.py:40-75):
// reflect-pad (n_fft-hop)/2 = 384 at BOTH ends of each utterance's own length, periodic hann,
// 1024-point DFT, centre=False, sqrt(re^2 + im^2 + 1e-6).  One CTA transforms 8 consecutive
// frames (radix-4 Stockham FFT in shared memory, twiddles and window from host-computed
// double-precision tables) and writes a [513][8] block so rows are stored 32 B at a time.
// Frames past an utterance's T = L / hop are written as zeros.
// ---------------------------------------------------------------------------------------------
constexpr int STFT_N = 1024, STFT_FR = 8;

__global__ void __launch_bounds__(256) stft_mag_kernel(const float* __restrict__ wav, long long wav_bs,
                                                       const long long* __restrict__ wav_len, int hop,
                                                       float* __restrict__ spec, long long spec_bs, int spec_pitch,
                                                       int Tmax, const float2* __restrict__ tw,
                                                       const float* __restrict__ win, long long* __restrict__ frames_out) {
  __shared__ float2 buf[2][STFT_N];
  __shared__ float mag[STFT_N / 2 + 1][STFT_FR + 1];
  __shared__ float2 tws[STFT_N];
  const int tid = threadIdx.x, b = blockIdx.y, t0 = blockIdx.x * STFT_FR;
  const int L = (int)wav_len[b];
  const int T = min(Tmax, L / hop);
  if (blockIdx.x == 0 && tid == 0 && frames_out) frames_out[b] = T;
  for (int i = tid; i < STFT_N; i += 256) tws[i] = tw[i];
  const float* w = wav + (size_t)b * wav_bs;
  const int pad = (STFT_N - hop) / 2;
  for (int fr = 0; fr < STFT_FR; ++fr) {
    const int t = t0 + fr;
    if (t < T) {   // block-uniform
      for (int n = tid; n < STFT_N; n += 256) {
        int idx = t * hop + n - pad;
        if (idx < 0) idx = -idx;
        if (idx >= L) idx = 2 * (L - 1) - idx;
        idx = max(0, min(idx, L - 1));   // (callers guarantee L > 384; never read out of bounds regardless)
        buf[0][n] = make_float2(w[idx] * win[n], 0.f);
      }
      __syncthreads();
      int src = 0;
#pragma unroll
      for (int Ns = 1; Ns < STFT_N; Ns *= 4) {
        const float2* in = buf[src];
        float2* out = buf[src ^ 1];
        const int k = tid & (Ns - 1);
        const int j0 = ((tid - k) << 2) + k;
        const int tstep = k * (256 / Ns);
        float2 u0 = in[tid], u1 = in[tid + 256], u2 = in[tid + 512], u3 = in[tid + 768];
        const float2 w1 = tws[tstep], w2 = tws[2 * tstep], w3 = tws[3 * tstep];
        u1 = make_float2(u1.x * w1.x - u1.y * w1.y, u1.x * w1.y + u1.y * w1.x);
        u2 = make_float2(u2.x * w2.x - u2.y * w2.y, u2.x * w2.y + u2.y * w2.x);
        u3 = make_float2(u3.x * w3.x - u3.y * w3.y, u3.x * w3.y + u3.y * w3.x);
        const float2 v0 = make_float2(u0.x + u2.x, u0.y + u2.y), v1 = make_float2(u0.x - u2.x, u0.y - u2.y);
        const float2 v2 = make_float2(u1.x + u3.x, u1.y + u3.y);
        const float2 d = make_float2(u1.x - u3.x, u1.y - u3.y);
        const float2 v3 = make_float2(d.y, -d.x);   // (u1 - u3) * (-i)
        out[j0] = make_float2(v0.x + v2.x, v0.y + v2.y);
        out[j0 + Ns] = make_float2(v1.x + v3.x, v1.y + v3.y);
        out[j0 + 2 * Ns] = make_float2(v0.x - v2.x, v0.y - v2.y);
        out[j0 + 3 * Ns] = make_float2(v1.x - v3.x, v1.y - v3.y);
        src ^= 1;
        __syncthreads();
      }
      for (int f = tid; f <= STFT_N / 2; f += 256) {
        const float2 c = buf[src][f];
        mag[f][fr] = sqrtf(c.x * c.x + c.y * c.y + 1e-6f);
      }
    } else {
      for (int f = tid; f <= STFT_N / 2; f += 256) mag[f][fr] = 0.f;
    }
    __syncthreads();
  }
  float* sp = spec + (size_t)b * spec_bs;
  for (int e = tid; e < (STFT_N / 2 + 1) * STFT_FR; e += 256) {
    const int f = e / STFT_FR, fr = e % STFT_FR;
    if (t0 + fr < Tmax) sp[(size_t)f * spec_pitch + t0 + fr] = mag[f][fr];
  }
}

}  // namespace ovc
