// g++ build of openvoice_b200/csrc/ovc_tts_ops.h: the same element functions the CUDA kernels wrap, looped on the
// CPU so that tests/test_tts_ops_host.py can check them against the oracle without a GPU.  TEST CODE ONLY -- it is
// never linked into libovc_b200.so.
#include "../../openvoice_b200/csrc/ovc_tts_ops.h"

using namespace ovc_tts;

extern "C" {

void hc_embed(const long long* tokens, const long long* lens, const float* emb, int B, int T, int H, float* out) {
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < T; ++t)
      for (int c = 0; c < H; ++c)
        out[((size_t)b * T + t) * H + c] = embed_at(tokens, emb, H, sqrtf((float)H), T, b, t, c, (int)lens[b]);
}

void hc_layer_norm(const float* a, const float* r, const float* res, const float* gamma, const float* beta, int rows, int C,
                   int pre, int post, float* out) {
  for (int i = 0; i < rows; ++i)
    layer_norm_row(a + (size_t)i * C, r ? r + (size_t)i * C : nullptr, res ? res + (size_t)i * C : nullptr, gamma, beta, C, pre,
                   post, out + (size_t)i * C);
}

// qkv [B][T][3H] -> out [B][T][H]; scratch scores [heads][T][T]
void hc_attention(const float* qkv, const long long* lens, const float* rel_k, const float* rel_v, int B, int T, int H,
                  int heads, int window, float* scores, float* out) {
  const int dk = H / heads, ld = 3 * H;
  for (int b = 0; b < B; ++b) {
    const float* q = qkv + (size_t)b * T * ld;
    const int len = (int)lens[b];
    for (int h = 0; h < heads; ++h)
      for (int i = 0; i < len; ++i)
        for (int j = 0; j < len; ++j) scores[((size_t)h * T + i) * T + j] = attn_score(q, ld, H, dk, h, i, j, len, rel_k, window);
    for (int i = 0; i < T; ++i)
      for (int h = 0; h < heads; ++h)
        for (int d = 0; d < dk; ++d)
          out[((size_t)b * T + i) * H + h * dk + d] =
              i < len ? attn_out(scores + ((size_t)h * T + i) * T, q, ld, H, dk, h, i, d, len, rel_v, window) : 0.f;
  }
}

void hc_dwconv(const float* x, const long long* lens, const float* w, const float* bias, int B, int T, int C, int dil,
               float* out) {
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < T; ++t)
      for (int c = 0; c < C; ++c)
        out[((size_t)b * T + t) * C + c] =
            t < lens[b] ? dwconv_at(x + (size_t)b * T * C, w, bias, C, t, c, (int)lens[b], dil) : 0.f;
}

// w [K][Cin][N]
void hc_dense(const float* x, const long long* lens, const float* w, const float* bias, int B, int T, int Cin, int K, int N,
              int relu_in, float* out) {
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < T; ++t)
      for (int n = 0; n < N; ++n)
        out[((size_t)b * T + t) * N + n] =
            t < lens[b] ? dense_at(x + (size_t)b * T * Cin, w, bias, Cin, K, N, t, n, (int)lens[b], relu_in) : 0.f;
}

void hc_spline_inverse(const float* x, const float* p, int n, float scale, float bound, float* out) {
  for (int i = 0; i < n; ++i) out[i] = rq_spline_inverse(x[i], p + (size_t)i * NP, scale, bound);
}

void hc_convflow_tail(const float* h, const long long* lens, const float* pw, const float* pb, const float* x1, int B, int T,
                      int C, float bound, float* out) {
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < T; ++t)
      out[(size_t)b * T + t] =
          t < lens[b] ? convflow_tail(h + ((size_t)b * T + t) * C, pw, pb, C, x1[(size_t)b * T + t], bound) : 0.f;
}

void hc_durations(const float* ls, const float* ld, const long long* lens, float ratio, float length_scale, int B, int T,
                  float* logw, float* w_ceil, int* cum, long long* y_len) {
  for (int b = 0; b < B; ++b)
    y_len[b] = durations_row(ls + (size_t)b * T, ld + (size_t)b * T, ratio, length_scale, T, (int)lens[b], logw + (size_t)b * T,
                             w_ceil + (size_t)b * T, cum + (size_t)b * T);
}

void hc_frame_tokens(const int* cum, const long long* y_len, int B, int T, int Ty, int* tok) {
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < Ty; ++y) tok[(size_t)b * Ty + y] = y < y_len[b] ? frame_token(cum + (size_t)b * T, T, y) : -1;
}

}  // extern "C"
