// Micro-benchmark: fp32 FMA issue rate on sm_100a for 3-register FFMA vs packed FFMA2
// (fma.rn.f32x2) when no operand can come from the reuse cache.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma_bench tools/ffma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
  float a[16], b[16], c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = out[(threadIdx.x * 16 + i) & 1023]; b[i] = out[(threadIdx.x * 16 + i + 7) & 2047]; c[i] = i; }
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      // 64 FFMA, operands rotate so consecutive instructions share no register
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fmaf(a[(i + r) & 15], b[(i * 3 + r) & 15], c[i]);
    } else if (MODE == 2) {
      // peak issue rate: one operand is shared by the 16 independent chains (register reuse cache), 64 FFMA
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fmaf(c[i], a[r], b[r]);
    } else if (MODE == 3) {
      // packed peak: 64 FFMA2 whose multiplier pair is shared by the 8 chains of a round
      unsigned long long* a2 = reinterpret_cast<unsigned long long*>(a);
      unsigned long long* b2 = reinterpret_cast<unsigned long long*>(b);
      unsigned long long* c2 = reinterpret_cast<unsigned long long*>(c);
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(c2[i]) : "l"(a2[r]), "l"(b2[r]));
    } else {
      unsigned long long* a2 = reinterpret_cast<unsigned long long*>(a);
      unsigned long long* b2 = reinterpret_cast<unsigned long long*>(b);
      unsigned long long* c2 = reinterpret_cast<unsigned long long*>(c);
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c2[i]) : "l"(a2[(i + r) & 7]), "l"(b2[(i * 3 + r) & 7]));
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int fma_per_iter) {
  float* d;
  cudaMalloc(&d, 148 * 8 * 256 * sizeof(float));
  const int iters = 20000;
  k<MODE><<<148 * 8, 256>>>(d, 100, 1.f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(d, iters, 1.f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double flops = 2.0 * fma_per_iter * (double)iters * 148 * 8 * 256;
  printf("%s: %.3f ms, %.2f TFLOP/s\n", name, ms, flops / ms / 1e9);
  cudaFree(d);
}

int main() {
  run<0>("FFMA  (3 fresh regs)", 64);
  run<1>("FFMA2 (3 fresh pairs)", 128);
  run<2>("FFMA  peak (shared multiplier, 16 chains)", 64);
  run<3>("FFMA2 peak (shared multiplier pair, 8 chains)", 128);
  return 0;
}
