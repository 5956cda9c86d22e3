// Validates the tcgen05 building blocks of ovc_tc.cuh on a B200:
//   D[128 x N] = A[shift .. shift+128) x B^T   (K-major, no swizzle, kind::tf32, fp32 accumulate in TMEM)
// (1) single-pass TF32, (2) 3xTF32 split precision, (3) row-shifted A (a convolution tap).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I openvoice_b200/csrc -o /tmp/tc_gemm_test tools/tc_gemm_test.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ovc_conv.cuh"
#include "ovc_tc.cuh"

using namespace ovc;

constexpr int M = 128, N = 64, K = 32, ROWS_A = 160;

__global__ void __launch_bounds__(128) gemm_test(const float* A, const float* B, float* D, int shift, int split) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* a_hi = reinterpret_cast<float*>(smem);                 // [K/4][ROWS_A][4]
  float* a_lo = a_hi + ROWS_A * K;
  float* b_hi = a_lo + ROWS_A * K;                              // [K/4][N][4]
  float* b_lo = b_hi + N * K;
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_lo + N * K);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int e = tid; e < ROWS_A * K; e += 128) {
    const int row = e / K, k = e % K;
    float hi, lo;
    tc::split_tf32(A[e], hi, lo);
    if (!split) { hi = A[e]; lo = 0.f; }
    a_hi[((k / 4) * ROWS_A + row) * 4 + (k % 4)] = hi;
    a_lo[((k / 4) * ROWS_A + row) * 4 + (k % 4)] = lo;
  }
  for (int e = tid; e < N * K; e += 128) {
    const int row = e / K, k = e % K;
    float hi, lo;
    tc::split_tf32(B[e], hi, lo);
    if (!split) { hi = B[e]; lo = 0.f; }
    b_hi[((k / 4) * N + row) * 4 + (k % 4)] = hi;
    b_lo[((k / 4) * N + row) * 4 + (k % 4)] = lo;
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 64);
  // generic-proxy smem writes must be visible to the tensor-core (async) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_tf32(M, N);
    const uint32_t lbo_a = ROWS_A * 16, lbo_b = N * 16, sbo = 128;
    bool acc = false;
    for (int pass = 0; pass < (split ? 3 : 1); ++pass) {
      const float* pa = (pass == 1) ? a_lo : a_hi;
      const float* pb = (pass == 2) ? b_lo : b_hi;
      for (int k8 = 0; k8 < K / 8; ++k8) {
        const uint64_t ad = tc::make_desc(tc::smem_addr(pa) + (2 * k8) * lbo_a + shift * 16, lbo_a, sbo);
        const uint64_t bd = tc::make_desc(tc::smem_addr(pb) + (2 * k8) * lbo_b, lbo_b, sbo);
        tc::mma_tf32(tmem_d, ad, bd, idesc, acc);
        acc = true;
      }
    }
    tc::mma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc::fence_after();
  for (int c0 = 0; c0 < N; c0 += 8) {
    float v[8];
    tc::tmem_ld8(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 8; ++i) D[(size_t)tid * N + c0 + i] = v[i];
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_d, 64);
}

static double run(int shift, int split, const std::vector<float>& A, const std::vector<float>& B) {
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, M * N * 4);
  const size_t smem = (2 * ROWS_A * K + 2 * N * K) * 4 + 64;
  cudaFuncSetAttribute(gemm_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  gemm_test<<<1, 128, smem>>>(dA, dB, dD, shift, split);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  std::vector<float> D(M * N);
  cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      double r = 0;
      for (int k = 0; k < K; ++k) r += (double)A[(size_t)(i + shift) * K + k] * (double)B[(size_t)j * K + k];
      maxerr = fmax(maxerr, fabs(r - D[(size_t)i * N + j]));
      maxref = fmax(maxref, fabs(r));
    }
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return maxerr / maxref;
}

int main() {
  std::vector<float> A(ROWS_A * K), B(N * K);
  srand(1);
  for (auto& v : A) v = (rand() / (float)RAND_MAX) * 2.f - 1.f;
  for (auto& v : B) v = (rand() / (float)RAND_MAX) * 2.f - 1.f;
  printf("tf32 single pass, shift 0 : rel err %.3e (expect ~1e-3)\n", run(0, 0, A, B));
  printf("3xTF32 split,     shift 0 : rel err %.3e (expect ~1e-6)\n", run(0, 1, A, B));
  printf("3xTF32 split,     shift 3 : rel err %.3e\n", run(3, 1, A, B));
  printf("3xTF32 split,     shift 29: rel err %.3e\n", run(29, 1, A, B));
  return 0;
}
