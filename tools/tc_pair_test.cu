// Probe for the next step of tcconv_kernel<128> (DESIGN.md "what comes next" 1a): tcgen05.mma.cta_group::2.
// A CTA PAIR (cluster of 2) computes D[256 x N] = A[256 x K] * B[N x K]^T with ONE MMA stream issued by the leader:
//   * each CTA stages ITS 128 rows of A in its own shared memory (same offsets in both CTAs),
//   * each CTA stages HALF of B -- under hypothesis `bmode` 0: columns [rank*N/2, +N/2) as an (N/2) x K K-major tile,
//   * the accumulator of CTA r (rows r*128 .. r*128+127, all N columns) lands in CTA r's TMEM.
// That halves the weight bytes each SM ingests, which is what bounds the C >= 256 layers today.
// The kernel cannot hang: every wait is bounded and reports a code instead.  The host prints, per CTA, the error vs a
// CPU product AND -- if the hypothesis is wrong -- which B column each output column actually matches, so one 10-second
// run tells the layout.  NOT part of the library; nothing here is used by the product yet.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I openvoice_b200/csrc -o /tmp/tc_pair_test tools/tc_pair_test.cu
//   /tmp/tc_pair_test            (run on a B200; exits 0 when hypothesis 0 holds)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ovc_conv.cuh"
#include "ovc_tc.cuh"

using namespace ovc;

constexpr int MP = 256, N = 128, K = 32, HN = N / 2;

__device__ __forceinline__ uint32_t ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_addr(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void commit2_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   tc::smem_addr(bar)),
               "h"(mask)
               : "memory");
}
// bounded wait: returns false after ~0.2 s instead of hanging the GPU
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t phase) {
  for (int spin = 0; spin < (1 << 22); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(tc::smem_addr(bar)), "r"(phase)
        : "memory");
    if (ok) return true;
  }
  return false;
}

// A [256][K], B [N][K] row-major fp32 in global; D [256][N]; status[2]
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) pair_test(const float* A, const float* B, float* D, int* status,
                                                                          int bmode) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* a_s = reinterpret_cast<float*>(smem);          // [K/4][128][4]   this CTA's 128 rows
  float* b_s = a_s + 128 * K;                            // [K/4][HN][4]    this CTA's half of B
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_s + HN * K);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = ctarank();

  for (int e = tid; e < 128 * K; e += 128) {
    const int row = e / K, k = e % K;
    float hi, lo;
    tc::split_tf32(A[(size_t)(rank * 128 + row) * K + k], hi, lo);        // tf32-exact operands: the product is exact in fp32
    a_s[((k / 4) * 128 + row) * 4 + (k % 4)] = hi;
  }
  for (int e = tid; e < HN * K; e += 128) {
    const int col = e / K, k = e % K;
    // bmode 0: CTA r holds B columns [r*HN, +HN).  bmode 1: interleaved (column 2*col + r) -- the alternative guess.
    const int src = bmode == 0 ? (int)rank * HN + col : 2 * col + (int)rank;
    float hi, lo;
    tc::split_tf32(B[(size_t)src * K + k], hi, lo);
    b_s[((k / 4) * HN + col) * 4 + (k % 4)] = hi;
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc2(tmem_slot, 128);            // both CTAs of the pair call it (cta_group::2)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc::fence_before();
  __syncthreads();
  cluster_sync();                                         // both CTAs' operands staged, both barriers initialised
  tc::fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (rank == 0 && tid == 0) {                            // the leader issues for the pair
    const uint32_t idesc = tc::make_idesc_tf32(MP, N);
    const uint32_t lbo_a = 128 * 16, lbo_b = HN * 16, sbo = 128;
    for (int k8 = 0; k8 < K / 8; ++k8) {
      const uint64_t ad = tc::make_desc(tc::smem_addr(a_s) + (2 * k8) * lbo_a, lbo_a, sbo);
      const uint64_t bd = tc::make_desc(tc::smem_addr(b_s) + (2 * k8) * lbo_b, lbo_b, sbo);
      mma2_tf32(tmem_d, ad, bd, idesc, k8 > 0);
    }
    commit2_mcast(bar, 0x3);                              // arrives on `bar` of BOTH CTAs
  }
  const bool ok = mbar_wait_bounded(bar, 0);
  if (tid == 0) status[rank] = ok ? 0 : 1;
  tc::fence_after();
  if (ok) {
    for (int c0 = 0; c0 < N; c0 += 8) {
      float v[8];
      tc::tmem_ld8(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int i = 0; i < 8; ++i) D[(size_t)(rank * 128 + tid) * N + c0 + i] = v[i];
    }
  }
  tc::fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) tmem_dealloc2(tmem_d, 128);
}

int main() {
  std::vector<float> A(MP * K), B(N * K);
  srand(3);
  auto tf32 = [](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; };
  for (auto& v : A) v = tf32((rand() / (float)RAND_MAX) * 2.f - 1.f);
  for (auto& v : B) v = tf32((rand() / (float)RAND_MAX) * 2.f - 1.f);
  std::vector<double> R((size_t)MP * N);
  for (int i = 0; i < MP; ++i)
    for (int j = 0; j < N; ++j) {
      double r = 0;
      for (int k = 0; k < K; ++k) r += (double)A[(size_t)i * K + k] * B[(size_t)j * K + k];
      R[(size_t)i * N + j] = r;
    }
  float *dA, *dB, *dD;
  int* dS;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, (size_t)MP * N * 4); cudaMalloc(&dS, 8);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (128 * K + HN * K) * 4 + 64;
  cudaFuncSetAttribute(pair_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int rc = 1;
  for (int bmode = 0; bmode < 2; ++bmode) {
    cudaMemset(dD, 0xFF, (size_t)MP * N * 4);
    int st[2] = {-1, -1};
    cudaMemcpy(dS, st, 8, cudaMemcpyHostToDevice);
    pair_test<<<2, 128, smem>>>(dA, dB, dD, dS, bmode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("bmode %d: CUDA error: %s\n", bmode, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(st, dS, 8, cudaMemcpyDeviceToHost);
    std::vector<float> Dh((size_t)MP * N);
    cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost);
    printf("bmode %d: wait status cta0 %d cta1 %d (0 = MMA completion arrived)\n", bmode, st[0], st[1]);
    if (st[0] || st[1]) continue;
    for (int r = 0; r < 2; ++r) {
      double maxerr = 0;
      for (int i = 0; i < 128; ++i)
        for (int j = 0; j < N; ++j) maxerr = fmax(maxerr, fabs(R[(size_t)(r * 128 + i) * N + j] - Dh[(size_t)(r * 128 + i) * N + j]));
      printf("  cta %d rows: max |D - A.B^T| = %.3e %s\n", r, maxerr, maxerr < 1e-4 ? "OK" : "MISMATCH");
      if (maxerr >= 1e-4) {   // which reference (row, column) does output (row r*128, column j) equal?
        for (int j = 0; j < N; j += 9) {
          int best = -1, brow = -1;
          for (int i2 = 0; i2 < MP && best < 0; ++i2)
            for (int j2 = 0; j2 < N; ++j2)
              if (fabs(R[(size_t)i2 * N + j2] - Dh[(size_t)(r * 128) * N + j]) < 1e-5) { best = j2; brow = i2; break; }
          printf("    D[%d][%d] = ref[%d][%d]\n", r * 128, j, brow, best);
        }
      } else if (bmode == 0 && r == 1) rc = 0;
    }
    if (rc == 0) { printf("hypothesis bmode %d holds: CTA r stages B columns [r*N/2, +N/2)\n", bmode); break; }
  }
  return rc;
}
