// tmem_ld_probe.cu -- prints which (TMEM lane, column) each register of tcgen05.ld.16x256b.x4 holds, against the layout
// the conv epilogue assumes (ovc_tc.cuh: tmem_ld16x256_x4_issue).  One CTA, 4 warps; TMEM is filled through
// tcgen05.st.32x32b (lane l, column c <- l * 1000 + c).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build_probe/tmem_ld_probe tools/tmem_ld_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../openvoice_b200/csrc/ovc_tc.cuh"

using namespace ovc;

__global__ void probe(float* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tc::tmem_alloc(&slot, 32);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t base = slot;
  // fill: thread of TMEM lane l writes columns 0..31
  {
    const int l = warp * 32 + lane;
    uint32_t v[32];
    for (int c = 0; c < 32; ++c) v[c] = __float_as_uint((float)(l * 1000 + c));
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(base + ((uint32_t)(warp * 32) << 16)),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
        "r"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  for (int h = 0; h < 2; ++h) {
    uint32_t r[16];
    tc::tmem_ld16x256_x4_issue(base + ((uint32_t)(warp * 32 + 16 * h) << 16), r);
    tc::tmem_ld_wait16(r);
    for (int i = 0; i < 16; ++i) out[((warp * 2 + h) * 32 + lane) * 16 + i] = __uint_as_float(r[i]);
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(base, 32);
}

int main() {
  float* d;
  cudaMalloc(&d, 4 * 2 * 32 * 16 * sizeof(float));
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  static float h[4 * 2 * 32 * 16];
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int w = 0; w < 4; ++w)
    for (int hh = 0; hh < 2; ++hh)
      for (int t = 0; t < 32; ++t)
        for (int i = 0; i < 16; ++i) {
          const int g = i / 4, e2 = i % 4;
          const int lane = w * 32 + 16 * hh + t / 4 + (e2 >= 2 ? 8 : 0);
          const int col = 8 * g + 2 * (t % 4) + (e2 & 1);
          const float want = (float)(lane * 1000 + col), got = h[((w * 2 + hh) * 32 + t) * 16 + i];
          if (want != got) {
            if (bad < 24) printf("warp %d half %d thread %2d reg %2d: got lane %d col %d, assumed lane %d col %d\n", w, hh, t, i,
                                 (int)got / 1000, (int)got % 1000, lane, col);
            ++bad;
          }
        }
  printf("tmem_ld_probe: %d mismatches against the assumed 16x256b.x4 fragment layout\n", bad);
  if (bad) {
    printf("warp 0, half 0 mapping (thread: reg -> lane.col):\n");
    for (int t = 0; t < 32; ++t) {
      printf("t%2d:", t);
      for (int i = 0; i < 16; ++i) printf(" %d.%d", (int)h[t * 16 + i] / 1000, (int)h[t * 16 + i] % 1000);
      printf("\n");
    }
  }
  return bad != 0;
}
