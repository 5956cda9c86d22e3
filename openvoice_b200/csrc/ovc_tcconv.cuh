// ovc_tcconv.cuh -- split-precision (3xTF32) tensor-core Conv1d for the generator's ResBlock convs.
//
//   Y[t, n] = (bias[n] + sum_{tap, c} W[n, c, tap] * lrelu(X[t + (tap - (K-1)/2) * DIL, c]) [+ R[t, n]] [+ Y_old[t, n]]) * scale
//
// Channels-last activations X[b][t][C] (time-major rows).  Every product is evaluated as
// a_hi*b_hi + a_lo*b_hi + a_hi*b_lo with tf32-exact high parts and fp32 remainders (error ~1e-6,
// i.e. fp32-grade; tools/tc_gemm_test.cu), accumulated in fp32 in TMEM by tcgen05.mma.kind::tf32.
//
// One CTA = MT MMA tiles of 128 time steps x TN output channels (two fp32 accumulators each in TMEM).
// K-major, no-swizzle operand tiles (see ovc_tc.cuh): a convolution tap is a 16-byte-per-row shift
// of the A descriptor's start address, so all taps (any dilation) read ONE staged halo tile.
// Warp roles (224 threads):
//   warp 0      : TMA bulk copies of pre-split, pre-laid-out weight slots [tap][hi|lo] into an 8-slot ring
//   warps 1, 6  : one tcgen05.mma-issuing thread each (half of the MMA tiles); tcgen05.commit releases ring
//                 slots / A buffers
//   warps 2..5  : A producers -- global (16 B, zero-filled past the utterance) -> lrelu -> hi/lo split
//                 -> shared (2 buffers, 8 input channels each); afterwards the epilogue warps:
//                 tcgen05.ld -> bias / residual / MRF accumulate / scale -> global
#pragma once
#include "ovc_conv.cuh"
#include "ovc_tc.cuh"

namespace ovc {

struct TcConvArgs {
  const float* x; long long x_bs;     // [B][Lpitch][Cin]
  const float* w;                      // packed [n_tiles][Cin/8][K][2 (hi|lo)][2 (k chunk)][TN][4]
  const float* bias; long long bias_bs; // [Ntot] (+ b * bias_bs: per-utterance speaker-conditioning bias of the WN gate)
  float* y; long long y_bs; int y_ld;  // [B][Lpitch][y_ld]
  const float* r;                      // residual, same geometry as y (nullable)
  float* s; long long s_bs;            // EPI 2: skip accumulator [B][Lpitch][y_ld]
  int epi;                             // 0 linear (bias, residual, accumulate, scale) | 1 WN gate | 2 WN res/skip
  int split; int first;                // EPI 2: columns < split update y (x += rs), the rest go to s[col - split] (= / +=)
  const long long* lens; int tmax; int mul;   // valid steps = min(tmax, lens[b]) * mul   (lens NULL -> tmax)
  int Cin; int Ntot; int K; int DIL;   // Ntot = output row width (C for a ResBlock conv, stride*Cout for a polyphase transposed conv)
  float slope; float scale; int accumulate;
  int dbg;      // ablation switches (timing experiments only): 1 skip A production, 2 skip epilogue, 4 skip MMAs
  int passes;   // 3: split precision (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo); 1: single-pass TF32 (what cuDNN does by default)
};

constexpr int TC_THREADS = 224;           // warps: 0 TMA, 1 + 6 MMA issuers, 2..5 A producers / epilogue
constexpr int TC_NISS = 2;                // MMA-issuing threads (each owns half of the CTA's MMA tiles)
#ifndef OVC_TC_CL
#define OVC_TC_CL 1
#endif
constexpr int TC_CL128 = OVC_TC_CL;      // CTAs per cluster of the wide variant (weight multicast across the cluster)

// The tensor core ADDS into the TMEM accumulator with truncation: over ~10^3 accumulation steps a single
// accumulator drifts by ~6e-5 of the output rms toward zero (tools/tc_acc_test.cu).  The two low-order passes
// (a_lo*b_hi, a_hi*b_lo; 2^-11 of the result) therefore get their OWN accumulator, so the main one sees a third
// of the steps and the small terms are summed at their own scale; the epilogue adds the two in fp32
// (measured 2.0e-5 vs 6.1e-5; a plain fp32 fmaf chain has 1.0e-5).  TMEM columns = 2 * MT * TN <= 512.
template <int TN>
struct TcCfg {
  // accumulation steps = 3 * Cin/8 * K: only the wide layers (C >= 128 -> TN = 128) are long enough to drift
  // Two ways to keep the low-order terms out of the long accumulation (wide variant only):
  //   LOACC    : their own TMEM accumulator (costs half the TMEM -> MT = 2, twice the weight re-streaming)
  //   TWOSWEEP : sweep 0 accumulates all a_lo*b_hi + a_hi*b_lo terms (tiny magnitudes, no drift), sweep 1 adds the
  //              a_hi*b_hi terms on top -- one accumulator, MT = 4, weights re-streamed 1.5x per 512 steps instead of
  //              2x per 256.  Same accuracy (3.1e-5), but measured SLOWER (wide kernels 116 vs 75 ms per call): the A
  //              tiles are produced twice and the short hi*hi sweep is issue / producer bound.  Kept for reference.
#ifndef OVC_TC_TWOSWEEP
#define OVC_TC_TWOSWEEP 0
#endif
  static constexpr bool TWOSWEEP = TN == 128 && OVC_TC_TWOSWEEP;
  static constexpr bool LOACC = TN == 128 && !TWOSWEEP;
  static constexpr int MT = LOACC ? 2 : 4;                       // MMA tiles (of 128 steps) per CTA
  static constexpr int ROWS = MT * 128 + 64;                     // staged rows per A buffer (tile + 2*25 halo, padded)
  // the producers are latency-bound (ncu: long_scoreboard ~70 %): the 1-CTA/SM wide variant gets a deep pipeline,
  // the narrow ones run 2 CTAs per SM and hide latency that way
  static constexpr int NABUF = 2;                                // A stages (8 input channels each)
  // weight streaming is latency-bound: throughput = ring bytes / L2 latency, so the wide variant (1 CTA per SM)
  // spends all the shared memory it can on the ring (ablation: A production is hidden, the ring is not)
  static constexpr int SLOTS = TN == 128 ? (TWOSWEEP ? 14 : 20) : 8;   // weight ring depth
  static constexpr int A_BUF_FLOATS = 2 * 2 * ROWS * 4;          // [hi|lo][k chunk][row][4]
  static constexpr int B_SLOT_FLOATS = 2 * 2 * TN * 4;           // [hi|lo][k chunk][n][4]
  // raw fp32 landing stages for the producers' cp.async prefetch (rows x 8 channels), RAWD chunks ahead
  static constexpr int RAWD = TN == 128 ? 1 : 0;
  static constexpr int RAW_FLOATS = ROWS * 8;
  static constexpr size_t SMEM_BYTES =
      512 + sizeof(float) * (NABUF * A_BUF_FLOATS + SLOTS * B_SLOT_FLOATS + (RAWD ? (RAWD + 1) * RAW_FLOATS : 0));
  static constexpr uint32_t TMEM_COLS = (LOACC ? 2 : 1) * MT * TN;   // 128 (TN 32) / 256 (TN 64) / 512 (TN 128)
};

// cluster helpers (2-CTA clusters share every weight slot: each CTA fetches half and multicasts it to both)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_mcast(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mma_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Epilogue of one CTA tile (MT x 128 steps x TN columns): TMEM -> registers -> fused ops -> global.
// Warp w may read TMEM lanes [32*(w%4), +32); tmem_d = main accumulators, low-order ones (LOACC) MT*TN columns after.
template <int TN, int MT, bool LOACC>
__device__ __forceinline__ void tc_epilogue(const TcConvArgs& a, uint32_t tmem_d, int b, int t0, int n0, int lim, int warp,
                                            int lane) {
    const int lane_base = (warp & 3) * 32;
    float* yb = a.y + (size_t)b * a.y_bs;
    const float* rb = a.r ? a.r + (size_t)b * a.y_bs : nullptr;
    float* sb = a.s ? a.s + (size_t)b * a.s_bs : nullptr;
    const float* bias = a.bias + (size_t)b * a.bias_bs;
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      const int t = t0 + mt * 128 + lane_base + lane;
      const bool ok = t < lim && !(a.dbg & 2);
      float* yp = yb + (size_t)t * a.y_ld + n0;
      const float* rp = rb ? rb + (size_t)t * a.y_ld + n0 : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < TN; c0 += 32) {
        // everything with latency is issued first: both TMEM reads and the residual / accumulate loads
        const bool two = LOACC && a.passes == 3;
        uint32_t rm[32], rl[32];
        tc::tmem_ld32_issue(tmem_d + ((uint32_t)lane_base << 16) + mt * TN + c0, rm);
        if (two) tc::tmem_ld32_issue(tmem_d + ((uint32_t)lane_base << 16) + (MT + mt) * TN + c0, rl);
        float4 rq[8], yq[8];
        if (a.epi != 0) {
          // ---- WaveNet epilogues (modules.py:185-210), channels-last
          tc::tmem_ld_wait(rm);
          if (two) tc::tmem_ld_wait(rl);
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i)
            v[i] = (two ? __uint_as_float(rm[i]) + __uint_as_float(rl[i]) : __uint_as_float(rm[i])) + bias[n0 + c0 + i];
          if (!ok) continue;
          const int col = n0 + c0;
          if (a.epi == 1) {
            // columns [0,16) of the group: tanh inputs of 16 channels, [16,32): their sigmoid partners
            float* op = yb + (size_t)t * a.y_ld + (col >> 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = tanhf(v[4 * q + e]) * sigmoidf_acc(v[16 + 4 * q + e]);
              *reinterpret_cast<float4*>(op + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
            }
          } else {
            const bool to_x = col < a.split;          // uniform per 32-column group (split is a multiple of 32)
            float* op = to_x ? yb + (size_t)t * a.y_ld + col : sb + (size_t)t * a.y_ld + (col - a.split);
            const bool add = to_x || !a.first;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 cur = make_float4(0.f, 0.f, 0.f, 0.f);
              if (add) cur = *reinterpret_cast<const float4*>(op + 4 * q);
              cur.x += v[4 * q]; cur.y += v[4 * q + 1]; cur.z += v[4 * q + 2]; cur.w += v[4 * q + 3];
              *reinterpret_cast<float4*>(op + 4 * q) = cur;
            }
          }
          continue;
        }
        if (ok && rp) {
#pragma unroll
          for (int q = 0; q < 8; ++q) rq[q] = *reinterpret_cast<const float4*>(rp + c0 + 4 * q);
        }
        if (ok && a.accumulate) {
#pragma unroll
          for (int q = 0; q < 8; ++q) yq[q] = *reinterpret_cast<const float4*>(yp + c0 + 4 * q);
        }
        tc::tmem_ld_wait(rm);
        if (two) tc::tmem_ld_wait(rl);
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = two ? __uint_as_float(rm[i]) + __uint_as_float(rl[i]) : __uint_as_float(rm[i]);
        if (ok) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bq = *reinterpret_cast<const float4*>(bias + n0 + c0 + 4 * q);
            v[4 * q] += bq.x; v[4 * q + 1] += bq.y; v[4 * q + 2] += bq.z; v[4 * q + 3] += bq.w;
          }
          if (rp) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              v[4 * q] += rq[q].x; v[4 * q + 1] += rq[q].y; v[4 * q + 2] += rq[q].z; v[4 * q + 3] += rq[q].w;
            }
          }
          if (a.accumulate) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              v[4 * q] = yq[q].x + v[4 * q]; v[4 * q + 1] = yq[q].y + v[4 * q + 1];
              v[4 * q + 2] = yq[q].z + v[4 * q + 2]; v[4 * q + 3] = yq[q].w + v[4 * q + 3];
            }
          }
          if (a.scale != 1.f) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= a.scale;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(yp + c0 + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
    }
}

template <int TN, int CL>
__global__ void __launch_bounds__(TC_THREADS, TN == 128 ? 1 : 2) tcconv_kernel(const TcConvArgs a) {
  using Cfg = TcCfg<TN>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NABUF = Cfg::NABUF, SLOTS = Cfg::SLOTS, ROWS = Cfg::ROWS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* a_full = bars, *a_empty = bars + NABUF, *b_full = bars + 2 * NABUF, *b_empty = b_full + SLOTS,
            *acc_full = b_empty + SLOTS;
  static_assert((2 * NABUF + 2 * SLOTS + 1) * 8 + 8 <= 512, "barrier area");
  static_assert((NABUF & (NABUF - 1)) == 0, "NABUF must be a power of two");
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* abuf = reinterpret_cast<float*>(smem_raw + 512);
  float* bring = abuf + NABUF * Cfg::A_BUF_FLOATS;

  const int b = blockIdx.z;
  constexpr int MT = Cfg::MT;
  const int t0 = blockIdx.x * (MT * 128);
  const int n0 = blockIdx.y * TN;
  const int lim = (a.lens ? (int)min((long long)a.tmax, a.lens[b]) : a.tmax) * a.mul;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
  if (t0 - (int)crank * (MT * 128) >= lim) return;   // the whole cluster lies past the utterance (cluster-uniform)
  const bool active = t0 < lim;                       // a padding CTA still takes part in the weight multicast
  constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = (a.K - 1) / 2 * a.DIL;
  const int rows = MT * 128 + 2 * H;
  const int nk8 = a.Cin / 8;
  const int n_sweeps = (Cfg::TWOSWEEP && a.passes == 3) ? 2 : 1;
  const int nq = n_sweeps * nk8;            // flattened (sweep, 8-channel chunk) sequence

  if (tid == 0) {
    for (int i = 0; i < NABUF; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], TC_NISS); }
    for (int i = 0; i < SLOTS; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], CL * TC_NISS); }
    mbar_init(acc_full, TC_NISS);
    fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc::fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // every CTA's barriers exist before any remote arrive / multicast
  tc::fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ weight producer (TMA bulk)
    if (lane == 0) {
      const float* wt = a.w + (size_t)blockIdx.y * nk8 * a.K * Cfg::B_SLOT_FLOATS;
      // a slot is [hi | lo]; the hi*hi sweep and single-pass TF32 need (and fetch) only the first half
      const int n_slots = nk8 * a.K;
      int slot = 0;
      uint32_t phase = 1;   // the first pass over the ring finds every slot free
      for (int sweep = 0; sweep < n_sweeps; ++sweep) {
        const bool full = a.passes == 3 && (!Cfg::TWOSWEEP || sweep == 0);
        const uint32_t BYTES = (full ? Cfg::B_SLOT_FLOATS : Cfg::B_SLOT_FLOATS / 2) * sizeof(float);
        const int PART = (int)(BYTES / sizeof(float)) / CL;
        const float* wp = wt;
        for (int it = 0; it < n_slots; ++it) {
          mbar_wait(&b_empty[slot], phase);
          mbar_expect_tx(&b_full[slot], BYTES);
          float* dst = bring + slot * Cfg::B_SLOT_FLOATS;
          if (CL == 1) tma_bulk_g2s(dst, wp, BYTES, &b_full[slot]);
          else tma_bulk_g2s_mcast(dst + crank * PART, wp + crank * PART, BYTES / CL, &b_full[slot], CMASK);
          wp += Cfg::B_SLOT_FLOATS;
          if (++slot == SLOTS) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 6) {
    // ------------------------------------------------------------ MMA issuers (one thread each)
    // The loop body is kept minimal on purpose: this single thread's instruction latency, not the tensor
    // pipe, bounded the first version (112 SASS instructions per tap incl. two integer divisions).
    // Descriptors differ only in their 14-bit start-address field, so they are advanced by plain adds.
    if (lane == 0) {
      const int mt_lo = (warp == 1 ? 0 : MT / TC_NISS), mt_hi = mt_lo + MT / TC_NISS;
      const uint32_t idesc = tc::make_idesc_tf32(128, TN);
      constexpr uint32_t LBO_A = ROWS * 16, LBO_B = TN * 16, SBO = 128;
      constexpr uint32_t A_LO16 = (2 * ROWS * 16) >> 4;          // hi -> lo inside an A buffer, in 16-byte units
      constexpr uint32_t B_LO16 = (2 * TN * 16) >> 4;
      constexpr uint32_t SLOT16 = (Cfg::B_SLOT_FLOATS * 4) >> 4;
      const uint64_t a_proto = tc::make_desc(0, LBO_A, SBO), b_proto = tc::make_desc(0, LBO_B, SBO);
      const uint64_t b_ring = b_proto + (tc::smem_addr(bring) >> 4);
      const uint32_t dil = (uint32_t)a.DIL;
      const bool three = a.passes == 3;
      int slot = 0;
      uint32_t bphase = 0;
      bool first = true;
      for (int q = 0; q < nq; ++q) {
        const int sweep = q >= nk8 ? 1 : 0;
        const int buf = q & (NABUF - 1);
        if (active) mbar_wait(&a_full[buf], (q / NABUF) & 1);
        tc::fence_after();
        uint64_t a_cur = a_proto + (tc::smem_addr(abuf + buf * Cfg::A_BUF_FLOATS) >> 4);
        for (int tap = 0; tap < a.K; ++tap) {
          mbar_wait(&b_full[slot], bphase);
          tc::fence_after();
          const uint64_t bd_hi = b_ring + (uint32_t)slot * SLOT16, bd_lo = bd_hi + B_LO16;
          if (active && !(a.dbg & 4)) {
#pragma unroll
            for (int mt = mt_lo; mt < mt_hi; ++mt) {
              // output step (t0 + mt*128 + i) reads staged row (mt*128 + i + tap*DIL): the halo tile starts at t0 - H
              const uint64_t ad_hi = a_cur + mt * 128, ad_lo = ad_hi + A_LO16;
              const uint32_t d = tmem_d + mt * TN;
              const uint32_t dl = Cfg::LOACC ? tmem_d + (MT + mt) * TN : d;   // low-order terms: own accumulator
              if (Cfg::TWOSWEEP && three) {
                if (sweep == 0) {              // low-order terms first, at their own scale
                  tc::mma_tf32(d, ad_lo, bd_hi, idesc, !first);
                  tc::mma_tf32(d, ad_hi, bd_lo, idesc, true);
                } else {
                  tc::mma_tf32(d, ad_hi, bd_hi, idesc, true);
                }
              } else {
                tc::mma_tf32(d, ad_hi, bd_hi, idesc, !first);
                if (three) {
                  tc::mma_tf32(dl, ad_lo, bd_hi, idesc, Cfg::LOACC ? !first : true);
                  tc::mma_tf32(dl, ad_hi, bd_lo, idesc, true);
                }
              }
            }
          }
          first = false;
          if (CL == 1) tc::mma_commit(&b_empty[slot]);       // slot reusable once these MMAs have read it
          else mma_commit_mcast(&b_empty[slot], CMASK);      // ... in every CTA of the cluster
          a_cur += dil;
          if (++slot == SLOTS) { slot = 0; bphase ^= 1; }
        }
        tc::mma_commit(&a_empty[buf]);
      }
      tc::mma_commit(acc_full);
    }
  } else if (active && warp >= 2 && warp <= 5) {
    // ------------------------------------------------------------ A producers, then epilogue
    const int pt = tid - 64;                                   // 0..127
    const float* xb = a.x + (size_t)b * a.x_bs;
    const int items = rows * 2;                                // (row, 16-byte half of the 8 channels)
    if constexpr (Cfg::RAWD > 0) {
      // deep prefetch: raw fp32 rows land by cp.async (zero-filled outside [0, lim)) RAWD chunks ahead; the
      // conversion (lrelu, hi/lo split, operand layout) then runs shared -> shared on data this thread staged
      constexpr int NST = Cfg::RAWD + 1;
      float* raw = bring + SLOTS * Cfg::B_SLOT_FLOATS;
      auto stage = [&](int q) {
        const int k8 = q >= nk8 ? q - nk8 : q;
        float* dst = raw + (q % NST) * Cfg::RAW_FLOATS;
        for (int i = pt; i < ((a.dbg & 1) ? 0 : items); i += 128) {
          const int row = i >> 1, kc = i & 1;
          const int t = t0 - H + row;
          const bool ok = (t >= 0 && t < lim);
          const float* src = ok ? xb + (size_t)t * a.Cin + k8 * 8 + kc * 4 : xb;
          cp_async16_zfill(dst + i * 4, src, ok ? 16 : 0);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      for (int q = 0; q < Cfg::RAWD; ++q) {
        if (q < nq) stage(q);
        else asm volatile("cp.async.commit_group;" ::: "memory");
      }
      for (int q = 0; q < nq; ++q) {
        if (q + Cfg::RAWD < nq) stage(q + Cfg::RAWD);
        else asm volatile("cp.async.commit_group;" ::: "memory");   // keep the group count uniform
        asm volatile("cp.async.wait_group %0;" ::"n"(Cfg::RAWD) : "memory");
        const int buf = q % NABUF;
        mbar_wait(&a_empty[buf], ((q / NABUF) & 1) ^ 1);
        float* ah = abuf + buf * Cfg::A_BUF_FLOATS;
        float* al = ah + 2 * ROWS * 4;
        const float* src = raw + (q % NST) * Cfg::RAW_FLOATS;
        for (int i = pt; i < ((a.dbg & 1) ? 0 : items); i += 128) {
          const int row = i >> 1, kc = i & 1;
          float4 q = *reinterpret_cast<const float4*>(src + i * 4);
          q.x = lrelu(q.x, a.slope); q.y = lrelu(q.y, a.slope); q.z = lrelu(q.z, a.slope); q.w = lrelu(q.w, a.slope);
          float4 hi, lo;
          tc::split_tf32(q.x, hi.x, lo.x); tc::split_tf32(q.y, hi.y, lo.y);
          tc::split_tf32(q.z, hi.z, lo.z); tc::split_tf32(q.w, hi.w, lo.w);
          *reinterpret_cast<float4*>(ah + (kc * ROWS + row) * 4) = hi;
          *reinterpret_cast<float4*>(al + (kc * ROWS + row) * 4) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&a_full[buf]);
      }
    } else
    for (int q = 0; q < nq; ++q) {
      const int k8 = q >= nk8 ? q - nk8 : q;
      const int buf = q % NABUF;
      mbar_wait(&a_empty[buf], ((q / NABUF) & 1) ^ 1);
      float* ah = abuf + buf * Cfg::A_BUF_FLOATS;
      float* al = ah + 2 * ROWS * 4;
      // all global loads of a batch are issued before any is consumed (the loop is latency-, not bandwidth-bound)
      constexpr int PB = 5;
      for (int i0 = pt; i0 < items; i0 += 128 * PB) {
        float4 v[PB];
#pragma unroll
        for (int u = 0; u < PB; ++u) {
          const int i = i0 + 128 * u;
          const int row = i >> 1, kc = i & 1;
          const int t = t0 - H + row;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < items && t >= 0 && t < lim) v[u] = *reinterpret_cast<const float4*>(xb + (size_t)t * a.Cin + k8 * 8 + kc * 4);
        }
#pragma unroll
        for (int u = 0; u < PB; ++u) {
          const int i = i0 + 128 * u;
          if (i >= items) break;
          const int row = i >> 1, kc = i & 1;
          float4 q = v[u];
          q.x = lrelu(q.x, a.slope); q.y = lrelu(q.y, a.slope); q.z = lrelu(q.z, a.slope); q.w = lrelu(q.w, a.slope);
          float4 hi, lo;
          tc::split_tf32(q.x, hi.x, lo.x); tc::split_tf32(q.y, hi.y, lo.y);
          tc::split_tf32(q.z, hi.z, lo.z); tc::split_tf32(q.w, hi.w, lo.w);
          *reinterpret_cast<float4*>(ah + (kc * ROWS + row) * 4) = hi;
          *reinterpret_cast<float4*>(al + (kc * ROWS + row) * 4) = lo;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tensor-core (async) proxy
      mbar_arrive(&a_full[buf]);
    }
    // epilogue: warp w may read TMEM lanes [32*(w%4), +32)
    mbar_wait(acc_full, 0);
    tc::fence_after();
    tc_epilogue<TN, MT, Cfg::LOACC>(a, tmem_d, b, t0, n0, lim, warp, lane);
  }
  tc::fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // peers may still multicast into / arrive on this CTA's shared memory
  if (warp == 1) tc::tmem_dealloc(tmem_d, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------
// tcconv_narrow_kernel: persistent variant for the narrow layers (TN = 32 / 64 output columns per CTA).
// Per 512-step tile these layers have only 1-5 us of MMA work, so the one-tile-per-CTA kernel above is dominated
// by fixed costs (barrier init, TMEM allocation, first-load latency, a serial epilogue).  Here one CTA per SM
// loops over tiles: barriers / TMEM live for the whole launch, the layer's weights stay resident in shared
// memory when they fit (C = 32: <= 88 KB; otherwise the ring streams them per tile), A staging runs ahead across
// tile boundaries, and the epilogue of tile i (own 4 warps, second TMEM accumulator set) overlaps the MMAs of
// tile i+1.  Warps: 0 TMA, 1-2 MMA issuers, 3-6 A producers, 7-10 epilogue.
// ---------------------------------------------------------------------------------------------------------
constexpr int TCN_THREADS = 352;

template <int TN>
struct TcnCfg {
  static constexpr int MT = 4;
  static constexpr int ROWS = MT * 128 + 64;
  static constexpr int NABUF = 2;
  static constexpr int RING = TN == 32 ? 44 : 16;                 // weight slots: 88 KB (TN 32) / 64 KB (TN 64)
  static constexpr int A_BUF_FLOATS = 2 * 2 * ROWS * 4;
  static constexpr int B_SLOT_FLOATS = 2 * 2 * TN * 4;
  static constexpr size_t SMEM_BYTES = 1024 + sizeof(float) * (NABUF * A_BUF_FLOATS + RING * B_SLOT_FLOATS);
  static constexpr uint32_t TMEM_COLS = 2 * MT * TN;              // two accumulator sets: 256 / 512
};

template <int TN>
__global__ void __launch_bounds__(TCN_THREADS, 1) tcconv_narrow_kernel(const TcConvArgs a, int n_tt, int total) {
  using Cfg = TcnCfg<TN>;
  constexpr int MT = Cfg::MT, ROWS = Cfg::ROWS, NABUF = Cfg::NABUF, RING = Cfg::RING;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* a_full = bars, *a_empty = bars + NABUF, *b_full = bars + 2 * NABUF, *b_empty = b_full + RING,
            *acc_full = b_empty + RING, *acc_empty = acc_full + 2;
  static_assert((2 * NABUF + 2 * RING + 4) * 8 + 8 <= 1024, "barrier area");
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* abuf = reinterpret_cast<float*>(smem_raw + 1024);
  float* bring = abuf + NABUF * Cfg::A_BUF_FLOATS;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * TN;
  const int H = (a.K - 1) / 2 * a.DIL;
  const int rows = MT * 128 + 2 * H;
  const int nk8 = a.Cin / 8;
  const int n_slots = nk8 * a.K;
  const bool resident = n_slots <= RING;
  const uint32_t SLOT_BYTES = (a.passes == 3 ? Cfg::B_SLOT_FLOATS : Cfg::B_SLOT_FLOATS / 2) * sizeof(float);

  if (tid == 0) {
    for (int i = 0; i < NABUF; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 2); }
    for (int i = 0; i < RING; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 2); mbar_init(&acc_empty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem_d = *tmem_slot;

  // every role walks the same tile sequence
#define TCN_FOR_TILES                                                                                   \
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {                                        \
    const int b = tile / n_tt, t0 = (tile % n_tt) * (MT * 128);                                         \
    const int lim = (a.lens ? (int)min((long long)a.tmax, a.lens[b]) : a.tmax) * a.mul;                 \
    if (t0 >= lim) continue;

  if (warp == 0) {
    // ------------------------------------------------------------ weights
    if (lane == 0) {
      const float* wt = a.w + (size_t)blockIdx.y * n_slots * Cfg::B_SLOT_FLOATS;
      if (resident) {
        for (int it = 0; it < n_slots; ++it) {
          mbar_expect_tx(&b_full[it], SLOT_BYTES);
          tma_bulk_g2s(bring + it * Cfg::B_SLOT_FLOATS, wt + (size_t)it * Cfg::B_SLOT_FLOATS, SLOT_BYTES, &b_full[it]);
        }
      } else {
        int slot = 0;
        uint32_t phase = 1;
        TCN_FOR_TILES
          (void)b; (void)lim;
          for (int it = 0; it < n_slots; ++it) {
            mbar_wait(&b_empty[slot], phase);
            mbar_expect_tx(&b_full[slot], SLOT_BYTES);
            tma_bulk_g2s(bring + slot * Cfg::B_SLOT_FLOATS, wt + (size_t)it * Cfg::B_SLOT_FLOATS, SLOT_BYTES, &b_full[slot]);
            if (++slot == RING) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------ MMA issuers (each owns 2 of the 4 MMA tiles)
    if (lane == 0) {
      const int mt_lo = (warp - 1) * 2, mt_hi = mt_lo + 2;
      const uint32_t idesc = tc::make_idesc_tf32(128, TN);
      constexpr uint32_t LBO_A = ROWS * 16, LBO_B = TN * 16, SBO = 128;
      constexpr uint32_t A_LO16 = (2 * ROWS * 16) >> 4, B_LO16 = (2 * TN * 16) >> 4, SLOT16 = (Cfg::B_SLOT_FLOATS * 4) >> 4;
      const uint64_t a_proto = tc::make_desc(0, LBO_A, SBO), b_proto = tc::make_desc(0, LBO_B, SBO);
      const uint64_t b_ring = b_proto + (tc::smem_addr(bring) >> 4);
      const uint32_t dil = (uint32_t)a.DIL;
      const bool three = a.passes == 3;
      int slot = 0, ka = 0, n = 0;
      uint32_t bphase = 0;
      TCN_FOR_TILES
        (void)b; (void)lim;
        const int set = n & 1;
        mbar_wait(&acc_empty[set], ((n >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator set
        tc::fence_after();
        const uint32_t acc = tmem_d + set * (MT * TN);
        bool first = true;
        if (resident) slot = 0;
        for (int k8 = 0; k8 < nk8; ++k8, ++ka) {
          const int buf = ka & (NABUF - 1);
          mbar_wait(&a_full[buf], (ka / NABUF) & 1);
          tc::fence_after();
          uint64_t a_cur = a_proto + (tc::smem_addr(abuf + buf * Cfg::A_BUF_FLOATS) >> 4);
          for (int tap = 0; tap < a.K; ++tap) {
            if (!resident || n == 0) {
              mbar_wait(&b_full[slot], resident ? 0u : bphase);
              tc::fence_after();
            }
            const uint64_t bd_hi = b_ring + (uint32_t)slot * SLOT16, bd_lo = bd_hi + B_LO16;
#pragma unroll
            for (int mt = mt_lo; mt < mt_hi; ++mt) {
              if (a.dbg & 4) continue;
              const uint64_t ad_hi = a_cur + mt * 128, ad_lo = ad_hi + A_LO16;
              const uint32_t d = acc + mt * TN;
              tc::mma_tf32(d, ad_hi, bd_hi, idesc, !first);
              if (three) {
                tc::mma_tf32(d, ad_lo, bd_hi, idesc, true);
                tc::mma_tf32(d, ad_hi, bd_lo, idesc, true);
              }
            }
            first = false;
            if (!resident) tc::mma_commit(&b_empty[slot]);
            a_cur += dil;
            if (++slot == RING) { slot = 0; bphase ^= 1; }
          }
          tc::mma_commit(&a_empty[buf]);
        }
        tc::mma_commit(&acc_full[set]);
        ++n;
      }
    }
  } else if (warp >= 3 && warp <= 6) {
    // ------------------------------------------------------------ A producers (run ahead across tiles)
    const int pt = tid - 96;
    const int items = (a.dbg & 1) ? 0 : rows * 2;
    int ka = 0;
    TCN_FOR_TILES
      const float* xb = a.x + (size_t)b * a.x_bs;
      for (int k8 = 0; k8 < nk8; ++k8, ++ka) {
        const int buf = ka & (NABUF - 1);
        mbar_wait(&a_empty[buf], ((ka / NABUF) & 1) ^ 1);
        float* ah = abuf + buf * Cfg::A_BUF_FLOATS;
        float* al = ah + 2 * ROWS * 4;
        constexpr int PB = 9;   // all of a thread's loads for one chunk in flight at once (<= 1124 items / 128 threads)
        for (int i0 = pt; i0 < items; i0 += 128 * PB) {
          float4 v[PB];
#pragma unroll
          for (int u = 0; u < PB; ++u) {
            const int i = i0 + 128 * u;
            const int row = i >> 1, kc = i & 1;
            const int t = t0 - H + row;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < items && t >= 0 && t < lim) v[u] = *reinterpret_cast<const float4*>(xb + (size_t)t * a.Cin + k8 * 8 + kc * 4);
          }
#pragma unroll
          for (int u = 0; u < PB; ++u) {
            const int i = i0 + 128 * u;
            if (i >= items) break;
            const int row = i >> 1, kc = i & 1;
            float4 q = v[u];
            q.x = lrelu(q.x, a.slope); q.y = lrelu(q.y, a.slope); q.z = lrelu(q.z, a.slope); q.w = lrelu(q.w, a.slope);
            float4 hi, lo;
            tc::split_tf32(q.x, hi.x, lo.x); tc::split_tf32(q.y, hi.y, lo.y);
            tc::split_tf32(q.z, hi.z, lo.z); tc::split_tf32(q.w, hi.w, lo.w);
            *reinterpret_cast<float4*>(ah + (kc * ROWS + row) * 4) = hi;
            *reinterpret_cast<float4*>(al + (kc * ROWS + row) * 4) = lo;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&a_full[buf]);
      }
    }
  } else if (warp >= 7) {
    // ------------------------------------------------------------ epilogue (overlaps the next tile's MMAs)
    int n = 0;
    TCN_FOR_TILES
      const int set = n & 1;
      mbar_wait(&acc_full[set], (n >> 1) & 1);
      tc::fence_after();
      tc_epilogue<TN, MT, false>(a, tmem_d + set * (MT * TN), b, t0, n0, lim, warp, lane);
      tc::fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[set]);
      ++n;
    }
  }
#undef TCN_FOR_TILES
  tc::fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_d, Cfg::TMEM_COLS);
}

// conv_post on channels-last input: y[b, t] = tanh(sum_{k<7, ci<C} w[ci, k] * lrelu_0.01(x[b, t+k-3, ci]))
// (models.py:287-289).  HBM-bound (132 B per sample): a CTA stages 256+6 rows with coalesced 16-byte loads into a
// transposed, conflict-free shared tile (leaky_relu applied once), then one thread per output sample.
template <int C>
__global__ void __launch_bounds__(256) conv_post_cl_kernel(const float* __restrict__ x, long long x_bs,
                                                           const float* __restrict__ w, float* __restrict__ y,
                                                           long long y_bs, int y_len, const long long* lens, int tmax,
                                                           int mul) {
  constexpr int TB = 256, ROWS = TB + 6, LD = ROWS + 3;   // LD odd: conflict-free transposed stores
  __shared__ float ws[7][C];
  __shared__ float xs[C][LD];
  for (int i = threadIdx.x; i < C * 7; i += blockDim.x) ws[i % 7][i / 7] = w[i];
  const int b = blockIdx.y;
  const int lim = (lens ? (int)min((long long)tmax, lens[b]) : tmax) * mul;
  const int t0 = blockIdx.x * TB;
  const float* xb = x + (size_t)b * x_bs;
  for (int idx = threadIdx.x; idx < ROWS * (C / 4); idx += blockDim.x) {
    const int row = idx / (C / 4), q = idx % (C / 4);
    const int tt = t0 - 3 + row;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tt >= 0 && tt < lim) v = *reinterpret_cast<const float4*>(xb + (size_t)tt * C + 4 * q);
    xs[4 * q + 0][row] = v.x > 0.f ? v.x : 0.01f * v.x;
    xs[4 * q + 1][row] = v.y > 0.f ? v.y : 0.01f * v.y;
    xs[4 * q + 2][row] = v.z > 0.f ? v.z : 0.01f * v.z;
    xs[4 * q + 3][row] = v.w > 0.f ? v.w : 0.01f * v.w;
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= y_len) return;
  float acc = 0.f;
  if (t < lim) {
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll 8
      for (int ci = 0; ci < C; ++ci) acc = fmaf(ws[k][ci], xs[ci][threadIdx.x + k], acc);
    acc = tanhf(acc);
  }
  y[(size_t)b * y_bs + t] = acc;
}

// [B][C][pitch] <-> [B][pitch][C] tiled transpose (32x32 through shared memory)
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows,
                                                        int cols, long long bs) {
  __shared__ float tile[32][33];
  const float* s = src + (size_t)blockIdx.z * bs;
  float* d = dst + (size_t)blockIdx.z * bs;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? s[(size_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) d[(size_t)c * rows + r] = tile[tx][i];
  }
}

}  // namespace ovc
