// tcconv_pair_kernel -- the wide (TN = 128) tensor-core convolution on CTA PAIRS (tcgen05 cta_group::2).
//
// STATUS: compiled only with -DOVC_TC_PAIR=1 (make EXTRA=-DOVC_TC_PAIR=1); NOT in the default build and not yet
// run on hardware.  The building block is proven (tools/tc_pair_test.cu on a B200: a pair MMA with each CTA staging its
// own 128 rows of A and B columns [rank*N/2, +N/2) gives the exact product in both CTAs' TMEM); this file is that block
// inside the tcconv pipeline, for the next round to validate (tests/test_gpu_parity.py run unchanged against it).
//
// Why: the ablation of tcconv_kernel<128> charges 29 of 62 ms per call to weight ingest: every
// SM streams the layer's full [Cin*K][128] weight tile through its shared memory once per 256 output steps.  A CTA pair
// (two SMs of one TPC, adjacent time tiles of the same utterance) issues ONE MMA stream of M = 256: each CTA keeps
// producing its own activation rows, but stages only HALF of the weight columns, so the bytes each SM ingests -- and
// the B-operand shared-memory reads per MMA -- halve, and the freed 80 KB double the ring depth in steps (40 slots).
//
// Protocol (per CTA unless noted; barriers live at the same offsets in both CTAs):
//   warp 0      TMA: its half of every weight slot  -> b_full[slot] (local)
//   warps 2-5   A producers (as in tcconv_kernel)   -> a_full[buf]  (local), then the epilogue on their own TMEM rows
//   leader warps 1, 6  MMA issuers: wait a_full / b_full AND the peer's copies (pa_full / pb_full, see below), issue
//               tcgen05.mma.cta_group::2, commit with .multicast to b_empty / a_empty / acc_full of BOTH CTAs
//   peer warps 1, 6    forwarders: mbarrier.try_wait exists for shared::cta only, so the peer watches its own
//               b_full / a_full and arrives remotely (mapa + mbarrier.arrive.shared::cluster) on the leader's pb_full /
//               pa_full; the slot cannot be refilled before the leader has consumed it (refill needs the leader's commit)
#pragma once
#include "ovc_tcconv.cuh"

namespace ovc {

struct TcPairCfg {
  static constexpr int TN = 128, HN = 64, MT = 2;
  static constexpr int ROWS = MT * 128 + 64;
  static constexpr int NABUF = 2;
  static constexpr int SLOTS = 40;                               // half-width slots: same bytes as 20 full ones
  static constexpr int A_BUF_FLOATS = 2 * 2 * ROWS * 4;          // [hi|lo][k chunk][row][4]
  static constexpr int B_SLOT_FLOATS = 2 * 2 * HN * 4;           // [hi|lo][k chunk][n in this CTA's half][4]
  static constexpr int RAWD = 1;
  static constexpr int RAW_FLOATS = ROWS * 8;
  static constexpr int BAR_BYTES = 1536;
  static constexpr size_t SMEM_BYTES =
      BAR_BYTES + sizeof(float) * (NABUF * A_BUF_FLOATS + SLOTS * B_SLOT_FLOATS + (RAWD + 1) * RAW_FLOATS);
  static constexpr uint32_t TMEM_COLS = 2 * MT * TN;             // main + low-order accumulators: all 512 columns
};

namespace tc {
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {   // one full warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(smem_dst)), "r"(ncols)
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
__device__ __forceinline__ void mma2_commit(uint64_t* bar) {   // arrives on `bar` of both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_addr(bar)),
               "h"((uint16_t)0x3)
               : "memory");
}
}  // namespace tc

__device__ __forceinline__ void mbar_arrive_remote(uint64_t* local_bar, uint32_t target_rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_bar)), "r"(target_rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// Every wait of the pair kernel is bounded: a protocol error traps (CUDA error, context lost) instead of hanging the
// GPU box.  2^24 try_waits is seconds -- far beyond any legitimate wait of this kernel (a whole launch is ~1 ms).
#ifndef OVC_TC_PAIR_SPIN_LIMIT
#define OVC_TC_PAIR_SPIN_LIMIT (1 << 24)
#endif
__device__ __forceinline__ void pair_wait(uint64_t* bar, uint32_t parity) {           // barrier arrived on locally / by commits
  for (int spin = 0; spin < OVC_TC_PAIR_SPIN_LIMIT; ++spin)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void pair_wait_cluster(uint64_t* bar, uint32_t parity) {   // barrier a peer CTA arrives on
  for (int spin = 0; spin < OVC_TC_PAIR_SPIN_LIMIT; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}

// weights: [n tile][Cin/8][K][rank 2][hi|lo][k chunk 2][64][4]  (pack_tc under OVC_TC_PAIR)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) tcconv_pair_kernel(const TcConvArgs a) {
  using Cfg = TcPairCfg;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NABUF = Cfg::NABUF, SLOTS = Cfg::SLOTS, ROWS = Cfg::ROWS, MT = Cfg::MT, TN = Cfg::TN, HN = Cfg::HN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t *a_full = bars, *a_empty = a_full + NABUF, *pa_full = a_empty + NABUF, *b_full = pa_full + NABUF,
           *b_empty = b_full + SLOTS, *pb_full = b_empty + SLOTS, *acc_full = pb_full + SLOTS;
  static_assert((3 * NABUF + 3 * SLOTS + 1) * 8 + 8 <= Cfg::BAR_BYTES, "barrier area");
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* abuf = reinterpret_cast<float*>(smem_raw + Cfg::BAR_BYTES);
  float* bring = abuf + NABUF * Cfg::A_BUF_FLOATS;
  float* raw = bring + SLOTS * Cfg::B_SLOT_FLOATS;

  const int b = blockIdx.z;
  const uint32_t rank = cluster_ctarank();                 // 0 = leader (earlier time tile)
  const int t0 = blockIdx.x * (MT * 128);
  const int n0 = blockIdx.y * TN;
  const int lim = (a.lens ? (int)min((long long)a.tmax, a.lens[b]) : a.tmax) * a.mul;
  if (t0 - (int)rank * (MT * 128) >= lim) return;          // the whole pair lies past the utterance (pair-uniform)
  const bool active = t0 < lim;                             // a padding peer still stages weights and keeps the protocol

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = (a.K - 1) / 2 * a.DIL;
  const int rows = MT * 128 + 2 * H;
  const int nk8 = a.Cin / 8;
  const int n_slots = nk8 * a.K;

  if (tid == 0) {
    for (int i = 0; i < NABUF; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], TC_NISS); mbar_init(&pa_full[i], 1); }
    for (int i = 0; i < SLOTS; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], TC_NISS); mbar_init(&pb_full[i], 1); }
    mbar_init(acc_full, TC_NISS);
    fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc2(tmem_slot, Cfg::TMEM_COLS);
  tc::fence_before();
  __syncthreads();
  cluster_sync_all();                                       // both CTAs' barriers exist before any remote arrive / multicast commit
  tc::fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ weight producer: this CTA's half of every slot
    if (lane == 0) {
      const float* wp = a.w + (size_t)blockIdx.y * n_slots * (2 * Cfg::B_SLOT_FLOATS) + rank * Cfg::B_SLOT_FLOATS;
      const uint32_t BYTES = (a.passes == 3 ? Cfg::B_SLOT_FLOATS : Cfg::B_SLOT_FLOATS / 2) * sizeof(float);
      int slot = 0;
      uint32_t phase = 1;
      for (int it = 0; it < n_slots; ++it) {
        pair_wait(&b_empty[slot], phase);
        mbar_expect_tx(&b_full[slot], BYTES);
        tma_bulk_g2s(bring + slot * Cfg::B_SLOT_FLOATS, wp, BYTES, &b_full[slot]);
        wp += 2 * Cfg::B_SLOT_FLOATS;
        if (++slot == SLOTS) { slot = 0; phase ^= 1; }
      }
    }
  } else if ((warp == 1 || warp == 6) && rank == 0) {
    // ------------------------------------------------------------ leader: MMA issuers for the pair
    if (lane == 0) {
      const int mt_lo = (warp == 1 ? 0 : MT / TC_NISS), mt_hi = mt_lo + MT / TC_NISS;
      const uint32_t idesc = tc::make_idesc_tf32(256, TN);
      constexpr uint32_t LBO_A = ROWS * 16, LBO_B = HN * 16, SBO = 128;
      constexpr uint32_t A_LO16 = (2 * ROWS * 16) >> 4;
      constexpr uint32_t B_LO16 = (2 * HN * 16) >> 4;
      constexpr uint32_t SLOT16 = (Cfg::B_SLOT_FLOATS * 4) >> 4;
      const uint64_t a_proto = tc::make_desc(0, LBO_A, SBO), b_proto = tc::make_desc(0, LBO_B, SBO);
      const uint64_t b_ring = b_proto + (tc::smem_addr(bring) >> 4);
      const uint32_t dil = (uint32_t)a.DIL;
      const bool three = a.passes == 3;
      int slot = 0;
      uint32_t bphase = 0;
      bool first = true;
      for (int q = 0; q < nk8; ++q) {
        const int buf = q & (NABUF - 1);
        const uint32_t aph = (q / NABUF) & 1;
        pair_wait(&a_full[buf], aph);
        pair_wait_cluster(&pa_full[buf], aph);
        tc::fence_after();
        uint64_t a_cur = a_proto + (tc::smem_addr(abuf + buf * Cfg::A_BUF_FLOATS) >> 4);
        for (int tap = 0; tap < a.K; ++tap) {
          pair_wait(&b_full[slot], bphase);
          pair_wait_cluster(&pb_full[slot], bphase);
          tc::fence_after();
          const uint64_t bd_hi = b_ring + (uint32_t)slot * SLOT16, bd_lo = bd_hi + B_LO16;
#pragma unroll
          for (int mt = mt_lo; mt < mt_hi; ++mt) {
            const uint64_t ad_hi = a_cur + mt * 128, ad_lo = ad_hi + A_LO16;
            const uint32_t d = tmem_d + mt * TN, dl = tmem_d + (MT + mt) * TN;   // low-order terms: own accumulator
            tc::mma2_tf32(d, ad_hi, bd_hi, idesc, !first);
            if (three) {
              tc::mma2_tf32(dl, ad_lo, bd_hi, idesc, !first);
              tc::mma2_tf32(dl, ad_hi, bd_lo, idesc, true);
            }
          }
          first = false;
          tc::mma2_commit(&b_empty[slot]);      // both CTAs' halves of the slot are reusable once these MMAs have read them
          a_cur += dil;
          if (++slot == SLOTS) { slot = 0; bphase ^= 1; }
        }
        tc::mma2_commit(&a_empty[buf]);
      }
      tc::mma2_commit(acc_full);
    }
  } else if (warp == 1 && rank == 1) {
    // ------------------------------------------------------------ peer: forward "my weight half has landed"
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_slots; ++it) {
        pair_wait(&b_full[slot], phase);
        mbar_arrive_remote(&pb_full[slot], 0);
        if (++slot == SLOTS) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 6 && rank == 1) {
    // ------------------------------------------------------------ peer: forward "my activation rows are staged"
    if (lane == 0) {
      for (int q = 0; q < nk8; ++q) {
        const int buf = q & (NABUF - 1);
        pair_wait(&a_full[buf], (q / NABUF) & 1);
        mbar_arrive_remote(&pa_full[buf], 0);
      }
    }
  } else if (warp >= 2 && warp <= 5) {
    // ------------------------------------------------------------ A producers (as tcconv_kernel<128>), then epilogue
    const int pt = tid - 64;
    if (!active) {
      // padding peer: no rows to stage (its accumulator rows are never read), but the leader waits on these barriers
      for (int q = 0; q < nk8; ++q) {
        const int buf = q % NABUF;
        pair_wait(&a_empty[buf], ((q / NABUF) & 1) ^ 1);
        mbar_arrive(&a_full[buf]);
      }
    } else {
      const float* xb = a.x + (size_t)b * a.x_bs;
      const int items = rows * 2;
      constexpr int NST = Cfg::RAWD + 1;
      auto stage = [&](int q) {
        float* dst = raw + (q % NST) * Cfg::RAW_FLOATS;
        for (int i = pt; i < items; i += 128) {
          const int row = i >> 1, kc = i & 1;
          const int t = t0 - H + row;
          const bool ok = (t >= 0 && t < lim);
          const float* src = ok ? xb + (size_t)t * a.Cin + q * 8 + kc * 4 : xb;
          cp_async16_zfill(dst + i * 4, src, ok ? 16 : 0);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      for (int q = 0; q < Cfg::RAWD; ++q) {
        if (q < nk8) stage(q);
        else asm volatile("cp.async.commit_group;" ::: "memory");
      }
      for (int q = 0; q < nk8; ++q) {
        if (q + Cfg::RAWD < nk8) stage(q + Cfg::RAWD);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(Cfg::RAWD) : "memory");
        const int buf = q % NABUF;
        pair_wait(&a_empty[buf], ((q / NABUF) & 1) ^ 1);
        float* ah = abuf + buf * Cfg::A_BUF_FLOATS;
        float* al = ah + 2 * ROWS * 4;
        const float* src = raw + (q % NST) * Cfg::RAW_FLOATS;
        for (int i = pt; i < items; i += 128) {
          const int row = i >> 1, kc = i & 1;
          float4 v = *reinterpret_cast<const float4*>(src + i * 4);
          v.x = lrelu(v.x, a.slope); v.y = lrelu(v.y, a.slope); v.z = lrelu(v.z, a.slope); v.w = lrelu(v.w, a.slope);
          float4 hi, lo;
          tc::split_tf32(v.x, hi.x, lo.x); tc::split_tf32(v.y, hi.y, lo.y);
          tc::split_tf32(v.z, hi.z, lo.z); tc::split_tf32(v.w, hi.w, lo.w);
          *reinterpret_cast<float4*>(ah + (kc * ROWS + row) * 4) = hi;
          *reinterpret_cast<float4*>(al + (kc * ROWS + row) * 4) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&a_full[buf]);
      }
      pair_wait(acc_full, 0);
      tc::fence_after();
      tc_epilogue<TN, MT, true>(a, tmem_d, b, t0, n0, lim, warp, lane);
    }
  }
  tc::fence_before();
  __syncthreads();
  cluster_sync_all();                                       // the peer's shared memory / TMEM are still being read until here
  if (warp == 1) tc::tmem_dealloc2(tmem_d, Cfg::TMEM_COLS);
}

}  // namespace ovc
