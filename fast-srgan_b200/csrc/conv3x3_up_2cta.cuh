// conv3x3_up_2cta.cuh - the CTA-pair form of the upsampling conv (default; GPU-validated bit-identical to the single-CTA kernel,
// tests/test_fused_chain_gpu.py).
//
// The 64 -> 256 upsampling conv (+ bias + PixelShuffle(2) + PReLU; reference model.py:30-40) as a CTA-PAIR kernel:
// tcgen05.mma.cta_group::2, M = 256 (two 128-pixel tiles, one per CTA), N = 256 (all output channels), K = 16.
//
// Why: the single-CTA kernel runs N = 128 per CTA - the A tile (4 KB) and the weight tile (4 KB)
// are both fetched from shared memory for every 64-cycle MMA = 128 B/clk, the whole smem port; measured 0.64-0.70 of the
// tensor peak, and every pixel tile is staged twice (once per 128-column half).  In a CTA pair each SM feeds the pair's
// MMA with its own A tile (4 KB) and HALF of the weights (128 of the 256 N rows, 4 KB) per 128-cycle MMA = 64 B/clk, the
// activations are staged once, and the weights still take 147 KB per CTA (9 taps x 128 rows x 128 B).
//
// Protocol (leader = cluster rank 0; barriers at identical smem offsets in both CTAs):
//   * both producers: wait LOCAL empty[s] -> remote arrive.expect_tx on the LEADER's full[s] (count 2) -> TMA box of
//     their own tile with .cta_group::2, completing on the leader's full[s];
//   * leader MMA warp: wait tempty[acc] (8 arrivals: 4 epilogue warps x 2 CTAs, CTA 1's are remote) and full[s], issue
//     the 36 MMAs of the tile pair, commit (multicast 0b11) to empty[s] and tfull[acc] of BOTH CTAs;
//   * each CTA's epilogue warps: wait LOCAL tfull[acc], drain their own 128 rows x 256 columns, arrive on the leader's
//     tempty[acc];
//   * cluster barrier before TMEM dealloc / exit (the leader's MMAs read CTA 1's shared memory).
#pragma once
#include "conv3x3_tc.cuh"

namespace fsr {

struct Up2Cfg {
  using Geo = ConvGeo<true>;
  static constexpr int kN = 256;                         // GEMM columns of the pair's MMA
  static constexpr int kWBytes = 9 * (kN / 2) * 128;     // this CTA's half of the weights: 147456
  static constexpr int kStages = 2;
  static constexpr int kEpiWarps = 8;                    // two per SM sub-partition (a lone warp cannot hide its own ALU latency)
  static constexpr int kThreads = 64 + 32 * kEpiWarps;   // 320
  static constexpr int kStagingBytes = kEpiWarps * 4096;
  static constexpr int kTmemCols = 512;                  // 2 accumulators x 256 columns
  static constexpr int kSmemBytes = kWBytes + kStages * Geo::kStageBytes + kStagingBytes + 2048 /*barriers + bias*/ + 1024 /*align*/;
  static_assert(kSmemBytes <= 232448, "smem");
};

FSR_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
FSR_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `rank` of this cluster
FSR_DEVINL uint32_t mapa_cluster(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
FSR_DEVINL void mbar_arrive_expect_tx_cluster(uint32_t cluster_bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_bar), "r"(bytes) : "memory");
}
FSR_DEVINL void mbar_arrive_cluster(uint32_t cluster_bar) {
  // default semantics like CUTLASS' ClusterBarrier::arrive(cta_id): the TMEM reads that must precede it are ordered by
  // tcgen05.wait::ld + tcgen05.fence::before_thread_sync, not by a memory fence (the .release.cluster form costs a
  // MEMBAR.ALL.GPU + ERRBAR per arrive: 13 % of all stall samples in profiles/r02/ncu_full_up_2cta_v1.md)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// TMA loads whose completion lands on a barrier that may live in the peer CTA (cluster address)
FSR_DEVINL void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
FSR_DEVINL void tma_load_4d_pair(void* dst, const CUtensorMap* m, uint32_t cluster_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
FSR_DEVINL void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs of the pair have completed) on the barrier at this offset in BOTH CTAs
FSR_DEVINL void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
template <uint32_t kCols>
FSR_DEVINL void tmem_alloc_pair(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
FSR_DEVINL void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

template <typename T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Up2Cfg::kThreads, 1)
conv3x3_up_2cta_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                       const __grid_constant__ CUtensorMap tm_out, const ConvParams p) {
  pdl_grid_sync();
  using Cfg = Up2Cfg;
  using Geo = Cfg::Geo;
  constexpr int TH = Geo::TH, TW = Geo::TW;
  extern __shared__ uint8_t smem_raw[];
  // identical layout in both CTAs (the pair's MMA applies ONE descriptor to both shared memories): the dynamic smem base
  // offset is the same for every CTA of a launch, so the same rounding gives the same offsets
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem_w + Cfg::kWBytes;
  uint8_t* smem_stg = smem_a + Cfg::kStages * Geo::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;                        // [kStages]  used in the leader only (count 2 + tx of both CTAs)
  uint64_t* empty_bar = bars + Cfg::kStages;        // [kStages]  local, signalled by the leader's multicast commit
  uint64_t* w_bar = bars + 2 * Cfg::kStages;        // [1]        leader only (count 2 + tx)
  uint64_t* tfull_bar = w_bar + 1;                  // [2]        local, multicast commit
  uint64_t* tempty_bar = tfull_bar + 2;             // [2]        leader only (count 2 * kEpiWarps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* smem_bias = reinterpret_cast<float*>(tmem_slot + 4);   // [256]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int num_pairs = (p.num_tiles + 1) >> 1;
  const int pr_begin = (int)(((long long)cluster_id * num_pairs) / num_clusters);
  const int pr_end = (int)(((long long)(cluster_id + 1) * num_pairs) / num_clusters);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_out);
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full_bar[i], 2); mbar_init(&empty_bar[i], 1); }
    mbar_init(w_bar, 2);
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 2 * Cfg::kEpiWarps); }   // 8 warps x 2 CTAs
    fence_mbar_init();
    fence_proxy_async();
  }
  for (int i = threadIdx.x; i < Cfg::kN; i += blockDim.x) smem_bias[i] = p.bias ? p.bias[i] : 0.f;
  __syncthreads();
  cluster_sync_all();                               // both CTAs' barriers are initialised before any remote arrive
  if (warp == 1) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_of = [&](int pair) {                    // this CTA's tile of the pair (clamped: an odd tail re-does the last tile)
    const int t = 2 * pair + (int)rank;
    return t < p.num_tiles ? t : p.num_tiles - 1;
  };

  if (warp == 0) {
    // =============================== TMA producer (both CTAs) ===============================
    const uint32_t w_bar_leader = mapa_cluster(smem_u32(w_bar), 0);
    if (elect_one()) {
      mbar_arrive_expect_tx_cluster(w_bar_leader, Cfg::kWBytes);
      for (int tap = 0; tap < 9; ++tap)             // this CTA's 128 of the 256 N rows of every tap
        tma_load_2d_pair(smem_w + tap * (Cfg::kN / 2) * 128, &tm_w, w_bar_leader, 0, tap * Cfg::kN + (int)rank * (Cfg::kN / 2));
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int pr = pr_begin; pr < pr_end; ++pr) {
      const int t = tile_of(pr);
      const int n = t / tiles_per_img;
      const int rem = t - n * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        const uint32_t full_leader = mapa_cluster(smem_u32(&full_bar[stage]), 0);
        mbar_arrive_expect_tx_cluster(full_leader, Geo::kTxBytes);
        tma_load_4d_pair(smem_a + stage * Geo::kStageBytes, &tm_x, full_leader, 0, tx * TW - 1, ty * TH - 1, n);
      }
      __syncwarp();
      if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_f16(256, Cfg::kN, std::is_same<T, __nv_bfloat16>::value);
      const uint32_t a_lo0 = desc_lo_sw128(smem_u32(smem_a));
      const uint32_t b_lo0 = desc_lo_sw128(smem_u32(smem_w));
      mbar_wait(w_bar, 0);
      tc_fence_after();
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int pr = pr_begin; pr < pr_end; ++pr, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + stage * (Geo::kStageBytes >> 4);
        const uint32_t d_tmem = tmem_base + acc * Cfg::kN;
        if (elect_one()) {
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              constexpr uint32_t kSbo = (uint32_t)((TW + 2) * 128) >> 4;
              const uint32_t hi = kSbo | (1u << 14) | (2u << 29);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t aoff = (uint32_t)(((r * (TW + 2) + s) * 128 + k * 32) >> 4);
                const uint64_t bdesc = desc_join(b_lo0 + (((r * 3 + s) * ((Cfg::kN / 2) * 128) + k * 32) >> 4), kDescHiSw128);
                if (p.pair_rows && ((s == 0 && k < 2) || (s == 2 && k >= 2))) continue;   // structural zeros (pairs.py)
                const uint32_t accf = (r | s) != 0 ? 1u : (p.pair_rows ? (k > 2 ? 1u : 0u) : (k != 0 ? 1u : 0u));
                umma_f16_pair(d_tmem, desc_join(a_lo + aoff, hi), bdesc, idesc, accf);
              }
            }
          }
          umma_commit_pair(&empty_bar[stage]);
          umma_commit_pair(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // =============================== epilogue warps (both CTAs, own 128 rows x 256 columns) ===============================
    // 8 warps: warp (q, g) drains TMEM lane quarter q = warp % 4 of column half g (GEMM column blocks 2g, 2g+1 = the
    // PixelShuffle positions (i, j) = (g, 0), (g, 1)).  Each 64-column block goes registers -> swizzled smem -> ONE TMA store:
    // the output tensor map (host: make_ps_out_map) is [N][H][i:2][W][(j,c):128], so the 32 pixels x 64 channels of a warp at
    // sub-position (i, j) are the box {64, 8, 1, 4, 1} at {j*64, x0, i, y0 + 4q, n}; image edges are clipped by the hardware.
    const int ew = warp - 2;
    const int q = warp & 3;
    const int g = ew >> 2;
    const uint32_t stg = smem_u32(smem_stg + ew * 4096);
    const float slope = __ldg(p.alpha);
    int it = 0;
    for (int pr = pr_begin; pr < pr_end; ++pr, ++it) {
      const int acc = it & 1;
      const int t_raw = 2 * pr + (int)rank;
      const bool tile_valid = t_raw < p.num_tiles;
      const int t = tile_valid ? t_raw : p.num_tiles - 1;
      const int n = t / tiles_per_img;
      const int rem = t - n * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int x0 = tx * TW, y0 = ty * TH;
      mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::kN + g * 128;
#pragma unroll 1
      for (int c2 = 0; c2 < 2; ++c2) {
        const int chunk = 2 * g + c2;             // GEMM column block = 2*i + j
        uint32_t pk[32];
        {
          uint32_t r0[32], r1[32];
          tmem_ld32(t_row + c2 * 64, r0);
          tmem_ld32(t_row + c2 * 64 + 32, r1);
          tmem_ld_wait();
          if (c2 == 1) {                          // this warp's TMEM reads of the accumulator are done
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_cluster(smem_u32(&tempty_bar[acc]), 0));
          }
          const float* bs = smem_bias + chunk * 64;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a0 = __uint_as_float(r0[2 * i]) + bs[2 * i];
            float a1 = __uint_as_float(r0[2 * i + 1]) + bs[2 * i + 1];
            float b0 = __uint_as_float(r1[2 * i]) + bs[32 + 2 * i];
            float b1 = __uint_as_float(r1[2 * i + 1]) + bs[32 + 2 * i + 1];
            a0 = a0 >= 0.f ? a0 : a0 * slope; a1 = a1 >= 0.f ? a1 : a1 * slope;
            b0 = b0 >= 0.f ? b0 : b0 * slope; b1 = b1 >= 0.f ? b1 : b1 * slope;
            pk[i] = Cvt<T>::pack2(a0, a1);
            pk[16 + i] = Cvt<T>::pack2(b0, b1);
          }
        }
        if (lane == 0) tma_store_wait_read();     // the previous TMA store has finished reading this warp's buffer
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          st_shared_v4(stg + lane * 128 + ((k ^ (lane & 7)) << 4), pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
        fence_proxy_async();                      // generic-proxy writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0 && tile_valid) {
          tma_store_5d(&tm_out, smem_stg + ew * 4096, (chunk & 1) * 64, x0, chunk >> 1, y0 + q * (32 / TW), n);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // the leader's MMAs have stopped reading the peer's shared memory
  if (warp == 1) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
}

}  // namespace fsr
