// conv3x3_gen_2cta.cuh - the general implicit-GEMM 3x3 convolution (conv3x3_gen.cuh: same GenParams / tap tables, same
// results) as a CTA-PAIR kernel with 128-wide output-channel slices:  tcgen05.mma.cta_group::2, M = 256 (one 128-pixel
// tile per CTA), N = 128, weight-stationary over groups of two tile pairs.
//
// Why (measured, profiles/r02): the 64-column kernels (conv3x3_gen_ws: VGG19 + discriminator forward / data gradient =
// 44 % of the GAN step) fetch 4 KB of A and 2 KB of B from shared memory per 32-cycle MMA = 192 B/clk against a
// ~128 B/clk port: tensor pipe 42 %, the MMA warp never waits.  Here each SM feeds the pair's MMA with its own A tile
// (4 KB) and HALF of a 128-row weight tile (64 rows, 2 KB) per 64-cycle MMA = 96 B/clk - the same change that took the
// 64->256 upsampling conv from 0.68 to 1.05 of the sustained GEMM peak (conv3x3_up_2cta.cuh: tensor pipe 99 %).
//
// Loop nest per cluster: group (2 tile pairs) > K step (kc, kind) > tile pair.  The K step's weight rows (taps x 64 rows
// of THIS CTA's half) are loaded once per group step and used by both tile pairs; 2 sets x 2 accumulators x 128 columns
// = all 512 TMEM columns.  Warp roles (11 warps): 0 = activation-box TMA producer, 1 = weight TMA producer,
// 2 = MMA issuer (leader CTA only) + TMEM alloc, 3..10 = epilogue (TMEM lane quarter = warp % 4, column half = (warp-3)/4).
// Protocol as in conv3x3_up_2cta.cuh: full barriers live in the leader (both producers arrive remotely, TMA
// .cta_group::2 completes on them), empty / tfull barriers are local and signalled by the leader's multicast commits,
// tempty lives in the leader (8 warps x 2 CTAs arrive).
#pragma once
#include "conv3x3_gen.cuh"
#include "conv3x3_up_2cta.cuh"

namespace fsr {

template <int MAXTAPS>
struct Gen2Cfg {
  static constexpr int NS = 128;                               // GEMM columns of the pair's MMA (64 weight rows per CTA)
  static constexpr int TG = 2;                                 // tile pairs per group = accumulators per TMEM set
  static constexpr int kABytes = 23552;                        // >= 10*18*128, 1024-aligned
  static constexpr int kAStages = MAXTAPS > 4 ? 2 : 4;
  static constexpr int kWBytes = MAXTAPS * 64 * 128;           // one K step's weight rows of this CTA: 73728 | 32768
  static constexpr int kWBufs = 2;
  static constexpr int kEpiWarps = 8;
  static constexpr int kThreads = 96 + 32 * kEpiWarps;         // 352
  static constexpr int kStagingBytes = kEpiWarps * 4096;
  static constexpr int kTmemCols = 512;                        // 2 sets x TG x 128
  static constexpr int kSmemBytes = kAStages * kABytes + kWBufs * kWBytes + kStagingBytes + 1024 + 1024;
  static_assert(kSmemBytes <= 232448, "smem");
};

template <int EPI, typename T_, int MAXTAPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Gen2Cfg<MAXTAPS>::kThreads, 1)
conv3x3_gen_2cta_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
                        const __grid_constant__ CUtensorMap tm_a2, const __grid_constant__ CUtensorMap tm_a3,
                        const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_out,
                        const __grid_constant__ GenParams p) {
  pdl_grid_sync();
  using Cfg = Gen2Cfg<MAXTAPS>;
  using T = T_;
  constexpr int TH = 16, TW = 8, TG = Cfg::TG;
  extern __shared__ uint8_t smem_raw[];
  // identical layout in both CTAs (one descriptor addresses both shared memories)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;                                            // activation boxes [kAStages]
  uint8_t* smem_w = smem_a + Cfg::kAStages * Cfg::kABytes;           // weight rows     [kWBufs]
  uint8_t* smem_stg = smem_w + Cfg::kWBufs * Cfg::kWBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + Cfg::kStagingBytes);
  uint64_t* afull = bars;                            // [kAStages] leader (count 2 + tx)
  uint64_t* aempty = afull + Cfg::kAStages;          // [kAStages] local, multicast commit
  uint64_t* wfull = aempty + Cfg::kAStages;          // [2] leader (count 2 + tx)
  uint64_t* wempty = wfull + 2;                      // [2] local, multicast commit
  uint64_t* tfull = wempty + 2;                      // [2] local, multicast commit
  uint64_t* tempty = tfull + 2;                      // [2] leader (count 2 * kEpiWarps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* smem_bias = reinterpret_cast<float*>(tmem_slot + 4);   // [128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int nsl = p.num_slices;                      // 128-wide slices
  const int cluster_id = blockIdx.x >> 1;
  const int slice = cluster_id % nsl;
  const int cl_in_slice = cluster_id / nsl;
  const int cls_per_slice = (gridDim.x >> 1) / nsl;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int num_pairs = (p.num_tiles + 1) >> 1;
  const int pr_begin = (int)(((long long)cl_in_slice * num_pairs) / cls_per_slice);
  const int pr_end = (int)(((long long)(cl_in_slice + 1) * num_pairs) / cls_per_slice);
  const int KC = p.cin >> 6;
  const int nsteps = KC * p.nkinds;                  // K steps per group; step = kc * nkinds + kind
  const bool w_resident = nsteps == 1;               // one K step: its weights are loaded once per CTA and never released

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_a0);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_out);
    for (int i = 0; i < Cfg::kAStages; ++i) { mbar_init(&afull[i], 2); mbar_init(&aempty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&wfull[i], 2); mbar_init(&wempty[i], 1);
      mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * Cfg::kEpiWarps);
    }
    fence_mbar_init();
    fence_proxy_async();
  }
  for (int i = threadIdx.x; i < Cfg::NS; i += blockDim.x)
    smem_bias[i] = (EPI != EPI_RAW_STATS && p.bias != nullptr) ? p.bias[slice * Cfg::NS + i] : 0.f;
  __syncthreads();
  cluster_sync_all();                                // both CTAs' barriers are initialised before any remote arrive
  if (warp == 2) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_of = [&](int pair) {                     // this CTA's tile of the pair (clamped: an odd tail re-does the last tile)
    const int t = 2 * pair + (int)rank;
    return t < p.num_tiles ? t : p.num_tiles - 1;
  };

  if (warp == 0) {
    // =============================== activation-box producer (both CTAs) ===============================
    int stage = 0; uint32_t phase = 0;
    for (int p0 = pr_begin; p0 < pr_end; p0 += TG) {
      const int np = min(TG, pr_end - p0);
      for (int step = 0; step < nsteps; ++step) {
        const int kc = step / p.nkinds, kd = step - kc * p.nkinds;
        const GenKind& K = p.kinds[kd];
        const CUtensorMap* tm = K.map == 0 ? &tm_a0 : K.map == 1 ? &tm_a1 : K.map == 2 ? &tm_a2 : &tm_a3;
        for (int i = 0; i < np; ++i) {
          const int t = tile_of(p0 + i);
          const int n = t / tiles_per_img;
          const int rem = t - n * tiles_per_img;
          const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
          mbar_wait(&aempty[stage], phase ^ 1);
          if (elect_one()) {
            const uint32_t full_leader = mapa_cluster(smem_u32(&afull[stage]), 0);
            mbar_arrive_expect_tx_cluster(full_leader, K.box_rows * 128);
            if (p.flat)        // rows [t*128 - lead, ...) of the flattened padded tensor; rows outside [0, Q) are zero-filled
              tma_load_2d_pair(smem_a + stage * Cfg::kABytes, tm, full_leader, kc * 64, t * 128 - p.flat_lead);
            else
              tma_load_4d_pair(smem_a + stage * Cfg::kABytes, tm, full_leader, kc * 64, tx * TW + K.dx, ty * TH + K.dy, n);
          }
          __syncwarp();
          if (++stage == Cfg::kAStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== weight producer (both CTAs: own 64 of the slice's 128 rows) ===============================
    int buf = 0; uint32_t phase = 0;
    for (int p0 = pr_begin; p0 < pr_end; p0 += TG) {
      if (w_resident && p0 != pr_begin) break;       // resident weights: loaded with the first group only
      for (int step = 0; step < nsteps; ++step) {
        const int kc = step / p.nkinds, kd = step - kc * p.nkinds;
        const GenKind& K = p.kinds[kd];
        mbar_wait(&wempty[buf], phase ^ 1);
        if (elect_one()) {
          const uint32_t full_leader = mapa_cluster(smem_u32(&wfull[buf]), 0);
          mbar_arrive_expect_tx_cluster(full_leader, K.ntaps * 64 * 128);
          for (int j = 0; j < K.ntaps; ++j)
            tma_load_2d_pair(smem_w + buf * Cfg::kWBytes + j * 64 * 128, &tm_w, full_leader, kc * 64,
                             K.taps[j].wrow * p.cout_total + slice * Cfg::NS + (int)rank * 64);
        }
        __syncwarp();
        if (++buf == Cfg::kWBufs) { buf = 0; phase ^= 1; }
      }
    }
  } else if (warp == 2) {
    // =============================== MMA issuer (leader CTA only) ===============================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_f16(256, Cfg::NS, std::is_same<T, __nv_bfloat16>::value);
      const uint32_t a_lo0 = desc_lo_sw128(smem_u32(smem_a));
      const uint32_t w_lo0 = desc_lo_sw128(smem_u32(smem_w));
      int stage = 0; uint32_t phase = 0;
      int wbuf = 0; uint32_t wphase = 0;
      int git = 0;
      for (int p0 = pr_begin; p0 < pr_end; p0 += TG, ++git) {
        const int np = min(TG, pr_end - p0);
        const int set = git & 1;
        mbar_wait(&tempty[set], ((git >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int step = 0; step < nsteps; ++step) {
          const int kd = step % p.nkinds;
          const GenKind& K = p.kinds[kd];
          if (!w_resident || p0 == pr_begin) {
            mbar_wait(&wfull[wbuf], wphase);
            tc_fence_after();
          }
          const uint32_t b_lo = w_lo0 + wbuf * (Cfg::kWBytes >> 4);
          const uint32_t a_hi = ((uint32_t)(K.box_w * 128) >> 4) | (1u << 14) | (2u << 29);
          for (int i = 0; i < np; ++i) {
            mbar_wait(&afull[stage], phase);
            tc_fence_after();
            const uint32_t a_lo = a_lo0 + stage * (Cfg::kABytes >> 4);
            const uint32_t d_tmem = tmem_base + (uint32_t)(set * TG + i) * Cfg::NS;
            if (elect_one()) {
              uint32_t accum = step > 0 ? 1u : 0u;    // first K step of the group overwrites the accumulator
              for (int j = 0; j < K.ntaps; ++j) {
                const uint32_t aj = a_lo + ((K.taps[j].a_off * 128) >> 4);
                const uint32_t bj = b_lo + ((j * 64 * 128) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_f16_pair(d_tmem, desc_join(aj + 2 * k, a_hi), desc_join(bj + 2 * k, kDescHiSw128), idesc, accum);
                  accum = 1;
                }
              }
              umma_commit_pair(&aempty[stage]);
              if (i == np - 1) {
                if (!w_resident) umma_commit_pair(&wempty[wbuf]);    // all tile pairs of this K step issued: weights reusable
                if (step == nsteps - 1) umma_commit_pair(&tfull[set]);
              }
            }
            __syncwarp();
            if (++stage == Cfg::kAStages) { stage = 0; phase ^= 1; }
          }
          if (!w_resident) {
            if (++wbuf == Cfg::kWBufs) { wbuf = 0; wphase ^= 1; }
          }
        }
      }
    }
  } else {
    // =============================== epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves ===============================
    const int ew = warp - 3;                // 0..7
    const int hcol = ew >> 2;               // 64-column half of the 128-wide slice
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const uint32_t stg = smem_u32(smem_stg + ew * 4096);
    const float slope = (p.act == ACT_PRELU) ? __ldg(p.alpha) : p.slope;
    const int col0 = slice * Cfg::NS + hcol * 64;       // first GEMM column of this warp's chunk
    const float* bs = smem_bias + hcol * 64;
    long long st_s0 = 0, st_q0 = 0, st_s1 = 0, st_q1 = 0;     // fixed point, see conv3x3_tc.cuh
    int st_n = -1;
    auto flush_stats = [&](int img) {
      if (EPI == EPI_RAW_STATS && img >= 0) {
        long long* st = p.stats + ((size_t)img * p.cout_total + col0 + 2 * lane) * 2;
        stat_atomic_add(st + 0, st_s0);
        stat_atomic_add(st + 1, st_q0);
        stat_atomic_add(st + 2, st_s1);
        stat_atomic_add(st + 3, st_q1);
      }
      st_s0 = st_q0 = st_s1 = st_q1 = 0;
    };
    const uint32_t tempty_leader0 = mapa_cluster(smem_u32(&tempty[0]), 0);
    int git = 0;
    for (int p0 = pr_begin; p0 < pr_end; p0 += TG, ++git) {
      const int np = min(TG, pr_end - p0);
      const int set = git & 1;
      mbar_wait(&tfull[set], (git >> 1) & 1);
      tc_fence_after();
      for (int i = 0; i < TG; ++i) {
        const bool last = i == TG - 1;
        if (i >= np) {                      // short last group: nothing to drain, but the set must still be released
          if (last) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader0 + set * 8);
          }
          continue;
        }
        const int t_raw = 2 * (p0 + i) + (int)rank;
        const bool tile_valid = t_raw < p.num_tiles;
        const int t = tile_valid ? t_raw : p.num_tiles - 1;
        const int n = t / tiles_per_img;
        const int rem = t - n * tiles_per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int x0 = tx * TW, y0 = ty * TH;
        const bool interior = (y0 + TH <= p.Ho) && (x0 + TW <= p.Wo);
        bool row_live = true;               // flat mode: this thread's position is an image pixel (not border / tail padding)
        if (p.flat) {
          const int pos = t * 128 + q * 32 + lane;
          const int W2 = p.Wo + 2, per = (p.Ho + 2) * W2;
          const int rm = pos % per, yy = rm / W2, xx = rm - yy * W2;
          row_live = pos < p.flat_q && yy >= 1 && yy <= p.Ho && xx >= 1 && xx <= p.Wo;
        }
        if (EPI == EPI_RAW_STATS && tile_valid && n != st_n) { flush_stats(st_n); st_n = n; }
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * TG + i) * Cfg::NS + hcol * 64;
        uint32_t pk[32];
        {
          uint32_t r0[32], r1[32];
          tmem_ld32(t_row, r0);
          tmem_ld32(t_row + 32, r1);
          tmem_ld_wait();
          if (last) {                       // this warp has read everything it needs from the set
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader0 + set * 8);
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            float a0 = __uint_as_float(r0[2 * c]), a1 = __uint_as_float(r0[2 * c + 1]);
            float b0 = __uint_as_float(r1[2 * c]), b1 = __uint_as_float(r1[2 * c + 1]);
            if constexpr (EPI != EPI_RAW_STATS) {
              a0 = apply_act(a0 + bs[2 * c], p.act, slope);
              a1 = apply_act(a1 + bs[2 * c + 1], p.act, slope);
              b0 = apply_act(b0 + bs[32 + 2 * c], p.act, slope);
              b1 = apply_act(b1 + bs[32 + 2 * c + 1], p.act, slope);
            }
            pk[c] = row_live ? Cvt<T>::pack2(a0, a1) : 0u;      // flat mode keeps the zero border of the padded layout
            pk[16 + c] = row_live ? Cvt<T>::pack2(b0, b1) : 0u;
          }
        }
        if (lane == 0) tma_store_wait_read();           // the previous TMA store has finished reading this warp's buffer
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          st_shared_v4(stg + lane * 128 + ((k ^ (lane & 7)) << 4), pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
        fence_proxy_async();                            // generic-proxy writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0 && tile_valid) {
          // NHWC [N][Ho][Wo][cout_total] (image stride may be a parity-plane stride): 32 pixels x 64 channels of this warp
          if (p.flat) {
            tma_store_2d(&tm_out, smem_stg + ew * 4096, col0, t * 128 + q * 32);   // [Q][cout_total], tail clipped
          } else if (p.ps) {
            // UpSamplingBlock with F = cout_total / 4 (a multiple of 128): GEMM column (2i+j)*F + c -> out[n, 2y+i, 2x+j, c];
            // the output is viewed as [N][Ho][i][Wo][(j,c): 2F] (host: make_ps_out_map), one box per (i, j, 64 channels)
            const int Fc = p.cout_total >> 2;
            const int qq = col0 / Fc, c0 = col0 - qq * Fc;
            tma_store_5d(&tm_out, smem_stg + ew * 4096, (qq & 1) * Fc + c0, x0, qq >> 1, y0 + (q * 32) / TW, n);
          } else {
            tma_store_4d(&tm_out, smem_stg + ew * 4096, col0, x0, y0 + (q * 32) / TW, n);
          }
          tma_store_commit();
        }
        if constexpr (EPI == EPI_RAW_STATS) {
          if (tile_valid) {
            const uint32_t colw = ((lane & 3) << 2);
            float t_s0 = 0.f, t_q0 = 0.f, t_s1 = 0.f, t_q1 = 0.f;      // this tile's partial sums (fixed order)
#pragma unroll 4
            for (int rr = 0; rr < 32; ++rr) {
              const int mm = q * 32 + rr;
              const bool ok = interior || ((y0 + mm / TW < p.Ho) && (x0 + mm % TW < p.Wo));
              const uint32_t w = ld_shared_u32(stg + rr * 128 + ((((lane >> 2) ^ (rr & 7))) << 4) + colw);
              const float2 f = Cvt<T>::unpack2(w);
              if (ok) {
                t_s0 += f.x; t_q0 = fmaf(f.x, f.x, t_q0);
                t_s1 += f.y; t_q1 = fmaf(f.y, f.y, t_q1);
              }
            }
            st_s0 += stat_fix(t_s0, kStatSumScale); st_q0 += stat_fix(t_q0, kStatSqScale);
            st_s1 += stat_fix(t_s1, kStatSumScale); st_q1 += stat_fix(t_q1, kStatSqScale);
          }
        }
        __syncwarp();
      }
    }
    if (lane == 0) tma_store_wait_all();
    if (EPI == EPI_RAW_STATS) flush_stats(st_n);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // the leader's MMAs have stopped reading the peer's shared memory
  if (warp == 2) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
}

}  // namespace fsr
