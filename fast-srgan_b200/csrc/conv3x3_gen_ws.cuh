// conv3x3_gen_ws.cuh - the general implicit-GEMM 3x3 convolution of conv3x3_gen.cuh (same GenParams / tap tables, same
// results) re-blocked WEIGHT-STATIONARY over a group of T = 4 pixel tiles.
//
// Why: conv3x3_gen_kernel streams, for every 128-pixel tile and every 64-channel input chunk, the activation box
// (23 KB) AND the chunk's weight rows (9 taps x 8 KB = 72 KB): 95 KB of L2->smem fill per 36 MMAs (1152 tensor cycles)
// = 82 B/clk/SM, ~10 TB/s over 148 SMs - the L2, not the tensor pipe, sets its speed (350-700 TFLOP/s measured on the
// VGG / discriminator layers).  Here the loop nest is  group(T tiles) > K step (kc, kind) > tile:  the weight rows of a
// K step are loaded ONCE and used by T tiles whose T fp32 accumulators (T x 64 TMEM columns, two sets = all 512
// columns) stay live across the whole K loop.  Fill per group step: 72 + 4 x 23 KB per 144 MMAs = 36 B/clk/SM (2.3x
// less).  With a single K step (Cin = 64, one kind) the weights are loaded once per CTA.
//
// Warp roles (11 warps): 0 = activation-box TMA producer, 1 = weight TMA producer (its own warp: a weight prefetch
// must not queue behind box loads), 2 = tcgen05.mma issuer (+ TMEM alloc), 3..10 = two epilogue groups of four warps
// (TMEM lane quarter = warp % 4); group e drains tiles e, e+2 of each set.
#pragma once
#include "conv3x3_gen.cuh"

namespace fsr {

template <int MAXTAPS>
struct GenWsCfg {
  static constexpr int NS = 64;
  static constexpr int T = 4;                                  // pixel tiles per group = accumulators per TMEM set
  static constexpr int kABytes = 23552;                        // >= 10*18*128, 1024-aligned
  static constexpr int kAStages = MAXTAPS > 4 ? 2 : 4;
  static constexpr int kWBytes = MAXTAPS * NS * 128;           // one K step's weight rows: 73728 | 32768
  static constexpr int kWBufs = 2;
  static constexpr int kEpiWarps = 8;
  static constexpr int kThreads = 96 + 32 * kEpiWarps;         // 352
  static constexpr int kStagingBytes = kEpiWarps * 4096;
  static constexpr int kTmemCols = 512;                        // 2 sets x T x 64
  static constexpr int kSmemBytes = kAStages * kABytes + kWBufs * kWBytes + kStagingBytes + 1024 + 1024;
  static_assert(kSmemBytes <= 232448, "smem");
};

template <int EPI, typename T_, int MAXTAPS>
__global__ void __launch_bounds__(GenWsCfg<MAXTAPS>::kThreads, 1)
conv3x3_gen_ws_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
                      const __grid_constant__ CUtensorMap tm_a2, const __grid_constant__ CUtensorMap tm_a3,
                      const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ GenParams p) {
  pdl_grid_sync();
  using Cfg = GenWsCfg<MAXTAPS>;
  using T = T_;
  constexpr int NS = 64, TH = 16, TW = 8, TG = Cfg::T;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;                                            // activation boxes [kAStages]
  uint8_t* smem_w = smem_a + Cfg::kAStages * Cfg::kABytes;           // weight rows     [kWBufs]
  uint8_t* smem_stg = smem_w + Cfg::kWBufs * Cfg::kWBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + Cfg::kStagingBytes);
  uint64_t* afull = bars;                            // [kAStages]
  uint64_t* aempty = afull + Cfg::kAStages;          // [kAStages]
  uint64_t* wfull = aempty + Cfg::kAStages;          // [2]
  uint64_t* wempty = wfull + 2;                      // [2]
  uint64_t* tfull = wempty + 2;                      // [2]
  uint64_t* tempty = tfull + 2;                      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* smem_bias = reinterpret_cast<float*>(tmem_slot + 4);   // [64]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int slice = blockIdx.x % p.num_slices;
  const int cta_in_slice = blockIdx.x / p.num_slices;
  const int ctas_per_slice = gridDim.x / p.num_slices;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int t_begin = (int)(((long long)cta_in_slice * p.num_tiles) / ctas_per_slice);
  const int t_end = (int)(((long long)(cta_in_slice + 1) * p.num_tiles) / ctas_per_slice);
  const int KC = p.cin >> 6;
  const int nsteps = KC * p.nkinds;                  // K steps per group; step = kc * nkinds + kind
  const bool w_resident = nsteps == 1;               // one K step: its weights are loaded once per CTA and never released

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_a0);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < Cfg::kAStages; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1);
      mbar_init(&tfull[i], 1); mbar_init(&tempty[i], Cfg::kEpiWarps);
    }
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  for (int i = threadIdx.x; i < NS; i += blockDim.x)
    smem_bias[i] = (EPI != EPI_RAW_STATS && p.bias != nullptr) ? p.bias[slice * NS + i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== activation-box producer ===============================
    int stage = 0; uint32_t phase = 0;
    for (int t0 = t_begin; t0 < t_end; t0 += TG) {
      const int nt = min(TG, t_end - t0);
      for (int step = 0; step < nsteps; ++step) {
        const int kc = step / p.nkinds, kd = step - kc * p.nkinds;
        const GenKind& K = p.kinds[kd];
        const CUtensorMap* tm = K.map == 0 ? &tm_a0 : K.map == 1 ? &tm_a1 : K.map == 2 ? &tm_a2 : &tm_a3;
        for (int i = 0; i < nt; ++i) {
          const int t = t0 + i;
          const int n = t / tiles_per_img;
          const int rem = t - n * tiles_per_img;
          const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
          mbar_wait(&aempty[stage], phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&afull[stage], K.box_rows * 128);
            tma_load_4d(smem_a + stage * Cfg::kABytes, tm, &afull[stage], kc * 64, tx * TW + K.dx, ty * TH + K.dy, n);
          }
          __syncwarp();
          if (++stage == Cfg::kAStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== weight producer ===============================
    int buf = 0; uint32_t phase = 0;
    for (int t0 = t_begin; t0 < t_end; t0 += TG) {
      if (w_resident && t0 != t_begin) break;        // resident weights: loaded with the first group only
      for (int step = 0; step < nsteps; ++step) {
        const int kc = step / p.nkinds, kd = step - kc * p.nkinds;
        const GenKind& K = p.kinds[kd];
        mbar_wait(&wempty[buf], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&wfull[buf], K.ntaps * NS * 128);
          for (int j = 0; j < K.ntaps; ++j)
            tma_load_2d(smem_w + buf * Cfg::kWBytes + j * NS * 128, &tm_w, &wfull[buf], kc * 64,
                        K.taps[j].wrow * p.cout_total + slice * NS);
        }
        __syncwarp();
        if (++buf == Cfg::kWBufs) { buf = 0; phase ^= 1; }
      }
    }
  } else if (warp == 2) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc = make_idesc_f16(128, NS, std::is_same<T, __nv_bfloat16>::value);
    const uint32_t a_lo0 = desc_lo_sw128(smem_u32(smem_a));
    const uint32_t w_lo0 = desc_lo_sw128(smem_u32(smem_w));
    int stage = 0; uint32_t phase = 0;
    int wbuf = 0; uint32_t wphase = 0;
    int git = 0;
    for (int t0 = t_begin; t0 < t_end; t0 += TG, ++git) {
      const int nt = min(TG, t_end - t0);
      const int set = git & 1;
      mbar_wait(&tempty[set], ((git >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int step = 0; step < nsteps; ++step) {
        const int kd = step % p.nkinds;
        const GenKind& K = p.kinds[kd];
        if (!w_resident || t0 == t_begin) {
          mbar_wait(&wfull[wbuf], wphase);
          tc_fence_after();
        }
        const uint32_t b_lo = w_lo0 + wbuf * (Cfg::kWBytes >> 4);
        const uint32_t a_hi = ((uint32_t)(K.box_w * 128) >> 4) | (1u << 14) | (2u << 29);
        for (int i = 0; i < nt; ++i) {
          mbar_wait(&afull[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (Cfg::kABytes >> 4);
          const uint32_t d_tmem = tmem_base + (uint32_t)(set * TG + i) * NS;
          if (elect_one()) {
            uint32_t accum = step > 0 ? 1u : 0u;      // first K step of the group overwrites the accumulator
            for (int j = 0; j < K.ntaps; ++j) {
              const uint32_t aj = a_lo + ((K.taps[j].a_off * 128) >> 4);
              const uint32_t bj = b_lo + ((j * NS * 128) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_f16(d_tmem, desc_join(aj + 2 * k, a_hi), desc_join(bj + 2 * k, kDescHiSw128), idesc, accum);
                accum = 1;
              }
            }
            umma_commit(&aempty[stage]);
            if (i == nt - 1) {
              if (!w_resident) umma_commit(&wempty[wbuf]);       // all tiles of this K step issued: weights reusable
              if (step == nsteps - 1) umma_commit(&tfull[set]);
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
  } else {
    // =============================== epilogue: 2 groups x 4 warps ===============================
    const int ew = warp - 3;                // 0..7
    const int eg = ew >> 2;                 // epilogue group: tiles eg, eg+2 of every set
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const uint32_t stg = smem_u32(smem_stg + ew * 4096);
    const float slope = (p.act == ACT_PRELU) ? __ldg(p.alpha) : p.slope;
    long long st_s0 = 0, st_q0 = 0, st_s1 = 0, st_q1 = 0;     // fixed point, see conv3x3_tc.cuh
    int st_n = -1;
    auto flush_stats = [&](int img) {
      if (EPI == EPI_RAW_STATS && img >= 0) {
        long long* st = p.stats + ((size_t)img * p.cout_total + slice * NS + 2 * lane) * 2;
        stat_atomic_add(st + 0, st_s0);
        stat_atomic_add(st + 1, st_q0);
        stat_atomic_add(st + 2, st_s1);
        stat_atomic_add(st + 3, st_q1);
      }
      st_s0 = st_q0 = st_s1 = st_q1 = 0;
    };
    int git = 0;
    for (int t0 = t_begin; t0 < t_end; t0 += TG, ++git) {
      const int nt = min(TG, t_end - t0);
      const int set = git & 1;
      mbar_wait(&tfull[set], (git >> 1) & 1);
      tc_fence_after();
      for (int i = eg; i < TG; i += 2) {
        const bool last_of_mine = i + 2 >= TG;
        if (i >= nt) {                      // short last group: nothing to drain, but the set must still be released
          if (last_of_mine) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[set]);
          }
          continue;
        }
        const int t = t0 + i;
        const int n = t / tiles_per_img;
        const int rem = t - n * tiles_per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int x0 = tx * TW, y0 = ty * TH;
        const bool interior = (y0 + TH <= p.Ho) && (x0 + TW <= p.Wo);
        if (EPI == EPI_RAW_STATS && n != st_n) { flush_stats(st_n); st_n = n; }
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * TG + i) * NS;
        uint32_t pk[32];
        {
          uint32_t r0[32], r1[32];
          tmem_ld32(t_row, r0);
          tmem_ld32(t_row + 32, r1);
          tmem_ld_wait();
          if (last_of_mine) {               // this warp has read everything it needs from the set
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[set]);
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            float a0 = __uint_as_float(r0[2 * c]), a1 = __uint_as_float(r0[2 * c + 1]);
            float b0 = __uint_as_float(r1[2 * c]), b1 = __uint_as_float(r1[2 * c + 1]);
            if constexpr (EPI != EPI_RAW_STATS) {
              a0 = apply_act(a0 + smem_bias[2 * c], p.act, slope);
              a1 = apply_act(a1 + smem_bias[2 * c + 1], p.act, slope);
              b0 = apply_act(b0 + smem_bias[32 + 2 * c], p.act, slope);
              b1 = apply_act(b1 + smem_bias[32 + 2 * c + 1], p.act, slope);
            }
            pk[c] = Cvt<T>::pack2(a0, a1);
            pk[16 + c] = Cvt<T>::pack2(b0, b1);
          }
        }
        const int col0 = slice * NS;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          st_shared_v4(stg + lane * 128 + ((k ^ (lane & 7)) << 4), pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
        __syncwarp();
        uint4 val[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rrow = j * 4 + (lane >> 3);
          val[j] = ld_shared_v4(stg + rrow * 128 + (((lane & 7) ^ (rrow & 7)) << 4));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rrow = j * 4 + (lane >> 3);
          const int mm = q * 32 + rrow;
          const int py = y0 + mm / TW, px = x0 + mm % TW;
          if (interior || (py < p.Ho && px < p.Wo)) {
            T* dst;
            if (p.ps) {
              const int F = p.cout_total >> 2;
              const int qq = col0 / F, c0 = col0 - qq * F;
              dst = reinterpret_cast<T*>(p.out) + (size_t)n * p.out_img_stride +
                    ((size_t)(2 * py + (qq >> 1)) * (2 * p.Wo) + 2 * px + (qq & 1)) * F + c0;
            } else {
              dst = reinterpret_cast<T*>(p.out) + (size_t)n * p.out_img_stride +
                    ((size_t)py * p.Wo + px) * p.cout_total + col0;
            }
            *reinterpret_cast<uint4*>(dst + (lane & 7) * 8) = val[j];
          }
        }
        if constexpr (EPI == EPI_RAW_STATS) {
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
        __syncwarp();                       // staging buffer is re-written by this warp's next tile
      }
    }
    if (EPI == EPI_RAW_STATS) flush_stats(st_n);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

}  // namespace fsr
