// conv3x3_gen.cuh - general implicit-GEMM 3x3 convolution on tcgen05: any Cin = 64*KC, any
// Cout = 64*slices, stride 1 or 2, forward or data-gradient, driven by a small "tap table".
//
// Replaces torch.nn.Conv2d at reference model.py:124-131 (SimpleBlock.conv, stride 1|2, widths 64..512),
// torchvision vgg19.features convs (model.py:8) and every `convolution_backward` data-gradient that
// autograd issues for trainer.py:180,195.
//
// One work item = (64-wide Cout slice, 16x8-pixel output tile).  K loop = for each 64-channel input
// chunk kc, for each "stage kind": one TMA box of activations (halo tile, <= 184 rows x 128 B) plus the
// weight rows of the taps that read this box ([NS][64] per tap, K-major); every tap is a UMMA descriptor
// into the box at row offset `a_off` with stride-byte-offset = box pitch (see conv3x3_tc.cuh for the
// unaligned-descriptor trick).
//   stride-1 forward / stride-1 dgrad : 1 kind  (box 10x18 at (x0-1,y0-1)), 9 taps at rows r*10+s.
//   stride-2 forward                  : 4 kinds = the 4 input parity planes (box 9x17 each),
//                                       tap (r,s) -> plane ((r!=1),(s!=1)), row (r!=0)*9 + (s!=0).
//   stride-2 dgrad, output parity (pr,ps): 1 kind (box 9x17 of dY at (x0,y0)), 1/2/2/4 taps.
// Activations are NHWC (fp16|bf16); a stride-2 conv reads its input in "parity-plane" layout
// [N][4][H/2][W/2][C] (written directly by the producing elementwise kernel), a stride-2 dgrad writes it.
#pragma once
#include "conv3x3_tc.cuh"

namespace fsr {

constexpr int kGenMaxKinds = 4;
constexpr int kGenMaxTaps = 9;

struct GenTap {
  int wrow;     // weight tap index 0..8 (row block in the packed [9][Cout][Cin] matrix)
  int a_off;    // first row (pixel) of this tap's A operand inside the kind's box
};
struct GenKind {
  int map;          // which activation tensor map
  int ntaps;
  int box_w;        // box pitch in pixels (stride-byte-offset = box_w * 128)
  int box_rows;     // box_w * box_h  (tx bytes = box_rows * 128)
  int dx, dy;       // box origin = (x0 + dx, y0 + dy)
  GenTap taps[kGenMaxTaps];
};

struct GenParams {
  int N, Ho, Wo;            // OUTPUT spatial size (tiles are over the output)
  int cin, cout_total;      // cin = 64*KC; cout_total = 64*num_slices
  int num_slices, tiles_x, tiles_y, num_tiles;
  int nkinds;
  GenKind kinds[kGenMaxKinds];
  void* out;                // NHWC T, pixel pitch cout_total
  long long out_img_stride; // elements between consecutive images of `out`
  const float* bias;        // [cout_total] or nullptr
  long long* stats;         // [N][cout_total][2] fixed-point int64 (EPI_RAW_STATS)
  const float* alpha;
  float slope;
  int act;
  int ps;                   // 1: PixelShuffle(2) scatter store (UpSamplingBlock with F != 64): GEMM columns are packed
                            //    (2i+j)*F + c, out = [N, 2Ho, 2Wo, F = cout_total/4]
  // FLAT mode (conv3x3_gen_2cta.cuh only): input and output are zero-bordered PADDED tensors [N][Ho+2][Wo+2][C]; an M tile
  // is 128 CONSECUTIVE positions of the flattened (n, y', x') index, whatever image they belong to (the <= 12x12 layers
  // of VGG19: 73 % / 56 % of every tile is image instead of 56 % / 28 % with 16x8 tiles).  Tap (r,s) is the row offset
  // r*(Wo+2)+s inside a plain 2-D box of the [Q][Cin] matrix; border positions are written as zeros.
  int flat;                 // 0 | 1
  int flat_q;               // Q = N * (Ho+2) * (Wo+2)
  int flat_lead;            // rows of the box before the tile's first position: (Wo+2) + 1
};

template <int MAXTAPS>
struct GenCfg {
  static constexpr int NS = 64;
  static constexpr int kABytes = 23552;                       // >= 10*18*128, 1024-aligned
  static constexpr int kBBytes = MAXTAPS * NS * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;       // 97280 | 56320
  static constexpr int kStages = MAXTAPS > 4 ? 2 : 3;
  static constexpr int kEpiWarps = 4;
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
  static constexpr int kStagingBytes = kEpiWarps * 4096;
  static constexpr int kTmemCols = 128;                       // 2 x 64
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 + 1024;
  static_assert(kSmemBytes <= 232448, "smem");
};

template <int EPI, typename T, int MAXTAPS>
__global__ void __launch_bounds__(GenCfg<MAXTAPS>::kThreads, 1)
conv3x3_gen_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
                   const __grid_constant__ CUtensorMap tm_a2, const __grid_constant__ CUtensorMap tm_a3,
                   const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ GenParams p) {
  pdl_grid_sync();
  using Cfg = GenCfg<MAXTAPS>;
  constexpr int NS = 64, TH = 16, TW = 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_st = smem;                                        // stages: [A | B taps]
  uint8_t* smem_stg = smem_st + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;     // [2]
  uint64_t* tempty_bar = tfull_bar + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
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

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_a0);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  for (int i = threadIdx.x; i < NS; i += blockDim.x)
    smem_bias[i] = (EPI != EPI_RAW_STATS && p.bias != nullptr) ? p.bias[slice * NS + i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    int stage = 0; uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int n = t / tiles_per_img;
      const int rem = t - n * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int x0 = tx * TW, y0 = ty * TH;
      for (int kc = 0; kc < KC; ++kc) {
        for (int kd = 0; kd < p.nkinds; ++kd) {
          const GenKind& K = p.kinds[kd];
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = smem_st + stage * Cfg::kStageBytes;
            mbar_arrive_expect_tx(&full_bar[stage], K.box_rows * 128 + K.ntaps * NS * 128);
            const CUtensorMap* tm = K.map == 0 ? &tm_a0 : K.map == 1 ? &tm_a1 : K.map == 2 ? &tm_a2 : &tm_a3;
            tma_load_4d(sa, tm, &full_bar[stage], kc * 64, x0 + K.dx, y0 + K.dy, n);
            for (int j = 0; j < K.ntaps; ++j)
              tma_load_2d(sa + Cfg::kABytes + j * NS * 128, &tm_w, &full_bar[stage], kc * 64,
                          K.taps[j].wrow * p.cout_total + slice * NS);
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc = make_idesc_f16(128, NS, std::is_same<T, __nv_bfloat16>::value);
    const uint32_t st_lo0 = desc_lo_sw128(smem_u32(smem_st));
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * NS;
      uint32_t first = 0;   // 0 -> overwrite accumulator
      for (int kc = 0; kc < KC; ++kc) {
        for (int kd = 0; kd < p.nkinds; ++kd) {
          const GenKind& K = p.kinds[kd];
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = st_lo0 + stage * (Cfg::kStageBytes >> 4);
          const uint32_t b_lo = a_lo + (Cfg::kABytes >> 4);
          const uint32_t a_hi = ((uint32_t)(K.box_w * 128) >> 4) | (1u << 14) | (2u << 29);
          if (elect_one()) {
            for (int j = 0; j < K.ntaps; ++j) {
              const uint32_t aj = a_lo + ((K.taps[j].a_off * 128) >> 4);
              const uint32_t bj = b_lo + ((j * NS * 128) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_f16(d_tmem, desc_join(aj + 2 * k, a_hi), desc_join(bj + 2 * k, kDescHiSw128), idesc, first);
                first = 1;
              }
            }
            umma_commit(&empty_bar[stage]);
            if (kc == KC - 1 && kd == p.nkinds - 1) umma_commit(&tfull_bar[acc]);
          }
          first = 1;
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // =============================== epilogue warps (4) ===============================
    const int ew = warp - 2;
    const int q = warp & 3;
    const uint32_t stg = smem_u32(smem_stg + ew * 4096);
    const int m = q * 32 + lane;
    const int yy = m / TW, xx = m % TW;
    (void)yy; (void)xx;
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
    int it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n = t / tiles_per_img;
      const int rem = t - n * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int x0 = tx * TW, y0 = ty * TH;
      const bool interior = (y0 + TH <= p.Ho) && (x0 + TW <= p.Wo);
      if (EPI == EPI_RAW_STATS && n != st_n) { flush_stats(st_n); st_n = n; }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * NS;
      uint32_t pk[32];
      {
        uint32_t r0[32], r1[32];
        tmem_ld32(t_row, r0);
        tmem_ld32(t_row + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a0 = __uint_as_float(r0[2 * i]), a1 = __uint_as_float(r0[2 * i + 1]);
          float b0 = __uint_as_float(r1[2 * i]), b1 = __uint_as_float(r1[2 * i + 1]);
          if constexpr (EPI != EPI_RAW_STATS) {
            a0 = apply_act(a0 + smem_bias[2 * i], p.act, slope);
            a1 = apply_act(a1 + smem_bias[2 * i + 1], p.act, slope);
            b0 = apply_act(b0 + smem_bias[32 + 2 * i], p.act, slope);
            b1 = apply_act(b1 + smem_bias[32 + 2 * i + 1], p.act, slope);
          }
          pk[i] = Cvt<T>::pack2(a0, a1);
          pk[16 + i] = Cvt<T>::pack2(b0, b1);
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
    }
    if (EPI == EPI_RAW_STATS) flush_stats(st_n);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

}  // namespace fsr
