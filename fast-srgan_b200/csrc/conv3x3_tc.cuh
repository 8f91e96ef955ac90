// conv3x3_tc.cuh - implicit-GEMM 3x3 / stride-1 / pad-1 convolution on tcgen05 tensor cores.
//
// Replaces torch.nn.Conv2d(F, F|4F|3, kernel_size=3, padding=1) at reference model.py:30-35,
// 47-54, 57-64, 87-93, 103-108 (and their fused neighbours: InstanceNorm statistics :55,65,94,
// PixelShuffle+PReLU :36-37, Tanh :109).
//
// Layout: activations NHWC, 64 input channels = one 128-byte row per pixel (fp16 or bf16).
// GEMM view: D[pixels, Cout] = sum over 9 taps of  A_tap[pixels, 64] * W_tap[64, Cout].
//   * M tile  = 8 x 16 output pixels (128 rows = one UMMA M),
//   * A       = TMA 4-D box {64ch, 16, 8+2, 1} of the input at (x0+s-1, y0-1): a halo tile per
//               COLUMN shift s (zero fill outside the image = the conv padding).  The three ROW
//               shifts r reuse the same smem tile through a descriptor offset of r*16 rows
//               (2048 B, a multiple of the 1024-B swizzle atom) -> 3 smem fills per tile, not 9.
//   * B       = the Cout-slice of the weights, resident in smem for the whole persistent CTA
//               (9 taps x NS rows x 128 B, K-major, 128B swizzle).
//   * D       = fp32 accumulators in TMEM, double buffered (2 x NS columns).
// Warp roles (192 threads): warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc),
// warps2-5 = epilogue (TMEM -> registers -> fused epilogue -> coalesced global stores).
#pragma once
#include "fsr_common.cuh"
#include <type_traits>

namespace fsr {

enum ConvEpilogue : int {
  EPI_RAW_STATS = 0,   // store raw conv output (NHWC) + per-(n,c) sum / sum-of-squares (InstanceNorm stats)
  EPI_BIAS_ACT = 1,    // store act(conv + bias) NHWC
  EPI_PS_PRELU = 2,    // bias + PReLU + PixelShuffle(2) scatter: out[N,2H,2W,64]
  EPI_HEAD_TANH = 3,   // NS=16 (3 real channels): tanh(conv + bias) -> fp32 NCHW or uint8 NHWC
};

enum ActMode : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_PRELU = 3 };

struct ConvParams {
  int N, H, W;              // conv input == output spatial size (stride 1, pad 1)
  int cout_total;           // GEMM N over all slices (multiple of NS)
  int num_slices;           // cout_total / NS
  int tiles_x, tiles_y;     // ceil(W/16), ceil(H/8)
  int num_tiles;            // N * tiles_y * tiles_x
  void* out;                // see ConvEpilogue
  const float* bias;        // [cout_total] in GEMM column order, or nullptr
  float* stats;             // [N][cout_total][2] fp32 (sum, sumsq), EPI_RAW_STATS
  const float* alpha;       // device pointer to the PReLU slope (ACT_PRELU)
  float slope;              // LeakyReLU slope (ACT_LRELU)
  int act;                  // ActMode (EPI_BIAS_ACT)
  int out_u8;               // EPI_HEAD_TANH: 0 -> fp32 NCHW [N,3,H,W]; 1 -> uint8 NHWC [N,H,W,3]
};

constexpr int kTileH = 8, kTileW = 16;
constexpr int kStageBytes = (kTileH + 2) * kTileW * 128;   // 20480
constexpr int kStagingBytes = 4 * 4096;                    // per-epilogue-warp transpose buffers
constexpr int kConvThreads = 192;

template <int NS>
struct ConvCfg {
  static constexpr int kWBytes = 9 * NS * 128;
  static constexpr int kStages = (NS >= 128) ? 3 : 6;
  static constexpr int kTmemCols = (2 * NS <= 32) ? 32 : (2 * NS <= 64 ? 64 : (2 * NS <= 128 ? 128 : 256));
  static constexpr int kSmemBytes = kWBytes + kStages * kStageBytes + kStagingBytes + 1024 /*barriers*/ + 1024 /*align*/;
};

FSR_DEVINL float apply_act(float v, int act, float slope) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_LRELU || act == ACT_PRELU) return v >= 0.f ? v : v * slope;
  return v;
}

// Sum 64 per-lane values across the 32 lanes of a warp with a halving butterfly
// (32+16+8+4+2 = 62 shuffles): afterwards lane L holds the totals of columns 2L and 2L+1.
FSR_DEVINL void warp_reduce64(float (&v)[64], int lane) {
#pragma unroll
  for (int step = 0; step < 5; ++step) {
    const int off = 16 >> step;
    const int half = 32 >> step;
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half];
      const float keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

template <int NS, int EPI, typename T>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_c64_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                   const ConvParams p) {
  using Cfg = ConvCfg<NS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem_w + Cfg::kWBytes;
  uint8_t* smem_stg = smem_a + Cfg::kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + kStagingBytes);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]
  uint64_t* w_bar = bars + 2 * Cfg::kStages;       // [1]
  uint64_t* tfull_bar = w_bar + 1;                 // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* smem_bias = reinterpret_cast<float*>(tmem_slot + 4);   // [NS]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int slice = blockIdx.x % p.num_slices;
  const int cta_in_slice = blockIdx.x / p.num_slices;
  const int ctas_per_slice = gridDim.x / p.num_slices;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(w_bar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  if (EPI != EPI_RAW_STATS && p.bias != nullptr) {
    for (int i = threadIdx.x; i < NS; i += blockDim.x) smem_bias[i] = p.bias[slice * NS + i];
  } else {
    for (int i = threadIdx.x; i < NS; i += blockDim.x) smem_bias[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      mbar_arrive_expect_tx(w_bar, Cfg::kWBytes);
      for (int tap = 0; tap < 9; ++tap)
        tma_load_2d(smem_w + tap * NS * 128, &tm_w, w_bar, 0, tap * p.cout_total + slice * NS);
      int stage = 0; uint32_t phase = 0;
      for (int t = cta_in_slice; t < p.num_tiles; t += ctas_per_slice) {
        const int n = t / tiles_per_img;
        const int rem = t - n * tiles_per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int x0 = tx * kTileW, y0 = ty * kTileH;
        for (int s = 0; s < 3; ++s) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
          tma_load_4d(smem_a + stage * kStageBytes, &tm_x, &full_bar[stage], 0, x0 + s - 1, y0 - 1, n);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, NS, std::is_same<T, __nv_bfloat16>::value);
      const uint32_t w_base = smem_u32(smem_w);
      mbar_wait(w_bar, 0);
      tc_fence_after();
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int t = cta_in_slice; t < p.num_tiles; t += ctas_per_slice, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * NS;
        for (int s = 0; s < 3; ++s) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem_a + stage * kStageBytes);
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adesc = make_kmajor_sw128_desc(a_base + r * (kTileW * 128) + k * 32);
              const uint64_t bdesc = make_kmajor_sw128_desc(w_base + (r * 3 + s) * (NS * 128) + k * 32);
              umma_f16(d_tmem, adesc, bdesc, idesc, (s | r | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue warps (2..5) ===============================
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    uint8_t* stg = smem_stg + (warp - 2) * 4096;
    const int m = q * 32 + lane;                 // accumulator row == pixel within the tile
    const int yy = m / kTileW, xx = m % kTileW;
    float prelu_a = 0.f;
    if (EPI == EPI_PS_PRELU || (EPI == EPI_BIAS_ACT && p.act == ACT_PRELU)) prelu_a = __ldg(p.alpha);
    const float slope = (EPI == EPI_PS_PRELU || p.act == ACT_PRELU) ? prelu_a : p.slope;
    int it = 0;
    for (int t = cta_in_slice; t < p.num_tiles; t += ctas_per_slice, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n = t / tiles_per_img;
      const int rem = t - n * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int x0 = tx * kTileW, y0 = ty * kTileH;
      const int y = y0 + yy, x = x0 + xx;
      const bool pvalid = (y < p.H) && (x < p.W);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * NS;

      if constexpr (EPI == EPI_HEAD_TANH) {
        uint32_t r[16];
        tmem_ld16(t_row, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        if (pvalid) {
          float o[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) o[c] = tanhf(__uint_as_float(r[c]) + smem_bias[c]);
          if (p.out_u8) {
            // reference inference.py:54-56: ((y+1)/2*255).astype(uint8)  (truncation)
            uint8_t* o8 = reinterpret_cast<uint8_t*>(p.out) + ((size_t)(n * p.H + y) * p.W + x) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              float f = (o[c] + 1.0f) / 2.0f * 255.0f;
              o8[c] = (uint8_t)(int)fminf(fmaxf(f, 0.f), 255.f);
            }
          } else {
            float* of = reinterpret_cast<float*>(p.out);
            const size_t plane = (size_t)p.H * p.W;
#pragma unroll
            for (int c = 0; c < 3; ++c) of[((size_t)n * 3 + c) * plane + (size_t)y * p.W + x] = o[c];
          }
        }
      } else {
#pragma unroll 1
        for (int chunk = 0; chunk < NS / 64; ++chunk) {
          float v[64];
          {
            uint32_t r0[32], r1[32];
            tmem_ld32(t_row + chunk * 64, r0);
            tmem_ld32(t_row + chunk * 64 + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) { v[i] = __uint_as_float(r0[i]); v[32 + i] = __uint_as_float(r1[i]); }
          }
          if (chunk == NS / 64 - 1) {           // all TMEM reads of this accumulator are done
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          }
          const int col0 = slice * NS + chunk * 64;   // first GEMM column of this chunk

          if constexpr (EPI != EPI_RAW_STATS) {
#pragma unroll
            for (int i = 0; i < 64; ++i) {
              float a = v[i] + smem_bias[chunk * 64 + i];
              v[i] = (EPI == EPI_PS_PRELU) ? (a >= 0.f ? a : a * slope) : apply_act(a, p.act, slope);
            }
          }

          // ---- registers -> swizzled smem (row = pixel, 8 x 16B chunks) -> coalesced global
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            uint4 pk;
            pk.x = Cvt<T>::pack2(v[8 * k + 0], v[8 * k + 1]);
            pk.y = Cvt<T>::pack2(v[8 * k + 2], v[8 * k + 3]);
            pk.z = Cvt<T>::pack2(v[8 * k + 4], v[8 * k + 5]);
            pk.w = Cvt<T>::pack2(v[8 * k + 6], v[8 * k + 7]);
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((k ^ (lane & 7)) << 4)) = pk;
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int rrow = j * 4 + (lane >> 3);
            const int c16 = lane & 7;
            const uint4 val = *reinterpret_cast<const uint4*>(stg + rrow * 128 + ((c16 ^ (rrow & 7)) << 4));
            const int mm = q * 32 + rrow;
            const int py = y0 + mm / kTileW, px = x0 + mm % kTileW;
            if (py < p.H && px < p.W) {
              T* dst;
              if constexpr (EPI == EPI_PS_PRELU) {
                const int qq = col0 >> 6;               // GEMM column block = 2*i + j
                const int oy = 2 * py + (qq >> 1), ox = 2 * px + (qq & 1);
                dst = reinterpret_cast<T*>(p.out) + ((size_t)(n * 2 * p.H + oy) * (2 * p.W) + ox) * 64;
              } else {
                dst = reinterpret_cast<T*>(p.out) + ((size_t)(n * p.H + py) * p.W + px) * p.cout_total + col0;
              }
              *reinterpret_cast<uint4*>(dst + c16 * 8) = val;
            }
          }

          if constexpr (EPI == EPI_RAW_STATS) {
            // InstanceNorm statistics from the fp32 accumulators (reference model.py:55,65,94,132)
            float sq[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) { v[i] = pvalid ? v[i] : 0.f; sq[i] = v[i] * v[i]; }
            warp_reduce64(v, lane);
            warp_reduce64(sq, lane);
            float* st = p.stats + ((size_t)n * p.cout_total + col0 + 2 * lane) * 2;
            atomicAdd(st + 0, v[0]);
            atomicAdd(st + 1, sq[0]);
            atomicAdd(st + 2, v[1]);
            atomicAdd(st + 3, sq[1]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

}  // namespace fsr
