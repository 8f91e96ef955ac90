// conv3x3_tc.cuh - implicit-GEMM 3x3 / stride-1 / pad-1 convolution on tcgen05 tensor cores.
//
// Replaces torch.nn.Conv2d(F, F|4F|3, kernel_size=3, padding=1) at reference model.py:30-35,
// 47-54, 57-64, 87-93, 103-108 (and their fused neighbours: InstanceNorm statistics :55,65,94,
// PixelShuffle+PReLU :36-37, Tanh :109).
//
// Layout: activations NHWC, 64 input channels = one 128-byte row per pixel (fp16 or bf16).
// GEMM view: D[pixels, Cout] = sum over 9 taps of  A_tap[pixels, 64] * W_tap[64, Cout].
//   * M tile  = 128 output pixels = one UMMA M.  Two geometries (template HALO1):
//       HALO1 = false: 8 x 16 pixels; A = one TMA 4-D box {64ch, 16, 8+2, 1} per COLUMN shift s
//               (3 smem fills per tile); the 3 ROW shifts r are descriptor offsets of r*16 rows
//               (2048 B = multiple of the 1024-B swizzle atom).
//       HALO1 = true : 16 x 8 pixels; A = ONE TMA box {64ch, 8+2, 16+2, 1} (the halo tile, 180 rows);
//               tap (r,s) = descriptor start at halo row r*10+s with stride-byte-offset 1280 B
//               (one 8-pixel tile row per 8-row core group; groups are 10 halo pixels apart).
//               1 smem fill per tile (2.6x less L2->SM traffic).
//     Zero fill outside the image (TMA OOB) = the conv padding in both cases.
//   * B       = the Cout-slice of the weights, resident in smem for the whole persistent CTA
//               (9 taps x NS rows x 128 B, K-major, 128B swizzle).
//   * D       = fp32 accumulators in TMEM, double buffered (2 x NS columns).
// XF (fused input transform, HALO1 + NS = 64 only): 0 = none; 1 = the input is the RAW output of the block's first conv
// and y = PReLU(InstanceNorm(raw)) (model.py:55-56) is applied to the staged halo tile before the MMAs read it;
// 2 = the input is the RAW conv2 output of the PREVIOUS block: x_next = InstanceNorm(raw) + x_prev (model.py:65+69) is
// formed on the staged tile, and the tile's own pixels of x_next are written back for the next skip connection.
// Warp roles: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc), then 4 or 8 epilogue warps
// (TMEM -> registers -> fused epilogue -> swizzled smem transpose -> coalesced global stores); with 8,
// the two groups of 4 alternate tiles (group g owns accumulator buffer g).
#pragma once
#include "fsr_common.cuh"
#include <type_traits>

namespace fsr {

enum ConvEpilogue : int {
  EPI_RAW_STATS = 0,   // store raw conv output (NHWC) + per-(n,c) sum / sum-of-squares (InstanceNorm stats)
  EPI_BIAS_ACT = 1,    // store act(conv + bias) NHWC
  EPI_PS_PRELU = 2,    // bias + PReLU + PixelShuffle(2) scatter: out[N,2H,2W,64]
  EPI_HEAD_TANH = 3,   // NS=16 (3 real channels): tanh(conv + bias) -> fp32 NCHW or uint8 NHWC
  EPI_F32 = 4,         // precise mode: out fp32 NHWC [N,H,W,cout_total] = (act ? out : 0) + conv  (store | accumulate)
};

enum ActMode : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_PRELU = 3 };

struct ConvParams {
  int N, H, W;              // conv input == output spatial size (stride 1, pad 1)
  int cout_total;           // GEMM N over all slices (multiple of NS)
  int num_slices;           // cout_total / NS
  int tiles_x, tiles_y;     // ceil(W/TW), ceil(H/TH)
  int num_tiles;            // N * tiles_y * tiles_x
  void* out;                // see ConvEpilogue
  const float* bias;        // [cout_total] in GEMM column order, or nullptr
  long long* stats;         // [N][cout_total][2] fixed-point int64 (sum * 2^24, sumsq * 2^20), EPI_RAW_STATS
  const float* alpha;       // device pointer to the PReLU slope (ACT_PRELU)
  float slope;              // LeakyReLU slope (ACT_LRELU)
  int act;                  // ActMode (EPI_BIAS_ACT)
  int ws;                   // 1: weight-stationary MMAs (B re-used from the collector across the two interleaved tiles)
  int pair_rows;            // 1: weights are pair-expanded (pairs.py): tap column 0 only has K in [32,64), column 2 only K in
                            //    [0,32) -> those 12 of 36 k-steps multiply structural zeros and are not issued
  int out_u8;               // EPI_HEAD_TANH: 0 -> tanh, fp32 NCHW [N,3,H,W]; 1 -> tanh, uint8 NHWC [N,H,W,3];
                            //                2 -> linear (no tanh) fp32 NCHW store; 3 -> linear, accumulate (+=)
  // XF (fused input transform): the conv input is the RAW output of the previous conv; its InstanceNorm + PReLU
  // (model.py:55-56) is applied to the staged halo tile in shared memory before the MMAs read it
  const long long* in_stats;  // [N][64][2] fixed-point statistics of the input tensor
  const float* in_alpha;      // PReLU slope (device pointer)                                   (XF == 1)
  float in_eps;
  // XF == 2: the input is the RAW conv2 output of the PREVIOUS residual block; x_next = InstanceNorm(raw) + in_res
  // (model.py:65 + :69) is formed on the staged tile and the tile's own pixels of x_next are written to x_out
  const void* in_res;         // x_prev [N,H,W,64] NHWC T
  void* x_out;                // x_next [N,H,W,64] NHWC T (aliases neither in_res nor the raw input)
  int backoff_ns;             // nanosleep between mbarrier polls of the producer / epilogue / transform warps (0 = spin)
};

template <bool HALO1>
struct ConvGeo {
  static constexpr int TH = HALO1 ? 16 : 8;
  static constexpr int TW = HALO1 ? 8 : 16;
  static constexpr int kBoxW = HALO1 ? TW + 2 : TW;
  static constexpr int kBoxH = TH + 2;
  static constexpr int kLoads = HALO1 ? 1 : 3;                       // TMA fills per tile
  static constexpr int kStageBytes = ((kBoxW * kBoxH * 128 + 1023) / 1024) * 1024;   // 23552 | 20480
  static constexpr int kTxBytes = kBoxW * kBoxH * 128;               // bytes one fill delivers
};

constexpr int kXfWarps = 4;   // input-transform warps of the XF variant (448 threads: 146 registers per thread)

template <int NS, bool HALO1>
struct ConvCfg {
  using Geo = ConvGeo<HALO1>;
  static constexpr int kWBytes = 9 * NS * 128;
  static constexpr int kEpiWarps = (NS >= 128) ? 4 : 8;
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
  static constexpr int kThreadsXf = kThreads + 32 * kXfWarps;
  static constexpr int kStagingBytes = kEpiWarps * 4096;
  static constexpr int kStages = (NS >= 128) ? (HALO1 ? 2 : 3) : (HALO1 ? 5 : 6);
  static constexpr bool kPair = HALO1 && kStages >= 4;   // interleave the MMAs of two tiles (needs both tiles staged)
  static constexpr int kAccs = kPair ? 4 : 2;
  static constexpr int kTmemCols = (kAccs * NS <= 32) ? 32 : (kAccs * NS <= 64 ? 64 : (kAccs * NS <= 128 ? 128 : (kAccs * NS <= 256 ? 256 : 512)));
  static constexpr int kSmemBytes = kWBytes + kStages * Geo::kStageBytes + kStagingBytes + 1024 /*barriers+bias*/ + 1024 /*align*/;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB per-CTA shared memory of sm_100");
};

FSR_DEVINL float apply_act(float v, int act, float slope) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_LRELU || act == ACT_PRELU) return v >= 0.f ? v : v * slope;
  return v;
}

// Sum 64 per-lane values across the 32 lanes of a warp with a halving butterfly (32+16+8+4+2 = 62 shuffles):
// afterwards lane L holds the totals of columns 2L and 2L+1.  Register-only: costs no shared-memory bandwidth
// (the conv is bound by the smem data pipe).
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

template <int NS, int EPI, typename T, bool HALO1, int XF = 0>
__global__ void __launch_bounds__(XF ? ConvCfg<NS, HALO1>::kThreadsXf : ConvCfg<NS, HALO1>::kThreads, 1)
conv3x3_c64_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                   const __grid_constant__ CUtensorMap tm_out, const ConvParams p) {
  pdl_grid_sync();
  // NHWC outputs leave through a TMA store of the staged (swizzled) tile: no smem read-back, hardware edge clipping
  constexpr bool kTmaStore = (EPI == EPI_RAW_STATS || EPI == EPI_BIAS_ACT);
  using Cfg = ConvCfg<NS, HALO1>;
  using Geo = ConvGeo<HALO1>;
  constexpr int TH = Geo::TH, TW = Geo::TW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem_w + Cfg::kWBytes;
  uint8_t* smem_stg = smem_a + Cfg::kStages * Geo::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]
  uint64_t* w_bar = bars + 2 * Cfg::kStages;       // [1]
  uint64_t* tfull_bar = w_bar + 1;                 // [4]
  uint64_t* tempty_bar = tfull_bar + 4;            // [4]
  uint64_t* xfull_bar = tempty_bar + 4;            // [kStages] (XF: the stage's halo tile has been transformed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xfull_bar + Cfg::kStages);
  float* smem_bias = reinterpret_cast<float*>(tmem_slot + 4);   // [NS]
  static_assert(!XF || (HALO1 && NS == 64), "the fused input transform exists for the single-halo-tile 64->64 conv");
  // the MMA warp consumes a stage once it is `ready`: filled by TMA, and in the XF variant transformed in place
  uint64_t* ready_bar = XF ? xfull_bar : full_bar;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int slice = blockIdx.x % p.num_slices;
  const int cta_in_slice = blockIdx.x / p.num_slices;
  const int ctas_per_slice = gridDim.x / p.num_slices;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  // contiguous tile range per CTA: consecutive tiles share halos (L2 locality) and mostly one image
  // (InstanceNorm statistics stay in registers across tiles, see the epilogue)
  const int t_begin = (int)(((long long)cta_in_slice * p.num_tiles) / ctas_per_slice);
  const int t_end = (int)(((long long)(cta_in_slice + 1) * p.num_tiles) / ctas_per_slice);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); mbar_init(&xfull_bar[i], kXfWarps);
    }
    mbar_init(w_bar, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  // accumulator ring: HALO1 issues the MMAs of TWO tiles interleaved (two independent accumulation chains keep the
  // tensor pipe busy while each chain waits on its own previous MMA) -> 4 accumulators, tile `it` uses
  // ((it>>1)&1)*2 + (it&1), reused every 4 tiles; HALO0 keeps the plain double buffer.
  auto acc_of = [](int it) { return Cfg::kPair ? (((it >> 1) & 1) * 2 + (it & 1)) : (it & 1); };
  auto phase_of = [](int it) { return (uint32_t)(Cfg::kPair ? ((it >> 2) & 1) : ((it >> 1) & 1)); };
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
    // =============================== TMA producer (whole warp runs the loop, one elected lane issues)
    if (elect_one()) {
      mbar_arrive_expect_tx(w_bar, Cfg::kWBytes);
      for (int tap = 0; tap < 9; ++tap)
        tma_load_2d(smem_w + tap * NS * 128, &tm_w, w_bar, 0, tap * p.cout_total + slice * NS);
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int n = t / tiles_per_img;
      const int rem = t - n * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int x0 = tx * TW, y0 = ty * TH;
#pragma unroll
      for (int s = 0; s < Geo::kLoads; ++s) {
        mbar_wait_backoff(&empty_bar[stage], phase ^ 1, p.backoff_ns);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], Geo::kTxBytes);
          tma_load_4d(smem_a + stage * Geo::kStageBytes, &tm_x, &full_bar[stage], 0,
                      HALO1 ? x0 - 1 : x0 + s - 1, y0 - 1, n);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer: warp-uniform control flow so that descriptors live in
    // uniform registers; only the tcgen05 instructions are predicated on the elected lane.
    constexpr uint32_t idesc = make_idesc_f16(128, NS, std::is_same<T, __nv_bfloat16>::value);
    const uint32_t a_lo0 = desc_lo_sw128(smem_u32(smem_a));
    const uint32_t b_lo0 = desc_lo_sw128(smem_u32(smem_w));
    mbar_wait(w_bar, 0);
    tc_fence_after();
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    if constexpr (HALO1) {
      constexpr int kStep = Cfg::kPair ? 2 : 1;
      for (int t = t_begin; t < t_end; t += kStep, it += kStep) {
        const bool two = Cfg::kPair && (t + 1 < t_end);
        const int acc_a = acc_of(it), acc_b = Cfg::kPair ? acc_of(it + 1) : acc_a;
        const uint32_t ph = phase_of(it);
        mbar_wait(&tempty_bar[acc_a], ph ^ 1);
        if (two) mbar_wait(&tempty_bar[acc_b], ph ^ 1);
        tc_fence_after();
        const int stage_a = stage;
        const uint32_t phase_a = phase;
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        const int stage_b = stage;
        const uint32_t phase_b = phase;
        if (two) { if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; } }
        mbar_wait(&ready_bar[stage_a], phase_a);
        if (two) mbar_wait(&ready_bar[stage_b], phase_b);
        tc_fence_after();
        const uint32_t a_lo_a = a_lo0 + stage_a * (Geo::kStageBytes >> 4);
        const uint32_t a_lo_b = a_lo0 + stage_b * (Geo::kStageBytes >> 4);
        const uint32_t d_a = tmem_base + acc_a * NS, d_b = tmem_base + acc_b * NS;
        if (elect_one()) {
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              // halo row r*10+s: the start address is only 128-B aligned and core groups are 1280 B apart.
              // The 128B swizzle XOR is a function of the absolute smem address bits [7,10) (measured on
              // B200: SBO = 1280 reads TMA-written data correctly), so the base-offset field stays 0.
              constexpr uint32_t kSbo = (uint32_t)((TW + 2) * 128) >> 4;     // 1280 B between core groups
              const uint32_t hi = kSbo | (1u << 14) | (2u << 29);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t aoff = (uint32_t)(((r * (TW + 2) + s) * 128 + k * 32) >> 4);
                const uint64_t bdesc = desc_join(b_lo0 + (((r * 3 + s) * (NS * 128) + k * 32) >> 4), kDescHiSw128);
                if (p.pair_rows && ((s == 0 && k < 2) || (s == 2 && k >= 2))) continue;   // structural zeros
                // first issued k-step overwrites the accumulator: (r,s,k) = (0,0,0), or (0,0,2) on pair rows
                const uint32_t accf = (r | s) != 0 ? 1u : (p.pair_rows ? (k > 2 ? 1u : 0u) : (k != 0 ? 1u : 0u));
                if (two && p.ws) {
                  umma_f16_ws_fill(d_a, desc_join(a_lo_a + aoff, hi), bdesc, idesc, accf);
                  umma_f16_ws_lastuse(d_b, desc_join(a_lo_b + aoff, hi), bdesc, idesc, accf);
                } else {
                  umma_f16(d_a, desc_join(a_lo_a + aoff, hi), bdesc, idesc, accf);
                  if (two) umma_f16(d_b, desc_join(a_lo_b + aoff, hi), bdesc, idesc, accf);
                }
              }
            }
          }
          umma_commit(&empty_bar[stage_a]);
          if (two) umma_commit(&empty_bar[stage_b]);
          umma_commit(&tfull_bar[acc_a]);
          if (two) umma_commit(&tfull_bar[acc_b]);
        }
        __syncwarp();
      }
    } else {
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int acc = acc_of(it);
        const uint32_t acc_phase = phase_of(it);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * NS;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (Geo::kStageBytes >> 4);
          if (elect_one()) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t adesc = desc_join(a_lo + ((r * (TW * 128) + k * 32) >> 4), kDescHiSw128);
                const uint64_t bdesc = desc_join(b_lo0 + (((r * 3 + s) * (NS * 128) + k * 32) >> 4), kDescHiSw128);
                umma_f16(d_tmem, adesc, bdesc, idesc, (s | r | k) != 0 ? 1u : 0u);
              }
            }
            umma_commit(&empty_bar[stage]);
            if (s == 2) umma_commit(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (XF && warp >= 2 + Cfg::kEpiWarps) {
    // =============================== input-transform warps (XF) ===============================
    // y = PReLU((raw - mean[n,c]) * rstd[n,c]) applied IN PLACE to the TMA-written halo tile (rows = 180 halo pixels of
    // 128 B; the 16-byte chunk of channel group g sits at chunk g ^ (row & 7): 128B swizzle on absolute address bits,
    // stage bases are 1024-B aligned).  Same fp32 operations as instnorm_apply_kernel -> bit-identical activations.
    // Rows outside the image stay zero (the conv's zero padding comes AFTER the normalisation).
    if constexpr (XF == 1) {
      const int tid = threadIdx.x - (64 + 32 * Cfg::kEpiWarps);      // 0 .. 32*kXfWarps-1
      const int g = tid & 7, r_first = tid >> 3;
      constexpr int kRowStep = 4 * kXfWarps;
      const float slope = __ldg(p.in_alpha);
      const double inv_hw = 1.0 / (double)(p.H * p.W);
      float mean[8], rstd[8];
      int cur_n = -1;
      int stage = 0; uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int n = t / tiles_per_img;
        const int rem = t - n * tiles_per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int x0 = tx * TW, y0 = ty * TH;
        if (n != cur_n) {
          // the four lanes that share channel group g (lane & 7) decode two channels each (fp64 divide + sqrt: ~0.4 us
          // apiece, serial) and exchange them - the small training planes change image every other tile
          float m2[2], r2[2];
#pragma unroll
          for (int j = 0; j < 2; ++j)
            stat_mean_rstd(p.in_stats + ((size_t)n * 64 + 8 * g + 2 * (lane >> 3) + j) * 2, inv_hw, p.in_eps, m2[j], r2[j]);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            mean[k] = __shfl_sync(0xffffffffu, m2[k & 1], (k >> 1) * 8 + g);
            rstd[k] = __shfl_sync(0xffffffffu, r2[k & 1], (k >> 1) * 8 + g);
          }
          cur_n = n;
        }
        const bool inside = (y0 >= 1) && (x0 >= 1) && (y0 + TH + 1 <= p.H) && (x0 + TW + 1 <= p.W);   // whole halo box in the image
        mbar_wait_backoff(&full_bar[stage], phase, p.backoff_ns);
        const uint32_t base = smem_u32(smem_a + stage * Geo::kStageBytes);
        // ROLLED loop (two rows in flight per thread): a fully unrolled, branch-free version measured no faster than two
        // warps with a branch per row (358 us either way) while the kernel's SASS grew to 64 KB and 27 % of the stall samples
        // became `no_instructions` - the transform's code footprint, not its arithmetic, is what has to stay small.
        constexpr int kRows = Geo::kBoxW * Geo::kBoxH;                    // 180
#pragma unroll 2
        for (int r = r_first; r < kRows; r += kRowStep) {
          const uint32_t addr = base + (uint32_t)r * 128u + (uint32_t)((g ^ (r & 7)) << 4);
          const uint4 v = ld_shared_v4(addr);
          bool ok = inside;
          if (!inside) {
            const int by = r / Geo::kBoxW, bx = r - by * Geo::kBoxW;
            const int gy = y0 - 1 + by, gx = x0 - 1 + bx;
            ok = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
          }
          const uint32_t vu[4] = {v.x, v.y, v.z, v.w};
          uint32_t ou[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = Cvt<T>::unpack2(vu[k]);
            float a = (f.x - mean[2 * k]) * rstd[2 * k];
            float b = (f.y - mean[2 * k + 1]) * rstd[2 * k + 1];
            a = apply_act(a, ACT_PRELU, slope);
            b = apply_act(b, ACT_PRELU, slope);
            ou[k] = ok ? Cvt<T>::pack2(a, b) : vu[k];                     // rows outside the image stay zero
          }
          st_shared_v4(addr, ou[0], ou[1], ou[2], ou[3]);
        }
        fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&xfull_bar[stage]);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
    // =============================== input-transform warps (XF2) ===============================
    // y = (raw - mean[n,c]) * rstd[n,c] + residual  (bn2 + skip, model.py:65+69 / bottleneck :94 fed by the chain) applied
    // IN PLACE to the TMA-written halo tile of the RAW conv2 output; the residual x comes straight from global memory
    // (16-byte vectors, 1.4x halo overhead, L2 hits for the halo) and the tile's OWN 16x8 pixels of y are written back to
    // x_out for the next block's skip connection.  Same fp32 operations as instnorm_apply_kernel (normalise, no
    // activation, add the residual as float, round once) -> bit-identical activations.
    if constexpr (XF == 2) {
      const int tid = threadIdx.x - (64 + 32 * Cfg::kEpiWarps);      // 0 .. 32*kXfWarps-1
      const int g = tid & 7, r_first = tid >> 3;
      constexpr int kRowStep = 4 * kXfWarps;
      const double inv_hw = 1.0 / (double)(p.H * p.W);
      float mean[8], rstd[8];
      int cur_n = -1;
      int stage = 0; uint32_t phase = 0;
      const T* res = reinterpret_cast<const T*>(p.in_res);
      T* xo = reinterpret_cast<T*>(p.x_out);
      for (int t = t_begin; t < t_end; ++t) {
        const int n = t / tiles_per_img;
        const int rem = t - n * tiles_per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int x0 = tx * TW, y0 = ty * TH;
        if (n != cur_n) {
          // the four lanes that share channel group g (lane & 7) decode two channels each (fp64 divide + sqrt: ~0.4 us
          // apiece, serial) and exchange them - the small training planes change image every other tile
          float m2[2], r2[2];
#pragma unroll
          for (int j = 0; j < 2; ++j)
            stat_mean_rstd(p.in_stats + ((size_t)n * 64 + 8 * g + 2 * (lane >> 3) + j) * 2, inv_hw, p.in_eps, m2[j], r2[j]);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            mean[k] = __shfl_sync(0xffffffffu, m2[k & 1], (k >> 1) * 8 + g);
            rstd[k] = __shfl_sync(0xffffffffu, r2[k & 1], (k >> 1) * 8 + g);
          }
          cur_n = n;
        }
        constexpr int kRows = Geo::kBoxW * Geo::kBoxH;                    // 180
        // coordinates of the NEXT tile: its residual rows are prefetched into L2 one tile (~1.5 us) ahead, so that the
        // loads of the rolled loop below are L2 hits (x_prev was written two launches ago: 236 MB, not L2 resident)
        const bool has_next = t + 1 < t_end;
        const int tn = has_next ? t + 1 : t;
        const int nn = tn / tiles_per_img;
        const int remn = tn - nn * tiles_per_img;
        const int tyn = remn / p.tiles_x, txn = remn - tyn * p.tiles_x;
        const int x0n = txn * TW, y0n = tyn * TH;
        mbar_wait(&full_bar[stage], phase);
        const uint32_t base = smem_u32(smem_a + stage * Geo::kStageBytes);
#pragma unroll 4
        for (int r = r_first; r < kRows; r += kRowStep) {
          const int by = r / Geo::kBoxW, bx = r - by * Geo::kBoxW;
          const int gy = y0 - 1 + by, gx = x0 - 1 + bx;
          const bool ok = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
          const size_t goff = (((size_t)n * p.H + (ok ? gy : 0)) * p.W + (ok ? gx : 0)) * 64 + 8 * g;
          const uint4 rv = *reinterpret_cast<const uint4*>(res + goff);
          if (has_next && (g & 1) == 0) {          // one prefetch per 32-byte sector of the next tile's row
            const int gyn = y0n - 1 + by, gxn = x0n - 1 + bx;
            if (gyn >= 0 && gyn < p.H && gxn >= 0 && gxn < p.W)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(res + (((size_t)nn * p.H + gyn) * p.W + gxn) * 64 + 8 * g));
          }
          const uint32_t addr = base + (uint32_t)r * 128u + (uint32_t)((g ^ (r & 7)) << 4);
          const uint4 v = ld_shared_v4(addr);
          const uint32_t vu[4] = {v.x, v.y, v.z, v.w};
          const uint32_t su[4] = {rv.x, rv.y, rv.z, rv.w};
          uint32_t ou[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = Cvt<T>::unpack2(vu[k]);
            const float2 s2 = Cvt<T>::unpack2(su[k]);
            float a = (f.x - mean[2 * k]) * rstd[2 * k];
            float b = (f.y - mean[2 * k + 1]) * rstd[2 * k + 1];
            a += s2.x;
            b += s2.y;
            ou[k] = ok ? Cvt<T>::pack2(a, b) : vu[k];                     // rows outside the image stay zero
          }
          st_shared_v4(addr, ou[0], ou[1], ou[2], ou[3]);
          // write-back of the tile's own pixels (each image pixel is interior to exactly one tile)
          if (ok && by >= 1 && by <= TH && bx >= 1 && bx <= TW)
            *reinterpret_cast<uint4*>(xo + goff) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
        }
        fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&xfull_bar[stage]);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // =============================== epilogue warps ===============================
    const int ew = warp - 2;
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int egroup = ew >> 2;                  // with 8 warps: group g serves tiles with (it & 1) == g
    const uint32_t stg = smem_u32(smem_stg + ew * 4096);   // warp-private 32 x 128 B transpose buffer
    const int m = q * 32 + lane;                 // accumulator row == pixel within the tile
    const int yy = m / TW, xx = m % TW;
    float prelu_a = 0.f;
    if (EPI == EPI_PS_PRELU || (EPI == EPI_BIAS_ACT && p.act == ACT_PRELU)) prelu_a = __ldg(p.alpha);
    const float slope = (EPI == EPI_PS_PRELU || p.act == ACT_PRELU) ? prelu_a : p.slope;
    // InstanceNorm statistics of channels (2*lane, 2*lane+1) of the current image, carried across tiles
    // (each TILE's partial sums are converted to fixed point before they are added up: the totals are independent of
    //  how tiles are distributed over warps / CTAs / launches -> batch-size and run-to-run invariant)
    long long st_s0 = 0, st_q0 = 0, st_s1 = 0, st_q1 = 0;
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
      if (Cfg::kEpiWarps == 8 && (it & 1) != egroup) continue;
      const int acc = acc_of(it);
      const uint32_t acc_phase = phase_of(it);
      const int n = t / tiles_per_img;
      const int rem = t - n * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int x0 = tx * TW, y0 = ty * TH;
      const int y = y0 + yy, x = x0 + xx;
      const bool pvalid = (y < p.H) && (x < p.W);
      const bool interior = (y0 + TH <= p.H) && (x0 + TW <= p.W);   // warp-uniform
      if (EPI == EPI_RAW_STATS && n != st_n) { flush_stats(st_n); st_n = n; }
      mbar_wait_backoff(&tfull_bar[acc], acc_phase, p.backoff_ns);
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
          // p.act == 2: "pixel-pair rows" (n_filters = 32 networks, engine docs DESIGN.md 3.7): one 128-byte row holds
          // two horizontally adjacent 32-channel pixels, the GEMM has 6 real columns (pixel parity, rgb) and the image
          // is 2*p.W wide
          const int npx = p.act == 2 ? 2 : 1;
          const int Wt = p.W * npx;
#pragma unroll
          for (int po = 0; po < 2; ++po) {
            if (po >= npx) break;
            const int xt = x * npx + po;
            float o[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float v = __uint_as_float(r[po * 3 + c]) + smem_bias[po * 3 + c];
              o[c] = p.out_u8 >= 2 ? v : tanhf(v);
            }
            if (p.out_u8 == 1) {
              // reference inference.py:54-56: ((y+1)/2*255).astype(uint8)  (truncation)
              uint8_t* o8 = reinterpret_cast<uint8_t*>(p.out) + ((size_t)(n * p.H + y) * Wt + xt) * 3;
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                float f = (o[c] + 1.0f) / 2.0f * 255.0f;
                o8[c] = (uint8_t)(int)fminf(fmaxf(f, 0.f), 255.f);
              }
            } else {
              float* of = reinterpret_cast<float*>(p.out);
              const size_t plane = (size_t)p.H * Wt;
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                float* dst = of + ((size_t)n * 3 + c) * plane + (size_t)y * Wt + xt;
                *dst = (p.out_u8 == 3) ? *dst + o[c] : o[c];
              }
            }
          }
        }
      } else {
#pragma unroll 1
        for (int chunk = 0; chunk < NS / 64; ++chunk) {
          uint32_t pk[32];                        // 64 output values packed to 2-byte pairs
          {
            uint32_t r0[32], r1[32];
            tmem_ld32(t_row + chunk * 64, r0);
            tmem_ld32(t_row + chunk * 64 + 32, r1);
            tmem_ld_wait();
            if (chunk == NS / 64 - 1) {           // all TMEM reads of this accumulator are done
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            }
            if constexpr (EPI == EPI_F32) {
              // precise mode (precise.cuh): the fp32 accumulators go to HBM as they are; the second and third operand
              // products of a split conv (a_lo*w_hi, a_hi*w_lo) are added onto the first (p.act = 1)
              if (pvalid) {
                float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) +
                                                        ((size_t)(n * p.H + y) * p.W + x) * p.cout_total + slice * NS + chunk * 64);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const uint32_t* r = i < 8 ? r0 : r1;
                  const int j = (i & 7) * 4;
                  float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                  if (p.act) {
                    const float4 o = dst[i];
                    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                  }
                  dst[i] = v;
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 16 && EPI != EPI_F32; ++i) {
              float a0 = __uint_as_float(r0[2 * i]), a1 = __uint_as_float(r0[2 * i + 1]);
              float b0 = __uint_as_float(r1[2 * i]), b1 = __uint_as_float(r1[2 * i + 1]);
              if constexpr (EPI != EPI_RAW_STATS) {
                a0 += smem_bias[chunk * 64 + 2 * i];      a1 += smem_bias[chunk * 64 + 2 * i + 1];
                b0 += smem_bias[chunk * 64 + 32 + 2 * i]; b1 += smem_bias[chunk * 64 + 32 + 2 * i + 1];
                if constexpr (EPI == EPI_PS_PRELU) {
                  a0 = a0 >= 0.f ? a0 : a0 * slope; a1 = a1 >= 0.f ? a1 : a1 * slope;
                  b0 = b0 >= 0.f ? b0 : b0 * slope; b1 = b1 >= 0.f ? b1 : b1 * slope;
                } else {
                  a0 = apply_act(a0, p.act, slope); a1 = apply_act(a1, p.act, slope);
                  b0 = apply_act(b0, p.act, slope); b1 = apply_act(b1, p.act, slope);
                }
              }
              pk[i] = Cvt<T>::pack2(a0, a1);
              pk[16 + i] = Cvt<T>::pack2(b0, b1);
            }
          }
          if constexpr (EPI == EPI_F32) continue;
          const int col0 = slice * NS + chunk * 64;   // first GEMM column of this chunk

          // ---- registers -> swizzled smem (row = pixel, 8 x 16B chunks) -> global
          if constexpr (kTmaStore) {
            if (lane == 0) tma_store_wait_read();       // the previous TMA store has finished reading this buffer
          }
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 8; ++k)
            st_shared_v4(stg + lane * 128 + ((k ^ (lane & 7)) << 4), pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
          if constexpr (kTmaStore) {
            fence_proxy_async();                         // generic-proxy writes -> visible to the TMA (async proxy)
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tm_out, smem_stg + ew * 4096, col0, x0, y0 + (q * 32) / TW, n);
              tma_store_commit();
            }
          } else {
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
              if (interior || (py < p.H && px < p.W)) {
                T* dst;
                if constexpr (EPI == EPI_PS_PRELU) {
                  const int qq = col0 >> 6;               // GEMM column block = 2*i + j
                  const int oy = 2 * py + (qq >> 1), ox = 2 * px + (qq & 1);
                  dst = reinterpret_cast<T*>(p.out) + ((size_t)(n * 2 * p.H + oy) * (2 * p.W) + ox) * 64;
                } else {
                  dst = reinterpret_cast<T*>(p.out) + ((size_t)(n * p.H + py) * p.W + px) * p.cout_total + col0;
                }
                *reinterpret_cast<uint4*>(dst + (lane & 7) * 8) = val[j];
              }
            }
          }

          if constexpr (EPI == EPI_RAW_STATS) {
            // InstanceNorm statistics (reference model.py:55,65,94,132) of the STORED (rounded) values, reduced in
            // registers: a butterfly over the warp's 32 pixels leaves channels (2L, 2L+1) in lane L.
            float v[64];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float2 f = Cvt<T>::unpack2(pk[i]);
              v[2 * i] = pvalid ? f.x : 0.f;
              v[2 * i + 1] = pvalid ? f.y : 0.f;
            }
            warp_reduce64(v, lane);
            const float sum0 = v[0], sum1 = v[1];
            if constexpr (XF) asm volatile("" ::: "memory");   // XF runs at 146 registers: keep the two butterflies apart
            float sq[64];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float2 f = Cvt<T>::unpack2(pk[i]);
              sq[2 * i] = pvalid ? f.x * f.x : 0.f;
              sq[2 * i + 1] = pvalid ? f.y * f.y : 0.f;
            }
            warp_reduce64(sq, lane);
            v[0] = sum0; v[1] = sum1;
            st_s0 += stat_fix(v[0], kStatSumScale); st_q0 += stat_fix(sq[0], kStatSqScale);
            st_s1 += stat_fix(v[1], kStatSumScale); st_q1 += stat_fix(sq[1], kStatSqScale);
          }
        }
      }
    }
    if (kTmaStore && lane == 0) tma_store_wait_all();
    if (EPI == EPI_RAW_STATS) flush_stats(st_n);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

}  // namespace fsr
