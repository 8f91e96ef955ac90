// conv3x3_head.cuh - 3x3 conv with 3 output channels (Generator.head, model.py:102-110, and the 64->3
// image-gradient convs of VGG conv1_1 / Discriminator.neck) as ONE 1x1 GEMM plus a shift-add epilogue.
//
// With N = 3 (padded 16) the implicit-GEMM form is bound by the A-operand smem port: 9 taps x 4 k-steps each
// re-read a 4 KB A tile for a 16-column MMA.  Instead:
//     Z[p, tap*3+c] = X[p, :] . W[c, :, tap]          one GEMM over the HALO tile: M = 180 (2 x 128), N = 32 (27 used), K = 64
//     out[p, c]     = bias[c] + sum_tap Z[p + tap_offset, tap*3+c]     (fp32, in the epilogue through smem)
// A is read 8 times per tile (2 M-halves x 4 k-steps) instead of 36, all products/accumulation stay fp32,
// and the kernel becomes HBM-bound on reading X (its roofline).
#pragma once
#include "conv3x3_tc.cuh"

namespace fsr {

struct HeadCfg {
  static constexpr int TH = 16, TW = 8, BW = 10, BH = 18;
  static constexpr int kHaloRows = BW * BH;                 // 180
  static constexpr int kStageBytes = 23552;                 // 184 rows, 1024-aligned
  static constexpr int kStages = 5;
  static constexpr int kMaxKC = 8;                          // Cin <= 512
  static constexpr int kWBytes = kMaxKC * 32 * 128;         // B tiles: per 64-channel chunk 32 rows (tap*3+c) x 64 ch
  static constexpr int kZPitch = 33;                        // floats per halo pixel (odd -> conflict-free column reads)
  static constexpr int kZBytes = ((kHaloRows * kZPitch * 4 + 1023) / 1024) * 1024;   // 24576
  static constexpr int kEpiWarps = 8;
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
  static constexpr int kTmemCols = 128;                     // 2 accumulator sets x (2 M-halves x 32 cols)
  static constexpr int kSmemBytes = kStages * kStageBytes + kWBytes + 2 * kZBytes + 1024 + 1024 + 12288 /*over-read pad*/;
};

FSR_DEVINL void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
FSR_DEVINL void st_shared_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
FSR_DEVINL float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

template <typename T>
__global__ void __launch_bounds__(HeadCfg::kThreads, 1)
conv3x3_head_kernel(const __grid_constant__ CUtensorMap tm_x, const T* __restrict__ w_packed /*[9][16][cin]*/,
                    const ConvParams p, const int cin) {
  pdl_grid_sync();
  const int KC = cin >> 6;
  using Cfg = HeadCfg;
  constexpr int TH = Cfg::TH, TW = Cfg::TW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;                                              // stages (the M=256 view over-reads 72 rows)
  uint8_t* smem_pad = smem_a + Cfg::kStages * Cfg::kStageBytes;         // 12 KB pad so the last stage's over-read stays inside
  uint8_t* smem_w = smem_pad + 12288;
  uint8_t* smem_z = smem_w + Cfg::kWBytes;                              // 2 x Z buffers (one per epilogue group)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_z + 2 * Cfg::kZBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* smem_bias = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int t_begin = (int)(((long long)blockIdx.x * p.num_tiles) / gridDim.x);
  const int t_end = (int)(((long long)(blockIdx.x + 1) * p.num_tiles) / gridDim.x);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_x);
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  // B tile: row j = tap*3 + c  <-  w_packed[tap][c][0..63]; 128B-swizzled like a TMA write (chunk ^= row & 7)
  for (int i = threadIdx.x; i < KC * 32 * 8; i += blockDim.x) {
    const int kc = i >> 8, row = (i >> 3) & 31, ch = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row < 27) {
      const int tap = row / 3, c = row % 3;
      v = *reinterpret_cast<const uint4*>(w_packed + ((size_t)tap * 16 + c) * cin + kc * 64 + ch * 8);
    }
    *reinterpret_cast<uint4*>(smem_w + kc * 4096 + row * 128 + ((ch ^ (row & 7)) << 4)) = v;
  }
  if (threadIdx.x < 4) smem_bias[threadIdx.x] = (p.bias != nullptr && threadIdx.x < 3) ? p.bias[threadIdx.x] : 0.f;
  fence_proxy_async();        // generic-proxy smem writes (B tile) -> visible to the tensor core (async proxy)
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
      for (int kc = 0; kc < KC; ++kc) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kHaloRows * 128);
          tma_load_4d(smem_a + stage * Cfg::kStageBytes, &tm_x, &full_bar[stage], kc * 64, tx * TW - 1, ty * TH - 1, n);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc = make_idesc_f16(128, 32, std::is_same<T, __nv_bfloat16>::value);
    const uint32_t a_lo0 = desc_lo_sw128(smem_u32(smem_a));
    const uint32_t b_lo0 = desc_lo_sw128(smem_u32(smem_w));
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      for (int kc = 0; kc < KC; ++kc) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + stage * (Cfg::kStageBytes >> 4);
        const uint32_t b_lo = b_lo0 + kc * (4096 >> 4);
        if (elect_one()) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {          // halo rows [0,128) and [128,256) (rows >= 180 are never read back)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adesc = desc_join(a_lo + ((h * 128 * 128 + k * 32) >> 4), kDescHiSw128);
              const uint64_t bdesc = desc_join(b_lo + ((k * 32) >> 4), kDescHiSw128);
              umma_f16(tmem_base + acc * 64 + h * 32, adesc, bdesc, idesc, (k | kc) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (kc == KC - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // =============================== epilogue: Z -> smem, shift-add, tanh, store ===============================
    const int ew = warp - 2;
    const int q = warp & 3;
    const int egroup = ew >> 2;                           // group g serves tiles with (it & 1) == g
    const uint32_t zs = smem_u32(smem_z + egroup * Cfg::kZBytes);
    const int m = q * 32 + lane;                          // output pixel of this thread in the 16x8 tile
    const int yy = m / TW, xx = m % TW;
    int it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const int acc = it & 1;
      if (acc != egroup) continue;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n = t / tiles_per_img;
      const int rem = t - n * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int y = ty * TH + yy, x = tx * TW + xx;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 64;
      tmem_ld32(t_row, r0);                               // halo row m
      tmem_ld32(t_row + 32, r1);                          // halo row 128 + m
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      named_bar_sync(1 + egroup, 128);                    // previous tile's Z reads are done
#pragma unroll
      for (int j = 0; j < 27; ++j) st_shared_f32(zs + (uint32_t)(m * Cfg::kZPitch + j) * 4, __uint_as_float(r0[j]));
      if (128 + m < Cfg::kHaloRows) {
#pragma unroll
        for (int j = 0; j < 27; ++j) st_shared_f32(zs + (uint32_t)((128 + m) * Cfg::kZPitch + j) * 4, __uint_as_float(r1[j]));
      }
      named_bar_sync(1 + egroup, 128);                    // Z complete
      float o[3] = {smem_bias[0], smem_bias[1], smem_bias[2]};
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const uint32_t base = zs + (uint32_t)(((yy + r) * Cfg::BW + xx + s) * Cfg::kZPitch + (r * 3 + s) * 3) * 4;
          o[0] += ld_shared_f32(base);
          o[1] += ld_shared_f32(base + 4);
          o[2] += ld_shared_f32(base + 8);
        }
      if (y < p.H && x < p.W) {
        if (p.out_u8 < 2) {
#pragma unroll
          for (int c = 0; c < 3; ++c) o[c] = tanhf(o[c]);
        }
        if (p.out_u8 == 1) {
          // reference inference.py:54-56: ((y+1)/2*255).astype(uint8)  (truncation)
          uint8_t* o8 = reinterpret_cast<uint8_t*>(p.out) + ((size_t)(n * p.H + y) * p.W + x) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float f = (o[c] + 1.0f) / 2.0f * 255.0f;
            o8[c] = (uint8_t)(int)fminf(fmaxf(f, 0.f), 255.f);
          }
        } else {
          float* of = reinterpret_cast<float*>(p.out);
          const size_t plane = (size_t)p.H * p.W;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float* dst = of + ((size_t)n * 3 + c) * plane + (size_t)y * p.W + x;
            *dst = (p.out_u8 == 3) ? *dst + o[c] : o[c];
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
