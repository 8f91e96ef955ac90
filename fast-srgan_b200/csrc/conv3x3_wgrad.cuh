// conv3x3_wgrad.cuh - weight gradient of a 3x3 convolution on tcgen05 tensor cores.
//
// Replaces autograd's `convolution_backward` (weight gradient) issued by trainer.py:180,195 for every
// Conv2d of model.py:47-64, 87-93, 30-35 (Generator) and :124-131 (Discriminator).
//
//   dW[co, ci, r, s] = sum over pixels  dY[n, y, x, co] * X[n, s_*y + r - 1, s_*x + s - 1, ci]
//
// GEMM view per tap: D_tap[ci, co] = sum_px  Xshift_tap[px, ci]^T * dY[px, co]   (K = pixels).
// Both operands sit in smem as [pixel rows][64 channels = 128 B] tiles (TMA, 128B swizzle), i.e. they
// are "MN-major" UMMA operands (the contraction index = the row index).  Two taps are stacked along M
// (M = 128 = 2 taps x 64 input channels): the second 64-row block of the A descriptor starts LBO bytes
// after the first = the distance between the two shifted views of the same halo tile.  9 taps = 5 pairs
// (the last pair duplicates tap 8) -> 5 accumulators [128 x 64] fp32 in TMEM (320 columns) that live for
// the CTA's whole pixel range (split-K over the CTAs of a (cin chunk, cout slice) pair).
// Out-of-image pixels need no masking: TMA zero-fills both operands.
//
// Two stages, no atomics (round 2): every CTA stores its partial [5][128][64] fp32 tile to a workspace slot with
// coalesced 256-byte rows; wgrad_reduce_kernel then sums the partials of each pair IN FIXED ORDER and adds the
// result into the fp32 OIHW gradient through a shared-memory transpose (576 contiguous floats per output channel).
// Round 1's epilogue issued 40960 scattered fp32 atomicAdds per CTA: ~50 us per launch whatever the layer size
// (the 17 generator layers: 0.17 GFLOP each, 50 us each) and run-to-run different sums.  The weight gradient is now
// bitwise reproducible.
// GROUPED form: `groups` independent problems of identical shape (the generator's 17 64->64 convs) run as ONE
// launch - their activations / output gradients sit in two arenas, addressed through the outermost (5th) tensor-map
// dimension, and the work item becomes (group, pair, tile range).
#pragma once
#include "conv3x3_tc.cuh"

namespace fsr {

struct WgradParams {
  int N, Ho, Wo;            // dY spatial size (tiles are over dY pixels)
  int cin, cout;            // channels of X / dY (multiples of 64)
  int tiles_x, tiles_y, num_tiles;
  int nplanes;              // 1 (stride 1) or 4 (stride 2, X in parity planes)
  int box_w, box_rows;      // X box pitch / rows per plane box
  int plane_bytes;          // smem bytes reserved per plane box (1024-aligned)
  int dx, dy;               // X box origin relative to the tile
  int tap_row[10];          // absolute first row (pixel) of tap t's view inside the stage's X region; [9] = dup of [8]
  int tap_id[10];           // r*3+s of the tap
  float* partial;           // workspace: [gridDim.x][5][128][64] fp32, one slot per CTA
  int groups;               // independent problems stacked along the 5th tensor-map dimension
  int ps_perm;              // dY columns are in pixel-shuffle-permuted order (UpSamplingBlock convs)
};

constexpr int kWgradMaxGroups = 40;
struct WgradReduceParams {
  const float* partial;     // [grid][5][128][64]
  float* dw[kWgradMaxGroups];   // per group: fp32 OIHW [cout][cin][3][3], accumulated (+=)
  int cin, cout;
  int npairs_all;           // groups * (cin/64) * (cout/64): partial slot of split c of pair q = q + c * npairs_all
  int ctas_per_pair;
  int slot_of_tap[9];       // accumulator slot (2*pr + half) holding tap t
  int ps_perm;
};

struct WgradCfg {
  static constexpr int kXBytes = 4 * 20480;          // up to 4 parity boxes of 9x17 rows, or one 10x18 box (23552)
  static constexpr int kDyBytes = 16384;             // 128 pixels x 128 B
  static constexpr int kStageBytes = kXBytes + kDyBytes;   // 98304
  static constexpr int kStages = 2;
  static constexpr int kThreads = 64 + 128;
  static constexpr int kTmemCols = 512;              // 5 x 64 used
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 1024;
};

template <typename T>
__global__ void __launch_bounds__(WgradCfg::kThreads, 1)
conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tm_x0, const __grid_constant__ CUtensorMap tm_x1,
                     const __grid_constant__ CUtensorMap tm_x2, const __grid_constant__ CUtensorMap tm_x3,
                     const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ WgradParams p) {
  pdl_grid_sync();
  using Cfg = WgradCfg;
  constexpr int TH = 16, TW = 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* done_bar = bars + 2 * Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // work item = (input-channel chunk kc, output-channel slice) pair; the CTAs of a pair split the pixel tiles
  const int KC = p.cin >> 6, NSL = p.cout >> 6;
  const int npairs = KC * NSL;
  const int npairs_all = npairs * p.groups;
  const int pair_all = blockIdx.x % npairs_all;
  const int grp = pair_all / npairs, pair = pair_all % npairs;
  const int cta_in_pair = blockIdx.x / npairs_all;
  const int ctas_per_pair = gridDim.x / npairs_all;
  const int kc = pair / NSL, sl = pair % NSL;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int t_begin = (int)(((long long)cta_in_pair * p.num_tiles) / ctas_per_pair);
  const int t_end = (int)(((long long)(cta_in_pair + 1) * p.num_tiles) / ctas_per_pair);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_x0);
    tma_prefetch_desc(&tm_dy);
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
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
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* sx = smem + stage * Cfg::kStageBytes;
        mbar_arrive_expect_tx(&full_bar[stage], p.nplanes * p.box_rows * 128 + Cfg::kDyBytes);
        for (int pl = 0; pl < p.nplanes; ++pl) {
          const CUtensorMap* tm = pl == 0 ? &tm_x0 : pl == 1 ? &tm_x1 : pl == 2 ? &tm_x2 : &tm_x3;
          tma_load_5d(sx + pl * p.plane_bytes, tm, &full_bar[stage], kc * 64, x0 + p.dx, y0 + p.dy, n, grp);
        }
        tma_load_5d(sx + Cfg::kXBytes, &tm_dy, &full_bar[stage], sl * 64, x0, y0, n, grp);
      }
      __syncwarp();
      if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // both operands MN-major (bit 15 / bit 16 of the instruction descriptor)
    constexpr uint32_t idesc = make_idesc_f16(128, 64, std::is_same<T, __nv_bfloat16>::value) | (1u << 15) | (1u << 16);
    const uint32_t base_lo = (smem_u32(smem) & 0x3FFFF) >> 4;
    const uint32_t a_hi = ((uint32_t)(p.box_w * 128) >> 4) | (1u << 14) | (2u << 29);   // SBO = one 8-pixel K group
    constexpr uint32_t b_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    int stage = 0; uint32_t phase = 0;
    uint32_t accumulate = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t x_lo = base_lo + stage * (Cfg::kStageBytes >> 4);
      const uint32_t dy_lo = x_lo + (Cfg::kXBytes >> 4);
      if (elect_one()) {
#pragma unroll 1
        for (int pr = 0; pr < 5; ++pr) {
          const int r0 = p.tap_row[2 * pr], r1 = p.tap_row[2 * pr + 1];
          const uint32_t lbo = (uint32_t)((r1 - r0) * 128) >> 4;     // second 64-row M block = the other tap's view
          const uint32_t a0 = x_lo + ((uint32_t)(r0 * 128) >> 4);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {       // 16 pixels = two 8-pixel tile rows per MMA
            const uint64_t adesc = desc_join((a0 + ((uint32_t)(kk * 2 * p.box_w * 128) >> 4)) | (lbo << 16), a_hi);
            const uint64_t bdesc = desc_join((dy_lo + ((uint32_t)(kk * 2048) >> 4)) | (1u << 16), b_hi);
            umma_f16(tmem_base + pr * 64, adesc, bdesc, idesc, (accumulate | (uint32_t)kk) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);
      }
      accumulate = 1;
      __syncwarp();
      if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else {
    // =============================== epilogue: TMEM -> this CTA's partial slot (coalesced rows, no atomics)
    const int q = warp & 3;
    const int m = q * 32 + lane;          // accumulator row: (tap within pair) * 64 + local input channel
    float4* dst = reinterpret_cast<float4*>(p.partial + ((size_t)blockIdx.x * 5 * 128 + m) * 64);
    if (t_end > t_begin) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int pr = 0; pr < 5; ++pr) {
      uint32_t r0[32], r1[32];
      if (t_end > t_begin) {
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + pr * 64;
        tmem_ld32(t_row, r0);
        tmem_ld32(t_row + 32, r1);
        tmem_ld_wait();
      } else {                             // more CTAs than tiles: this slot still takes part in the fixed-order sum
#pragma unroll
        for (int j = 0; j < 32; ++j) { r0[j] = 0u; r1[j] = 0u; }
      }
      float4* d = dst + (size_t)pr * 128 * 16;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        d[j] = make_float4(__uint_as_float(r0[4 * j]), __uint_as_float(r0[4 * j + 1]), __uint_as_float(r0[4 * j + 2]), __uint_as_float(r0[4 * j + 3]));
        d[8 + j] = make_float4(__uint_as_float(r1[4 * j]), __uint_as_float(r1[4 * j + 1]), __uint_as_float(r1[4 * j + 2]), __uint_as_float(r1[4 * j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

// Stage 2: dw[group][co][ci][tap] += sum over the pair's CTAs (fixed order) of their partial tiles.
// grid (npairs_all, 8, 9): block = one (group, cin chunk, cout slice) pair x 8 output channels x one tap.  256 threads =
// 64 input channels x 4 split slices: slice s sums partials s, s+4, s+8, ... four at a time (independent loads in
// flight: a 1-pair layer has up to 148 partials), the four slice sums are then added in fixed order through smem.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const __grid_constant__ WgradReduceParams p) {
  pdl_grid_sync();
  __shared__ float part[4][64][9];
  const int pair_all = blockIdx.x, co0 = blockIdx.y * 8, tap = blockIdx.z;
  const int KC = p.cin >> 6, NSL = p.cout >> 6;
  const int npairs = KC * NSL;
  const int grp = pair_all / npairs, pair = pair_all % npairs;
  const int kc = pair / NSL, sl = pair % NSL;
  const int ci = threadIdx.x & 63, s4 = threadIdx.x >> 6;
  const size_t slot_stride = (size_t)5 * 128 * 64;
  const int slot = p.slot_of_tap[tap];
  const float* src = p.partial + (size_t)pair_all * slot_stride + (((size_t)(slot >> 1)) * 128 + (slot & 1) * 64 + ci) * 64 + co0;
  const size_t cstride = (size_t)p.npairs_all * slot_stride;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int c = s4;
  for (; c + 12 < p.ctas_per_pair; c += 16) {          // 4 partials per trip: 8 independent 16-byte loads in flight
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4* q = reinterpret_cast<const float4*>(src + (size_t)(c + 4 * u) * cstride);
      v[2 * u] = q[0]; v[2 * u + 1] = q[1];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {                        // fixed order: c, c+4, c+8, c+12
      acc[0] += v[2 * u].x; acc[1] += v[2 * u].y; acc[2] += v[2 * u].z; acc[3] += v[2 * u].w;
      acc[4] += v[2 * u + 1].x; acc[5] += v[2 * u + 1].y; acc[6] += v[2 * u + 1].z; acc[7] += v[2 * u + 1].w;
    }
  }
  for (; c < p.ctas_per_pair; c += 4) {
    const float4* q = reinterpret_cast<const float4*>(src + (size_t)c * cstride);
    const float4 a = q[0], b = q[1];
    acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) part[s4][ci][j] = acc[j];
  __syncthreads();
  // 512 results (64 ci x 8 co) by 256 threads: two each; dw[(co * cin + ci) * 9 + tap] is owned by exactly one thread
  float* dw = p.dw[grp];
  for (int idx = threadIdx.x; idx < 512; idx += 256) {
    const int j = idx >> 6, c2 = idx & 63;
    const float r = ((part[0][c2][j] + part[1][c2][j]) + part[2][c2][j]) + part[3][c2][j];
    const int col = sl * 64 + co0 + j;                // GEMM column = dY channel
    int co = col;
    if (p.ps_perm) { const int cq = p.cout >> 2; co = 4 * (col % cq) + col / cq; }
    dw[((size_t)co * p.cin + kc * 64 + c2) * 9 + tap] += r;
  }
}

}  // namespace fsr
