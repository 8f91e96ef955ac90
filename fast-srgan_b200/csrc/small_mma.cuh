// small_mma.cuh - the two 3-channel-sided 3x3 convs of the path on warp-level tensor-core MMAs (mma.sync m16n8k16).
//   neck_conv3x3_mma_kernel : Conv2d(3->64k, k3, p1) + bias + activation   (reference model.py:75-78, 143-146; VGG conv1_1
//                             model.py:20-23 with the renormalisation folded into the load)
//   wgrad_c3_mma_kernel     : weight gradient of a 3x3 conv with one 3-channel side (G/D neck, G head; autograd of
//                             model.py:76, 103-108, 144 called at trainer.py:180,195)
// Both are K = 27 (padded to 32) contractions: far too thin for tcgen05 tiles, and HBM-bound on the 64-channel side
// (128 B per pixel).  The CUDA-core versions (elementwise.cuh / train_kernels.cuh, kept as the A/B switch
// fsr_set_small_mma(0)) were FMA-issue bound (neck: 355 us at b32 180x320 against 45 us of HBM time) resp. latency bound
// (wgrad_c3: 573 us per launch at b64 96x96 against 15 us).  The fp32 3-channel operand is split into two 16-bit terms
// (hi + lo) so that the result keeps fp32-input accuracy: hi*W + lo*W (+ hi*Wlo for the neck weights).
#pragma once
#include "fsr_common.cuh"
#include "conv3x3_tc.cuh"    // ActMode / apply_act
#include "elementwise.cuh"   // NeckParams

namespace fsr {

template <typename T>
struct Mma16816;
template <>
struct Mma16816<__half> {
  FSR_DEVINL static void run(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};
template <>
struct Mma16816<__nv_bfloat16> {
  FSR_DEVINL static void run(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};

// v0, v1 (fp32) -> packed 16-bit pair of the leading terms and of the remainders
template <typename T>
FSR_DEVINL void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const T h0 = Cvt<T>::from_f(v0), h1 = Cvt<T>::from_f(v1);
  hi = Cvt<T>::pack2(Cvt<T>::to_f(h0), Cvt<T>::to_f(h1));
  lo = Cvt<T>::pack2(v0 - Cvt<T>::to_f(h0), v1 - Cvt<T>::to_f(h1));
}

// ------------------------------------------------------------------ neck: 3 -> 64 conv, one warp = 16 pixels of a row
// GEMM per warp strip: D[16 px, 64 ch] = A[16 px, 32] * B[32, 64], A = im2col of a 3x3x18 fp32 strip staged in the warp's
// private smem (no block barrier anywhere), B = weights held in registers as fp16 hi/lo fragments for the whole kernel.
// The arithmetic is fp16 hi/lo regardless of the output type T (inputs are images in [-1,1] / ImageNet-normalised).
// Column c of n-tile j is output channel 16*(c>>1) + 2j + (c&1): every lane then owns 16 CONTIGUOUS channels of its two
// pixels and stores them as two 16-byte vectors (a warp writes whole 128-byte pixel rows).
constexpr int kNeckWarps = 4;

// One lane's share (6 of the 162 values) of the 3 x 3 x 18 input strip of (image n, row y, columns x0-1 .. x0+16):
// branch-free predicated loads, so that all six are in flight together (a first version with the uint8 / renormalise /
// bounds logic as branches serialised the six load latencies: 245 us at b32 180x320).
template <bool IN_U8, bool VGG>
FSR_DEVINL void neck_load_strip(const NeckParams& p, int n, int y, int x0, int lane, float (&v)[6]) {
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int e = lane + 32 * i;
    const int ci = e / 54, r = (e % 54) / 18, c = e % 18;
    const int yy = y + r - 1, xx = x0 + c - 1;
    const bool ok = e < 162 && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
    float val;
    if (IN_U8) {
      const size_t idx = ok ? ((size_t)(n * p.H + yy) * p.W + xx) * 3 + ci : 0;
      const uint8_t u = __ldg(reinterpret_cast<const uint8_t*>(p.x) + idx);
      val = (float)u / 127.5f - 1.0f;                                         // inference.py:50
    } else {
      const size_t idx = ok ? ((size_t)(n * 3 + ci) * p.H + yy) * p.W + xx : 0;
      val = __ldg(reinterpret_cast<const float*>(p.x) + idx);
    }
    if (VGG) {   // model.py:21-22 (same operation order as the CUDA-core kernel)
      const float mu = ci == 0 ? 0.485f : (ci == 1 ? 0.456f : 0.406f);
      const float sd = ci == 0 ? 0.229f : (ci == 1 ? 0.224f : 0.225f);
      val = ((val + 1.0f) / 2.0f - mu) / sd;
    }
    v[i] = ok ? val : 0.f;   // zero outside the image: the conv's padding comes after any renormalisation
  }
}

template <typename T, bool IN_U8, bool VGG>
__global__ void __launch_bounds__(kNeckWarps * 32, 3) neck_conv3x3_mma_kernel(const NeckParams p) {
  pdl_grid_sync();
  __shared__ float s_strip[kNeckWarps][3 * 3 * 18 + 30];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int cg = blockIdx.y;
  float* strip = s_strip[wib];

  // ---- B fragments: b[ks][j] = {rows k = 16ks+2t,+1 | rows 16ks+8+2t,+1} x column g of n-tile j
  uint32_t bh[2][8][2], bl[2][8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = cg * 64 + 16 * (g >> 1) + 2 * j + (g & 1);
    const float* wr = p.w + (size_t)ch * 27;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = 16 * ks + 8 * h + 2 * t;
        const float w0 = __ldg(wr + (k < 27 ? k : 0));
        const float w1 = __ldg(wr + (k + 1 < 27 ? k + 1 : 0));
        split2<__half>(k < 27 ? w0 : 0.f, k + 1 < 27 ? w1 : 0.f, bh[ks][j][h], bl[ks][j][h]);
      }
  }
  // ---- per-lane im2col offsets into the strip: k = (ci, r, s) -> ci*54 + r*18 + s (k >= 27: B row is zero, any finite A)
  int koff[2][2][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 16 * ks + 8 * h + 2 * t + e;
        koff[ks][h][e] = k < 27 ? (k / 9) * 54 + ((k % 9) / 3) * 18 + (k % 3) : 0;
      }
  const float* bias_l = p.bias ? p.bias + cg * 64 + 16 * t : nullptr;   // lane's 16 channels (views of a flat buffer: 4-B aligned only)
  const float slope = (p.act == ACT_PRELU) ? __ldg(p.alpha) : p.slope;
  const int pitch = p.pitch ? p.pitch : p.cout;

  const int tx = (p.W + 15) >> 4;
  const int rows = p.N * p.H;                    // host guarantees N*H*tx < 2^31
  const int nstrips = rows * tx;
  const int stride = gridDim.x * kNeckWarps;
  int sidx = blockIdx.x * kNeckWarps + wib;
  float v[6];
  if (sidx < nstrips) {
    const int row = sidx / tx;
    neck_load_strip<IN_U8, VGG>(p, row / p.H, row % p.H, (sidx - row * tx) << 4, lane, v);
  }
  for (; sidx < nstrips; sidx += stride) {
    const int row = sidx / tx;
    const int n = row / p.H, y = row - n * p.H, x0 = (sidx - row * tx) << 4;
#pragma unroll
    for (int i = 0; i < 6; ++i) strip[lane + 32 * i] = v[i];       // 192 slots (162 used)
    __syncwarp();
    if (sidx + stride < nstrips) {   // next strip's loads fly during this strip's MMAs and stores
      const int nrow = (sidx + stride) / tx;
      neck_load_strip<IN_U8, VGG>(p, nrow / p.H, nrow % p.H, (sidx + stride - nrow * tx) << 4, lane, v);
    }
    // A fragments (row = pixel g / g+8, col = k pair), hi/lo
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int m = g + 8 * rr;
          split2<__half>(strip[koff[ks][h][0] + m], strip[koff[ks][h][1] + m], ah[ks][2 * h + rr], al[ks][2 * h + rr]);
        }
    __syncwarp();   // strip is free for the next iteration's staging
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j][0] = acc[j][2] = bias_l ? __ldg(bias_l + 2 * j) : 0.f;       // L1-resident after the first strip
      acc[j][1] = acc[j][3] = bias_l ? __ldg(bias_l + 2 * j + 1) : 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        Mma16816<__half>::run(acc[j], al[ks], bh[ks][j][0], bh[ks][j][1]);
        Mma16816<__half>::run(acc[j], ah[ks], bl[ks][j][0], bl[ks][j][1]);
        Mma16816<__half>::run(acc[j], ah[ks], bh[ks][j][0], bh[ks][j][1]);
      }
    // epilogue: lane owns channels cg*64 + 16t .. +15 of pixels x0+g and x0+g+8
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int x = x0 + g + 8 * rr;
      if (x < p.W && cg * 64 + 16 * t < pitch) {
        T* o = reinterpret_cast<T*>(p.out) + ((size_t)(n * p.H + y) * p.W + x) * pitch + cg * 64 + 16 * t;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint4 pk;
          pk.x = Cvt<T>::pack2(apply_act(acc[4 * q + 0][2 * rr], p.act, slope), apply_act(acc[4 * q + 0][2 * rr + 1], p.act, slope));
          pk.y = Cvt<T>::pack2(apply_act(acc[4 * q + 1][2 * rr], p.act, slope), apply_act(acc[4 * q + 1][2 * rr + 1], p.act, slope));
          pk.z = Cvt<T>::pack2(apply_act(acc[4 * q + 2][2 * rr], p.act, slope), apply_act(acc[4 * q + 2][2 * rr + 1], p.act, slope));
          pk.w = Cvt<T>::pack2(apply_act(acc[4 * q + 3][2 * rr], p.act, slope), apply_act(acc[4 * q + 3][2 * rr + 1], p.act, slope));
          reinterpret_cast<uint4*>(o)[q] = pk;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ 3-channel weight gradient
//   dW[k27][c64] = sum_px imgcol[px][k27] * act[px][c64],  k27 = c3*9 + r*3 + s,  imgcol = img[n, c3, y+dy, x+dx],
//   (dy,dx) = (r-1,s-1) (img = conv input) or (1-r,1-s) when `flip` (img = output gradient; train_kernels.cuh wgrad_c3).
// Per warp step of 16 consecutive pixels: A = imgcol^T [32 (27) x 16 px] built from L1-cached scalar gathers and split
// hi/lo in T; B = act [16 px x 64 ch] straight from global: lane (g,t) loads the 16-byte vectors (channels 8g..8g+7) of
// pixels 2t, 2t+1, 2t+8, 2t+9 - whole 128-byte rows per warp - and byte-permutes the two pixels of each channel into
// the k-pair registers; column g of n-tile j is channel 8g + j.  64 fp32 accumulators per lane live for the warp's whole
// pixel range; warps combine through shared atomics, blocks through 27x64 global atomics.
constexpr int kWgc3Warps = 4;
template <typename T>
__global__ void __launch_bounds__(kWgc3Warps * 32, 3) wgrad_c3_mma_kernel(const float* __restrict__ img, const T* __restrict__ act,
                                                                          float* __restrict__ out, int N, int H, int W, int C64,
                                                                          int flip, int layout, DetRed red) {
  pdl_grid_sync();
  __shared__ float s_red[32 * 65];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int cbase = blockIdx.y * 64;
  for (int i = threadIdx.x; i < 32 * 65; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();

  // the four A rows of this lane: k27 = g, g+8, g+16, g+24
  int roff[4], rdy[4], rdx[4];
  bool rok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = g + 8 * i;
    rok[i] = k < 27;
    const int kk = rok[i] ? k : 0;
    const int c3 = kk / 9, r = (kk % 9) / 3, s = kk % 3;
    rdy[i] = flip ? 1 - r : r - 1;
    rdx[i] = flip ? 1 - s : s - 1;
    roff[i] = (c3 * H + rdy[i]) * W + rdx[i];
  }
  const int total = N * H * W;                   // host guarantees N*H*W < 2^31 - 16
  const int nsteps = (total + 15) >> 4;
  const int nwarps = gridDim.x * kWgc3Warps;
  const int per = (nsteps + nwarps - 1) / nwarps;
  const int s0 = (blockIdx.x * kWgc3Warps + wib) * per;
  const int s1 = (s0 + per < nsteps) ? s0 + per : nsteps;

  float acc[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][j][e] = 0.f;

  const int HW = H * W;
  // (n, y, x) of this lane's first pixel p0 + 2t, carried across the warp's contiguous steps (one division per warp, not
  // eight per step); the other three pixels (+1, +8, +9) are derived by wrapping increments
  int bn, by, bx;
  {
    const int pb = (s0 < nsteps ? s0 : 0) * 16 + 2 * t;
    bn = pb / HW;
    const int rem0 = pb - bn * HW;
    by = rem0 / W;
    bx = rem0 - by * W;
  }
  auto wrap = [&](int& n, int& y, int& x) {
    while (x >= W) {
      x -= W;
      if (++y == H) { y = 0; ++n; }
    }
  };
  const int dq[4] = {0, 1, 8, 9};
  for (int st = s0; st < s1; ++st) {
    const int p0 = (st << 4) + 2 * t;
    uint4 v[4];
    float a[4][4];   // [row i][pixel q]
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int n = bn, y = by, x = bx + dq[q];
      wrap(n, y, x);
      const bool in = p0 + dq[q] < total;
      const int rem = in ? y * W + x : 0;
      n = in ? n : 0;
      const int pc = n * HW + rem;
      // unconditional loads from clamped addresses + selects: all 20 loads of a step are in flight together
      const uint4 av = *reinterpret_cast<const uint4*>(act + (size_t)pc * C64 + cbase + 8 * g);
      v[q] = in ? av : make_uint4(0, 0, 0, 0);
      const float* ib = img + (size_t)n * 3 * HW;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int yy = y + rdy[i], xx = x + rdx[i];
        const bool ok = in && rok[i] && yy >= 0 && yy < H && xx >= 0 && xx < W;
        const float iv = __ldg(ib + (ok ? rem + roff[i] : 0));
        a[i][q] = ok ? iv : 0.f;
      }
    }
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      split2<T>(a[2 * mt][0], a[2 * mt][1], ah[mt][0], al[mt][0]);          // row g(+16),   px 2t, 2t+1
      split2<T>(a[2 * mt + 1][0], a[2 * mt + 1][1], ah[mt][1], al[mt][1]);  // row g+8(+16), px 2t, 2t+1
      split2<T>(a[2 * mt][2], a[2 * mt][3], ah[mt][2], al[mt][2]);          // row g(+16),   px 2t+8, 2t+9
      split2<T>(a[2 * mt + 1][2], a[2 * mt + 1][3], ah[mt][3], al[mt][3]);  // row g+8(+16), px 2t+8, 2t+9
    }
    const uint32_t w0[4] = {v[0].x, v[0].y, v[0].z, v[0].w}, w1[4] = {v[1].x, v[1].y, v[1].z, v[1].w};
    const uint32_t w2[4] = {v[2].x, v[2].y, v[2].z, v[2].w}, w3[4] = {v[3].x, v[3].y, v[3].z, v[3].w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t sel = (j & 1) ? 0x7632u : 0x5410u;
      const uint32_t b0 = __byte_perm(w0[j >> 1], w1[j >> 1], sel);   // (px 2t, px 2t+1) of channel 8g + j
      const uint32_t b1 = __byte_perm(w2[j >> 1], w3[j >> 1], sel);   // (px 2t+8, px 2t+9)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        Mma16816<T>::run(acc[mt][j], al[mt], b0, b1);
        Mma16816<T>::run(acc[mt][j], ah[mt], b0, b1);
      }
    }
    bx += 16;
    wrap(bn, by, bx);
  }
  // C fragment: c0,c1 = (row g, cols 2t, 2t+1), c2,c3 = (row g+8, ...); column c of n-tile j = channel 8c + j.
  // The warps add their fragments into s_red one after the other (a warp's 32 x 64 (k, ch) cells are distinct, so no
  // atomics: shared fp32 atomicAdd is a CAS loop).
  for (int w = 0; w < kWgc3Warps; ++w) {
    if (wib == w) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = 16 * mt + g + 8 * (e >> 1);
            const int ch = 8 * (2 * t + (e & 1)) + j;
            s_red[k * 65 + ch] += acc[mt][j][e];
          }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) {
    int k, c;
    size_t idx;
    if (layout == 1) {          // OIHW [3][C64][9] (head)
      const int c3 = i / 576, tap = i % 9;
      c = (i / 9) % 64;
      k = c3 * 9 + tap;
      idx = ((size_t)c3 * C64 + cbase + c) * 9 + tap;
    } else if (layout == 2) {   // OIHW [C64][3][9] (neck)
      c = i / 27;
      k = i % 27;
      idx = (size_t)(cbase + c) * 27 + k;
    } else {                    // [27][C64]
      k = i / 64;
      c = i % 64;
      idx = (size_t)k * C64 + cbase + c;
    }
    const float vsum = s_red[k * 65 + c];
    if (vsum != 0.f) det_add(red, (int)idx, vsum);       // fixed-point: the total does not depend on block order
  }
  det_finish(red, out, 27 * C64);
}

}  // namespace fsr
