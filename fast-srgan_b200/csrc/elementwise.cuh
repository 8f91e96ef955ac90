// elementwise.cuh - HBM-bound kernels of the Fast-SRGAN hot path (CUDA cores, vectorised 16-B access).
//   neck_conv3x3_kernel   : Conv2d(3->F,k3,p1)+bias+PReLU|LeakyReLU   (reference model.py:75-78, 143-146)
//   instnorm_apply_kernel : InstanceNorm2d normalise (+PReLU|LeakyReLU) (+residual add)
//                           (reference model.py:55-56, 65+69, 94+115, 132-133)
//   pixel_shuffle2_kernel : torch.nn.PixelShuffle(2) on NHWC          (reference model.py:36)
//   layout kernels        : NCHW fp32 <-> NHWC fp16/bf16 at the module boundary
#pragma once
#include "fsr_common.cuh"
#include "conv3x3_tc.cuh"   // ActMode / apply_act

namespace fsr {

// ------------------------------------------------------------------ neck: direct 3->COUT conv
// Two threads = one output pixel x 64 output channels (blockIdx.y selects the 64-channel group).
// Weights [27][64] fp32 in smem, read as broadcast float4.  Input: fp32 NCHW or uint8 NHWC
// (uint8 path folds reference inference.py:50  x/127.5 - 1).  VGG mode folds model.py:21-22
// ((x+1)/2 - mean)/std applied to in-image pixels only (zero padding comes AFTER the renorm).
struct NeckParams {
  const void* x;
  const float* w;      // [COUT,3,3,3] OIHW fp32
  const float* bias;   // [COUT]
  const float* alpha;  // PReLU slope pointer (ACT_PRELU)
  void* out;           // NHWC T [N,H,W,COUT]
  int N, H, W, cout;
  int act;
  float slope;
  int in_u8;
  int vgg_norm;
  int pitch;          // channels per stored pixel (0 = cout); < cout: only the first `pitch` channels are written (pair rows)
};

template <typename T, int TPP>
__global__ void __launch_bounds__(TPP == 2 ? 256 : 128, TPP == 2 ? 3 : 1) neck_conv3x3_kernel(const NeckParams p) {
  pdl_grid_sync();
  // TPP threads per output pixel, 64/TPP output channels each.  TPP = 2 (fewer registers, more resident warps) wins on
  // the small training images, TPP = 1 (one input gather per pixel) on large inference batches (measured).
  // weights [27][64] fp32 in smem, read as broadcast float4
  constexpr int CPT = 64 / TPP;
  __shared__ __align__(16) float sw[27 * 64];
  __shared__ float sb[64];
  const int cg = blockIdx.y;   // 64-channel group
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) {
    const int tap_ci = i / 64, co = i % 64;          // tap_ci = ci*9 + r*3 + s  (OIHW inner order)
    sw[i] = p.w[(size_t)(cg * 64 + co) * 27 + tap_ci];
  }
  if (threadIdx.x < 64) sb[threadIdx.x] = p.bias ? p.bias[cg * 64 + threadIdx.x] : 0.f;
  __syncthreads();
  const float slope = (p.act == ACT_PRELU) ? __ldg(p.alpha) : p.slope;
  const size_t total = (size_t)p.N * p.H * p.W;
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t pix = gt / TPP;
  const int half = (int)(gt % TPP);
  if (pix >= total) return;
  const int x = (int)(pix % p.W);
  const int y = (int)((pix / p.W) % p.H);
  const int n = (int)(pix / ((size_t)p.W * p.H));

  float in[27];
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int yy = y + r - 1, xx = x + s - 1;
        float v = 0.f;
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
          if (p.in_u8) {
            v = (float)reinterpret_cast<const uint8_t*>(p.x)[((size_t)(n * p.H + yy) * p.W + xx) * 3 + ci] / 127.5f - 1.0f;
          } else {
            v = __ldg(reinterpret_cast<const float*>(p.x) + ((size_t)(n * 3 + ci) * p.H + yy) * p.W + xx);
          }
          if (p.vgg_norm) v = ((v + 1.0f) / 2.0f - mean[ci]) / stdv[ci];
        }
        in[ci * 9 + r * 3 + s] = v;
      }

  float acc[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) acc[c] = sb[half * CPT + c];
#pragma unroll
  for (int t = 0; t < 27; ++t) {
    const float4* wr = reinterpret_cast<const float4*>(sw + t * 64 + half * CPT);
#pragma unroll
    for (int c4 = 0; c4 < CPT / 4; ++c4) {
      const float4 w4 = wr[c4];
      acc[4 * c4 + 0] = fmaf(in[t], w4.x, acc[4 * c4 + 0]);
      acc[4 * c4 + 1] = fmaf(in[t], w4.y, acc[4 * c4 + 1]);
      acc[4 * c4 + 2] = fmaf(in[t], w4.z, acc[4 * c4 + 2]);
      acc[4 * c4 + 3] = fmaf(in[t], w4.w, acc[4 * c4 + 3]);
    }
  }
  T* o = reinterpret_cast<T*>(p.out) + pix * p.cout + cg * 64 + half * CPT;
#pragma unroll
  for (int k = 0; k < CPT / 8; ++k) {
    uint4 pk;
    pk.x = Cvt<T>::pack2(apply_act(acc[8 * k + 0], p.act, slope), apply_act(acc[8 * k + 1], p.act, slope));
    pk.y = Cvt<T>::pack2(apply_act(acc[8 * k + 2], p.act, slope), apply_act(acc[8 * k + 3], p.act, slope));
    pk.z = Cvt<T>::pack2(apply_act(acc[8 * k + 4], p.act, slope), apply_act(acc[8 * k + 5], p.act, slope));
    pk.w = Cvt<T>::pack2(apply_act(acc[8 * k + 6], p.act, slope), apply_act(acc[8 * k + 7], p.act, slope));
    reinterpret_cast<uint4*>(o)[k] = pk;
  }
}

// ------------------------------------------------------------------ InstanceNorm apply
// out = act((raw - mean[n,c]) * rstd[n,c]) (+ residual);  stats = [N][C][2] (sum, sumsq) fp32
// accumulated by the producing conv's epilogue.  grid = (blocks_per_image, N).
struct InApplyParams {
  const void* raw;
  const long long* stats;  // [N][C][2] fixed-point (see stat_atomic_add)
  const void* residual;  // nullable
  void* out;
  const float* alpha;    // PReLU slope pointer
  float slope;
  int act;
  int HW, C;
  float eps;
  int parity_w;          // > 0: write `out` in the parity-plane layout [N][4][H/2][W/2][C] the stride-2 convs read
                         //      (image width W = parity_w; saves the separate re-layout pass of the discriminator)
};

template <typename T>
__global__ void __launch_bounds__(256) instnorm_apply_kernel(const InApplyParams p) {
  pdl_grid_sync();
  extern __shared__ float s_ms[];  // mean[C], rstd[C]
  float* s_mean = s_ms;
  float* s_rstd = s_ms + p.C;
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    stat_mean_rstd(p.stats + ((size_t)n * p.C + c) * 2, 1.0 / (double)p.HW, p.eps, s_mean[c], s_rstd[c]);
  }
  __syncthreads();
  const float slope = (p.act == ACT_PRELU) ? __ldg(p.alpha) : p.slope;
  const int vec_per_pix = p.C / 8;
  const size_t nvec = (size_t)p.HW * vec_per_pix;
  const uint4* raw = reinterpret_cast<const uint4*>(p.raw) + (size_t)n * nvec;
  const uint4* res = p.residual ? reinterpret_cast<const uint4*>(p.residual) + (size_t)n * nvec : nullptr;
  uint4* out = reinterpret_cast<uint4*>(p.out) + (size_t)n * nvec;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % vec_per_pix) * 8;
    const uint4 r = raw[i];
    uint4 rs = make_uint4(0, 0, 0, 0);
    if (res) rs = res[i];
    const uint32_t ru[4] = {r.x, r.y, r.z, r.w};
    const uint32_t su[4] = {rs.x, rs.y, rs.z, rs.w};
    uint32_t ou[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = Cvt<T>::unpack2(ru[k]);
      float a = (f.x - s_mean[c0 + 2 * k]) * s_rstd[c0 + 2 * k];
      float b = (f.y - s_mean[c0 + 2 * k + 1]) * s_rstd[c0 + 2 * k + 1];
      a = apply_act(a, p.act, slope);
      b = apply_act(b, p.act, slope);
      if (res) {
        const float2 g = Cvt<T>::unpack2(su[k]);
        a += g.x;
        b += g.y;
      }
      ou[k] = Cvt<T>::pack2(a, b);
    }
    size_t o = i;
    if (p.parity_w > 0) {
      const int pix = (int)(i / vec_per_pix), v = (int)(i - (size_t)pix * vec_per_pix);
      const int y = pix / p.parity_w, x = pix - y * p.parity_w;
      const int W2 = p.parity_w >> 1, H2 = (p.HW / p.parity_w) >> 1;
      o = ((size_t)(((y & 1) * 2 + (x & 1)) * H2 + (y >> 1)) * W2 + (x >> 1)) * vec_per_pix + v;
    }
    out[o] = make_uint4(ou[0], ou[1], ou[2], ou[3]);
  }
}

// ------------------------------------------------------------------ standalone PixelShuffle(2), NHWC
// out[n, 2y+i, 2x+j, c] = in[n, y, x, 4c + 2i + j]   (reference channel order, model.py:36)
// One thread: one input pixel x 8 output channels -> reads 32 consecutive input channels (64 B),
// writes one 16-B vector to each of the 4 output pixels.
template <typename T>
__global__ void __launch_bounds__(256) pixel_shuffle2_kernel(const T* __restrict__ in, T* __restrict__ out, int N,
                                                             int H, int W, int C /*output channels*/) {
  pdl_grid_sync();
  const int groups = C / 8;
  const size_t total = (size_t)N * H * W * groups;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(idx % groups);
    const size_t pix = idx / groups;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const int n = (int)(pix / ((size_t)W * H));
    const uint4* src = reinterpret_cast<const uint4*>(in + pix * (size_t)(4 * C) + (size_t)g * 32);
    uint32_t w[16];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint4 v = src[k];
      w[4 * k + 0] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
    }
    // w[] holds 32 halfs: element e = 4*cl + 2i + j  (cl = local out channel 0..7)
    const uint16_t* h = reinterpret_cast<const uint16_t*>(w);
#pragma unroll
    for (int q = 0; q < 4; ++q) {   // q = 2i + j
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        o[k] = (uint32_t)h[4 * (2 * k) + q] | ((uint32_t)h[4 * (2 * k + 1) + q] << 16);
      const int oy = 2 * y + (q >> 1), ox = 2 * x + (q & 1);
      uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)(n * 2 * H + oy) * (2 * W) + ox) * C + (size_t)g * 8);
      *dst = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ------------------------------------------------------------------ layout conversion
template <typename T>
__global__ void __launch_bounds__(256) nchw_f32_to_nhwc_kernel(const float* __restrict__ in, T* __restrict__ out, int N,
                                                               int C, int HW) {
  pdl_grid_sync();
  // tile transpose through smem: 32 pixels x 32 channels
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, pp = p0 + tx;
    tile[j][tx] = (c < C && pp < HW) ? in[((size_t)n * C + c) * HW + pp] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int pp = p0 + j, c = c0 + tx;
    if (pp < HW && c < C) out[((size_t)n * HW + pp) * C + c] = Cvt<T>::from_f(tile[tx][j]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_f32_kernel(const T* __restrict__ in, float* __restrict__ out, int N,
                                                               int C, int HW) {
  pdl_grid_sync();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int pp = p0 + j, c = c0 + tx;
    tile[j][tx] = (pp < HW && c < C) ? Cvt<T>::to_f(in[((size_t)n * HW + pp) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, pp = p0 + tx;
    if (c < C && pp < HW) out[((size_t)n * C + c) * HW + pp] = tile[tx][j];
  }
}

// ------------------------------------------------------------------ weight packing
// OIHW fp32 [Cout, Cin, 3, 3]  ->  [9][cout_pad][Cin] T  (tap-major, each row = Cin contiguous = K-major B operand)
// ps_perm != 0: GEMM row n' = (2i+j)*(Cout/4) + c  holds reference output channel oc = 4c + 2i + j, so that
// the PixelShuffle(2) of model.py:36 becomes a contiguous 64-channel store per (i,j).
template <typename T>
__global__ void pack_conv3x3_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin,
                                           int cout_pad, int ps_perm) {
  pdl_grid_sync();
  const size_t total = (size_t)9 * cout_pad * cin;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(idx % cin);
    const int row = (int)((idx / cin) % cout_pad);
    const int tap = (int)(idx / ((size_t)cin * cout_pad));
    float v = 0.f;
    if (row < cout) {
      int oc = row;
      if (ps_perm) {
        const int cq = cout / 4;
        const int q = row / cq, c = row % cq;
        oc = 4 * c + q;
      }
      v = w[((size_t)oc * cin + ci) * 9 + tap];
    }
    out[idx] = Cvt<T>::from_f(v);
  }
}

// Data-gradient pack: out[tap'][ci (rows, padded to row_pad)][col] = scale[ci] * W[oc(col)][ci][flip ? 8-tap' : tap']
// (rows = forward INPUT channel, K = forward OUTPUT channel).  flip=0 for the general conv (its tap table
// carries the flip), flip=1 for the resident-weight 64-channel / head-style kernels.  ps_perm: col = q*(cout/4)+c
// <-> oc = 4c+q.  row_scale (nullable, [cin]) folds a per-input-channel factor (VGG renorm chain rule).
template <typename T>
__global__ void pack_conv3x3_weight_t_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin, int ps_perm,
                                             int flip, int row_pad, const float* __restrict__ row_scale) {
  pdl_grid_sync();
  const size_t total = (size_t)9 * row_pad * cout;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int col = (int)(idx % cout);
    const int ci = (int)((idx / cout) % row_pad);
    const int tap = (int)(idx / ((size_t)cout * row_pad));
    float v = 0.f;
    if (ci < cin) {
      int oc = col;
      if (ps_perm) { const int cq = cout / 4; oc = 4 * (col % cq) + col / cq; }
      v = w[((size_t)oc * cin + ci) * 9 + (flip ? 8 - tap : tap)];
      if (row_scale) v *= row_scale[ci];
    }
    out[idx] = Cvt<T>::from_f(v);
  }
}

__global__ void permute_bias_ps_kernel(const float* __restrict__ b, float* __restrict__ out, int cout, int cout_pad,
                                       int ps_perm) {
  pdl_grid_sync();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= cout_pad) return;
  float v = 0.f;
  if (row < cout) {
    int oc = row;
    if (ps_perm) {
      const int cq = cout / 4;
      oc = 4 * (row % cq) + row / cq;
    }
    v = b[oc];
  }
  out[row] = v;
}

// All weight packs of a network in ONE launch (the training step re-packs every conv of G and D after each AdamW step:
// 54 launches of ~4 us each).  Task = one fsr_pack_conv3x3_weight / _t call; a block finds its task by a linear scan.
constexpr int kPackMaxTasks = 48;
struct PackTaskDev {
  const float* w; void* out; const float* bias; float* bias_out; const float* row_scale;
  int cout, cin, pad, flags;        // flags: 1 = transposed (data-gradient) pack, 2 = ps_perm, 4 = flip; pad = cout_pad | row_pad
};
struct PackMultiParams {
  PackTaskDev t[kPackMaxTasks];
  int block_begin[kPackMaxTasks + 1];
  int n;
};

template <typename T>
__global__ void __launch_bounds__(256) pack_multi_kernel(const __grid_constant__ PackMultiParams p) {
  pdl_grid_sync();
  int ti = 0;
  while (ti + 1 < p.n && (int)blockIdx.x >= p.block_begin[ti + 1]) ++ti;
  const PackTaskDev& k = p.t[ti];
  const int nb = p.block_begin[ti + 1] - p.block_begin[ti], b = blockIdx.x - p.block_begin[ti];
  const int cout = k.cout, cin = k.cin;
  const bool ps = k.flags & 2, flip = k.flags & 4;
  T* out = reinterpret_cast<T*>(k.out);
  if (k.flags & 1) {
    const int row_pad = k.pad;
    const size_t total = (size_t)9 * row_pad * cout;
    for (size_t idx = (size_t)b * 256 + threadIdx.x; idx < total; idx += (size_t)nb * 256) {
      const int col = (int)(idx % cout), ci = (int)((idx / cout) % row_pad), tap = (int)(idx / ((size_t)cout * row_pad));
      float v = 0.f;
      if (ci < cin) {
        int oc = col;
        if (ps) { const int cq = cout / 4; oc = 4 * (col % cq) + col / cq; }
        v = k.w[((size_t)oc * cin + ci) * 9 + (flip ? 8 - tap : tap)];
        if (k.row_scale) v *= k.row_scale[ci];
      }
      out[idx] = Cvt<T>::from_f(v);
    }
  } else {
    const int cout_pad = k.pad;
    const size_t total = (size_t)9 * cout_pad * cin;
    for (size_t idx = (size_t)b * 256 + threadIdx.x; idx < total; idx += (size_t)nb * 256) {
      const int ci = (int)(idx % cin), row = (int)((idx / cin) % cout_pad), tap = (int)(idx / ((size_t)cin * cout_pad));
      float v = 0.f;
      if (row < cout) {
        int oc = row;
        if (ps) { const int cq = cout / 4; oc = 4 * (row % cq) + row / cq; }
        v = k.w[((size_t)oc * cin + ci) * 9 + tap];
      }
      out[idx] = Cvt<T>::from_f(v);
    }
    if (k.bias && k.bias_out) {
      for (int row = b * 256 + threadIdx.x; row < cout_pad; row += nb * 256) {
        float v = 0.f;
        if (row < cout) {
          int oc = row;
          if (ps) { const int cq = cout / 4; oc = 4 * (row % cq) + row / cq; }
          v = k.bias[oc];
        }
        k.bias_out[row] = v;
      }
    }
  }
}

}  // namespace fsr
