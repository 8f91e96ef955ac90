// precise.cuh - elementwise kernels of the PRECISE generator forward (compute_dtype = float32).
//
// Why it exists (measured): on the reference's shipped checkpoint models/model.pt the residual
// stream grows to |x| ~ 17 and every 16-bit rounding - of the stored skip stream, of the raw conv outputs, of the conv
// operands, of the weights - contributes 1.0e-3 ... 1.9e-3 of output error on its own (fp16; 3.4e-3 together, bf16
// 3e-2): north_star's 1e-3 on that fixture needs ~fp32 arithmetic.  The tensor cores still do the work:
//   * activations are STORED in fp32 NHWC and split into two fp16 planes  a = a_hi + a_lo  (a_hi = fp16(a),
//     a_lo = fp16(a - a_hi): 22 significant bits), weights likewise  w = w_hi + w_lo;
//   * every conv is three tcgen05 launches of the same implicit-GEMM kernel accumulating in fp32:
//     a_hi*w_hi + a_lo*w_hi + a_hi*w_lo  (the dropped a_lo*w_lo term is 2^-22 relative), epilogue EPI_F32;
//   * InstanceNorm statistics come from the fp32 values (same order-independent fixed-point integers as the fast path).
// ~1/4 of the fast path's speed; a parity mode, not the benchmarked one.
#pragma once
#include "fsr_common.cuh"
#include "conv3x3_tc.cuh"

namespace fsr {

FSR_DEVINL void split_store8(const float (&v)[8], __half* hi, __half* lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __half2 hh = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(v[2 * k] - back.x, v[2 * k + 1] - back.y);
    h[k] = *reinterpret_cast<const uint32_t*>(&hh);
    l[k] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo) = make_uint4(l[0], l[1], l[2], l[3]);
}

// fp32 [n8 * 8] -> fp16 hi / lo planes (same layout)
__global__ void split_f32_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, size_t n8) {
  pdl_grid_sync();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(x)[2 * i], b = reinterpret_cast<const float4*>(x)[2 * i + 1];
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    split_store8(v, hi + 8 * i, lo + 8 * i);
  }
}

// neck (model.py:75-78) in fp32: direct 3 -> 64 conv + bias + PReLU; fp32 NCHW in -> fp32 NHWC out [N,H,W,64].
// Two threads per pixel, 32 channels each; weights [27][64] in smem (broadcast float4 reads).
__global__ void __launch_bounds__(256) neck_conv3x3_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ bias, const float* __restrict__ alpha,
                                                                float* __restrict__ out, int N, int H, int W) {
  pdl_grid_sync();
  __shared__ __align__(16) float sw[27 * 64];
  __shared__ float sb[64];
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) sw[i] = w[(size_t)(i % 64) * 27 + i / 64];   // OIHW -> [tap_ci][co]
  if (threadIdx.x < 64) sb[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const float slope = __ldg(alpha);
  const size_t total = (size_t)N * H * W;
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t pix = gt >> 1;
  const int half = (int)(gt & 1);
  if (pix >= total) return;
  const int px = (int)(pix % W), py = (int)((pix / W) % H), n = (int)(pix / ((size_t)W * H));
  float in[27];
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int yy = py + r - 1, xx = px + s - 1;
        in[ci * 9 + r * 3 + s] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(x + ((size_t)(n * 3 + ci) * H + yy) * W + xx) : 0.f;
      }
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = sb[half * 32 + c];
#pragma unroll
  for (int t = 0; t < 27; ++t) {
    const float4* wr = reinterpret_cast<const float4*>(sw + t * 64 + half * 32);
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      const float4 w4 = wr[c4];
      acc[4 * c4 + 0] = fmaf(in[t], w4.x, acc[4 * c4 + 0]);
      acc[4 * c4 + 1] = fmaf(in[t], w4.y, acc[4 * c4 + 1]);
      acc[4 * c4 + 2] = fmaf(in[t], w4.z, acc[4 * c4 + 2]);
      acc[4 * c4 + 3] = fmaf(in[t], w4.w, acc[4 * c4 + 3]);
    }
  }
  float4* o = reinterpret_cast<float4*>(out + pix * 64 + half * 32);
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4) {
    float4 v;
    v.x = apply_act(acc[4 * c4 + 0], ACT_PRELU, slope); v.y = apply_act(acc[4 * c4 + 1], ACT_PRELU, slope);
    v.z = apply_act(acc[4 * c4 + 2], ACT_PRELU, slope); v.w = apply_act(acc[4 * c4 + 3], ACT_PRELU, slope);
    o[c4] = v;
  }
}

// InstanceNorm statistics of an fp32 NHWC tensor [N][HW][64]: stats [N][64][2] += fixed point (sum * 2^24, sumsq * 2^20).
// grid (blocks_per_image, N), 256 threads: thread = (pixel lane 0..15, 4-channel group 0..15).
__global__ void __launch_bounds__(256) in_stats_f32_kernel(const float* __restrict__ x, long long* __restrict__ stats, int HW) {
  pdl_grid_sync();
  const int n = blockIdx.y;
  const int cg = threadIdx.x & 15, pl = threadIdx.x >> 4;
  const float4* base = reinterpret_cast<const float4*>(x + (size_t)n * HW * 64) + cg;
  long long s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  for (int p = blockIdx.x * 16 + pl; p < HW; p += gridDim.x * 16) {
    const float4 v = base[(size_t)p * 16];
    const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s[k] += stat_fix(f[k], kStatSumScale);
      q[k] += stat_fix(f[k] * f[k], kStatSqScale);
    }
  }
  __shared__ long long red[16][16][8];
#pragma unroll
  for (int k = 0; k < 4; ++k) { red[pl][cg][k] = s[k]; red[pl][cg][4 + k] = q[k]; }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int c16 = threadIdx.x >> 3, k = threadIdx.x & 7;
    long long t = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += red[i][c16][k];
    const int ch = c16 * 4 + (k & 3);
    stat_atomic_add(stats + ((size_t)n * 64 + ch) * 2 + (k >> 2), t);
  }
}

// out = act((x - mean) * rstd) (+ residual), all fp32 NHWC [N][HW][64]; optional fp16 hi / lo planes of the result.
__global__ void __launch_bounds__(256) in_apply_f32_kernel(const float* __restrict__ x, const long long* __restrict__ stats,
                                                            const float* __restrict__ residual, float* __restrict__ out,
                                                            __half* __restrict__ hi, __half* __restrict__ lo,
                                                            const float* __restrict__ alpha, int act, int HW, float eps) {
  pdl_grid_sync();
  const int n = blockIdx.y;
  __shared__ float smean[64], srstd[64];
  if (threadIdx.x < 64) stat_mean_rstd(stats + ((size_t)n * 64 + threadIdx.x) * 2, 1.0 / (double)HW, eps, smean[threadIdx.x], srstd[threadIdx.x]);
  __syncthreads();
  const float slope = act == ACT_PRELU ? __ldg(alpha) : 0.f;
  const size_t img = (size_t)n * HW * 64;
  const size_t nvec = (size_t)HW * 8;                       // 8-channel vectors per image
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i & 7) * 8;
    const float4 a = reinterpret_cast<const float4*>(x + img)[2 * i], b = reinterpret_cast<const float4*>(x + img)[2 * i + 1];
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = apply_act((v[k] - smean[c0 + k]) * srstd[c0 + k], act, slope);
    if (residual) {
      const float4 ra = reinterpret_cast<const float4*>(residual + img)[2 * i], rb = reinterpret_cast<const float4*>(residual + img)[2 * i + 1];
      v[0] += ra.x; v[1] += ra.y; v[2] += ra.z; v[3] += ra.w; v[4] += rb.x; v[5] += rb.y; v[6] += rb.z; v[7] += rb.w;
    }
    reinterpret_cast<float4*>(out + img)[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(out + img)[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
    if (hi) split_store8(v, hi + img + 8 * i, lo + img + 8 * i);
  }
}

// UpSamplingBlock tail (model.py:39-40) on the fp32 conv output: conv [N,H,W,256] in packed column order
// (column (2i+j)*64 + c <- reference channel 4c+2i+j) + bias_packed -> PixelShuffle(2) -> PReLU -> out fp32 [N,2H,2W,64]
// (+ hi / lo planes).
__global__ void __launch_bounds__(256) ps_prelu_f32_kernel(const float* __restrict__ conv, const float* __restrict__ bias_packed,
                                                            const float* __restrict__ alpha, float* __restrict__ out,
                                                            __half* __restrict__ hi, __half* __restrict__ lo, int N, int H, int W) {
  pdl_grid_sync();
  const float slope = __ldg(alpha);
  const size_t nvec = (size_t)N * H * W * 32;               // 8-column vectors of the conv output
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    const int v8 = (int)(i & 31);                            // vector within the pixel's 256 columns
    const size_t pix = i >> 5;
    const int px = (int)(pix % W), py = (int)((pix / W) % H), n = (int)(pix / ((size_t)W * H));
    const int q = v8 >> 3, c0 = (v8 & 7) * 8;                // q = 2i + j
    const float4 a = reinterpret_cast<const float4*>(conv)[2 * i], b = reinterpret_cast<const float4*>(conv)[2 * i + 1];
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = apply_act(v[k] + __ldg(bias_packed + q * 64 + c0 + k), ACT_PRELU, slope);
    const size_t o = (((size_t)n * 2 * H + 2 * py + (q >> 1)) * (2 * W) + 2 * px + (q & 1)) * 64 + c0;
    reinterpret_cast<float4*>(out + o)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(out + o)[1] = make_float4(v[4], v[5], v[6], v[7]);
    if (hi) split_store8(v, hi + o, lo + o);
  }
}

// head tail (model.py:109 / inference.py:54-56): y = tanh(pre) in place (fp32 NCHW), or -> uint8 NHWC (truncating cast)
__global__ void tanh_f32_kernel(float* __restrict__ y, size_t n) {
  pdl_grid_sync();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = tanhf(y[i]);
}
__global__ void tanh_u8_kernel(const float* __restrict__ pre, uint8_t* __restrict__ out, int N, int HW) {
  pdl_grid_sync();
  const size_t total = (size_t)N * HW * 3;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % 3);
    const size_t pix = i / 3;
    const int n = (int)(pix / HW);
    const size_t p = pix - (size_t)n * HW;
    const float f = (tanhf(pre[((size_t)n * 3 + c) * HW + p]) + 1.0f) / 2.0f * 255.0f;
    out[i] = (uint8_t)(int)fminf(fmaxf(f, 0.f), 255.f);
  }
}

}  // namespace fsr
