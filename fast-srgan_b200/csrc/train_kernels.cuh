// train_kernels.cuh - HBM-bound kernels of the GAN training step (reference trainer.py:168-196):
// layout (parity planes), max-pool, the discriminator's 1x1 logit conv, fused BCE-with-logits /
// SmoothL1 losses (forward + gradient), InstanceNorm / activation / pixel-shuffle / tanh backward,
// small-channel weight gradients, bias gradients and a flat fused AdamW.
#pragma once
#include "fsr_common.cuh"
#include "conv3x3_tc.cuh"

namespace fsr {

FSR_DEVINL float block_sum(float v, float* smem /* >= 32 floats */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? smem[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  return r;   // valid in warp 0
}

// ------------------------------------------------------------------ parity-plane layout
// NHWC [N,H,W,C]  <->  [N][4][H/2][W/2][C]  (plane = (y&1)*2 + (x&1)); 16-byte vectors.
template <bool TO_PARITY>
__global__ void __launch_bounds__(256) parity_layout_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int N,
                                                            int H, int W, int CV /*C/8*/) {
  pdl_grid_sync();
  const size_t total = (size_t)N * H * W * CV;
  const int H2 = H >> 1, W2 = W >> 1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    size_t pix = i / CV;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const int n = (int)(pix / ((size_t)W * H));
    const size_t j = ((((size_t)n * 4 + ((y & 1) * 2 + (x & 1))) * H2 + (y >> 1)) * W2 + (x >> 1)) * CV + cv;
    if (TO_PARITY) out[j] = in[i]; else out[i] = in[j];
  }
}

// ------------------------------------------------------------------ 2x2 max-pool (VGG idx 4,9,18,27), NHWC
template <typename T>
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const T* __restrict__ in, T* __restrict__ out, int N, int H,
                                                           int W, int C, int ip = 0, int op = 0) {
  pdl_grid_sync();
  // ip / op = 1: `in` / `out` are in the zero-bordered PADDED layout [N][H+2][W+2][C] of the flat conv kernels
  // (conv3x3_gen_2cta.cuh, flat mode); only interior pixels are read / written (the caller zero-fills a padded `out`)
  const int Ho = H >> 1, Wo = W >> 1, CV = C / 8;
  const int Wi = W + 2 * ip, Hi = H + 2 * ip, Wq = Wo + 2 * op, Hq = Ho + 2 * op;
  const size_t total = (size_t)N * Ho * Wo * CV;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    size_t pix = i / CV;
    const int x = (int)(pix % Wo);
    const int y = (int)((pix / Wo) % Ho);
    const int n = (int)(pix / ((size_t)Wo * Ho));
    const T* base = in + (((size_t)n * Hi + 2 * y + ip) * Wi + 2 * x + ip) * C + cv * 8;
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const uint4 v = *reinterpret_cast<const uint4*>(base + ((size_t)dy * Wi + dx) * C);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = Cvt<T>::unpack2(u[k]);
          m[2 * k] = fmaxf(m[2 * k], f.x);
          m[2 * k + 1] = fmaxf(m[2 * k + 1], f.y);
        }
      }
    uint4 o;
    o.x = Cvt<T>::pack2(m[0], m[1]); o.y = Cvt<T>::pack2(m[2], m[3]);
    o.z = Cvt<T>::pack2(m[4], m[5]); o.w = Cvt<T>::pack2(m[6], m[7]);
    *reinterpret_cast<uint4*>(out + (((size_t)n * Hq + y + op) * Wq + x + op) * C + cv * 8) = o;
  }
}

// backward of [conv+ReLU -> maxpool]: dIn[window] = dOut if (in == max and first such) else 0, and the ReLU
// mask of the pre-pool activation (in > 0) is applied in the same pass.
template <typename T>
__global__ void __launch_bounds__(256) maxpool2_relu_bwd_kernel(const T* __restrict__ in, const T* __restrict__ dout,
                                                                T* __restrict__ din, int N, int H, int W, int C, int ip = 0, int op = 0) {
  pdl_grid_sync();
  // ip = 1: `in` and `din` are in the padded layout [N][H+2][W+2][C]; op = 1: `dout` is [N][H/2+2][W/2+2][C]
  const int Ho = H >> 1, Wo = W >> 1;
  const int Wi = W + 2 * ip, Hi = H + 2 * ip, Wq = Wo + 2 * op, Hq = Ho + 2 * op;
  const size_t total = (size_t)N * Ho * Wo * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    size_t pix = i / C;
    const int x = (int)(pix % Wo);
    const int y = (int)((pix / Wo) % Ho);
    const int n = (int)(pix / ((size_t)Wo * Ho));
    const size_t b = (((size_t)n * Hi + 2 * y + ip) * Wi + 2 * x + ip) * C + c;
    const size_t idx[4] = {b, b + C, b + (size_t)Wi * C, b + (size_t)Wi * C + C};
    float v[4];
    int am = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = Cvt<T>::to_f(in[idx[k]]);
      if (v[k] > v[am]) am = k;
    }
    const float g = Cvt<T>::to_f(dout[(((size_t)n * Hq + y + op) * Wq + x + op) * C + c]);
#pragma unroll
    for (int k = 0; k < 4; ++k) din[idx[k]] = Cvt<T>::from_f((k == am && v[k] > 0.f) ? g : 0.f);
  }
}

// dx = dy * (y > 0)   (ReLU backward from the stored post-ReLU activation)
template <typename T>
__global__ void __launch_bounds__(256) relu_bwd_kernel(const uint4* __restrict__ y, const uint4* __restrict__ dy,
                                                       uint4* __restrict__ dx, size_t nvec) {
  pdl_grid_sync();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 a = y[i], g = dy[i];
    const uint32_t au[4] = {a.x, a.y, a.z, a.w}, gu[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = Cvt<T>::unpack2(au[k]), d = Cvt<T>::unpack2(gu[k]);
      o[k] = Cvt<T>::pack2(f.x > 0.f ? d.x : 0.f, f.y > 0.f ? d.y : 0.f);
    }
    dx[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// out = a + b
template <typename T>
__global__ void __launch_bounds__(256) add_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                                  uint4* __restrict__ out, size_t nvec) {
  pdl_grid_sync();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 x = a[i], y = b[i];
    const uint32_t xu[4] = {x.x, x.y, x.z, x.w}, yu[4] = {y.x, y.y, y.z, y.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = Cvt<T>::unpack2(xu[k]), g = Cvt<T>::unpack2(yu[k]);
      o[k] = Cvt<T>::pack2(f.x + g.x, f.y + g.y);
    }
    out[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------ D's final Conv2d(512 -> 1, k1) (model.py:184-186)
// logits[p] = sum_c x[p,c] * w[c] + b   (one warp per pixel, fp32 out)
template <typename T>
__global__ void __launch_bounds__(256) conv1x1_to1_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ b, float* __restrict__ z,
                                                              int npix, int C) {
  pdl_grid_sync();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= npix) return;
  float acc = 0.f;
  for (int c = lane * 2; c < C; c += 64) {
    const float2 f = Cvt<T>::unpack2(*reinterpret_cast<const uint32_t*>(x + (size_t)warp * C + c));
    acc = fmaf(f.x, w[c], acc);
    acc = fmaf(f.y, w[c + 1], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) z[warp] = acc + b[0];
}
// dx[p,c] = dz[p] * w[c];   dw[c] += sum_p dz[p] x[p,c];   db += sum_p dz[p]   (grid-stride over pixels)
template <typename T>
__global__ void __launch_bounds__(256) conv1x1_to1_bwd_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ dz, T* __restrict__ dx,
                                                              float* __restrict__ dw, float* __restrict__ db,
                                                              int npix, int C, DetRed red) {
  pdl_grid_sync();
  // thread = channel (C <= 1024 handled by stride), block handles a pixel range
  const int per = (npix + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per, p1 = min(npix, p0 + per);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float wc = w[c];
    float acc = 0.f;
    for (int p = p0; p < p1; ++p) {
      const float g = dz[p];
      acc = fmaf(g, Cvt<T>::to_f(x[(size_t)p * C + c]), acc);
      if (dx) dx[(size_t)p * C + c] = Cvt<T>::from_f(g * wc);
    }
    if (dw) det_add(red, c, acc);
  }
  if (db && threadIdx.x == 0) {
    float s = 0.f;
    for (int p = p0; p < p1; ++p) s += dz[p];
    det_add(red, C, s);
  }
  if ((dw || db) && det_arrive(red)) {            // deterministic cross-block sums (fsr_common.cuh DetRed)
    if (dw) det_collect(red, 0, C, dw);
    if (db) det_collect(red, C, 1, db);
    det_release(red);
  }
}

// ------------------------------------------------------------------ losses (trainer.py:41,43,175-179,187-188,192)
// BCEWithLogitsLoss(mean): loss = mean(max(z,0) - z*t + log1p(exp(-|z|))), t = lab_scale*noise + lab_shift;
// dz = grad_scale * (sigmoid(z) - t) / n.   Single block (n = B*36 is tiny).  loss_out += weight * loss.
__global__ void __launch_bounds__(256) bce_logits_kernel(const float* __restrict__ z, const float* __restrict__ noise,
                                                         float lab_scale, float lab_shift, int n, float* __restrict__ loss_out,
                                                         float* __restrict__ dz, float grad_scale) {
  pdl_grid_sync();
  __shared__ float sm[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float zi = z[i], t = lab_scale * noise[i] + lab_shift;
    acc += fmaxf(zi, 0.f) - zi * t + log1pf(expf(-fabsf(zi)));
    if (dz) dz[i] = grad_scale * (1.0f / (1.0f + expf(-zi)) - t) / (float)n;
  }
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) loss_out[0] = acc / (float)n;
}

// SmoothL1Loss(beta=1, mean) between features a (fake) and b (real): partial sums + gradient wrt a
template <typename T>
__global__ void __launch_bounds__(256) smooth_l1_kernel(const T* __restrict__ a, const T* __restrict__ b, size_t n,
                                                        float* __restrict__ loss_acc /* += sum */, T* __restrict__ da,
                                                        float grad_scale /* = weight / n */, DetRed red) {
  pdl_grid_sync();
  __shared__ float sm[32];
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float d = Cvt<T>::to_f(a[i]) - Cvt<T>::to_f(b[i]);
    const float ad = fabsf(d);
    acc += ad < 1.f ? 0.5f * d * d : ad - 0.5f;
    if (da) da[i] = Cvt<T>::from_f(grad_scale * (ad < 1.f ? d : (d > 0.f ? 1.f : -1.f)));
  }
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) det_add(red, 0, acc);
  det_finish(red, loss_acc, 1);
}
// same with fp32 NCHW operands (pretrain step, trainer.py:109): a = generator output, b = hr images
__global__ void __launch_bounds__(256) smooth_l1_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n,
                                                            float* __restrict__ loss_acc, float* __restrict__ da, float grad_scale,
                                                            DetRed red) {
  pdl_grid_sync();
  __shared__ float sm[32];
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    const float ad = fabsf(d);
    acc += ad < 1.f ? 0.5f * d * d : ad - 0.5f;
    if (da) da[i] = grad_scale * (ad < 1.f ? d : (d > 0.f ? 1.f : -1.f));
  }
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) det_add(red, 0, acc);
  det_finish(red, loss_acc, 1);
}

// ------------------------------------------------------------------ InstanceNorm (+activation) backward
// y = act(xhat), xhat = (raw - mean) * rstd.   g = dY * act'(xhat);   dRaw = rstd * (g - mean(g) - xhat * mean(g xhat))
// pass 1: per (n,c) sums S1 = sum g, S2 = sum g*xhat  (+ PReLU slope gradient sum_{xhat<0} dY*xhat)
struct InBwdParams {
  const void* raw;       // conv output (pre-norm) NHWC
  const long long* stats; // [N][C][2] fixed-point sum, sumsq from the forward
  const void* dy;        // gradient w.r.t. the block output NHWC
  float* red;            // [N][C][2] S1, S2 (zeroed by the caller)
  void* draw;            // pass 2 output NHWC
  const float* alpha;    // PReLU slope pointer (act == ACT_PRELU)
  float* dalpha;         // += sum (act == ACT_PRELU), may be null
  float slope;
  int act;
  int HW, C;
  float eps;
  int dy_parity_w;       // > 0: `dy` is in the parity-plane layout [N][4][H/2][W/2][C] a stride-2 data gradient writes
                         //      (image width W = dy_parity_w): folds fsr_parity_layout(from parity) into this pass (fused kernel)
  DetRed det;            // fused kernel: slot for the deterministic PReLU-slope sum
};

template <typename T, int PASS>
__global__ void __launch_bounds__(256) instnorm_bwd_kernel(const InBwdParams p) {
  pdl_grid_sync();
  extern __shared__ float s_all[];   // mean[C], rstd[C], (pass 1: acc1[C], acc2[C]) (pass 2: m1[C], m2[C])
  float* s_mean = s_all;
  float* s_rstd = s_all + p.C;
  float* s_a = s_all + 2 * p.C;
  float* s_b = s_all + 3 * p.C;
  const int n = blockIdx.y;
  const float inv_hw = 1.0f / (float)p.HW;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    stat_mean_rstd(p.stats + ((size_t)n * p.C + c) * 2, 1.0 / (double)p.HW, p.eps, s_mean[c], s_rstd[c]);
    if (PASS == 1) { s_a[c] = 0.f; s_b[c] = 0.f; }
    else {
      s_a[c] = p.red[((size_t)n * p.C + c) * 2 + 0] * inv_hw;
      s_b[c] = p.red[((size_t)n * p.C + c) * 2 + 1] * inv_hw;
    }
  }
  __syncthreads();
  const float slope = (p.act == ACT_PRELU) ? __ldg(p.alpha) : p.slope;
  const bool has_act = (p.act == ACT_PRELU || p.act == ACT_LRELU);
  const int vpp = p.C / 8;
  const size_t nvec = (size_t)p.HW * vpp;
  const uint4* raw = reinterpret_cast<const uint4*>(p.raw) + (size_t)n * nvec;
  const uint4* dy = reinterpret_cast<const uint4*>(p.dy) + (size_t)n * nvec;
  uint4* draw = reinterpret_cast<uint4*>(p.draw) + (size_t)n * nvec;
  // each thread keeps a fixed channel group: stride over pixels so that c0 is loop-invariant
  const int threads = gridDim.x * blockDim.x;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  float a1[8], a2[8], da = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { a1[k] = 0.f; a2[k] = 0.f; }
  // threads is a multiple of vpp (host guarantees) -> channel group fixed per thread
  const int c0 = (gtid % vpp) * 8;
  for (size_t i = gtid; i < nvec; i += threads) {
    const uint4 r = raw[i], g4 = dy[i];
    const uint32_t ru[4] = {r.x, r.y, r.z, r.w}, gu[4] = {g4.x, g4.y, g4.z, g4.w};
    uint32_t ou[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = Cvt<T>::unpack2(ru[k]), d = Cvt<T>::unpack2(gu[k]);
      const float xh0 = (f.x - s_mean[c0 + 2 * k]) * s_rstd[c0 + 2 * k];
      const float xh1 = (f.y - s_mean[c0 + 2 * k + 1]) * s_rstd[c0 + 2 * k + 1];
      float g0 = d.x, g1 = d.y;
      if (has_act) {
        if (xh0 < 0.f) { if (PASS == 1) da += d.x * xh0; g0 *= slope; }
        if (xh1 < 0.f) { if (PASS == 1) da += d.y * xh1; g1 *= slope; }
      }
      if (PASS == 1) {
        a1[2 * k] += g0; a2[2 * k] = fmaf(g0, xh0, a2[2 * k]);
        a1[2 * k + 1] += g1; a2[2 * k + 1] = fmaf(g1, xh1, a2[2 * k + 1]);
      } else {
        const float o0 = s_rstd[c0 + 2 * k] * (g0 - s_a[c0 + 2 * k] - xh0 * s_b[c0 + 2 * k]);
        const float o1 = s_rstd[c0 + 2 * k + 1] * (g1 - s_a[c0 + 2 * k + 1] - xh1 * s_b[c0 + 2 * k + 1]);
        ou[k] = Cvt<T>::pack2(o0, o1);
      }
    }
    if (PASS == 2) draw[i] = make_uint4(ou[0], ou[1], ou[2], ou[3]);
  }
  if (PASS == 1) {
    // block reduction without shared atomics (fp32 shared atomicAdd is a CAS loop; 32 threads per channel contended):
    // lanes with equal (lane % vpp) hold the same channel group when vpp divides 32 -> xor-shuffle over the other lane
    // bits, then one lane per group adds; for vpp > 32 (C > 256) every lane owns a distinct group already.
    const int lane = threadIdx.x & 31;
    if (vpp <= 32) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        for (int o = 16; o >= vpp; o >>= 1) {
          a1[k] += __shfl_xor_sync(0xffffffffu, a1[k], o);
          a2[k] += __shfl_xor_sync(0xffffffffu, a2[k], o);
        }
      }
    }
    if (vpp > 32 || lane < vpp) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { atomicAdd(&s_a[c0 + k], a1[k]); atomicAdd(&s_b[c0 + k], a2[k]); }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
      atomicAdd(&p.red[((size_t)n * p.C + c) * 2 + 0], s_a[c]);
      atomicAdd(&p.red[((size_t)n * p.C + c) * 2 + 1], s_b[c]);
    }
    if (p.act == ACT_PRELU && p.dalpha) {
      da = warp_sum(da);
      if ((threadIdx.x & 31) == 0) atomicAdd(p.dalpha, da);
    }
  }
}

// Single-launch form for the small planes of the training shapes (HW <= 4096): one block owns one sample's 16-channel
// group (a full 32-byte sector per pixel) - pass 1 sums, pass 2 writes dRaw.  Replaces 2 launches + 1 memset per call and
// keeps the per-(n,c) sums inside the block (fixed-order reduction, no atomics): the InstanceNorm backward is bitwise
// reproducible.  RES > 0 (HW <= 128 * RES, the 24x24 / 12x12 / 6x6 planes): every thread's pixels are loaded ONCE, all
// loads in flight together before the statistics are even decoded, and both passes run from registers - these launches
// sit on the critical path of the backward chains and were latency-bound (two dependent sweeps of 4-5 serial loads).
// RES = 0: two streaming sweeps (the second hits L1/L2).  Same arithmetic in the same order either way.
template <typename T, int RES>
__global__ void __launch_bounds__(256) instnorm_bwd_fused_kernel(const InBwdParams p) {
  pdl_grid_sync();
  __shared__ float s_red[8][2][16];                  // [warp][vector of the pixel][a1 0..7, a2 0..7]
  __shared__ float s_m[2][16];                       // mean(g), mean(g*xhat) per (vector, channel)
  __shared__ float s_mr[2][16];                      // mean, rstd of the block's 16 channels
  __shared__ float s_da[8];
  const int n = blockIdx.y, cg = blockIdx.x;          // sample, 16-channel group
  const int vec = threadIdx.x & 1, pl = threadIdx.x >> 1;   // 128 pixel lanes x 2 vectors of 8 channels
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = cg * 16 + vec * 8;
  const size_t base = ((size_t)n * p.HW) * p.C + c0;
  const T* raw = reinterpret_cast<const T*>(p.raw) + base;
  const T* dy = reinterpret_cast<const T*>(p.dy) + base;
  T* draw = reinterpret_cast<T*>(p.draw) + base;
  // pixel index of `dy` for NHWC pixel px (identity, or the parity-plane position of (y, x))
  const int pw = p.dy_parity_w, pw2 = pw >> 1, ph2 = pw ? (p.HW / pw) >> 1 : 0;
  auto dy_px = [&](int px) -> size_t {
    if (!pw) return (size_t)px;
    const int y = px / pw, x = px - y * pw;
    return (size_t)(((y & 1) * 2 + (x & 1)) * ph2 + (y >> 1)) * pw2 + (x >> 1);
  };
  constexpr int NR = RES > 0 ? RES : 1;
  uint4 rr[NR], gg[NR];
  if constexpr (RES > 0) {
#pragma unroll
    for (int u = 0; u < RES; ++u) {
      const int px = pl + u * 128;
      const bool ok = px < p.HW;
      rr[u] = ok ? *reinterpret_cast<const uint4*>(raw + (size_t)px * p.C) : make_uint4(0, 0, 0, 0);
      gg[u] = ok ? *reinterpret_cast<const uint4*>(dy + dy_px(px) * p.C) : make_uint4(0, 0, 0, 0);
    }
  }
  if (threadIdx.x < 16)                               // one fp64 decode per channel of the block instead of 8 per thread
    stat_mean_rstd(p.stats + ((size_t)n * p.C + cg * 16 + threadIdx.x) * 2, 1.0 / (double)p.HW, p.eps, s_mr[0][threadIdx.x],
                   s_mr[1][threadIdx.x]);
  __syncthreads();
  float mean[8], rstd[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { mean[k] = s_mr[0][vec * 8 + k]; rstd[k] = s_mr[1][vec * 8 + k]; }
  const float slope = (p.act == ACT_PRELU) ? __ldg(p.alpha) : p.slope;
  const bool has_act = (p.act == ACT_PRELU || p.act == ACT_LRELU);
  float a1[8], a2[8], da = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { a1[k] = 0.f; a2[k] = 0.f; }
  auto accum = [&](const uint4& r, const uint4& g4) {
    const uint32_t ru[4] = {r.x, r.y, r.z, r.w}, gu[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = Cvt<T>::unpack2(ru[k]), d = Cvt<T>::unpack2(gu[k]);
      const float xh0 = (f.x - mean[2 * k]) * rstd[2 * k], xh1 = (f.y - mean[2 * k + 1]) * rstd[2 * k + 1];
      float g0 = d.x, g1 = d.y;
      if (has_act) {
        if (xh0 < 0.f) { da += d.x * xh0; g0 *= slope; }
        if (xh1 < 0.f) { da += d.y * xh1; g1 *= slope; }
      }
      a1[2 * k] += g0; a2[2 * k] = fmaf(g0, xh0, a2[2 * k]);
      a1[2 * k + 1] += g1; a2[2 * k + 1] = fmaf(g1, xh1, a2[2 * k + 1]);
    }
  };
  if constexpr (RES > 0) {
#pragma unroll
    for (int u = 0; u < RES; ++u)
      if (pl + u * 128 < p.HW) accum(rr[u], gg[u]);
  } else {
    for (int px = pl; px < p.HW; px += 128)
      accum(*reinterpret_cast<const uint4*>(raw + (size_t)px * p.C), *reinterpret_cast<const uint4*>(dy + dy_px(px) * p.C));
  }
  // warp: lanes of equal parity hold the same channel vector -> butterfly over lane bits 1..4, then 8 warps through smem
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int o = 16; o >= 2; o >>= 1) {
      a1[k] += __shfl_xor_sync(0xffffffffu, a1[k], o);
      a2[k] += __shfl_xor_sync(0xffffffffu, a2[k], o);
    }
  }
  if (lane < 2) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { s_red[warp][lane][k] = a1[k]; s_red[warp][lane][8 + k] = a2[k]; }
  }
  if (p.act == ACT_PRELU && p.dalpha) {
    da = warp_sum(da);
    if (lane == 0) s_da[warp] = da;
  }
  __syncthreads();
  if (threadIdx.x < 32) {                              // thread = (vector, which sum, channel): fixed-order sum over the 8 warps
    const int v = threadIdx.x >> 4, i = threadIdx.x & 15;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[w][v][i];
    s_m[v][i] = t / (float)p.HW;
  }
  if (threadIdx.x == 32 && p.act == ACT_PRELU && p.dalpha) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s_da[i];
    det_add(p.det, 0, t);
  }
  __syncthreads();
  float m1[8], m2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { m1[k] = s_m[vec][k]; m2[k] = s_m[vec][8 + k]; }
  auto emit = [&](const uint4& r, const uint4& g4, int px) {
    const uint32_t ru[4] = {r.x, r.y, r.z, r.w}, gu[4] = {g4.x, g4.y, g4.z, g4.w};
    uint32_t ou[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = Cvt<T>::unpack2(ru[k]), d = Cvt<T>::unpack2(gu[k]);
      const float xh0 = (f.x - mean[2 * k]) * rstd[2 * k], xh1 = (f.y - mean[2 * k + 1]) * rstd[2 * k + 1];
      float g0 = d.x, g1 = d.y;
      if (has_act) {
        if (xh0 < 0.f) g0 *= slope;
        if (xh1 < 0.f) g1 *= slope;
      }
      ou[k] = Cvt<T>::pack2(rstd[2 * k] * (g0 - m1[2 * k] - xh0 * m2[2 * k]), rstd[2 * k + 1] * (g1 - m1[2 * k + 1] - xh1 * m2[2 * k + 1]));
    }
    *reinterpret_cast<uint4*>(draw + (size_t)px * p.C) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
  };
  if constexpr (RES > 0) {
#pragma unroll
    for (int u = 0; u < RES; ++u)
      if (pl + u * 128 < p.HW) emit(rr[u], gg[u], pl + u * 128);
  } else {
    for (int px = pl; px < p.HW; px += 128)
      emit(*reinterpret_cast<const uint4*>(raw + (size_t)px * p.C), *reinterpret_cast<const uint4*>(dy + dy_px(px) * p.C), px);
  }
  if (p.act == ACT_PRELU && p.dalpha) det_finish(p.det, p.dalpha, 1);   // order-independent slope-gradient total
}

// ------------------------------------------------------------------ plain activation backward (no norm)
// y = act(v) with v = conv + bias stored post-activation: sign(v) == sign(y) for slope > 0.
// dv = dy * (y >= 0 ? 1 : slope);  PReLU: dalpha += sum_{y<0} dy * y / slope
template <typename T>
__global__ void __launch_bounds__(256) act_bwd_kernel(const uint4* __restrict__ y, const uint4* __restrict__ dy,
                                                      uint4* __restrict__ dv, size_t nvec, const float* alpha, float slope_in,
                                                      int act, float* dalpha, DetRed red) {
  pdl_grid_sync();
  const float slope = (act == ACT_PRELU) ? __ldg(alpha) : slope_in;
  const float inv = slope != 0.f ? 1.0f / slope : 0.f;
  float da = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 a = y[i], g = dy[i];
    const uint32_t au[4] = {a.x, a.y, a.z, a.w}, gu[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = Cvt<T>::unpack2(au[k]), d = Cvt<T>::unpack2(gu[k]);
      float g0 = d.x, g1 = d.y;
      if (f.x < 0.f) { da += d.x * f.x * inv; g0 *= slope; }
      if (f.y < 0.f) { da += d.y * f.y * inv; g1 *= slope; }
      o[k] = Cvt<T>::pack2(g0, g1);
    }
    dv[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
  if (act == ACT_PRELU && dalpha) {
    da = warp_sum(da);
    if ((threadIdx.x & 31) == 0) det_add(red, 0, da);
    det_finish(red, dalpha, 1);
  }
}

// ------------------------------------------------------------------ UpSamplingBlock backward glue (model.py:39-40)
// U = PReLU(PixelShuffle(conv + bias)) stored NHWC [N,2H,2W,F]; dU same shape.
// dConv[n,y,x,q*F+c] = dU[n,2y+i,2x+j,c] * (U >= 0 ? 1 : alpha)   (q = 2i+j: the packed/permuted column order)
// (the sign of the pre-activation is taken from the stored output: valid for a positive slope, which is what
//  torch.nn.PReLU's 0.25 init stays at in practice; DESIGN.md section 9 lists this limitation)
template <typename T>
__global__ void __launch_bounds__(256) ps_prelu_bwd_kernel(const T* __restrict__ U, const T* __restrict__ dU,
                                                           T* __restrict__ dconv, int N, int H, int W, int F,
                                                           const float* __restrict__ alpha, float* dalpha, DetRed red) {
  pdl_grid_sync();
  const float slope = __ldg(alpha);
  const float inv = slope != 0.f ? 1.0f / slope : 0.f;
  const int vpc = F >> 3;                           // 8-channel vectors per output pixel
  const size_t total = (size_t)N * H * W * 4 * vpc;   // (pixel, q, 8-channel vector)
  float da = 0.f;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % vpc);
    const int q = (int)((idx / vpc) & 3);
    const size_t pix = idx / (4 * (size_t)vpc);
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const int n = (int)(pix / ((size_t)W * H));
    const size_t src = (((size_t)n * 2 * H + 2 * y + (q >> 1)) * (2 * W) + 2 * x + (q & 1)) * F + cv * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(U + src), g = *reinterpret_cast<const uint4*>(dU + src);
    const uint32_t au[4] = {a.x, a.y, a.z, a.w}, gu[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = Cvt<T>::unpack2(au[k]), d = Cvt<T>::unpack2(gu[k]);
      float g0 = d.x, g1 = d.y;
      if (f.x < 0.f) { da += d.x * f.x * inv; g0 *= slope; }
      if (f.y < 0.f) { da += d.y * f.y * inv; g1 *= slope; }
      o[k] = Cvt<T>::pack2(g0, g1);
    }
    *reinterpret_cast<uint4*>(dconv + pix * 4 * F + q * F + cv * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
  if (dalpha) {
    da = warp_sum(da);
    if ((threadIdx.x & 31) == 0) det_add(red, 0, da);
    det_finish(red, dalpha, 1);
  }
}

// ------------------------------------------------------------------ head backward glue (model.py:102-110)
// y = tanh(conv + b) fp32 NCHW [N,3,H,W];  dpre = dy * (1 - y^2)  (fp32 NCHW, feeds the 3->64 direct dgrad)
__global__ void __launch_bounds__(256) tanh_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                       float* __restrict__ dpre, size_t n) {
  pdl_grid_sync();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dpre[i] = dy[i] * (1.0f - y[i] * y[i]);
}

// ------------------------------------------------------------------ small-channel weight gradients
// A 3x3 conv with one 3-channel side (G/D neck: Cin = 3; G head: Cout = 3):
//   dW[c3][c64][r][s] = sum_{n,y,x} img[n,c3,y+dy,x+dx] * act[n,y,x,c64]
// with img fp32 NCHW (3 channels), act NHWC T (64 channels per group), (dy,dx) = (r-1,s-1) when the image is
// the conv INPUT (neck: img = x, act = dOut) and (1-r,1-s) when it is the OUTPUT gradient (head: img = dpre,
// act = x).  out[(c3*9 + tap)*C64 + c64] fp32 (+=).  Block = 64 threads-channels x pixel chunks.
template <typename T>
__global__ void __launch_bounds__(224) wgrad_c3_kernel(const float* __restrict__ img, const T* __restrict__ act,
                                                       float* __restrict__ out, int N, int H, int W, int C64, int flip,
                                                       int layout /*0: [27][C64]; 1: OIHW [3][C64][9] (head); 2: OIHW [C64][3][9] (neck)*/,
                                                       DetRed red) {
  pdl_grid_sync();
  // thread = (k = c3*9 + tap, 8-channel group): one image value + one 16-B activation vector -> 8 FMAs per pixel;
  // a block walks a contiguous pixel range and issues 8 atomics per thread at the end.
  const int k = threadIdx.x >> 3, cg = threadIdx.x & 7;
  if (k < 27) {
  const int c3 = k / 9, r = (k % 9) / 3, s = k % 3;
  const int dy = flip ? 1 - r : r - 1, dx = flip ? 1 - s : s - 1;
  const int cbase = blockIdx.y * 64 + cg * 8;
  const size_t total = (size_t)N * H * W;
  const size_t per = (total + gridDim.x - 1) / gridDim.x;
  const size_t p0 = (size_t)blockIdx.x * per, p1 = min(total, p0 + per);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  // 4 pixels per iteration, all 8 loads issued before the FMAs (latency bound loop); 32-bit incremental coordinates
  // (64-bit div/mod per pixel made the first version 5x slower than its memory traffic)
  const int ip0 = (int)p0, ip1 = (int)p1;
  int x = ip0 % W, y = (ip0 / W) % H, n = ip0 / (W * H);
  for (int pb = ip0; pb < ip1; pb += 4) {
    float v[4];
    uint4 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool in_range = pb + u < ip1;
      const int yy = y + dy, xx = x + dx;
      const bool ok = in_range && yy >= 0 && yy < H && xx >= 0 && xx < W;
      const int ii = ok ? ((n * 3 + c3) * H + yy) * W + xx : 0;
      v[u] = ok ? __ldg(img + ii) : 0.f;
      a[u] = *reinterpret_cast<const uint4*>(act + (size_t)(in_range ? pb + u : ip0) * C64 + cbase);
      if (++x == W) { x = 0; if (++y == H) { y = 0; ++n; } }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t au[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = Cvt<T>::unpack2(au[j]);
        acc[2 * j] = fmaf(v[u], f.x, acc[2 * j]);
        acc[2 * j + 1] = fmaf(v[u], f.y, acc[2 * j + 1]);
      }
    }
  }
  const int tap = k % 9;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c64 = cbase + j;
    size_t idx;
    if (layout == 1) idx = ((size_t)c3 * C64 + c64) * 9 + tap;
    else if (layout == 2) idx = ((size_t)c64 * 3 + c3) * 9 + tap;
    else idx = (size_t)k * C64 + c64;
    det_add(red, (int)idx, acc[j]);
  }
  }
  det_finish(red, out, 27 * C64);
}

// bias gradient: db[c] += sum over pixels of g[pix, c]  (g NHWC T, C channels)
template <typename T>
__global__ void __launch_bounds__(256) bias_grad_kernel(const T* __restrict__ g, float* __restrict__ db, size_t npix, int C,
                                                        int ps_perm /* g columns pixel-shuffle-permuted: col q*C/4+c <-> channel 4c+q */,
                                                        DetRed red) {
  pdl_grid_sync();
  // thread = (8-channel vector, pixel lane): 16-B loads, register accumulation, one smem + one global FIXED-POINT atomic
  // per channel (integer adds: the total does not depend on the order of lanes or blocks)
  extern __shared__ unsigned long long s_acc[];           // [C]
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_acc[c] = 0ull;
  __syncthreads();
  const int vpp = C / 8;                     // vectors per pixel
  const int lanes = blockDim.x / vpp;        // pixel lanes per block (host guarantees blockDim.x % vpp == 0)
  const int cv = threadIdx.x % vpp, pl = threadIdx.x / vpp;
  const size_t per = (npix + gridDim.x - 1) / gridDim.x;
  const size_t p0 = (size_t)blockIdx.x * per, p1 = min(npix, p0 + per);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (pl < lanes) {
    for (size_t p = p0 + pl; p < p1; p += lanes) {
      const uint4 v = *reinterpret_cast<const uint4*>(g + p * C + cv * 8);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = Cvt<T>::unpack2(u[k]);
        acc[2 * k] += f.x;
        acc[2 * k + 1] += f.y;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_acc[cv * 8 + k], det_fix(acc[k]));
  }
  __syncthreads();
  const int cq = C >> 2;
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(red.acc + (ps_perm ? 4 * (c % cq) + c / cq : c), s_acc[c]);
  det_finish(red, db, C);
}
// fp32 NCHW variant (head bias: g = dpre [N,3,H,W])
__global__ void __launch_bounds__(256) bias_grad_nchw_kernel(const float* __restrict__ g, float* __restrict__ db, int N, int C,
                                                             size_t HW, DetRed red) {
  pdl_grid_sync();
  __shared__ float sm[32];
  const int c = blockIdx.y;
  float acc = 0.f;
  for (int n = 0; n < N; ++n)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (size_t)gridDim.x * blockDim.x)
      acc += g[((size_t)n * C + c) * HW + i];
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) det_add(red, c, acc);
  det_finish(red, db, C);
}

// ------------------------------------------------------------------ fused flat AdamW (trainer.py:33-38,181,196)
// torch.optim.AdamW defaults: p *= (1 - lr*wd); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps).  g is pre-scaled by grad_scale (1/world for DDP means).
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps,
                                                    float wd, float bc1, float bc2_sqrt, float grad_scale) {
  pdl_grid_sync();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    pi -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
    p[i] = pi;
  }
}

// Device-side step counter variant (CUDA-graph friendly: nothing step-dependent is baked into launch parameters)
__global__ void step_inc_kernel(int* step) {
  pdl_grid_sync(); if (threadIdx.x == 0 && blockIdx.x == 0) *step += 1; }
__global__ void __launch_bounds__(256) adamw_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps,
                                                        float wd, const int* __restrict__ step, float grad_scale) {
  pdl_grid_sync();
  const float t = (float)(*step);
  const float bc1 = 1.0f - powf(b1, t);
  const float bc2_sqrt = sqrtf(1.0f - powf(b2, t));
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    pi -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
    p[i] = pi;
  }
}

}  // namespace fsr
