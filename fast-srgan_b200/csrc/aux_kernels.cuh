// aux_kernels.cuh - the rows SURVEY.md section 8 marks "next", both HBM/latency-bound CUDA-core kernels:
//   psnr_ssim_kernel      : validation metrics of reference trainer.py:46-51, 60-68 (torchmetrics 1.4.0 PSNR / SSIM,
//                           data_range 1, 11x11 gaussian sigma 1.5) fused into one pass over (prediction, target)
//   crop_resize_aa_kernel : NumpyImagesDataset.__getitem__ of reference dataloader.py:24-38 on a device-resident uint8
//                           image cache: HR crop + antialiased bicubic downscale (torch `_upsample_bicubic2d_aa`, the op
//                           torchvision v2.Resize dispatches to) + x/127.5 - 1, one block per (sample, channel)
#pragma once
#include "fsr_common.cuh"

namespace fsr {

FSR_DEVINL double block_sum_256(double v, double* s_tmp /*[8]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_tmp[w] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < 8; ++i) r += s_tmp[i];
  return r;   // valid in thread 0
}

// ------------------------------------------------------------------ PSNR + SSIM
// SSIM keeps only windows that lie fully inside the image (the reflect padding torchmetrics adds is cropped off again),
// i.e. a VALID 11x11 gaussian filter of x, y, x^2, y^2, xy -> one block = 16x16 window positions of one (image, channel)
// plane: 26x26 input tile in smem, separable filter (rows then columns), SSIM map value per thread, block sum -> one
// double atomic per block into ssim_sum[n].  The squared error of the pixels the block "owns" (its 16x16 input rows/cols,
// extended to the image edge for the last tile row / column) goes to sse[0] the same way.
struct MetricParams {
  const float* pred;     // fp32 NCHW
  const float* target;   // fp32 NCHW
  int N, C, H, W;
  float scale, shift;    // v -> scale*v + shift applied to both (trainer.py:64-66: (1 + v)/2)
  float c1, c2;
  float g[11];           // normalised 1-D gaussian taps
  double* sse;           // [1]  += sum (p - t)^2
  double* ssim_sum;      // [N]  += sum of the SSIM map over C x (H-10) x (W-10)
};

__global__ void __launch_bounds__(256) psnr_ssim_kernel(const MetricParams p) {
  pdl_grid_sync();
  __shared__ float sp[26][27], st[26][27];
  __shared__ float rw[5][26][17];
  __shared__ double s_tmp[8];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int plane = blockIdx.z, n = plane / p.C;
  const int ox0 = blockIdx.x * 16, oy0 = blockIdx.y * 16;
  const bool last_x = blockIdx.x == gridDim.x - 1, last_y = blockIdx.y == gridDim.y - 1;
  const float* P = p.pred + (size_t)plane * p.H * p.W;
  const float* T = p.target + (size_t)plane * p.H * p.W;
  double se = 0.0;
  for (int i = threadIdx.x; i < 26 * 26; i += 256) {
    const int r = i / 26, c = i % 26;
    const int iy = oy0 + r, ix = ox0 + c;
    float a = 0.f, b = 0.f;
    if (iy < p.H && ix < p.W) {
      a = fmaf(p.scale, __ldg(P + (size_t)iy * p.W + ix), p.shift);
      b = fmaf(p.scale, __ldg(T + (size_t)iy * p.W + ix), p.shift);
      if ((r < 16 || last_y) && (c < 16 || last_x)) {
        const double d = (double)a - (double)b;
        se += d * d;
      }
    }
    sp[r][c] = a;
    st[r][c] = b;
  }
  __syncthreads();
  // row pass: 26 rows x 16 columns x 5 maps
  for (int i = threadIdx.x; i < 26 * 16; i += 256) {
    const int r = i >> 4, c = i & 15;
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float a = sp[r][c + k], b = st[r][c + k], w = p.g[k];
      m0 = fmaf(w, a, m0);
      m1 = fmaf(w, b, m1);
      m2 = fmaf(w, a * a, m2);
      m3 = fmaf(w, b * b, m3);
      m4 = fmaf(w, a * b, m4);
    }
    rw[0][r][c] = m0; rw[1][r][c] = m1; rw[2][r][c] = m2; rw[3][r][c] = m3; rw[4][r][c] = m4;
  }
  __syncthreads();
  double ss = 0.0;
  if (oy0 + ty < p.H - 10 && ox0 + tx < p.W - 10) {
    float m[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) v = fmaf(p.g[k], rw[q][ty + k][tx], v);
      m[q] = v;
    }
    const float mu_pp = m[0] * m[0], mu_tt = m[1] * m[1], mu_pt = m[0] * m[1];
    const float s_p = fmaxf(m[2] - mu_pp, 0.f), s_t = fmaxf(m[3] - mu_tt, 0.f), s_pt = m[4] - mu_pt;
    ss = (double)(((2.f * mu_pt + p.c1) * (2.f * s_pt + p.c2)) / ((mu_pp + mu_tt + p.c1) * (s_p + s_t + p.c2)));
  }
  const double ss_b = block_sum_256(ss, s_tmp);
  const double se_b = block_sum_256(se, s_tmp);
  if (threadIdx.x == 0) {
    atomicAdd(p.ssim_sum + n, ss_b);
    atomicAdd(p.sse, se_b);
  }
}

// ------------------------------------------------------------------ HR crop + antialiased bicubic downscale
// grid = (B, 3), 256 threads.  dynamic smem: crop[hr][hr] uint8 | tmp[hr][lr] fp32.
// The tap table (first input index, tap count, normalised fp32 taps per output index; identical for rows and columns of a
// square crop) is built on the host exactly as ATen's `_compute_indices_min_size_weights_aa` does.
struct CropResizeParams {
  const uint8_t* cache;       // all images, uint8 CHW, back to back
  const long long* img_off;   // [n_images] byte offset of each image
  const int* img_h;           // [n_images]
  const int* img_w;           // [n_images]
  const int* samples;         // [B][3] = image index, crop_y, crop_x
  const int* tap_min;         // [lr]
  const int* tap_size;        // [lr]
  const float* tap_w;         // [lr][K]
  float* lr;                  // fp32 NCHW [B,3,lr,lr]
  float* hr;                  // fp32 NCHW [B,3,hr,hr]
  int B, lr_size, scale, K;
};

__global__ void __launch_bounds__(256) crop_resize_aa_kernel(const CropResizeParams p) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t s_raw[];
  const int hr = p.lr_size * p.scale, lr = p.lr_size;
  uint8_t* crop = s_raw;
  float* tmp = reinterpret_cast<float*>(s_raw + (((size_t)hr * hr + 15) / 16) * 16);
  const int b = blockIdx.x, c = blockIdx.y;
  const int idx = p.samples[3 * b];
  const int H = p.img_h[idx], W = p.img_w[idx];
  int cy = p.samples[3 * b + 1], cx = p.samples[3 * b + 2];
  cy = max(0, min(cy, H - hr));
  cx = max(0, min(cx, W - hr));
  const uint8_t* src = p.cache + p.img_off[idx] + ((size_t)c * H + cy) * W + cx;
  float* hro = p.hr + ((size_t)b * 3 + c) * hr * hr;
  for (int i = threadIdx.x; i < hr * hr; i += 256) {
    const int y = i / hr, x = i - y * hr;
    const uint8_t v = __ldg(src + (size_t)y * W + x);
    crop[i] = v;
    hro[i] = (float)v / 127.5f - 1.0f;           // dataloader.py:36
  }
  __syncthreads();
  for (int i = threadIdx.x; i < hr * lr; i += 256) {   // horizontal pass
    const int y = i / lr, j = i - y * lr;
    const int x0 = p.tap_min[j], ns = p.tap_size[j];
    const float* w = p.tap_w + (size_t)j * p.K;
    float acc = 0.f;
    for (int k = 0; k < ns; ++k) acc = fmaf(__ldg(w + k), (float)crop[y * hr + x0 + k], acc);
    tmp[i] = acc;
  }
  __syncthreads();
  float* lro = p.lr + ((size_t)b * 3 + c) * lr * lr;
  for (int i = threadIdx.x; i < lr * lr; i += 256) {   // vertical pass
    const int oy = i / lr, j = i - oy * lr;
    const int y0 = p.tap_min[oy], ns = p.tap_size[oy];
    const float* w = p.tap_w + (size_t)oy * p.K;
    float acc = 0.f;
    for (int k = 0; k < ns; ++k) acc = fmaf(__ldg(w + k), tmp[(y0 + k) * lr + j], acc);
    lro[i] = acc / 127.5f - 1.0f;                  // dataloader.py:37
  }
}

}  // namespace fsr
