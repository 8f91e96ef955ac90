// experimental.cu - kernels that compile for sm_100a but have NOT run on hardware yet, built into their OWN library
// (libfsr_b200_experimental.so) so that libfsr_b200.so stays exactly the binary the GPU suite validated.
// Nothing in the product loads this library; tests/test_experimental_gpu.py does, when FSR_TEST_EXPERIMENTAL=1.
#include "conv3x3_up_2cta.cuh"
#include "conv3x3_res_xf2.cuh"

#include <cudaTypedefs.h>

using namespace fsr;

namespace {

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

CUtensorMapDataType tm_type(int dtype) { return dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16; }

// NHWC activation [N,H,W,64], box {64, bw, bh, 1}, 128B swizzle, zero OOB fill (same as capi.cu make_act_map)
int act_map(CUtensorMap* tm, const void* ptr, int N, int H, int W, int bw, int bh, int dtype) {
  auto enc = encode_fn();
  if (!enc) return FSR_ERR_NO_DRIVER;
  cuuint64_t gdim[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(tm, tm_type(dtype), 4, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? FSR_OK : FSR_ERR_TENSORMAP;
}

// packed weights [rows][64], box {64, box_rows}
int w_map(CUtensorMap* tm, const void* ptr, int rows, int box_rows, int dtype) {
  auto enc = encode_fn();
  if (!enc) return FSR_ERR_NO_DRIVER;
  cuuint64_t gdim[2] = {64, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, tm_type(dtype), 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? FSR_OK : FSR_ERR_TENSORMAP;
}

template <typename T>
int launch_up_2cta(const void* x, const void* w_packed, ConvParams p, int dtype, cudaStream_t st) {
  using Cfg = Up2Cfg;
  using Geo = Cfg::Geo;
  auto kern = conv3x3_up_2cta_kernel<T>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  if (e != cudaSuccess) return FSR_ERR_CUDA_BASE - (int)e;
  p.tiles_x = (p.W + Geo::TW - 1) / Geo::TW;
  p.tiles_y = (p.H + Geo::TH - 1) / Geo::TH;
  p.num_tiles = p.N * p.tiles_x * p.tiles_y;
  CUtensorMap tmx, tmw;
  int rc = act_map(&tmx, x, p.N, p.H, p.W, Geo::kBoxW, Geo::kBoxH, dtype);
  if (rc) return rc;
  if ((rc = w_map(&tmw, w_packed, 9 * Cfg::kN, Cfg::kN / 2, dtype))) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int pairs = (p.num_tiles + 1) / 2;
  int clusters = sms / 2;
  if (clusters > pairs) clusters = pairs;
  if (clusters < 1) clusters = 1;
  kern<<<2 * clusters, Cfg::kThreads, Cfg::kSmemBytes, st>>>(tmx, tmw, p);   // __cluster_dims__(2,1,1)
  e = cudaGetLastError();
  return e == cudaSuccess ? FSR_OK : FSR_ERR_CUDA_BASE - (int)e;
}

template <typename T>
int launch_res_xf2(const void* x_raw, const void* w_packed, ConvParamsXf2 p, int dtype, cudaStream_t st) {
  using Cfg = ConvCfg<64, true>;
  using Geo = ConvGeo<true>;
  auto kern = conv3x3_c64_xf2_kernel<T>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  if (e != cudaSuccess) return FSR_ERR_CUDA_BASE - (int)e;
  p.tiles_x = (p.W + Geo::TW - 1) / Geo::TW;
  p.tiles_y = (p.H + Geo::TH - 1) / Geo::TH;
  p.num_tiles = p.N * p.tiles_x * p.tiles_y;
  p.ws = 1;
  CUtensorMap tmx, tmw, tmo;
  int rc = act_map(&tmx, x_raw, p.N, p.H, p.W, Geo::kBoxW, Geo::kBoxH, dtype);
  if (rc) return rc;
  if ((rc = w_map(&tmw, w_packed, 9 * 64, 64, dtype))) return rc;
  if ((rc = act_map(&tmo, p.out, p.N, p.H, p.W, Geo::TW, 32 / Geo::TW, dtype))) return rc;   // one epilogue warp's TMA store
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = sms < p.num_tiles ? sms : p.num_tiles;
  kern<<<grid, Cfg::kThreadsXf, Cfg::kSmemBytes, st>>>(tmx, tmw, tmo, p);
  e = cudaGetLastError();
  return e == cudaSuccess ? FSR_OK : FSR_ERR_CUDA_BASE - (int)e;
}

}  // namespace

extern "C" {

/* x_next = InstanceNorm(x_raw; in_stats) + res (written to x_out, NHWC [N,H,W,64]);  out, stats = conv3x3(x_next) RAW_STATS.
 * model.py:65+69 of block l folded into :47-54 of block l+1 (or the bottleneck :87-93).  x_out must alias neither res nor x_raw;
 * stats [N,64,2] int64 must be zeroed by the caller. */
int fsrx_conv3x3_c64_res_in(const void* x_raw, const int64_t* in_stats, float in_eps, const void* res, void* x_out,
                            const void* w_packed, void* out, int64_t* stats, int N, int H, int W, int dtype, void* stream) {
  if (!x_raw || !in_stats || !res || !x_out || !w_packed || !out || !stats) return FSR_ERR_BAD_ARG;
  if (x_out == res || x_out == x_raw || out == x_raw || N <= 0 || H <= 0 || W <= 0) return FSR_ERR_BAD_ARG;
  ConvParamsXf2 p{};
  p.N = N; p.H = H; p.W = W; p.out = out; p.stats = reinterpret_cast<long long*>(stats);
  p.cout_total = 64; p.num_slices = 1;
  p.in_stats = reinterpret_cast<const long long*>(in_stats); p.in_eps = in_eps; p.in_alpha = nullptr;
  p.in_res = res; p.x_out = x_out;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 1) return launch_res_xf2<__nv_bfloat16>(x_raw, w_packed, p, dtype, st);
  return launch_res_xf2<__half>(x_raw, w_packed, p, dtype, st);
}

/* Same contract as fsr_conv3x3_c64(..., FSR_EPI_PS_PRELU): x [N,H,W,64] NHWC, w_packed [9][256][64] (pixel-shuffle column
 * order), bias_packed [256], alpha device pointer -> out [N,2H,2W,64] = PReLU(PixelShuffle2(conv + bias)).  dtype 0 = fp16, 1 = bf16. */
int fsrx_conv3x3_up_2cta(const void* x, const void* w_packed, void* out, const float* bias_packed, const float* alpha,
                         int N, int H, int W, int dtype, void* stream) {
  if (!x || !w_packed || !out || !alpha || N <= 0 || H <= 0 || W <= 0) return FSR_ERR_BAD_ARG;
  ConvParams p{};
  p.N = N; p.H = H; p.W = W; p.out = out; p.bias = bias_packed; p.alpha = alpha; p.cout_total = 256; p.num_slices = 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 1) return launch_up_2cta<__nv_bfloat16>(x, w_packed, p, dtype, st);
  return launch_up_2cta<__half>(x, w_packed, p, dtype, st);
}

}  // extern "C"
