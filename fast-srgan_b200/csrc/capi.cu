// capi.cu - extern "C" entry points of libfsr_b200.so (see include/fsr_b200.h) and host launchers.
#include "../../include/fsr_b200.h"
#include "conv3x3_tc.cuh"
#include "elementwise.cuh"

#include <cudaTypedefs.h>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

using namespace fsr;

namespace {

std::atomic<unsigned long long> g_launches{0};
int g_prof_kernel = FSR_K_NONE;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;

// Brackets one launch with an event pair when profiling of `kernel_id` is enabled.
struct LaunchScope {
  cudaStream_t st;
  cudaEvent_t stop = nullptr;
  LaunchScope(int kernel_id, cudaStream_t s) : st(s) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (kernel_id == g_prof_kernel && g_prof_events.size() < FSR_PROFILE_MAX) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess) {
        cudaEventRecord(a, st);
        stop = b;
        g_prof_events.emplace_back(a, b);
      }
    }
  }
  ~LaunchScope() {
    if (stop) cudaEventRecord(stop, st);
  }
};

inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? FSR_OK : FSR_ERR_CUDA_BASE - (int)e; }
#define FSR_CUDA(expr)                         \
  do {                                         \
    cudaError_t _e = (expr);                   \
    if (_e != cudaSuccess) return cuda_rc(_e); \
  } while (0)

// cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time libcuda dependency.
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

inline CUtensorMapDataType tm_dtype(int dtype) {
  return dtype == FSR_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
}

// NHWC activation [N,H,W,C] (2-byte elements), box {64 ch, bw, bh, 1}, 128B swizzle, zero OOB fill.
int make_act_map(CUtensorMap* tm, const void* ptr, int N, int H, int W, int C, int bw, int bh, int dtype) {
  auto enc = get_encode_fn();
  if (!enc) return FSR_ERR_NO_DRIVER;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, tm_dtype(dtype), 4, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FSR_OK : FSR_ERR_TENSORMAP;
}

// packed weights [rows][64] (2-byte elements), box {64, box_rows}
int make_w_map(CUtensorMap* tm, const void* ptr, int rows, int box_rows, int dtype) {
  auto enc = get_encode_fn();
  if (!enc) return FSR_ERR_NO_DRIVER;
  cuuint64_t gdim[2] = {64, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, tm_dtype(dtype), 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FSR_OK : FSR_ERR_TENSORMAP;
}

int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

int g_halo1 = -1;   // A-operand staging: 0 = 3 column-shifted halo tiles, 1 = single halo tile (see conv3x3_tc.cuh)
int halo_mode() {
  if (g_halo1 < 0) {
    const char* e = getenv("FSR_HALO1");
    g_halo1 = (e && e[0] == '0') ? 0 : 1;   // default: single halo tile
  }
  return g_halo1;
}

template <int NS, int EPI, typename T, bool HALO1>
int launch_conv(const void* x, const void* w_packed, int w_rows, ConvParams p, int dtype, cudaStream_t st) {
  using Cfg = ConvCfg<NS, HALO1>;
  using Geo = ConvGeo<HALO1>;
  auto kern = conv3x3_c64_kernel<NS, EPI, T, HALO1>;
  static bool attr_done = false;   // per template instantiation
  if (!attr_done) {
    FSR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  p.tiles_x = (p.W + Geo::TW - 1) / Geo::TW;
  p.tiles_y = (p.H + Geo::TH - 1) / Geo::TH;
  p.num_tiles = p.N * p.tiles_x * p.tiles_y;
  CUtensorMap tmx, tmw;
  int rc = make_act_map(&tmx, x, p.N, p.H, p.W, 64, Geo::kBoxW, Geo::kBoxH, dtype);
  if (rc) return rc;
  if ((rc = make_w_map(&tmw, w_packed, w_rows, NS, dtype))) return rc;
  int ctas_per_slice = num_sms() / p.num_slices;
  if (ctas_per_slice < 1) ctas_per_slice = 1;
  if (ctas_per_slice > p.num_tiles) ctas_per_slice = p.num_tiles;
  const int grid = ctas_per_slice * p.num_slices;
  constexpr int kid = EPI == EPI_RAW_STATS ? FSR_K_CONV_RES : EPI == EPI_PS_PRELU ? FSR_K_CONV_UP
                    : EPI == EPI_HEAD_TANH ? FSR_K_CONV_HEAD : FSR_K_CONV_BIAS_ACT;
  {
    LaunchScope scope(kid, st);
    kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(tmx, tmw, p);
  }
  return cuda_rc(cudaGetLastError());
}

template <int NS, int EPI, typename T>
int launch_conv_mode(const void* x, const void* w_packed, int w_rows, const ConvParams& p, int dtype, cudaStream_t st) {
  if (halo_mode()) return launch_conv<NS, EPI, T, true>(x, w_packed, w_rows, p, dtype, st);
  return launch_conv<NS, EPI, T, false>(x, w_packed, w_rows, p, dtype, st);
}

template <typename T>
int conv_dispatch(const void* x, const void* w_packed, void* out, const float* bias, float* stats, const float* alpha,
                  int N, int H, int W, int cout, int epilogue, int act, float slope, int out_u8, int dtype,
                  cudaStream_t st) {
  if (N <= 0 || H <= 0 || W <= 0) return FSR_ERR_BAD_SHAPE;
  ConvParams p{};
  p.N = N; p.H = H; p.W = W;
  p.out = out; p.bias = bias; p.stats = stats; p.alpha = alpha; p.slope = slope; p.act = act; p.out_u8 = out_u8;
  switch (epilogue) {
    case FSR_EPI_RAW_STATS: {
      if (cout % 64 || !stats) return FSR_ERR_BAD_ARG;
      p.cout_total = cout; p.num_slices = cout / 64;
      return launch_conv_mode<64, EPI_RAW_STATS, T>(x, w_packed, 9 * cout, p, dtype, st);
    }
    case FSR_EPI_BIAS_ACT: {
      if (cout % 64) return FSR_ERR_BAD_ARG;
      if (act == FSR_ACT_PRELU && !alpha) return FSR_ERR_BAD_ARG;
      p.cout_total = cout; p.num_slices = cout / 64;
      return launch_conv_mode<64, EPI_BIAS_ACT, T>(x, w_packed, 9 * cout, p, dtype, st);
    }
    case FSR_EPI_PS_PRELU: {
      if (cout != 256 || !alpha) return FSR_ERR_BAD_ARG;
      p.cout_total = 256; p.num_slices = 2;
      return launch_conv_mode<128, EPI_PS_PRELU, T>(x, w_packed, 9 * 256, p, dtype, st);
    }
    case FSR_EPI_HEAD_TANH: {
      if (cout != 16) return FSR_ERR_BAD_ARG;   // padded head: 3 real + 13 zero rows
      p.cout_total = 16; p.num_slices = 1;
      return launch_conv_mode<16, EPI_HEAD_TANH, T>(x, w_packed, 9 * 16, p, dtype, st);
    }
  }
  return FSR_ERR_BAD_ARG;
}

}  // namespace

extern "C" {

int fsr_abi_version(void) { return FSR_ABI_VERSION; }

const char* fsr_error_string(int code) {
  switch (code) {
    case FSR_OK: return "ok";
    case FSR_ERR_BAD_SHAPE: return "bad shape";
    case FSR_ERR_BAD_ARG: return "bad argument";
    case FSR_ERR_TENSORMAP: return "cuTensorMapEncodeTiled failed";
    case FSR_ERR_WORKSPACE: return "workspace too small";
    case FSR_ERR_NO_DRIVER: return "CUDA driver entry point cuTensorMapEncodeTiled unavailable";
    default:
      if (code <= FSR_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(FSR_ERR_CUDA_BASE - code));
      return "unknown error";
  }
}

int fsr_pack_conv3x3_weight(const float* w_oihw, const float* bias, void* w_packed, float* bias_packed, int cout,
                            int cin, int cout_pad, int ps_perm, int dtype, void* stream) {
  if (cout <= 0 || cin <= 0 || cout_pad < cout || (ps_perm && cout % 4)) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)9 * cout_pad * cin;
  const int blocks = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
  LaunchScope scope(FSR_K_NONE - 1, st);
  if (dtype == FSR_BF16)
    pack_conv3x3_weight_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(w_oihw, (__nv_bfloat16*)w_packed, cout, cin, cout_pad, ps_perm);
  else
    pack_conv3x3_weight_kernel<__half><<<blocks, 256, 0, st>>>(w_oihw, (__half*)w_packed, cout, cin, cout_pad, ps_perm);
  if (bias && bias_packed) permute_bias_ps_kernel<<<(cout_pad + 127) / 128, 128, 0, st>>>(bias, bias_packed, cout, cout_pad, ps_perm);
  return cuda_rc(cudaGetLastError());
}

int fsr_conv3x3_c64(const void* x, const void* w_packed, void* out, const float* bias, float* stats,
                    const float* alpha, int N, int H, int W, int cout, int epilogue, int act, float slope,
                    int out_u8, int dtype, void* stream) {
  if (!x || !w_packed || !out) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FSR_BF16)
    return conv_dispatch<__nv_bfloat16>(x, w_packed, out, bias, stats, alpha, N, H, W, cout, epilogue, act, slope, out_u8, dtype, st);
  return conv_dispatch<__half>(x, w_packed, out, bias, stats, alpha, N, H, W, cout, epilogue, act, slope, out_u8, dtype, st);
}

int fsr_neck_conv3x3(const void* x, const float* w, const float* bias, const float* alpha, void* out, int N, int H,
                     int W, int cout, int act, float slope, int in_u8, int vgg_norm, int dtype, void* stream) {
  if (!x || !w || !out || cout % 64 || N <= 0 || H <= 0 || W <= 0) return FSR_ERR_BAD_ARG;
  if (act == FSR_ACT_PRELU && !alpha) return FSR_ERR_BAD_ARG;
  NeckParams p{x, w, bias, alpha, out, N, H, W, cout, act, slope, in_u8, vgg_norm};
  const size_t total = (size_t)N * H * W;
  dim3 grid((unsigned)((total + 127) / 128), cout / 64);
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope scope(FSR_K_NECK, st);
  if (dtype == FSR_BF16) neck_conv3x3_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(p);
  else neck_conv3x3_kernel<__half><<<grid, 128, 0, st>>>(p);
  return cuda_rc(cudaGetLastError());
}

int fsr_instnorm_apply(const void* raw, const float* stats, const void* residual, void* out, const float* alpha,
                       int N, int HW, int C, int act, float slope, float eps, int dtype, void* stream) {
  if (!raw || !stats || !out || C % 8 || N <= 0 || HW <= 0) return FSR_ERR_BAD_ARG;
  if (act == FSR_ACT_PRELU && !alpha) return FSR_ERR_BAD_ARG;
  InApplyParams p{raw, stats, residual, out, alpha, slope, act, HW, C, eps};
  const size_t nvec = (size_t)HW * (C / 8);
  int bpi = (int)((nvec + 256 * 4 - 1) / (256 * 4));   // ~4 vectors per thread
  const int cap = (num_sms() * 8 + N - 1) / N;
  if (bpi > cap) bpi = cap;
  if (bpi < 1) bpi = 1;
  dim3 grid(bpi, N);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t sm = (size_t)2 * C * sizeof(float);
  LaunchScope scope(FSR_K_IN_APPLY, st);
  if (dtype == FSR_BF16) instnorm_apply_kernel<__nv_bfloat16><<<grid, 256, sm, st>>>(p);
  else instnorm_apply_kernel<__half><<<grid, 256, sm, st>>>(p);
  return cuda_rc(cudaGetLastError());
}

int fsr_pixel_shuffle2(const void* in, void* out, int N, int H, int W, int C, int dtype, void* stream) {
  if (!in || !out || C % 8 || N <= 0) return FSR_ERR_BAD_ARG;
  const size_t total = (size_t)N * H * W * (C / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope scope(FSR_K_NONE - 1, st);
  if (dtype == FSR_BF16)
    pixel_shuffle2_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, N, H, W, C);
  else
    pixel_shuffle2_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)in, (__half*)out, N, H, W, C);
  return cuda_rc(cudaGetLastError());
}

int fsr_nchw_f32_to_nhwc(const float* in, void* out, int N, int C, int HW, int dtype, void* stream) {
  if (!in || !out || N <= 0) return FSR_ERR_BAD_ARG;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N);
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope scope(FSR_K_NONE - 1, st);
  if (dtype == FSR_BF16) nchw_f32_to_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(in, (__nv_bfloat16*)out, N, C, HW);
  else nchw_f32_to_nhwc_kernel<__half><<<grid, 256, 0, st>>>(in, (__half*)out, N, C, HW);
  return cuda_rc(cudaGetLastError());
}

int fsr_nhwc_to_nchw_f32(const void* in, float* out, int N, int C, int HW, int dtype, void* stream) {
  if (!in || !out || N <= 0) return FSR_ERR_BAD_ARG;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N);
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope scope(FSR_K_NONE - 1, st);
  if (dtype == FSR_BF16) nhwc_to_nchw_f32_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, out, N, C, HW);
  else nhwc_to_nchw_f32_kernel<__half><<<grid, 256, 0, st>>>((const __half*)in, out, N, C, HW);
  return cuda_rc(cudaGetLastError());
}

int fsr_profile_enable(int kernel_id) {
  for (auto& e : g_prof_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  g_prof_events.clear();
  g_prof_kernel = kernel_id;
  return FSR_OK;
}

int fsr_profile_read(float* ms_out, int capacity) {
  int n = 0;
  for (auto& e : g_prof_events) {
    float ms = 0.f;
    if (cudaEventSynchronize(e.second) == cudaSuccess && cudaEventElapsedTime(&ms, e.first, e.second) == cudaSuccess &&
        ms_out && n < capacity)
      ms_out[n++] = ms;
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  g_prof_events.clear();
  return n;
}

unsigned long long fsr_launch_count(void) { return g_launches.load(); }

int fsr_set_halo_mode(int single_halo_tile) {
  g_halo1 = single_halo_tile ? 1 : 0;
  return FSR_OK;
}

// ------------------------------------------------------------------ Generator.forward (model.py:112-117)
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t fsr_generator_workspace_bytes(int N, int H, int W, int n_filters, int n_layers) {
  const size_t P = align_up((size_t)N * H * W * n_filters * 2, 1024);
  const size_t stats = align_up((size_t)(2 * n_layers + 1) * N * n_filters * 2 * sizeof(float), 1024);
  return 4 * P + 4 * P + 16 * P + stats + 4096;
}

int fsr_generator_forward(const FsrGeneratorParams* prm, const void* x, void* y, void* workspace, size_t ws_bytes,
                          int N, int H, int W, int in_u8, int out_u8, int group, void* stream) {
  if (!prm || !x || !y || !workspace) return FSR_ERR_BAD_ARG;
  if (prm->n_filters != 64 || prm->n_layers < 0 || prm->n_layers > FSR_MAX_LAYERS) return FSR_ERR_BAD_SHAPE;
  const int F = 64, L = prm->n_layers, dt = prm->dtype;
  if (ws_bytes < fsr_generator_workspace_bytes(N, H, W, F, L)) return FSR_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t P = align_up((size_t)N * H * W * F * 2, 1024);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  uint8_t* b_res = base;            // neck output (long-skip source)
  uint8_t* b_x = base + P;          // running residual-chain activation
  uint8_t* b_raw = base + 2 * P;    // raw conv output (pre InstanceNorm)
  uint8_t* b_y = base + 3 * P;      // normalised + PReLU intermediate
  uint8_t* b_u0 = base + 4 * P;     // [N,2H,2W,64]
  uint8_t* b_u1 = base + 8 * P;     // [N,4H,4W,64]
  float* b_stats = reinterpret_cast<float*>(base + 24 * P);
  const size_t stats_per_conv = (size_t)N * F * 2;
  FSR_CUDA(cudaMemsetAsync(b_stats, 0, (size_t)(2 * L + 1) * stats_per_conv * sizeof(float), st));

  if (group <= 0 || group > N) group = N;
  const size_t img_bytes = (size_t)H * W * F * 2;
  const size_t in_img = in_u8 ? (size_t)H * W * 3 : (size_t)H * W * 3 * sizeof(float);
  int rc;
  for (int n0 = 0; n0 < N; n0 += group) {
    const int nb = (N - n0 < group) ? (N - n0) : group;
    uint8_t* res = b_res + n0 * img_bytes;
    uint8_t* xb = b_x + n0 * img_bytes;
    uint8_t* raw = b_raw + n0 * img_bytes;
    uint8_t* yb = b_y + n0 * img_bytes;
    const uint8_t* xin = reinterpret_cast<const uint8_t*>(x) + n0 * in_img;
    // neck (model.py:75-78)
    if ((rc = fsr_neck_conv3x3(xin, prm->neck_w, prm->neck_b, prm->neck_alpha, res, nb, H, W, F, FSR_ACT_PRELU, 0.f,
                               in_u8, 0, dt, st)))
      return rc;
    const uint8_t* cur = res;
    for (int l = 0; l < L; ++l) {   // ResidualBlock.forward (model.py:67-69)
      float* s1 = b_stats + (size_t)(2 * l) * stats_per_conv + (size_t)n0 * F * 2;
      float* s2 = b_stats + (size_t)(2 * l + 1) * stats_per_conv + (size_t)n0 * F * 2;
      if ((rc = fsr_conv3x3_c64(cur, prm->stem_w1[l], raw, nullptr, s1, nullptr, nb, H, W, F, FSR_EPI_RAW_STATS, 0, 0.f, 0, dt, st))) return rc;
      if ((rc = fsr_instnorm_apply(raw, s1, nullptr, yb, prm->stem_alpha[l], nb, H * W, F, FSR_ACT_PRELU, 0.f, 1e-5f, dt, st))) return rc;
      if ((rc = fsr_conv3x3_c64(yb, prm->stem_w2[l], raw, nullptr, s2, nullptr, nb, H, W, F, FSR_EPI_RAW_STATS, 0, 0.f, 0, dt, st))) return rc;
      if ((rc = fsr_instnorm_apply(raw, s2, cur, xb, nullptr, nb, H * W, F, FSR_ACT_NONE, 0.f, 1e-5f, dt, st))) return rc;
      cur = xb;
    }
    {   // bottleneck + long skip (model.py:86-95, 115)
      float* sb = b_stats + (size_t)(2 * L) * stats_per_conv + (size_t)n0 * F * 2;
      if ((rc = fsr_conv3x3_c64(cur, prm->bott_w, raw, nullptr, sb, nullptr, nb, H, W, F, FSR_EPI_RAW_STATS, 0, 0.f, 0, dt, st))) return rc;
      if ((rc = fsr_instnorm_apply(raw, sb, res, xb, nullptr, nb, H * W, F, FSR_ACT_NONE, 0.f, 1e-5f, dt, st))) return rc;
    }
  }
  // upsampling x2 (model.py:39-40) and head (model.py:102-110) over the whole batch
  if ((rc = fsr_conv3x3_c64(b_x, prm->up_w[0], b_u0, prm->up_b[0], nullptr, prm->up_alpha[0], N, H, W, 256, FSR_EPI_PS_PRELU, 0, 0.f, 0, dt, st))) return rc;
  if ((rc = fsr_conv3x3_c64(b_u0, prm->up_w[1], b_u1, prm->up_b[1], nullptr, prm->up_alpha[1], N, 2 * H, 2 * W, 256, FSR_EPI_PS_PRELU, 0, 0.f, 0, dt, st))) return rc;
  if ((rc = fsr_conv3x3_c64(b_u1, prm->head_w, y, prm->head_b, nullptr, nullptr, N, 4 * H, 4 * W, 16, FSR_EPI_HEAD_TANH, 0, 0.f, out_u8, dt, st))) return rc;
  return FSR_OK;
}

}  // extern "C"
