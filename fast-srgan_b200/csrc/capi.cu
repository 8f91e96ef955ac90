// capi.cu - extern "C" entry points of libfsr_b200.so (see include/fsr_b200.h) and host launchers.
#include "../../include/fsr_b200.h"
#include "conv3x3_tc.cuh"
#include "conv3x3_up_2cta.cuh"
#include "conv3x3_gen.cuh"
#include "conv3x3_gen_ws.cuh"
#include "conv3x3_gen_2cta.cuh"
#include "conv3x3_head.cuh"
#include "conv3x3_wgrad.cuh"
#include "train_kernels.cuh"
#include "elementwise.cuh"
#include "small_mma.cuh"
#include "aux_kernels.cuh"
#include "nccl_comm.cuh"
#include "precise.cuh"

#include <cudaTypedefs.h>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <mutex>
#include <new>
#include <vector>

using namespace fsr;

namespace {

// Caller-owned context (include/fsr_b200.h "contexts"): option overrides + the side streams / events of the sub-batch
// overlap.  Bound per THREAD (fsr_ctx_bind); an unbound thread uses the process-wide defaults below.
enum { kOptHalo1 = 0, kOptWs, kOptFuseIn, kOptFuseRes, kOptUp2Cta, kOptGenWs, kOptGen2Cta, kOptSmallMma, kOptInBwdFused,
       kOptOverlapStreams, kOptCount };
struct FsrCtxImpl {
  int opt[kOptCount];
  cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[4] = {nullptr, nullptr, nullptr, nullptr};
  FsrCtxImpl() { for (int& o : opt) o = -1; }
};
thread_local FsrCtxImpl* tl_ctx = nullptr;
inline int ctx_opt(int key) { return tl_ctx ? tl_ctx->opt[key] : -1; }

// Every kernel launch of the library goes through here.  With programmatic dependent launch on (FSR_PDL=1 or
// fsr_set_pdl(1); default OFF) the launch carries the programmatic-stream-serialization attribute: the grid may become
// resident while its predecessor drains and blocks in pdl_grid_sync() (fsr_common.cuh) until that grid has completed.
// Measured (same box, profiles/r02/pdl_ab.md): results identical, but every launch got ~1.7 us SLOWER - GAN step b64
// 6.57 -> 7.04 ms in one CUDA graph (284 programmatic edges), generator 7.90 -> 8.01 ms - so it stays an opt-in switch.
int g_pdl = -1;
inline int pdl_mode() {
  if (g_pdl < 0) {
    const char* e = getenv("FSR_PDL");
    g_pdl = (e && e[0] == '1') ? 1 : 0;
  }
  return g_pdl;
}
struct PdlLaunch {
  dim3 grid, block;
  size_t smem;
  cudaStream_t st;
  PdlLaunch(dim3 g, dim3 b, size_t s, cudaStream_t stream) : grid(g), block(b), smem(s), st(stream) {}
  template <typename... KArgs, typename... Args>
  void operator()(void (*kern)(KArgs...), Args&&... args) const {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_mode() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);   // errors surface through cudaGetLastError()
  }
};

std::atomic<unsigned long long> g_launches{0};
unsigned g_prof_mask = 0;            // bit k set: launches of kernel id k are bracketed with an event pair
struct ProfRec { int id; cudaEvent_t a, b; double flops; };
std::vector<ProfRec> g_prof_events;

// Brackets one launch with an event pair when profiling of `kernel_id` is enabled.
struct LaunchScope {
  cudaStream_t st;
  cudaEvent_t stop = nullptr;
  // flops: ALGORITHMIC work of the launch (2 * N * Ho * Wo * Cout * Cin * taps), reported with its time
  LaunchScope(int kernel_id, cudaStream_t s, double flops = 0.0) : st(s) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (kernel_id >= 0 && kernel_id < 32 && ((g_prof_mask >> kernel_id) & 1u) && g_prof_events.size() < FSR_PROFILE_MAX) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess) {
        cudaEventRecord(a, st);
        stop = b;
        g_prof_events.push_back(ProfRec{kernel_id, a, b, flops});
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

// same, but with an explicit element stride between images (parity-plane layouts)
int make_act_map_strided(CUtensorMap* tm, const void* ptr, int N, int H, int W, int C, long long img_stride_elems,
                         int bw, int bh, int dtype) {
  auto enc = get_encode_fn();
  if (!enc) return FSR_ERR_NO_DRIVER;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)img_stride_elems * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, tm_dtype(dtype), 4, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FSR_OK : FSR_ERR_TENSORMAP;
}

// PixelShuffle(2) output [N,2H,2W,64] viewed as [N][H][i:2][W][(j,c):128] (the up conv's store target): one box
// {64 ch, 8 px, 1, 4 rows, 1} = the 32 pixels of an epilogue warp at sub-position (i, j), start coordinate j*64 in dim 0
int make_ps_out_map(CUtensorMap* tm, const void* ptr, int N, int H, int W, int dtype, int F = 64) {
  auto enc = get_encode_fn();
  if (!enc) return FSR_ERR_NO_DRIVER;
  const cuuint64_t px = (cuuint64_t)F * 2;      // bytes of one output pixel (F channels)
  cuuint64_t gdim[5] = {(cuuint64_t)2 * F, (cuuint64_t)W, 2, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[4] = {2 * px, (cuuint64_t)2 * W * px, (cuuint64_t)4 * W * px, (cuuint64_t)2 * H * 2 * W * px};
  cuuint32_t box[5] = {64, 8, 1, 4, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, tm_dtype(dtype), 5, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FSR_OK : FSR_ERR_TENSORMAP;
}

// packed weights [rows][cin] (2-byte elements), box {64, box_rows}
int make_w_map(CUtensorMap* tm, const void* ptr, int rows, int box_rows, int dtype, int cin = 64) {
  auto enc = get_encode_fn();
  if (!enc) return FSR_ERR_NO_DRIVER;
  cuuint64_t gdim[2] = {(cuuint64_t)cin, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cin * 2};
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

// Pool of zeroed reduction slots for the deterministic cross-block sums (fsr_common.cuh DetRed).  Every launch that
// reduces takes the next slot round-robin; a slot is left zeroed by the launch that used it.  512 slots: a slot is only
// handed out again 512 reducing launches later (a training step has ~25), long after its previous user has finished -
// also when a step is a captured graph (the slot index is baked in; replays of one graph serialise).  Allocated per
// device on first use; the weight-pack entry points touch it too, so that it exists before any stream capture.
constexpr int kDetSlots = 512;
struct DetPool { unsigned long long* acc = nullptr; unsigned int* tickets = nullptr; };
DetPool g_det_pool[16];
std::atomic<unsigned> g_det_next{0};
std::mutex g_det_mutex;
int det_pool_ensure(DetPool** out = nullptr) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return FSR_ERR_BAD_ARG;
  DetPool& P = g_det_pool[dev];
  if (!P.acc) {
    std::lock_guard<std::mutex> lock(g_det_mutex);
    if (!P.acc) {
      unsigned long long* a = nullptr;
      unsigned int* t = nullptr;
      FSR_CUDA(cudaMalloc(&a, (size_t)kDetSlots * kDetSlotLen * sizeof(unsigned long long)));
      FSR_CUDA(cudaMalloc(&t, (size_t)kDetSlots * sizeof(unsigned int)));
      FSR_CUDA(cudaMemset(a, 0, (size_t)kDetSlots * kDetSlotLen * sizeof(unsigned long long)));
      FSR_CUDA(cudaMemset(t, 0, (size_t)kDetSlots * sizeof(unsigned int)));
      FSR_CUDA(cudaDeviceSynchronize());
      P.tickets = t;
      P.acc = a;
    }
  }
  if (out) *out = &P;
  return FSR_OK;
}
int det_slot(DetRed* r) {
  DetPool* P = nullptr;
  const int rc = det_pool_ensure(&P);
  if (rc) return rc;
  const unsigned s = g_det_next.fetch_add(1, std::memory_order_relaxed) % kDetSlots;
  r->acc = P->acc + (size_t)s * kDetSlotLen;
  r->ticket = P->tickets + s;
  return FSR_OK;
}

int g_halo1 = -1;   // A-operand staging: 0 = 3 column-shifted halo tiles, 1 = single halo tile (see conv3x3_tc.cuh)
int halo_mode() {
  if (ctx_opt(kOptHalo1) >= 0) return ctx_opt(kOptHalo1);
  if (g_halo1 < 0) {
    const char* e = getenv("FSR_HALO1");
    g_halo1 = (e && e[0] == '0') ? 0 : 1;   // default: single halo tile
  }
  return g_halo1;
}

int g_ws = -1;
int ws_mode() {
  if (ctx_opt(kOptWs) >= 0) return ctx_opt(kOptWs);
  if (g_ws < 0) {
    const char* e = getenv("FSR_WS");
    g_ws = (e && e[0] == '0') ? 0 : 1;   // default on: B200-measured ~5 % on the 64->64 conv, bit-identical results
  }
  return g_ws;
}

int g_fuse_in = -1;   // Generator.forward: 1 = the res-block's first InstanceNorm + PReLU is applied inside conv2's load path
int fuse_in_mode() {
  if (ctx_opt(kOptFuseIn) >= 0) return ctx_opt(kOptFuseIn);
  if (g_fuse_in < 0) {
    const char* e = getenv("FSR_FUSE_IN");
    g_fuse_in = (e && e[0] == '0') ? 0 : 1;   // default ON: 3832 vs 3707 frames/s on the same box
  }
  return g_fuse_in;
}

int g_fuse_res = -1;   // Generator.forward: 1 = bn2 + skip of block l is applied inside conv1 of block l+1 (XF == 2)
int fuse_res_mode() {
  if (ctx_opt(kOptFuseRes) >= 0) return ctx_opt(kOptFuseRes);
  if (g_fuse_res < 0) {
    const char* e = getenv("FSR_FUSE_RES");
    g_fuse_res = (e && e[0] == '1') ? 1 : 0;   // default OFF: measured 578 us vs 172 + 109 us unfused (profiles/r02)
  }
  return g_fuse_res;
}

int g_up_2cta = -1;    // 64 -> 256 upsampling conv: 1 = CTA-pair kernel (tcgen05 cta_group::2, conv3x3_up_2cta.cuh)
int up_2cta_mode() {
  if (ctx_opt(kOptUp2Cta) >= 0) return ctx_opt(kOptUp2Cta);
  if (g_up_2cta < 0) {
    const char* e = getenv("FSR_UP_2CTA");
    g_up_2cta = (e && e[0] == '0') ? 0 : 1;
  }
  return g_up_2cta;
}

int g_in_bwd_fused = -1;   // InstanceNorm backward: 1 = single-launch kernel for planes <= 64x64 (train_kernels.cuh)
int in_bwd_fused_mode() {
  if (ctx_opt(kOptInBwdFused) >= 0) return ctx_opt(kOptInBwdFused);
  if (g_in_bwd_fused < 0) {
    const char* e = getenv("FSR_IN_BWD_FUSED");
    g_in_bwd_fused = (e && e[0] == '0') ? 0 : 1;
  }
  return g_in_bwd_fused;
}

int g_small_mma = -1;   // 3-channel-sided convs (neck / wgrad_c3): 1 = mma.sync tensor-core kernels (small_mma.cuh), 0 = CUDA cores
int small_mma_mode() {
  if (ctx_opt(kOptSmallMma) >= 0) return ctx_opt(kOptSmallMma);
  if (g_small_mma < 0) {
    const char* e = getenv("FSR_SMALL_MMA");
    g_small_mma = (e && e[0] == '0') ? 0 : 1;
  }
  return g_small_mma;
}

// fsr_set_pair_rows: the calling thread's 64-channel convs run on pair-expanded weights (pairs.py) until switched off
thread_local int tl_pair_rows = 0;

template <int NS, int EPI, typename T, bool HALO1, int XF = 0>
int launch_conv(const void* x, const void* w_packed, int w_rows, ConvParams p, int dtype, cudaStream_t st) {
  using Cfg = ConvCfg<NS, HALO1>;
  using Geo = ConvGeo<HALO1>;
  auto kern = conv3x3_c64_kernel<NS, EPI, T, HALO1, XF>;
  static bool attr_done = false;   // per template instantiation
  if (!attr_done) {
    FSR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  p.tiles_x = (p.W + Geo::TW - 1) / Geo::TW;
  p.tiles_y = (p.H + Geo::TH - 1) / Geo::TH;
  p.num_tiles = p.N * p.tiles_x * p.tiles_y;
  p.pair_rows = tl_pair_rows;
  p.ws = NS >= 64 ? ws_mode() : 0;   // tcgen05.mma.ws needs N >= 64 (the 16-column head variant faults with it)
  {
    static int backoff = -1;
    if (backoff < 0) { const char* e = getenv("FSR_BACKOFF_NS"); backoff = e ? atoi(e) : 0; }
    p.backoff_ns = backoff;
  }
  CUtensorMap tmx, tmw, tmo;
  int rc = make_act_map(&tmx, x, p.N, p.H, p.W, 64, Geo::kBoxW, Geo::kBoxH, dtype);
  if (rc) return rc;
  if ((rc = make_w_map(&tmw, w_packed, w_rows, NS, dtype))) return rc;
  if (EPI == EPI_RAW_STATS || EPI == EPI_BIAS_ACT) {
    // one epilogue warp stores its 32 pixels (32/TW tile rows x TW pixels) x 64 channels per TMA store
    if ((rc = make_act_map(&tmo, p.out, p.N, p.H, p.W, p.cout_total, Geo::TW, 32 / Geo::TW, dtype))) return rc;
  } else {
    tmo = tmx;
  }
  int ctas_per_slice = num_sms() / p.num_slices;
  if (ctas_per_slice < 1) ctas_per_slice = 1;
  if (ctas_per_slice > p.num_tiles) ctas_per_slice = p.num_tiles;
  const int grid = ctas_per_slice * p.num_slices;
  constexpr int kid = EPI == EPI_RAW_STATS ? FSR_K_CONV_RES : EPI == EPI_PS_PRELU ? FSR_K_CONV_UP
                    : EPI == EPI_HEAD_TANH ? FSR_K_CONV_HEAD : FSR_K_CONV_BIAS_ACT;
  {
    LaunchScope scope(kid, st, 2.0 * p.N * p.H * p.W * (double)(EPI == EPI_HEAD_TANH ? 3 : p.cout_total) * 64 * 9);
    PdlLaunch(grid, XF ? Cfg::kThreadsXf : Cfg::kThreads, Cfg::kSmemBytes, st)(kern, tmx, tmw, tmo, p);
  }
  return cuda_rc(cudaGetLastError());
}

template <int NS, int EPI, typename T>
int launch_conv_mode(const void* x, const void* w_packed, int w_rows, const ConvParams& p, int dtype, cudaStream_t st) {
  if (halo_mode()) return launch_conv<NS, EPI, T, true>(x, w_packed, w_rows, p, dtype, st);
  return launch_conv<NS, EPI, T, false>(x, w_packed, w_rows, p, dtype, st);
}

// 64 -> 256 upsampling conv as a CTA-pair kernel (conv3x3_up_2cta.cuh): M = 256 (one 128-pixel tile per CTA), N = 256
template <typename T>
int launch_up_2cta(const void* x, const void* w_packed, ConvParams p, int dtype, cudaStream_t st) {
  using Cfg = Up2Cfg;
  using Geo = Cfg::Geo;
  auto kern = conv3x3_up_2cta_kernel<T>;
  static bool attr_done = false;
  if (!attr_done) {
    FSR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  p.pair_rows = tl_pair_rows;
  p.tiles_x = (p.W + Geo::TW - 1) / Geo::TW;
  p.tiles_y = (p.H + Geo::TH - 1) / Geo::TH;
  p.num_tiles = p.N * p.tiles_x * p.tiles_y;
  CUtensorMap tmx, tmw;
  int rc = make_act_map(&tmx, x, p.N, p.H, p.W, 64, Geo::kBoxW, Geo::kBoxH, dtype);
  if (rc) return rc;
  if ((rc = make_w_map(&tmw, w_packed, 9 * Cfg::kN, Cfg::kN / 2, dtype))) return rc;
  CUtensorMap tmo;
  if ((rc = make_ps_out_map(&tmo, p.out, p.N, p.H, p.W, dtype))) return rc;
  const int pairs = (p.num_tiles + 1) / 2;
  int clusters = num_sms() / 2;
  if (clusters > pairs) clusters = pairs;
  if (clusters < 1) clusters = 1;
  {
    LaunchScope scope(FSR_K_CONV_UP, st, 2.0 * p.N * p.H * p.W * 256.0 * 64 * 9);
    PdlLaunch(2 * clusters, Cfg::kThreads, Cfg::kSmemBytes, st)(kern, tmx, tmw, tmo, p);   // __cluster_dims__(2,1,1)
  }
  return cuda_rc(cudaGetLastError());
}

// 3-output-channel conv as 1x1 GEMM + shift-add epilogue (conv3x3_head.cuh)
template <typename T>
int launch_head(const void* x, const void* w_packed, ConvParams p, int dtype, cudaStream_t st, int cin = 64) {
  auto kern = conv3x3_head_kernel<T>;
  static bool attr_done = false;
  if (!attr_done) {
    FSR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, HeadCfg::kSmemBytes));
    attr_done = true;
  }
  p.tiles_x = (p.W + HeadCfg::TW - 1) / HeadCfg::TW;
  p.tiles_y = (p.H + HeadCfg::TH - 1) / HeadCfg::TH;
  p.num_tiles = p.N * p.tiles_x * p.tiles_y;
  CUtensorMap tmx;
  if (cin % 64 || cin > 64 * HeadCfg::kMaxKC) return FSR_ERR_BAD_SHAPE;
  int rc = make_act_map(&tmx, x, p.N, p.H, p.W, cin, HeadCfg::BW, HeadCfg::BH, dtype);
  if (rc) return rc;
  int grid = num_sms();
  if (grid > p.num_tiles) grid = p.num_tiles;
  {
    LaunchScope scope(FSR_K_CONV_HEAD, st, 2.0 * p.N * p.H * p.W * 3.0 * cin * 9);
    PdlLaunch(grid, HeadCfg::kThreads, HeadCfg::kSmemBytes, st)(kern, tmx, reinterpret_cast<const T*>(w_packed), p, cin);
  }
  return cuda_rc(cudaGetLastError());
}

template <typename T>
int conv_dispatch(const void* x, const void* w_packed, void* out, const float* bias, long long* stats, const float* alpha,
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
      if (up_2cta_mode() && halo_mode()) { p.cout_total = 256; p.num_slices = 1; return launch_up_2cta<T>(x, w_packed, p, dtype, st); }
      p.cout_total = 256; p.num_slices = 2;
      return launch_conv_mode<128, EPI_PS_PRELU, T>(x, w_packed, 9 * 256, p, dtype, st);
    }
    case FSR_EPI_F32: {
      if (cout % 64) return FSR_ERR_BAD_ARG;
      p.cout_total = cout; p.num_slices = cout / 64; p.bias = nullptr;
      if (!halo_mode()) return FSR_ERR_BAD_ARG;
      return launch_conv<64, EPI_F32, T, true>(x, w_packed, 9 * cout, p, dtype, st);
    }
    case FSR_EPI_HEAD_TANH: {
      if (cout != 16 || out_u8 < 0 || out_u8 > 3) return FSR_ERR_BAD_ARG;   // padded head: 3 real + 13 zero rows
      p.cout_total = 16; p.num_slices = 1;
      if (halo_mode()) return launch_head<T>(x, w_packed, p, dtype, st);
      return launch_conv_mode<16, EPI_HEAD_TANH, T>(x, w_packed, 9 * 16, p, dtype, st);
    }
  }
  return FSR_ERR_BAD_ARG;
}


// ------------------------------------------------------------------ general conv (conv3x3_gen.cuh)
int g_gen_ws = -1;   // general conv: 1 = weight-stationary over 4-tile groups (conv3x3_gen_ws.cuh), 0 = per-tile weight streaming
int gen_ws_mode() {
  if (ctx_opt(kOptGenWs) >= 0) return ctx_opt(kOptGenWs);
  if (g_gen_ws < 0) {
    const char* e = getenv("FSR_GEN_WS");
    g_gen_ws = (e && e[0] == '0') ? 0 : 1;
  }
  return g_gen_ws;
}

int g_gen_2cta = -1;   // general conv, Cout % 128 == 0: 1 = CTA-pair kernel with 128-wide slices (conv3x3_gen_2cta.cuh)
int gen_2cta_mode() {
  if (ctx_opt(kOptGen2Cta) >= 0) return ctx_opt(kOptGen2Cta);
  if (g_gen_2cta < 0) {
    const char* e = getenv("FSR_GEN_2CTA");
    g_gen_2cta = (e && e[0] == '0') ? 0 : 1;
  }
  return g_gen_2cta;
}

template <int EPI, typename T, int MAXTAPS>
int launch_gen(const CUtensorMap* maps, const CUtensorMap& tmw, GenParams& p, cudaStream_t st, int dtype) {
  using Cfg = GenCfg<MAXTAPS>;
  if (gen_2cta_mode() && p.cout_total % 128 == 0 && (!p.ps || (p.cout_total / 4) % 128 == 0)) {
    using C2 = Gen2Cfg<MAXTAPS>;
    auto k2 = conv3x3_gen_2cta_kernel<EPI, T, MAXTAPS>;
    static bool attr2_done = false;
    if (!attr2_done) {
      FSR_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, C2::kSmemBytes));
      attr2_done = true;
    }
    GenParams p2 = p;
    p2.num_slices = p.cout_total / 128;
    CUtensorMap tmo;
    int rc = p.ps ? make_ps_out_map(&tmo, p.out, p.N, p.Ho, p.Wo, dtype, p.cout_total / 4)
                  : make_act_map_strided(&tmo, p.out, p.N, p.Ho, p.Wo, p.cout_total, p.out_img_stride, 8, 4, dtype);
    if (rc) return rc;
    const int pairs = (p.num_tiles + 1) / 2;
    int cps = (num_sms() / 2) / p2.num_slices;      // clusters per 128-wide slice
    if (cps < 1) cps = 1;
    if (cps > pairs) cps = pairs;
    int ntaps2 = 0;
    for (int k = 0; k < p.nkinds; ++k) ntaps2 += p.kinds[k].ntaps;
    LaunchScope scope(FSR_K_CONV_GEN, st, 2.0 * p.N * p.Ho * p.Wo * (double)p.cout_total * p.cin * ntaps2);
    PdlLaunch(2 * cps * p2.num_slices, C2::kThreads, C2::kSmemBytes, st)(k2, maps[0], maps[1], maps[2], maps[3], tmw, tmo, p2);
    return cuda_rc(cudaGetLastError());
  }
  int ctas_per_slice = num_sms() / p.num_slices;
  if (ctas_per_slice < 1) ctas_per_slice = 1;
  if (ctas_per_slice > p.num_tiles) ctas_per_slice = p.num_tiles;
  const int grid = ctas_per_slice * p.num_slices;
  int ntaps = 0;
  for (int k = 0; k < p.nkinds; ++k) ntaps += p.kinds[k].ntaps;
  const double flops = 2.0 * p.N * p.Ho * p.Wo * (double)p.cout_total * p.cin * ntaps;
  if (gen_ws_mode()) {
    using WCfg = GenWsCfg<MAXTAPS>;
    auto wkern = conv3x3_gen_ws_kernel<EPI, T, MAXTAPS>;
    static bool wattr_done = false;
    if (!wattr_done) {
      FSR_CUDA(cudaFuncSetAttribute(wkern, cudaFuncAttributeMaxDynamicSharedMemorySize, WCfg::kSmemBytes));
      wattr_done = true;
    }
    LaunchScope scope(FSR_K_CONV_GEN, st, flops);
    PdlLaunch(grid, WCfg::kThreads, WCfg::kSmemBytes, st)(wkern, maps[0], maps[1], maps[2], maps[3], tmw, p);
    return cuda_rc(cudaGetLastError());
  }
  auto kern = conv3x3_gen_kernel<EPI, T, MAXTAPS>;
  static bool attr_done = false;
  if (!attr_done) {
    FSR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  {
    LaunchScope scope(FSR_K_CONV_GEN, st, flops);
    PdlLaunch(grid, Cfg::kThreads, Cfg::kSmemBytes, st)(kern, maps[0], maps[1], maps[2], maps[3], tmw, p);
  }
  return cuda_rc(cudaGetLastError());
}

template <typename T>
int gen_dispatch(const void* x, const void* w_packed, void* out, const float* bias, long long* stats, const float* alpha,
                 int N, int H, int W, int cin, int cout, int stride, int mode, int epilogue, int act, float slope,
                 int dtype, cudaStream_t st) {
  if (N <= 0 || H <= 0 || W <= 0 || cin % 64 || cout % 64) return FSR_ERR_BAD_SHAPE;
  if (stride != 1 && stride != 2) return FSR_ERR_BAD_ARG;
  if (stride == 2 && ((H | W) & 1)) return FSR_ERR_BAD_SHAPE;
  if (epilogue == FSR_EPI_RAW_STATS && !stats) return FSR_ERR_BAD_ARG;
  if (epilogue != FSR_EPI_RAW_STATS && epilogue != FSR_EPI_BIAS_ACT && epilogue != FSR_EPI_PS_PRELU) return FSR_ERR_BAD_ARG;
  GenParams p{};
  if (epilogue == FSR_EPI_PS_PRELU) {          // UpSamplingBlock with F != 64: bias + PReLU + pixel-shuffle scatter
    if (stride != 1 || mode != 0 || cout % 256 || !alpha) return FSR_ERR_BAD_ARG;
    p.ps = 1; act = FSR_ACT_PRELU; epilogue = FSR_EPI_BIAS_ACT;
  }
  p.N = N; p.cin = cin; p.cout_total = cout; p.num_slices = cout / 64;
  p.bias = bias; p.stats = stats; p.alpha = alpha; p.slope = slope; p.act = act;
  CUtensorMap maps[4], tmw;
  int rc;
  if ((rc = make_w_map(&tmw, w_packed, 9 * cout, 64, dtype, cin))) return rc;
  const int H2 = H / 2, W2 = W / 2;
  auto finish = [&](int Ho, int Wo, void* o, long long img_stride, bool taps9) -> int {
    p.Ho = Ho; p.Wo = Wo; p.out = o; p.out_img_stride = img_stride;
    p.tiles_x = (Wo + 7) / 8; p.tiles_y = (Ho + 15) / 16; p.num_tiles = N * p.tiles_x * p.tiles_y;
    if (epilogue == FSR_EPI_RAW_STATS)
      return taps9 ? launch_gen<EPI_RAW_STATS, T, 9>(maps, tmw, p, st, dtype) : launch_gen<EPI_RAW_STATS, T, 4>(maps, tmw, p, st, dtype);
    return taps9 ? launch_gen<EPI_BIAS_ACT, T, 9>(maps, tmw, p, st, dtype) : launch_gen<EPI_BIAS_ACT, T, 4>(maps, tmw, p, st, dtype);
  };
  if (stride == 1) {
    // forward: out[y,x] += W[r,s] x[y+r-1, x+s-1];  dgrad (transposed, unflipped pack): dX[y,x] += W[r,s]^T dY[y+1-r, x+1-s]
    if ((rc = make_act_map(&maps[0], x, N, H, W, cin, 10, 18, dtype))) return rc;
    maps[1] = maps[2] = maps[3] = maps[0];
    p.nkinds = 1;
    GenKind& K = p.kinds[0];
    K.map = 0; K.ntaps = 9; K.box_w = 10; K.box_rows = 180; K.dx = -1; K.dy = -1;
    for (int r = 0; r < 3; ++r)
      for (int s2 = 0; s2 < 3; ++s2) {
        K.taps[r * 3 + s2].wrow = r * 3 + s2;
        K.taps[r * 3 + s2].a_off = mode == 0 ? r * 10 + s2 : (2 - r) * 10 + (2 - s2);
      }
    return finish(H, W, out, p.ps ? (long long)4 * H * W * (cout / 4) : (long long)H * W * cout, true);
  }
  if (mode == 0) {
    // stride-2 forward: x is in parity-plane layout [N][4][H2][W2][cin]; input row 2y+r-1:
    //   r=0 -> odd plane row y-1, r=1 -> even plane row y, r=2 -> odd plane row y   (same for columns)
    const long long plane = (long long)H2 * W2 * cin;
    for (int pl = 0; pl < 4; ++pl)
      if ((rc = make_act_map_strided(&maps[pl], (const uint8_t*)x + (size_t)pl * plane * 2, N, H2, W2, cin, 4 * plane, 9, 17, dtype)))
        return rc;
    p.nkinds = 4;
    for (int pr = 0; pr < 2; ++pr)
      for (int ps = 0; ps < 2; ++ps) {
        GenKind& K = p.kinds[pr * 2 + ps];
        K.map = pr * 2 + ps; K.ntaps = 0; K.box_w = 9; K.box_rows = 153; K.dx = -1; K.dy = -1;
        for (int r = 0; r < 3; ++r)
          for (int s2 = 0; s2 < 3; ++s2)
            if ((r != 1) == pr && (s2 != 1) == ps) {
              K.taps[K.ntaps].wrow = r * 3 + s2;
              K.taps[K.ntaps].a_off = (r != 0) * 9 + (s2 != 0);
              ++K.ntaps;
            }
      }
    return finish(H2, W2, out, (long long)H2 * W2 * cout, false);
  }
  // stride-2 dgrad: x = dY NHWC [N,H2,W2,cin]; out = dX parity planes [N][4][H2][W2][cout] (H, W = dX size)
  //   dX[2y'+pr] : pr=0 -> r=1 reads dY[y'];  pr=1 -> r=0 reads dY[y'+1], r=2 reads dY[y']
  if ((rc = make_act_map(&maps[0], x, N, H2, W2, cin, 9, 17, dtype))) return rc;
  maps[1] = maps[2] = maps[3] = maps[0];
  const long long oplane = (long long)H2 * W2 * cout;
  for (int pr = 0; pr < 2; ++pr)
    for (int ps = 0; ps < 2; ++ps) {
      p.nkinds = 1;
      GenKind& K = p.kinds[0];
      K.map = 0; K.ntaps = 0; K.box_w = 9; K.box_rows = 153; K.dx = 0; K.dy = 0;
      for (int r = 0; r < 3; ++r)
        for (int s2 = 0; s2 < 3; ++s2)
          if ((r != 1) == pr && (s2 != 1) == ps) {
            K.taps[K.ntaps].wrow = r * 3 + s2;
            K.taps[K.ntaps].a_off = (r == 0) * 9 + (s2 == 0);
            ++K.ntaps;
          }
      T* o = reinterpret_cast<T*>(out) + (size_t)(pr * 2 + ps) * oplane;
      if ((rc = finish(H2, W2, o, 4 * oplane, false))) return rc;
    }
  return FSR_OK;
}


// ------------------------------------------------------------------ flat (padded-flattened) general conv
// 2-D map of a [rows][C] matrix (2-byte elements), box {64, box_rows}, 128B swizzle, zero OOB fill
int make_rows_map(CUtensorMap* tm, const void* ptr, long long rows, int C, int box_rows, int dtype) {
  auto enc = get_encode_fn();
  if (!enc) return FSR_ERR_NO_DRIVER;
  cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, tm_dtype(dtype), 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FSR_OK : FSR_ERR_TENSORMAP;
}

template <typename T>
int gen_flat_dispatch(const void* x, const void* w_packed, void* out, const float* bias, int N, int H, int W, int cin, int cout,
                      int mode, int act, float slope, int dtype, cudaStream_t st) {
  if (N <= 0 || H <= 0 || W <= 0 || cin % 64 || cout % 128) return FSR_ERR_BAD_SHAPE;
  const int W2 = W + 2, lead = W2 + 1, box_rows = 128 + 2 * W2 + 2;
  if (box_rows * 128 > Gen2Cfg<9>::kABytes || box_rows > 256) return FSR_ERR_BAD_SHAPE;      // W <= 25
  const long long Q = (long long)N * (H + 2) * W2;
  if (Q >= ((long long)1 << 31) - 256) return FSR_ERR_BAD_SHAPE;
  GenParams p{};
  p.N = N; p.Ho = H; p.Wo = W; p.cin = cin; p.cout_total = cout; p.num_slices = cout / 128;
  p.bias = bias; p.act = act; p.slope = slope; p.out = out; p.out_img_stride = (long long)(H + 2) * W2 * cout;
  p.tiles_x = 1; p.tiles_y = 1; p.num_tiles = (int)((Q + 127) / 128);
  p.flat = 1; p.flat_q = (int)Q; p.flat_lead = lead;
  p.nkinds = 1;
  GenKind& K = p.kinds[0];
  K.map = 0; K.ntaps = 9; K.box_w = 8; K.box_rows = box_rows; K.dx = 0; K.dy = 0;      // box_w 8: SBO = 1024 B (contiguous rows)
  for (int r = 0; r < 3; ++r)
    for (int s2 = 0; s2 < 3; ++s2) {
      K.taps[r * 3 + s2].wrow = r * 3 + s2;
      K.taps[r * 3 + s2].a_off = mode == 0 ? r * W2 + s2 : (2 - r) * W2 + (2 - s2);
    }
  CUtensorMap maps[4], tmw, tmo;
  int rc;
  if ((rc = make_rows_map(&maps[0], x, Q, cin, box_rows, dtype))) return rc;
  maps[1] = maps[2] = maps[3] = maps[0];
  if ((rc = make_w_map(&tmw, w_packed, 9 * cout, 64, dtype, cin))) return rc;
  if ((rc = make_rows_map(&tmo, out, Q, cout, 32, dtype))) return rc;
  using C2 = Gen2Cfg<9>;
  auto k2 = conv3x3_gen_2cta_kernel<EPI_BIAS_ACT, T, 9>;
  static bool attr_done = false;
  if (!attr_done) {
    FSR_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, C2::kSmemBytes));
    attr_done = true;
  }
  const int pairs = (p.num_tiles + 1) / 2;
  int cps = (num_sms() / 2) / p.num_slices;
  if (cps < 1) cps = 1;
  if (cps > pairs) cps = pairs;
  LaunchScope scope(FSR_K_CONV_GEN, st, 2.0 * N * H * W * (double)cout * cin * 9);
  PdlLaunch(2 * cps * p.num_slices, C2::kThreads, C2::kSmemBytes, st)(k2, maps[0], maps[1], maps[2], maps[3], tmw, tmo, p);
  return cuda_rc(cudaGetLastError());
}

// ------------------------------------------------------------------ weight gradient (conv3x3_wgrad.cuh)
// 5-D activation map [groups][N][H][W][C] (2-byte elements), box {64, bw, bh, 1, 1}; img / group strides in elements
int make_act_map5(CUtensorMap* tm, const void* ptr, int G, int N, int H, int W, int C, long long img_stride, long long grp_stride,
                  int bw, int bh, int dtype) {
  auto enc = get_encode_fn();
  if (!enc) return FSR_ERR_NO_DRIVER;
  cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)G};
  cuuint64_t gstr[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)img_stride * 2, (cuuint64_t)(G > 1 ? grp_stride : img_stride * N) * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, tm_dtype(dtype), 5, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FSR_OK : FSR_ERR_TENSORMAP;
}

size_t wgrad_workspace_bytes() { return (size_t)num_sms() * 5 * 128 * 64 * sizeof(float); }

// groups problems of identical shape: x / dy point at group 0, consecutive groups are x_grp_stride / dy_grp_stride elements
// apart; dw[g] = fp32 OIHW gradient of group g (accumulated).
template <typename T>
int wgrad_dispatch(const void* x, const void* dy, float* const* dw, int groups, long long x_grp_stride, long long dy_grp_stride,
                   int N, int H, int W, int cin, int cout, int stride, int ps_perm, float* workspace, size_t ws_bytes, int dtype,
                   cudaStream_t st) {
  if (N <= 0 || H <= 0 || W <= 0 || cin % 64 || cout % 64 || groups < 1 || groups > kWgradMaxGroups) return FSR_ERR_BAD_SHAPE;
  if (stride != 1 && stride != 2) return FSR_ERR_BAD_ARG;
  if (stride == 2 && (((H | W) & 1) || groups != 1)) return FSR_ERR_BAD_SHAPE;
  if (!workspace || ws_bytes < wgrad_workspace_bytes()) return FSR_ERR_WORKSPACE;
  WgradParams p{};
  const int Ho = H / stride, Wo = W / stride;
  p.N = N; p.Ho = Ho; p.Wo = Wo; p.cin = cin; p.cout = cout; p.partial = workspace; p.groups = groups; p.ps_perm = ps_perm;
  p.tiles_x = (Wo + 7) / 8; p.tiles_y = (Ho + 15) / 16; p.num_tiles = N * p.tiles_x * p.tiles_y;
  CUtensorMap mx[4], mdy;
  int rc;
  if ((rc = make_act_map5(&mdy, dy, groups, N, Ho, Wo, cout, (long long)Ho * Wo * cout, dy_grp_stride, 8, 16, dtype))) return rc;
  if (stride == 1) {
    p.nplanes = 1; p.box_w = 10; p.box_rows = 180; p.plane_bytes = 23552; p.dx = -1; p.dy = -1;
    if ((rc = make_act_map5(&mx[0], x, groups, N, H, W, cin, (long long)H * W * cin, x_grp_stride, 10, 18, dtype))) return rc;
    mx[1] = mx[2] = mx[3] = mx[0];
    for (int t = 0; t < 9; ++t) { p.tap_row[t] = (t / 3) * 10 + (t % 3); p.tap_id[t] = t; }
  } else {
    // X in parity planes [N][4][Ho][Wo][cin]; tap (r,s) -> plane ((r!=1),(s!=1)), row (r!=0)*9 + (s!=0)
    p.nplanes = 4; p.box_w = 9; p.box_rows = 153; p.plane_bytes = 20480; p.dx = -1; p.dy = -1;
    const long long plane = (long long)Ho * Wo * cin;
    for (int pl = 0; pl < 4; ++pl)
      if ((rc = make_act_map5(&mx[pl], (const uint8_t*)x + (size_t)pl * plane * 2, 1, N, Ho, Wo, cin, 4 * plane, 0, 9, 17, dtype)))
        return rc;
    // order taps by their absolute row inside the stage so that pair distances (LBO) are non-negative
    int n = 0;
    for (int pl = 0; pl < 4; ++pl)
      for (int r = 0; r < 3; ++r)
        for (int s2 = 0; s2 < 3; ++s2)
          if ((r != 1) * 2 + (s2 != 1) == pl) {
            p.tap_row[n] = pl * (20480 / 128) + (r != 0) * 9 + (s2 != 0);
            p.tap_id[n] = r * 3 + s2;
            ++n;
          }
  }
  p.tap_row[9] = p.tap_row[8]; p.tap_id[9] = p.tap_id[8];
  auto kern = conv3x3_wgrad_kernel<T>;
  static bool attr_done = false;
  if (!attr_done) {
    FSR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WgradCfg::kSmemBytes));
    attr_done = true;
  }
  const int npairs_all = groups * (cin / 64) * (cout / 64);
  int cpp = num_sms() / npairs_all;          // split-K factor: CTAs per (group, cin chunk, cout slice)
  if (cpp < 1) cpp = 1;
  if (cpp > p.num_tiles) cpp = p.num_tiles;
  // more pairs than SMs (512x512: 64 pairs x ...): the partial slots are indexed by blockIdx.x < npairs_all * cpp, and the
  // workspace holds num_sms slots -> run the pairs in waves of at most num_sms CTAs
  WgradReduceParams rp{};
  rp.partial = workspace; rp.cin = cin; rp.cout = cout; rp.ps_perm = ps_perm;
  for (int g = 0; g < groups; ++g) rp.dw[g] = dw[g];
  for (int t = 0; t < 9; ++t) rp.slot_of_tap[p.tap_id[t]] = t;
  if (npairs_all * cpp > num_sms()) return FSR_ERR_BAD_SHAPE;   // cannot happen for cin, cout <= 512 (64 pairs)
  rp.npairs_all = npairs_all; rp.ctas_per_pair = cpp;
  {
    LaunchScope scope(FSR_K_CONV_WGRAD, st, 2.0 * groups * N * Ho * Wo * (double)cin * cout * 9);
    PdlLaunch(npairs_all * cpp, WgradCfg::kThreads, WgradCfg::kSmemBytes, st)(kern, mx[0], mx[1], mx[2], mx[3], mdy, p);
  }
  FSR_CUDA(cudaGetLastError());
  {
    LaunchScope scope(FSR_K_NONE - 1, st);
    PdlLaunch(dim3(npairs_all, 8, 9), 256, 0, st)(wgrad_reduce_kernel, rp);
  }
  return cuda_rc(cudaGetLastError());
}

inline int ew_blocks(size_t n, int per_block = 256) {
  size_t b = (n + per_block - 1) / per_block;
  const size_t cap = (size_t)num_sms() * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
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
    case FSR_ERR_NO_NCCL: return "libnccl.so.2 could not be bound (dlopen)";
    case FSR_ERR_NCCL: return "NCCL call failed";
    default:
      if (code <= FSR_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(FSR_ERR_CUDA_BASE - code));
      return "unknown error";
  }
}

int fsr_pack_conv3x3_weight(const float* w_oihw, const float* bias, void* w_packed, float* bias_packed, int cout,
                            int cin, int cout_pad, int ps_perm, int dtype, void* stream) {
  { const int rc_pool = det_pool_ensure(); if (rc_pool) return rc_pool; }   // before any stream capture
  if (cout <= 0 || cin <= 0 || cout_pad < cout || (ps_perm && cout % 4)) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)9 * cout_pad * cin;
  const int blocks = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
  LaunchScope scope(FSR_K_NONE - 1, st);
  if (dtype == FSR_BF16)
    PdlLaunch(blocks, 256, 0, st)(pack_conv3x3_weight_kernel<__nv_bfloat16>, w_oihw, (__nv_bfloat16*)w_packed, cout, cin, cout_pad, ps_perm);
  else
    PdlLaunch(blocks, 256, 0, st)(pack_conv3x3_weight_kernel<__half>, w_oihw, (__half*)w_packed, cout, cin, cout_pad, ps_perm);
  if (bias && bias_packed) PdlLaunch((cout_pad + 127) / 128, 128, 0, st)(permute_bias_ps_kernel, bias, bias_packed, cout, cout_pad, ps_perm);
  return cuda_rc(cudaGetLastError());
}

int fsr_pack_multi(const FsrPackTask* tasks, int n, int dtype, void* stream) {
  { const int rc_pool = det_pool_ensure(); if (rc_pool) return rc_pool; }   // before any stream capture
  if (!tasks || n <= 0 || n > kPackMaxTasks) return FSR_ERR_BAD_ARG;
  PackMultiParams p{};
  p.n = n;
  int blocks = 0;
  for (int i = 0; i < n; ++i) {
    const FsrPackTask& t = tasks[i];
    if (!t.w || !t.out || t.cout <= 0 || t.cin <= 0 || t.pad < ((t.flags & 1) ? t.cin : t.cout)) return FSR_ERR_BAD_ARG;
    p.t[i] = PackTaskDev{t.w, t.out, t.bias, t.bias_out, t.row_scale, t.cout, t.cin, t.pad, t.flags};
    const size_t total = (size_t)9 * t.pad * ((t.flags & 1) ? t.cout : t.cin);
    int nb = (int)((total + 256 * 8 - 1) / (256 * 8));       // ~8 elements per thread
    if (nb < 1) nb = 1;
    if (nb > 64) nb = 64;
    p.block_begin[i] = blocks;
    blocks += nb;
  }
  p.block_begin[n] = blocks;
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope scope(FSR_K_NONE - 1, st);
  if (dtype == FSR_BF16) PdlLaunch(blocks, 256, 0, st)(pack_multi_kernel<__nv_bfloat16>, p);
  else PdlLaunch(blocks, 256, 0, st)(pack_multi_kernel<__half>, p);
  return cuda_rc(cudaGetLastError());
}

int fsr_conv3x3_c64(const void* x, const void* w_packed, void* out, const float* bias, int64_t* stats,
                    const float* alpha, int N, int H, int W, int cout, int epilogue, int act, float slope,
                    int out_u8, int dtype, void* stream) {
  if (!x || !w_packed || !out) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FSR_BF16)
    return conv_dispatch<__nv_bfloat16>(x, w_packed, out, bias, reinterpret_cast<long long*>(stats), alpha, N, H, W, cout, epilogue, act, slope, out_u8, dtype, st);
  return conv_dispatch<__half>(x, w_packed, out, bias, reinterpret_cast<long long*>(stats), alpha, N, H, W, cout, epilogue, act, slope, out_u8, dtype, st);
}

// ---- n_filters = 32 networks on "pixel-pair rows" (DESIGN.md 3.7): [N,H,W,32] viewed as [N,H,W/2,64] --------------
// The 64-channel kernels run unchanged on the pair grid with weights expanded on the host (fast_srgan_b200/pairs.py);
// the two entries below are what the pair view adds: folding the per-(parity, channel) InstanceNorm sums into
// per-channel ones, and a head that writes two rgb pixels per row.
__global__ void in_stats_fold_pair_kernel(long long* __restrict__ stats, int total) {
  pdl_grid_sync();
  // stats [N][64][2] of the pair grid; channel c of the 32-channel network lives in slots c and 32+c.  The consumers
  // divide by the PAIR count H*W/2, so both slots receive half the merged sums (exact: fixed-point integers, floor).
  const int i = blockIdx.x * blockDim.x + threadIdx.x;          // (n, c<32, k<2)
  if (i >= total) return;
  const int k = i & 1, c = (i >> 1) & 31, n = i >> 6;
  long long* a = stats + ((size_t)n * 64 + c) * 2 + k;
  long long* b = a + 64;
  const long long m = (*a + *b) >> 1;
  *a = m; *b = m;
}

int fsr_set_pdl(int on) {
  g_pdl = on ? 1 : 0;
  return FSR_OK;
}

int fsr_set_pair_rows(int on) {
  tl_pair_rows = on ? 1 : 0;
  return FSR_OK;
}

int fsr_in_stats_fold_pair(int64_t* stats, int N, void* stream) {
  if (!stats) return FSR_ERR_BAD_ARG;
  if (N <= 0) return FSR_ERR_BAD_SHAPE;
  const int total = N * 64;
  PdlLaunch((total + 127) / 128, 128, 0, (cudaStream_t)stream)(in_stats_fold_pair_kernel, reinterpret_cast<long long*>(stats), total);
  return cuda_rc(cudaGetLastError());
}

}  // extern "C"
template <typename T>
static int head_pair_dispatch(const void* x, const void* w_packed, void* out, const float* bias, int N, int H, int Wp, int out_u8,
                              int dtype, cudaStream_t st) {
  ConvParams p{};
  p.N = N; p.H = H; p.W = Wp;
  p.out = out; p.bias = bias; p.act = 2; p.out_u8 = out_u8;
  p.cout_total = 16; p.num_slices = 1;
  return launch_conv<16, EPI_HEAD_TANH, T, true>(x, w_packed, 9 * 16, p, dtype, st);
}
extern "C" {

int fsr_conv3x3_c64_head_pair(const void* x, const void* w_packed, void* out, const float* bias, int N, int H, int Wp,
                              int out_u8, int dtype, void* stream) {
  if (!x || !w_packed || !out) return FSR_ERR_BAD_ARG;
  if (N <= 0 || H <= 0 || Wp <= 0) return FSR_ERR_BAD_SHAPE;
  if (out_u8 < 0 || out_u8 > 1) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FSR_BF16) return head_pair_dispatch<__nv_bfloat16>(x, w_packed, out, bias, N, H, Wp, out_u8, dtype, st);
  return head_pair_dispatch<__half>(x, w_packed, out, bias, N, H, Wp, out_u8, dtype, st);
}

int fsr_conv3x3_c64_in(const void* x_raw, const int64_t* in_stats, const float* in_alpha, float in_eps, const void* w_packed,
                       void* out, int64_t* stats, int N, int H, int W, int dtype, void* stream) {
  if (!x_raw || !in_stats || !in_alpha || !w_packed || !out || !stats || x_raw == out) return FSR_ERR_BAD_ARG;
  if (N <= 0 || H <= 0 || W <= 0) return FSR_ERR_BAD_SHAPE;
  if (!halo_mode()) return FSR_ERR_BAD_ARG;      // exists for the single-halo-tile staging only
  cudaStream_t st = (cudaStream_t)stream;
  ConvParams p{};
  p.N = N; p.H = H; p.W = W; p.out = out; p.stats = reinterpret_cast<long long*>(stats);
  p.cout_total = 64; p.num_slices = 1;
  p.in_stats = reinterpret_cast<const long long*>(in_stats); p.in_alpha = in_alpha; p.in_eps = in_eps;
  if (dtype == FSR_BF16) return launch_conv<64, EPI_RAW_STATS, __nv_bfloat16, true, 1>(x_raw, w_packed, 9 * 64, p, dtype, st);
  return launch_conv<64, EPI_RAW_STATS, __half, true, 1>(x_raw, w_packed, 9 * 64, p, dtype, st);
}

int fsr_conv3x3_c64_res_in(const void* x_raw, const int64_t* in_stats, float in_eps, const void* res, void* x_out,
                           const void* w_packed, void* out, int64_t* stats, int N, int H, int W, int dtype, void* stream) {
  if (!x_raw || !in_stats || !res || !x_out || !w_packed || !out || !stats) return FSR_ERR_BAD_ARG;
  if (x_out == res || x_out == x_raw || out == x_raw || out == x_out || out == res) return FSR_ERR_BAD_ARG;
  if (N <= 0 || H <= 0 || W <= 0) return FSR_ERR_BAD_SHAPE;
  if (!halo_mode()) return FSR_ERR_BAD_ARG;      // exists for the single-halo-tile staging only
  cudaStream_t st = (cudaStream_t)stream;
  ConvParams p{};
  p.N = N; p.H = H; p.W = W; p.out = out; p.stats = reinterpret_cast<long long*>(stats);
  p.cout_total = 64; p.num_slices = 1;
  p.in_stats = reinterpret_cast<const long long*>(in_stats); p.in_eps = in_eps;
  p.in_res = res; p.x_out = x_out;
  if (dtype == FSR_BF16) return launch_conv<64, EPI_RAW_STATS, __nv_bfloat16, true, 2>(x_raw, w_packed, 9 * 64, p, dtype, st);
  return launch_conv<64, EPI_RAW_STATS, __half, true, 2>(x_raw, w_packed, 9 * 64, p, dtype, st);
}

int fsr_set_fuse_res(int on) {
  g_fuse_res = on < 0 ? -1 : (on ? 1 : 0);   // -1: back to the environment default (FSR_FUSE_RES)
  return FSR_OK;
}

int fsr_set_up_2cta(int on) {
  g_up_2cta = on < 0 ? -1 : (on ? 1 : 0);    // -1: back to the environment default (FSR_UP_2CTA)
  return FSR_OK;
}

int fsr_set_fuse_in(int on) {
  g_fuse_in = on < 0 ? -1 : (on ? 1 : 0);   // -1: back to the environment default (FSR_FUSE_IN)
  return FSR_OK;
}

int fsr_conv3x3_gen(const void* x, const void* w_packed, void* out, const float* bias, int64_t* stats,
                    const float* alpha, int N, int H, int W, int cin, int cout, int stride, int mode, int epilogue,
                    int act, float slope, int dtype, void* stream) {
  if (!x || !w_packed || !out) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FSR_BF16)
    return gen_dispatch<__nv_bfloat16>(x, w_packed, out, bias, reinterpret_cast<long long*>(stats), alpha, N, H, W, cin, cout, stride, mode, epilogue, act, slope, dtype, st);
  return gen_dispatch<__half>(x, w_packed, out, bias, reinterpret_cast<long long*>(stats), alpha, N, H, W, cin, cout, stride, mode, epilogue, act, slope, dtype, st);
}

int fsr_conv3x3_gen_flat(const void* x_padded, const void* w_packed, void* out_padded, const float* bias, int N, int H, int W,
                         int cin, int cout, int mode, int act, float slope, int dtype, void* stream) {
  if (!x_padded || !w_packed || !out_padded || x_padded == out_padded || (mode != 0 && mode != 1)) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FSR_BF16) return gen_flat_dispatch<__nv_bfloat16>(x_padded, w_packed, out_padded, bias, N, H, W, cin, cout, mode, act, slope, dtype, st);
  return gen_flat_dispatch<__half>(x_padded, w_packed, out_padded, bias, N, H, W, cin, cout, mode, act, slope, dtype, st);
}

int fsr_conv3x3_head(const void* x, const void* w_packed, void* out, const float* bias, int N, int H, int W, int cin,
                     int out_mode, int dtype, void* stream) {
  if (!x || !w_packed || !out || out_mode < 0 || out_mode > 3) return FSR_ERR_BAD_ARG;
  ConvParams p{};
  p.N = N; p.H = H; p.W = W; p.out = out; p.bias = bias; p.out_u8 = out_mode; p.cout_total = 16; p.num_slices = 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FSR_BF16) return launch_head<__nv_bfloat16>(x, w_packed, p, dtype, st, cin);
  return launch_head<__half>(x, w_packed, p, dtype, st, cin);
}

static int neck_impl(const void* x, const float* w, const float* bias, const float* alpha, void* out, int N, int H,
                     int W, int cout, int act, float slope, int in_u8, int vgg_norm, int pitch, int dtype, void* stream) {
  if (!x || !w || !out || cout % 64 || N <= 0 || H <= 0 || W <= 0) return FSR_ERR_BAD_ARG;
  if (act == FSR_ACT_PRELU && !alpha) return FSR_ERR_BAD_ARG;
  if (pitch && !small_mma_mode()) return FSR_ERR_BAD_ARG;
  NeckParams p{x, w, bias, alpha, out, N, H, W, cout, act, slope, in_u8, vgg_norm, pitch};
  const size_t total = (size_t)N * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope scope(FSR_K_NECK, st);
  if (small_mma_mode()) {            // one warp per 16-pixel row strip, >= 4 strips per warp, <= 3 blocks per SM
    const long long strips = (long long)N * H * ((W + 15) / 16);
    if (strips >= ((long long)1 << 31) - 1) return FSR_ERR_BAD_SHAPE;
    long long bx = (strips + 4 * kNeckWarps - 1) / (4 * kNeckWarps);
    if (bx > (long long)num_sms() * 3) bx = (long long)num_sms() * 3;
    dim3 grid((unsigned)bx, cout / 64);
#define FSR_NECK_MMA(T)                                                                                       \
  do {                                                                                                        \
    if (in_u8) PdlLaunch(grid, kNeckWarps * 32, 0, st)(neck_conv3x3_mma_kernel<T, true, false>, p);                  \
    else if (vgg_norm) PdlLaunch(grid, kNeckWarps * 32, 0, st)(neck_conv3x3_mma_kernel<T, false, true>, p);          \
    else PdlLaunch(grid, kNeckWarps * 32, 0, st)(neck_conv3x3_mma_kernel<T, false, false>, p);                       \
  } while (0)
    if (in_u8 && vgg_norm) return FSR_ERR_BAD_ARG;
    if (dtype == FSR_BF16) FSR_NECK_MMA(__nv_bfloat16);
    else FSR_NECK_MMA(__half);
#undef FSR_NECK_MMA
    return cuda_rc(cudaGetLastError());
  }
  if (total > (size_t)1 << 20) {     // large frames/batches: one thread per pixel
    dim3 grid((unsigned)((total + 127) / 128), cout / 64);
    if (dtype == FSR_BF16) PdlLaunch(grid, 128, 0, st)(neck_conv3x3_kernel<__nv_bfloat16, 1>, p);
    else PdlLaunch(grid, 128, 0, st)(neck_conv3x3_kernel<__half, 1>, p);
  } else {
    dim3 grid((unsigned)((2 * total + 255) / 256), cout / 64);
    if (dtype == FSR_BF16) PdlLaunch(grid, 256, 0, st)(neck_conv3x3_kernel<__nv_bfloat16, 2>, p);
    else PdlLaunch(grid, 256, 0, st)(neck_conv3x3_kernel<__half, 2>, p);
  }
  return cuda_rc(cudaGetLastError());
}

int fsr_neck_conv3x3(const void* x, const float* w, const float* bias, const float* alpha, void* out, int N, int H,
                     int W, int cout, int act, float slope, int in_u8, int vgg_norm, int dtype, void* stream) {
  return neck_impl(x, w, bias, alpha, out, N, H, W, cout, act, slope, in_u8, vgg_norm, 0, dtype, stream);
}

int fsr_neck_conv3x3_c32(const void* x, const float* w64, const float* bias64, const float* alpha, void* out, int N, int H,
                         int W, int act, float slope, int in_u8, int dtype, void* stream) {
  return neck_impl(x, w64, bias64, alpha, out, N, H, W, 64, act, slope, in_u8, 0, 32, dtype, stream);
}

static int instnorm_apply_impl(const void* raw, const int64_t* stats, const void* residual, void* out, const float* alpha,
                               int N, int HW, int C, int act, float slope, float eps, int parity_w, int dtype, void* stream) {
  if (!raw || !stats || !out || C % 8 || N <= 0 || HW <= 0) return FSR_ERR_BAD_ARG;
  if (act == FSR_ACT_PRELU && !alpha) return FSR_ERR_BAD_ARG;
  if (parity_w < 0 || (parity_w > 0 && ((parity_w & 1) || HW % parity_w || ((HW / parity_w) & 1) || out == raw))) return FSR_ERR_BAD_SHAPE;
  InApplyParams p{raw, reinterpret_cast<const long long*>(stats), residual, out, alpha, slope, act, HW, C, eps, parity_w};
  const size_t nvec = (size_t)HW * (C / 8);
  int bpi = (int)((nvec + 256 * 4 - 1) / (256 * 4));   // ~4 vectors per thread
  const int cap = (num_sms() * 8 + N - 1) / N;
  if (bpi > cap) bpi = cap;
  if (bpi < 1) bpi = 1;
  dim3 grid(bpi, N);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t sm = (size_t)2 * C * sizeof(float);
  LaunchScope scope(FSR_K_IN_APPLY, st);
  if (dtype == FSR_BF16) PdlLaunch(grid, 256, sm, st)(instnorm_apply_kernel<__nv_bfloat16>, p);
  else PdlLaunch(grid, 256, sm, st)(instnorm_apply_kernel<__half>, p);
  return cuda_rc(cudaGetLastError());
}

int fsr_instnorm_apply(const void* raw, const int64_t* stats, const void* residual, void* out, const float* alpha,
                       int N, int HW, int C, int act, float slope, float eps, int dtype, void* stream) {
  return instnorm_apply_impl(raw, stats, residual, out, alpha, N, HW, C, act, slope, eps, 0, dtype, stream);
}

int fsr_instnorm_apply_parity(const void* raw, const int64_t* stats, void* out, const float* alpha, int N, int H, int W, int C,
                              int act, float slope, float eps, int dtype, void* stream) {
  return instnorm_apply_impl(raw, stats, nullptr, out, alpha, N, H * W, C, act, slope, eps, W, dtype, stream);
}

int fsr_pixel_shuffle2(const void* in, void* out, int N, int H, int W, int C, int dtype, void* stream) {
  if (!in || !out || C % 8 || N <= 0) return FSR_ERR_BAD_ARG;
  const size_t total = (size_t)N * H * W * (C / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope scope(FSR_K_NONE - 1, st);
  if (dtype == FSR_BF16)
    PdlLaunch(blocks, 256, 0, st)(pixel_shuffle2_kernel<__nv_bfloat16>, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, N, H, W, C);
  else
    PdlLaunch(blocks, 256, 0, st)(pixel_shuffle2_kernel<__half>, (const __half*)in, (__half*)out, N, H, W, C);
  return cuda_rc(cudaGetLastError());
}

int fsr_nchw_f32_to_nhwc(const float* in, void* out, int N, int C, int HW, int dtype, void* stream) {
  if (!in || !out || N <= 0) return FSR_ERR_BAD_ARG;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N);
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope scope(FSR_K_NONE - 1, st);
  if (dtype == FSR_BF16) PdlLaunch(grid, 256, 0, st)(nchw_f32_to_nhwc_kernel<__nv_bfloat16>, in, (__nv_bfloat16*)out, N, C, HW);
  else PdlLaunch(grid, 256, 0, st)(nchw_f32_to_nhwc_kernel<__half>, in, (__half*)out, N, C, HW);
  return cuda_rc(cudaGetLastError());
}

int fsr_nhwc_to_nchw_f32(const void* in, float* out, int N, int C, int HW, int dtype, void* stream) {
  if (!in || !out || N <= 0) return FSR_ERR_BAD_ARG;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N);
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope scope(FSR_K_NONE - 1, st);
  if (dtype == FSR_BF16) PdlLaunch(grid, 256, 0, st)(nhwc_to_nchw_f32_kernel<__nv_bfloat16>, (const __nv_bfloat16*)in, out, N, C, HW);
  else PdlLaunch(grid, 256, 0, st)(nhwc_to_nchw_f32_kernel<__half>, (const __half*)in, out, N, C, HW);
  return cuda_rc(cudaGetLastError());
}

static void prof_clear() {
  for (auto& e : g_prof_events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  g_prof_events.clear();
}

int fsr_profile_enable(int kernel_id) {
  prof_clear();
  g_prof_mask = (kernel_id >= 0 && kernel_id < 32) ? (1u << kernel_id) : 0u;
  return FSR_OK;
}

int fsr_profile_enable_mask(unsigned mask) {
  prof_clear();
  g_prof_mask = mask;
  return FSR_OK;
}

int fsr_profile_read_ex(float* ms_out, int* ids_out, double* flops_out, int capacity) {
  int n = 0;
  for (auto& e : g_prof_events) {
    float ms = 0.f;
    if (cudaEventSynchronize(e.b) == cudaSuccess && cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess && ms_out && n < capacity) {
      if (ids_out) ids_out[n] = e.id;
      if (flops_out) flops_out[n] = e.flops;
      ms_out[n++] = ms;
    }
  }
  prof_clear();
  return n;
}

int fsr_profile_read_ids(float* ms_out, int* ids_out, int capacity) { return fsr_profile_read_ex(ms_out, ids_out, nullptr, capacity); }

int fsr_profile_read(float* ms_out, int capacity) { return fsr_profile_read_ids(ms_out, nullptr, capacity); }

unsigned long long fsr_launch_count(void) { return g_launches.load(); }

int fsr_set_ws_mode(int weight_stationary) {
  g_ws = weight_stationary < 0 ? -1 : (weight_stationary ? 1 : 0);   // -1: back to the environment default
  return FSR_OK;
}

int fsr_set_gen_2cta(int on) {
  g_gen_2cta = on < 0 ? -1 : (on ? 1 : 0);   // -1: back to the environment default (FSR_GEN_2CTA)
  return FSR_OK;
}

int fsr_set_gen_ws(int on) {
  g_gen_ws = on < 0 ? -1 : (on ? 1 : 0);   // -1: back to the environment default (FSR_GEN_WS)
  return FSR_OK;
}

int fsr_set_small_mma(int on) {
  g_small_mma = on < 0 ? -1 : (on ? 1 : 0);   // -1: back to the environment default (FSR_SMALL_MMA)
  return FSR_OK;
}

int fsr_set_halo_mode(int single_halo_tile) {
  g_halo1 = single_halo_tile ? 1 : 0;
  return FSR_OK;
}

// ------------------------------------------------------------------ Generator.forward (model.py:112-117)
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t fsr_generator_workspace_bytes(int N, int H, int W, int n_filters, int n_layers) {
  const size_t P = align_up((size_t)N * H * W * n_filters * 2, 1024);
  const size_t stats = align_up((size_t)(2 * n_layers + 1) * N * n_filters * 2 * sizeof(int64_t), 1024);
  return 5 * P + 4 * P + 16 * P + stats + 4096;
}

// Sub-batches of the forward run on internal side streams so that the HBM-bound kernels of one sub-batch
// (InstanceNorm apply, head) overlap the tensor-bound convolutions of the other: the persistent conv CTAs
// leave ~10 K registers and 5 KB smem per SM, enough for one elementwise block to co-reside.
int g_overlap = -1;
cudaStream_t g_side[4] = {nullptr, nullptr, nullptr, nullptr};
cudaEvent_t g_fork = nullptr, g_join[4] = {nullptr, nullptr, nullptr, nullptr};
int overlap_parts() {
  if (ctx_opt(kOptOverlapStreams) >= 1) return ctx_opt(kOptOverlapStreams) > 4 ? 4 : ctx_opt(kOptOverlapStreams);
  if (g_overlap < 0) {
    const char* e = getenv("FSR_STREAMS");
    g_overlap = e ? atoi(e) : 1;   // measured on B200 (power-capped): 2-4 sub-batches overlap but do not shorten the step
    if (g_overlap < 1) g_overlap = 1;
    if (g_overlap > 4) g_overlap = 4;
  }
  return g_overlap;
}

int fsr_set_overlap_streams(int parts) {
  g_overlap = parts < 1 ? 1 : (parts > 4 ? 4 : parts);
  return FSR_OK;
}

static int generator_chain(const FsrGeneratorParams* prm, const uint8_t* xin, uint8_t* yout, uint8_t* res, uint8_t* xa,
                           uint8_t* xb2, uint8_t* raw, uint8_t* yb, uint8_t* u0, uint8_t* u1, int64_t* stats, size_t stats_per_conv,
                           int nb, int H, int W, int in_u8, int out_u8, cudaStream_t st) {
  const int F = 64, L = prm->n_layers, dt = prm->dtype;
  int rc;
  // neck (model.py:75-78)
  if ((rc = fsr_neck_conv3x3(xin, prm->neck_w, prm->neck_b, prm->neck_alpha, res, nb, H, W, F, FSR_ACT_PRELU, 0.f, in_u8, 0, dt, st)))
    return rc;
  const bool fuse_in = fuse_in_mode() && halo_mode();
  const bool fuse_res = fuse_in && fuse_res_mode();
  int64_t* sb = stats + (size_t)(2 * L) * stats_per_conv;
  uint8_t* xlast;
  if (fuse_res && L > 0) {
    // Fully fused residual chain: per block TWO launches and no elementwise pass.
    //   conv1 of block 0 reads x_0 = neck output;  conv2 applies bn1 + relu1 in its load path (XF 1);
    //   conv1 of block l+1 - and the bottleneck conv - form x_{l+1} = bn2(c2_l) + x_l in their load path (XF 2) and write
    //   x_{l+1} back (ping-pong xa / xb2; x_0 stays in `res` for the long skip, model.py:115).
    const uint8_t* xprev = res;
    uint8_t* xnext = xa;
    for (int l = 0; l < L; ++l) {   // ResidualBlock.forward (model.py:67-69)
      int64_t* s1 = stats + (size_t)(2 * l) * stats_per_conv;
      int64_t* s2 = stats + (size_t)(2 * l + 1) * stats_per_conv;
      if (l == 0) {
        if ((rc = fsr_conv3x3_c64(res, prm->stem_w1[0], raw, nullptr, s1, nullptr, nb, H, W, F, FSR_EPI_RAW_STATS, 0, 0.f, 0, dt, st))) return rc;
      } else {
        int64_t* s2p = stats + (size_t)(2 * l - 1) * stats_per_conv;
        if ((rc = fsr_conv3x3_c64_res_in(yb, s2p, 1e-5f, xprev, xnext, prm->stem_w1[l], raw, s1, nb, H, W, dt, st))) return rc;
        xprev = xnext;
        xnext = (xnext == xa) ? xb2 : xa;
      }
      if ((rc = fsr_conv3x3_c64_in(raw, s1, prm->stem_alpha[l], 1e-5f, prm->stem_w2[l], yb, s2, nb, H, W, dt, st))) return rc;
    }
    // bottleneck (model.py:86-95): its input x_L = bn2(c2_{L-1}) + x_{L-1} is formed in the load path as well
    int64_t* s2p = stats + (size_t)(2 * L - 1) * stats_per_conv;
    if ((rc = fsr_conv3x3_c64_res_in(yb, s2p, 1e-5f, xprev, xnext, prm->bott_w, raw, sb, nb, H, W, dt, st))) return rc;
    xlast = (xnext == xa) ? xb2 : xa;      // free buffer for the bottleneck's normalised output
  } else {
    const uint8_t* cur = res;
    for (int l = 0; l < L; ++l) {   // ResidualBlock.forward (model.py:67-69)
      int64_t* s1 = stats + (size_t)(2 * l) * stats_per_conv;
      int64_t* s2 = stats + (size_t)(2 * l + 1) * stats_per_conv;
      if ((rc = fsr_conv3x3_c64(cur, prm->stem_w1[l], raw, nullptr, s1, nullptr, nb, H, W, F, FSR_EPI_RAW_STATS, 0, 0.f, 0, dt, st))) return rc;
      if (fuse_in) {
        // bn1 + relu1 (model.py:55-56) applied to the halo tile inside conv2's load path: one HBM round trip less per block
        if ((rc = fsr_conv3x3_c64_in(raw, s1, prm->stem_alpha[l], 1e-5f, prm->stem_w2[l], yb, s2, nb, H, W, dt, st))) return rc;
        if ((rc = fsr_instnorm_apply(yb, s2, cur, xa, nullptr, nb, H * W, F, FSR_ACT_NONE, 0.f, 1e-5f, dt, st))) return rc;
      } else {
        if ((rc = fsr_instnorm_apply(raw, s1, nullptr, yb, prm->stem_alpha[l], nb, H * W, F, FSR_ACT_PRELU, 0.f, 1e-5f, dt, st))) return rc;
        if ((rc = fsr_conv3x3_c64(yb, prm->stem_w2[l], raw, nullptr, s2, nullptr, nb, H, W, F, FSR_EPI_RAW_STATS, 0, 0.f, 0, dt, st))) return rc;
        if ((rc = fsr_instnorm_apply(raw, s2, cur, xa, nullptr, nb, H * W, F, FSR_ACT_NONE, 0.f, 1e-5f, dt, st))) return rc;
      }
      cur = xa;
    }
    // bottleneck (model.py:86-95)
    if ((rc = fsr_conv3x3_c64(cur, prm->bott_w, raw, nullptr, sb, nullptr, nb, H, W, F, FSR_EPI_RAW_STATS, 0, 0.f, 0, dt, st))) return rc;
    xlast = xb2;
  }
  // + long skip (model.py:115)
  if ((rc = fsr_instnorm_apply(raw, sb, res, xlast, nullptr, nb, H * W, F, FSR_ACT_NONE, 0.f, 1e-5f, dt, st))) return rc;
  // upsampling x2 (model.py:39-40) and head (model.py:102-110)
  if ((rc = fsr_conv3x3_c64(xlast, prm->up_w[0], u0, prm->up_b[0], nullptr, prm->up_alpha[0], nb, H, W, 256, FSR_EPI_PS_PRELU, 0, 0.f, 0, dt, st))) return rc;
  if ((rc = fsr_conv3x3_c64(u0, prm->up_w[1], u1, prm->up_b[1], nullptr, prm->up_alpha[1], nb, 2 * H, 2 * W, 256, FSR_EPI_PS_PRELU, 0, 0.f, 0, dt, st))) return rc;
  return fsr_conv3x3_c64(u1, prm->head_w, yout, prm->head_b, nullptr, nullptr, nb, 4 * H, 4 * W, 16, FSR_EPI_HEAD_TANH, 0, 0.f, out_u8, dt, st);
}

int fsr_generator_forward(const FsrGeneratorParams* prm, const void* x, void* y, void* workspace, size_t ws_bytes,
                          int N, int H, int W, int in_u8, int out_u8, int group, void* stream) {
  if (!prm || !x || !y || !workspace) return FSR_ERR_BAD_ARG;
  if (prm->n_filters != 64 || prm->n_layers < 0 || prm->n_layers > FSR_MAX_LAYERS) return FSR_ERR_BAD_SHAPE;
  const int F = 64, L = prm->n_layers;
  if (ws_bytes < fsr_generator_workspace_bytes(N, H, W, F, L)) return FSR_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t P = align_up((size_t)N * H * W * F * 2, 1024);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  uint8_t* b_res = base;            // neck output (long-skip source)
  uint8_t* b_x = base + P;          // running residual-chain activation
  uint8_t* b_raw = base + 2 * P;    // raw conv output (pre InstanceNorm)
  uint8_t* b_y = base + 3 * P;      // normalised + PReLU intermediate / raw conv2 output of the fused chain
  uint8_t* b_x2 = base + 4 * P;     // second residual-chain buffer (ping-pong of the fused chain)
  uint8_t* b_u0 = base + 5 * P;     // [N,2H,2W,64]
  uint8_t* b_u1 = base + 9 * P;     // [N,4H,4W,64]
  int64_t* b_stats = reinterpret_cast<int64_t*>(base + 25 * P);
  const size_t stats_per_conv = (size_t)N * F * 2;
  FSR_CUDA(cudaMemsetAsync(b_stats, 0, (size_t)(2 * L + 1) * stats_per_conv * sizeof(int64_t), st));

  // group > 0: that many images per sequential chunk (kept for A/B: L2-resident groups); else `parts` concurrent
  // sub-batches on side streams
  int parts = 1, per;
  if (group > 0 && group < N) { per = group; }
  else { parts = overlap_parts(); if (parts > N) parts = N; per = (N + parts - 1) / parts; }
  const bool concurrent = !(group > 0 && group < N) && parts > 1;
  // the bound context owns its side streams / events (two generators on two streams never share them); unbound callers
  // share the process-wide set and must not run fsr_generator_forward concurrently with overlap_streams > 1
  cudaStream_t* side = tl_ctx ? tl_ctx->side : g_side;
  cudaEvent_t* join = tl_ctx ? tl_ctx->join : g_join;
  cudaEvent_t& fork = tl_ctx ? tl_ctx->fork : g_fork;
  if (concurrent) {
    if (!fork) {
      FSR_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
      for (int i = 0; i < 4; ++i) {
        FSR_CUDA(cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking));
        FSR_CUDA(cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming));
      }
    }
    FSR_CUDA(cudaEventRecord(fork, st));
  }
  const size_t img_bytes = (size_t)H * W * F * 2;
  const size_t in_img = in_u8 ? (size_t)H * W * 3 : (size_t)H * W * 3 * sizeof(float);
  const size_t out_img = out_u8 == 1 ? (size_t)16 * H * W * 3 : (size_t)16 * H * W * 3 * sizeof(float);
  int rc, part = 0;
  for (int n0 = 0; n0 < N; n0 += per, ++part) {
    const int nb = (N - n0 < per) ? (N - n0) : per;
    cudaStream_t s = st;
    if (concurrent && part > 0) {
      s = side[part - 1];
      FSR_CUDA(cudaStreamWaitEvent(s, fork, 0));
    }
    rc = generator_chain(prm, reinterpret_cast<const uint8_t*>(x) + n0 * in_img, reinterpret_cast<uint8_t*>(y) + n0 * out_img,
                         b_res + n0 * img_bytes, b_x + n0 * img_bytes, b_x2 + n0 * img_bytes, b_raw + n0 * img_bytes, b_y + n0 * img_bytes,
                         b_u0 + n0 * 4 * img_bytes, b_u1 + n0 * 16 * img_bytes, b_stats + (size_t)n0 * F * 2, stats_per_conv,
                         nb, H, W, in_u8, out_u8, s);
    if (rc) return rc;
    if (concurrent && part > 0) {
      FSR_CUDA(cudaEventRecord(join[part - 1], s));
      FSR_CUDA(cudaStreamWaitEvent(st, join[part - 1], 0));
    }
  }
  return FSR_OK;
}


// ====================================================================== training-step entry points
#define FSR_T(expr_h, expr_b) do { if (dtype == FSR_BF16) { expr_b; } else { expr_h; } } while (0)

int fsr_pack_conv3x3_weight_t(const float* w_oihw, void* w_packed, int cout, int cin, int ps_perm, int flip, int row_pad,
                              const float* row_scale, int dtype, void* stream) {
  { const int rc_pool = det_pool_ensure(); if (rc_pool) return rc_pool; }   // before any stream capture
  // transposed pack for data-gradient convs (rows = forward input channel, K = forward output channel)
  if (!w_oihw || !w_packed || cout <= 0 || cin <= 0 || row_pad < cin) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)9 * cout * row_pad;
  LaunchScope scope(FSR_K_NONE - 1, st);
  FSR_T((PdlLaunch(ew_blocks(total), 256, 0, st)(pack_conv3x3_weight_t_kernel<__half>, w_oihw, (__half*)w_packed, cout, cin, ps_perm, flip, row_pad, row_scale)),
        (PdlLaunch(ew_blocks(total), 256, 0, st)(pack_conv3x3_weight_t_kernel<__nv_bfloat16>, w_oihw, (__nv_bfloat16*)w_packed, cout, cin, ps_perm, flip, row_pad, row_scale)));
  return cuda_rc(cudaGetLastError());
}

size_t fsr_wgrad_workspace_bytes(void) { return wgrad_workspace_bytes(); }

int fsr_conv3x3_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int cin, int cout, int stride,
                      int ps_perm, void* workspace, size_t ws_bytes, int dtype, void* stream) {
  if (!x || !dy || !dw) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  float* dws[1] = {dw};
  if (dtype == FSR_BF16) return wgrad_dispatch<__nv_bfloat16>(x, dy, dws, 1, 0, 0, N, H, W, cin, cout, stride, ps_perm, (float*)workspace, ws_bytes, dtype, st);
  return wgrad_dispatch<__half>(x, dy, dws, 1, 0, 0, N, H, W, cin, cout, stride, ps_perm, (float*)workspace, ws_bytes, dtype, st);
}

int fsr_conv3x3_wgrad_grouped(const void* x_arena, const void* dy_arena, float* const* dw_list_host, int groups,
                              long long x_group_stride, long long dy_group_stride, int N, int H, int W, int cin, int cout,
                              void* workspace, size_t ws_bytes, int dtype, void* stream) {
  if (!x_arena || !dy_arena || !dw_list_host) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FSR_BF16)
    return wgrad_dispatch<__nv_bfloat16>(x_arena, dy_arena, dw_list_host, groups, x_group_stride, dy_group_stride, N, H, W, cin, cout, 1, 0,
                                         (float*)workspace, ws_bytes, dtype, st);
  return wgrad_dispatch<__half>(x_arena, dy_arena, dw_list_host, groups, x_group_stride, dy_group_stride, N, H, W, cin, cout, 1, 0,
                                (float*)workspace, ws_bytes, dtype, st);
}

int fsr_parity_layout(const void* in, void* out, int N, int H, int W, int C, int to_parity, int dtype, void* stream) {
  if (!in || !out || C % 8 || ((H | W) & 1)) return FSR_ERR_BAD_ARG;
  (void)dtype;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)N * H * W * (C / 8);
  LaunchScope scope(FSR_K_NONE - 1, st);
  if (to_parity) PdlLaunch(ew_blocks(total), 256, 0, st)(parity_layout_kernel<true>, (const uint4*)in, (uint4*)out, N, H, W, C / 8);
  else PdlLaunch(ew_blocks(total), 256, 0, st)(parity_layout_kernel<false>, (const uint4*)in, (uint4*)out, N, H, W, C / 8);
  return cuda_rc(cudaGetLastError());
}

static int maxpool2_impl(const void* in, void* out, int N, int H, int W, int C, int in_pad, int out_pad, int dtype, void* stream) {
  if (!in || !out || C % 8 || ((H | W) & 1) || (in_pad & ~1) || (out_pad & ~1)) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)N * (H / 2) * (W / 2) * (C / 8);
  LaunchScope scope(FSR_K_NONE - 1, st);
  FSR_T((PdlLaunch(ew_blocks(total), 256, 0, st)(maxpool2_fwd_kernel<__half>, (const __half*)in, (__half*)out, N, H, W, C, in_pad, out_pad)),
        (PdlLaunch(ew_blocks(total), 256, 0, st)(maxpool2_fwd_kernel<__nv_bfloat16>, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, N, H, W, C, in_pad, out_pad)));
  return cuda_rc(cudaGetLastError());
}

static int maxpool2_relu_bwd_impl(const void* in, const void* dout, void* din, int N, int H, int W, int C, int in_pad, int out_pad,
                                  int dtype, void* stream) {
  if (!in || !dout || !din || ((H | W) & 1) || (in_pad & ~1) || (out_pad & ~1)) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)N * (H / 2) * (W / 2) * C;
  LaunchScope scope(FSR_K_NONE - 1, st);
  FSR_T((PdlLaunch(ew_blocks(total), 256, 0, st)(maxpool2_relu_bwd_kernel<__half>, (const __half*)in, (const __half*)dout, (__half*)din, N, H, W, C, in_pad, out_pad)),
        (PdlLaunch(ew_blocks(total), 256, 0, st)(maxpool2_relu_bwd_kernel<__nv_bfloat16>, (const __nv_bfloat16*)in, (const __nv_bfloat16*)dout, (__nv_bfloat16*)din, N, H, W, C, in_pad, out_pad)));
  return cuda_rc(cudaGetLastError());
}

int fsr_maxpool2(const void* in, void* out, int N, int H, int W, int C, int dtype, void* stream) {
  return maxpool2_impl(in, out, N, H, W, C, 0, 0, dtype, stream);
}
int fsr_maxpool2_relu_bwd(const void* in, const void* dout, void* din, int N, int H, int W, int C, int dtype, void* stream) {
  return maxpool2_relu_bwd_impl(in, dout, din, N, H, W, C, 0, 0, dtype, stream);
}
int fsr_maxpool2_padded(const void* in, void* out, int N, int H, int W, int C, int in_pad, int out_pad, int dtype, void* stream) {
  return maxpool2_impl(in, out, N, H, W, C, in_pad, out_pad, dtype, stream);
}
int fsr_maxpool2_relu_bwd_padded(const void* in, const void* dout, void* din, int N, int H, int W, int C, int in_pad, int out_pad,
                                 int dtype, void* stream) {
  return maxpool2_relu_bwd_impl(in, dout, din, N, H, W, C, in_pad, out_pad, dtype, stream);
}

int fsr_relu_bwd(const void* y, const void* dy, void* dx, size_t n_elems, int dtype, void* stream) {
  if (!y || !dy || !dx || n_elems % 8) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope scope(FSR_K_NONE - 1, st);
  FSR_T((PdlLaunch(ew_blocks(n_elems / 8), 256, 0, st)(relu_bwd_kernel<__half>, (const uint4*)y, (const uint4*)dy, (uint4*)dx, n_elems / 8)),
        (PdlLaunch(ew_blocks(n_elems / 8), 256, 0, st)(relu_bwd_kernel<__nv_bfloat16>, (const uint4*)y, (const uint4*)dy, (uint4*)dx, n_elems / 8)));
  return cuda_rc(cudaGetLastError());
}

int fsr_add(const void* a, const void* b, void* out, size_t n_elems, int dtype, void* stream) {
  if (!a || !b || !out || n_elems % 8) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope scope(FSR_K_NONE - 1, st);
  FSR_T((PdlLaunch(ew_blocks(n_elems / 8), 256, 0, st)(add_kernel<__half>, (const uint4*)a, (const uint4*)b, (uint4*)out, n_elems / 8)),
        (PdlLaunch(ew_blocks(n_elems / 8), 256, 0, st)(add_kernel<__nv_bfloat16>, (const uint4*)a, (const uint4*)b, (uint4*)out, n_elems / 8)));
  return cuda_rc(cudaGetLastError());
}

int fsr_conv1x1_to1_fwd(const void* x, const float* w, const float* b, float* z, int npix, int C, int dtype, void* stream) {
  if (!x || !w || !b || !z || C % 64) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (npix * 32 + 255) / 256;
  LaunchScope scope(FSR_K_NONE - 1, st);
  FSR_T((PdlLaunch(blocks, 256, 0, st)(conv1x1_to1_fwd_kernel<__half>, (const __half*)x, w, b, z, npix, C)),
        (PdlLaunch(blocks, 256, 0, st)(conv1x1_to1_fwd_kernel<__nv_bfloat16>, (const __nv_bfloat16*)x, w, b, z, npix, C)));
  return cuda_rc(cudaGetLastError());
}

int fsr_conv1x1_to1_bwd(const void* x, const float* w, const float* dz, void* dx, float* dw, float* db, int npix, int C,
                        int dtype, void* stream) {
  if (!x || !w || !dz) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (npix + 15) / 16;
  if (blocks > num_sms() * 4) blocks = num_sms() * 4;
  if (C + 1 > kDetSlotLen) return FSR_ERR_BAD_SHAPE;
  DetRed red;
  { const int rc = det_slot(&red); if (rc) return rc; }
  LaunchScope scope(FSR_K_NONE - 1, st);
  FSR_T((PdlLaunch(blocks, 256, 0, st)(conv1x1_to1_bwd_kernel<__half>, (const __half*)x, w, dz, (__half*)dx, dw, db, npix, C, red)),
        (PdlLaunch(blocks, 256, 0, st)(conv1x1_to1_bwd_kernel<__nv_bfloat16>, (const __nv_bfloat16*)x, w, dz, (__nv_bfloat16*)dx, dw, db, npix, C, red)));
  return cuda_rc(cudaGetLastError());
}

int fsr_bce_logits(const float* z, const float* noise, float lab_scale, float lab_shift, int n, float* loss_out, float* dz,
                   float grad_scale, void* stream) {
  if (!z || !noise || !loss_out || n <= 0) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope scope(FSR_K_NONE - 1, st);
  PdlLaunch(1, 256, 0, st)(bce_logits_kernel, z, noise, lab_scale, lab_shift, n, loss_out, dz, grad_scale);
  return cuda_rc(cudaGetLastError());
}

int fsr_smooth_l1(const void* a, const void* b, size_t n, float* loss_acc, void* da, float grad_scale, int dtype, void* stream) {
  if (!a || !b || !loss_acc) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  DetRed red;
  { const int rc = det_slot(&red); if (rc) return rc; }
  LaunchScope scope(FSR_K_NONE - 1, st);
  if (dtype == 2) PdlLaunch(ew_blocks(n), 256, 0, st)(smooth_l1_f32_kernel, (const float*)a, (const float*)b, n, loss_acc, (float*)da, grad_scale, red);
  else FSR_T((PdlLaunch(ew_blocks(n), 256, 0, st)(smooth_l1_kernel<__half>, (const __half*)a, (const __half*)b, n, loss_acc, (__half*)da, grad_scale, red)),
             (PdlLaunch(ew_blocks(n), 256, 0, st)(smooth_l1_kernel<__nv_bfloat16>, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, n, loss_acc, (__nv_bfloat16*)da, grad_scale, red)));
  return cuda_rc(cudaGetLastError());
}

static int instnorm_bwd_impl(const void* raw, const int64_t* stats, const void* dy, float* red, void* draw, const float* alpha,
                             float* dalpha, int N, int HW, int C, int act, float slope, float eps, int dy_parity_w, int dtype, void* stream) {
  if (!raw || !stats || !dy || !draw || C % 8 || 256 % (C / 8)) return FSR_ERR_BAD_ARG;
  if (act == FSR_ACT_PRELU && !alpha) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  InBwdParams p{raw, reinterpret_cast<const long long*>(stats), dy, red, draw, alpha, dalpha, slope, act, HW, C, eps, dy_parity_w};
  const bool fused = C % 16 == 0 && HW <= 4096 && in_bwd_fused_mode();
  if (dy_parity_w && (!fused || (dy_parity_w & 1) || HW % dy_parity_w || ((HW / dy_parity_w) & 1))) return FSR_ERR_BAD_SHAPE;
  if (fused) {
    // training shapes: one launch, the per-(n,c) sums stay inside the block (no `red` scratch, no memset, no atomics)
    if (act == FSR_ACT_PRELU && dalpha) { const int rc = det_slot(&p.det); if (rc) return rc; }
    LaunchScope scope(FSR_K_NONE - 1, st);
    static const int res_on = [] { const char* e = getenv("FSR_IN_BWD_RES"); return (e && e[0] == '1') ? 1 : 0; }();
    if (res_on && HW <= 128 * 5) {           // register-resident variant (see the kernel): measured SLOWER (6.21 -> 6.28 ms step), opt-in
      FSR_T((PdlLaunch(dim3(C / 16, N), 256, 0, st)(instnorm_bwd_fused_kernel<__half, 5>, p)),
            (PdlLaunch(dim3(C / 16, N), 256, 0, st)(instnorm_bwd_fused_kernel<__nv_bfloat16, 5>, p)));
    } else {
      FSR_T((PdlLaunch(dim3(C / 16, N), 256, 0, st)(instnorm_bwd_fused_kernel<__half, 0>, p)),
            (PdlLaunch(dim3(C / 16, N), 256, 0, st)(instnorm_bwd_fused_kernel<__nv_bfloat16, 0>, p)));
    }
    return cuda_rc(cudaGetLastError());
  }
  if (!red) return FSR_ERR_BAD_ARG;
  const size_t nvec = (size_t)HW * (C / 8);
  int bpi = (int)((nvec + 256 * 8 - 1) / (256 * 8));
  const int cap = (num_sms() * 4 + N - 1) / N;
  if (bpi > cap) bpi = cap;
  if (bpi < 1) bpi = 1;
  dim3 grid(bpi, N);
  const size_t sm = (size_t)4 * C * sizeof(float);
  FSR_CUDA(cudaMemsetAsync(red, 0, (size_t)N * C * 2 * sizeof(float), st));
  {
    LaunchScope scope(FSR_K_NONE - 1, st);
    FSR_T((PdlLaunch(grid, 256, sm, st)(instnorm_bwd_kernel<__half, 1>, p)), (PdlLaunch(grid, 256, sm, st)(instnorm_bwd_kernel<__nv_bfloat16, 1>, p)));
  }
  {
    LaunchScope scope(FSR_K_NONE - 1, st);
    FSR_T((PdlLaunch(grid, 256, sm, st)(instnorm_bwd_kernel<__half, 2>, p)), (PdlLaunch(grid, 256, sm, st)(instnorm_bwd_kernel<__nv_bfloat16, 2>, p)));
  }
  return cuda_rc(cudaGetLastError());
}

int fsr_instnorm_bwd(const void* raw, const int64_t* stats, const void* dy, float* red, void* draw, const float* alpha,
                     float* dalpha, int N, int HW, int C, int act, float slope, float eps, int dtype, void* stream) {
  return instnorm_bwd_impl(raw, stats, dy, red, draw, alpha, dalpha, N, HW, C, act, slope, eps, 0, dtype, stream);
}

int fsr_instnorm_bwd_parity(const void* raw, const int64_t* stats, const void* dy_parity, void* draw, const float* alpha, float* dalpha,
                            int N, int H, int W, int C, int act, float slope, float eps, int dtype, void* stream) {
  return instnorm_bwd_impl(raw, stats, dy_parity, nullptr, draw, alpha, dalpha, N, H * W, C, act, slope, eps, W, dtype, stream);
}

int fsr_act_bwd(const void* y, const void* dy, void* dv, size_t n_elems, const float* alpha, float slope, int act,
                float* dalpha, int dtype, void* stream) {
  if (!y || !dy || !dv || n_elems % 8) return FSR_ERR_BAD_ARG;
  if (act == FSR_ACT_PRELU && !alpha) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  DetRed red{nullptr, nullptr};
  if (act == FSR_ACT_PRELU && dalpha) { const int rc = det_slot(&red); if (rc) return rc; }
  LaunchScope scope(FSR_K_NONE - 1, st);
  FSR_T((PdlLaunch(ew_blocks(n_elems / 8), 256, 0, st)(act_bwd_kernel<__half>, (const uint4*)y, (const uint4*)dy, (uint4*)dv, n_elems / 8, alpha, slope, act, dalpha, red)),
        (PdlLaunch(ew_blocks(n_elems / 8), 256, 0, st)(act_bwd_kernel<__nv_bfloat16>, (const uint4*)y, (const uint4*)dy, (uint4*)dv, n_elems / 8, alpha, slope, act, dalpha, red)));
  return cuda_rc(cudaGetLastError());
}

int fsr_ps_prelu_bwd(const void* U, const void* dU, void* dconv, int N, int H, int W, int F, const float* alpha, float* dalpha,
                     int dtype, void* stream) {
  if (!U || !dU || !dconv || !alpha || F <= 0 || F % 8) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)N * H * W * 4 * (F / 8);
  DetRed red{nullptr, nullptr};
  if (dalpha) { const int rc = det_slot(&red); if (rc) return rc; }
  LaunchScope scope(FSR_K_NONE - 1, st);
  FSR_T((PdlLaunch(ew_blocks(total), 256, 0, st)(ps_prelu_bwd_kernel<__half>, (const __half*)U, (const __half*)dU, (__half*)dconv, N, H, W, F, alpha, dalpha, red)),
        (PdlLaunch(ew_blocks(total), 256, 0, st)(ps_prelu_bwd_kernel<__nv_bfloat16>, (const __nv_bfloat16*)U, (const __nv_bfloat16*)dU, (__nv_bfloat16*)dconv, N, H, W, F, alpha, dalpha, red)));
  return cuda_rc(cudaGetLastError());
}

int fsr_tanh_bwd(const float* y, const float* dy, float* dpre, size_t n, void* stream) {
  if (!y || !dy || !dpre) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope scope(FSR_K_NONE - 1, st);
  PdlLaunch(ew_blocks(n), 256, 0, st)(tanh_bwd_kernel, y, dy, dpre, n);
  return cuda_rc(cudaGetLastError());
}

int fsr_wgrad_c3(const float* img, const void* act, float* out, int N, int H, int W, int C64, int flip, int layout, int dtype,
                 void* stream) {
  if (!img || !act || !out || C64 % 64) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)N * H * W;
  int bx = (int)((total + 255) / 256);           // >= 256 pixels per block
  if (bx > num_sms() * 4) bx = num_sms() * 4;
  if (bx < 1) bx = 1;
  dim3 grid(bx, C64 / 64);
  if (27 * C64 > kDetSlotLen) return FSR_ERR_BAD_SHAPE;
  DetRed red;
  { const int rc = det_slot(&red); if (rc) return rc; }
  LaunchScope scope(FSR_K_NONE - 1, st);
  if (small_mma_mode()) {
    if (total >= ((size_t)1 << 31) - 16) return FSR_ERR_BAD_SHAPE;
    const long long steps = (long long)((total + 15) / 16);
    long long mb = (steps + 8 * kWgc3Warps - 1) / (8 * kWgc3Warps);      // >= 8 steps per warp
    if (mb > (long long)num_sms() * 3) mb = (long long)num_sms() * 3;
    dim3 mgrid((unsigned)mb, C64 / 64);
    FSR_T((PdlLaunch(mgrid, kWgc3Warps * 32, 0, st)(wgrad_c3_mma_kernel<__half>, img, (const __half*)act, out, N, H, W, C64, flip, layout, red)),
          (PdlLaunch(mgrid, kWgc3Warps * 32, 0, st)(wgrad_c3_mma_kernel<__nv_bfloat16>, img, (const __nv_bfloat16*)act, out, N, H, W, C64, flip, layout, red)));
    return cuda_rc(cudaGetLastError());
  }
  FSR_T((PdlLaunch(grid, 224, 0, st)(wgrad_c3_kernel<__half>, img, (const __half*)act, out, N, H, W, C64, flip, layout, red)),
        (PdlLaunch(grid, 224, 0, st)(wgrad_c3_kernel<__nv_bfloat16>, img, (const __nv_bfloat16*)act, out, N, H, W, C64, flip, layout, red)));
  return cuda_rc(cudaGetLastError());
}

int fsr_bias_grad(const void* g, float* db, size_t npix, int C, int ps_perm, int dtype, void* stream) {
  if (!g || !db || C % 8 || 256 % (C / 8)) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)((npix + 255) / 256);
  if (blocks > num_sms() * 4) blocks = num_sms() * 4;
  if (blocks < 1) blocks = 1;
  const size_t sm = (size_t)C * sizeof(unsigned long long);
  if (C > kDetSlotLen) return FSR_ERR_BAD_SHAPE;
  DetRed red;
  { const int rc = det_slot(&red); if (rc) return rc; }
  LaunchScope scope(FSR_K_NONE - 1, st);
  FSR_T((PdlLaunch(blocks, 256, sm, st)(bias_grad_kernel<__half>, (const __half*)g, db, npix, C, ps_perm, red)),
        (PdlLaunch(blocks, 256, sm, st)(bias_grad_kernel<__nv_bfloat16>, (const __nv_bfloat16*)g, db, npix, C, ps_perm, red)));
  return cuda_rc(cudaGetLastError());
}

int fsr_bias_grad_nchw(const float* g, float* db, int N, int C, size_t HW, void* stream) {
  if (!g || !db) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int bx = (int)((HW + 255) / 256);
  if (bx > 64) bx = 64;
  if (C > kDetSlotLen) return FSR_ERR_BAD_SHAPE;
  DetRed red;
  { const int rc = det_slot(&red); if (rc) return rc; }
  LaunchScope scope(FSR_K_NONE - 1, st);
  PdlLaunch(dim3(bx, C), 256, 0, st)(bias_grad_nchw_kernel, g, db, N, C, HW, red);
  return cuda_rc(cudaGetLastError());
}

int fsr_adamw(float* p, const float* g, float* m, float* v, size_t n, float lr, float b1, float b2, float eps, float wd,
              int step, float grad_scale, void* stream) {
  if (!p || !g || !m || !v || step < 1) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const float bc1 = 1.0f - powf(b1, (float)step);
  const float bc2s = sqrtf(1.0f - powf(b2, (float)step));
  LaunchScope scope(FSR_K_NONE - 1, st);
  PdlLaunch(ew_blocks(n), 256, 0, st)(adamw_kernel, p, g, m, v, n, lr, b1, b2, eps, wd, bc1, bc2s, grad_scale);
  return cuda_rc(cudaGetLastError());
}

int fsr_adamw_dev(float* p, const float* g, float* m, float* v, size_t n, float lr, float b1, float b2, float eps, float wd,
                  int* step_dev, float grad_scale, void* stream) {
  if (!p || !g || !m || !v || !step_dev) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  {
    LaunchScope scope(FSR_K_NONE - 1, st);
    PdlLaunch(1, 32, 0, st)(step_inc_kernel, step_dev);
  }
  LaunchScope scope(FSR_K_NONE - 1, st);
  PdlLaunch(ew_blocks(n), 256, 0, st)(adamw_dev_kernel, p, g, m, v, n, lr, b1, b2, eps, wd, step_dev, grad_scale);
  return cuda_rc(cudaGetLastError());
}

// ------------------------------------------------------------------ section-8 "next" rows: validation metrics, data path
int fsr_psnr_ssim(const float* pred, const float* target, int N, int C, int H, int W, float scale, float shift,
                  float data_range, const float* taps11_host, double* sse, double* ssim_sum, void* stream) {
  if (!pred || !target || !taps11_host || !sse || !ssim_sum) return FSR_ERR_BAD_ARG;
  if (N <= 0 || C <= 0 || H < 11 || W < 11 || (long long)N * C > 65535) return FSR_ERR_BAD_SHAPE;
  MetricParams p{};
  p.pred = pred; p.target = target; p.N = N; p.C = C; p.H = H; p.W = W; p.scale = scale; p.shift = shift;
  p.c1 = (0.01f * data_range) * (0.01f * data_range);
  p.c2 = (0.03f * data_range) * (0.03f * data_range);
  for (int i = 0; i < 11; ++i) p.g[i] = taps11_host[i];
  p.sse = sse; p.ssim_sum = ssim_sum;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((W - 10 + 15) / 16), (unsigned)((H - 10 + 15) / 16), (unsigned)(N * C));
  LaunchScope scope(FSR_K_NONE - 1, st);
  PdlLaunch(grid, 256, 0, st)(psnr_ssim_kernel, p);
  return cuda_rc(cudaGetLastError());
}

int fsr_crop_resize_aa(const uint8_t* cache, const int64_t* img_off, const int32_t* img_h, const int32_t* img_w,
                       const int32_t* samples, int B, int lr_size, int scale, const int32_t* tap_min,
                       const int32_t* tap_size, const float* tap_w, int K, float* lr, float* hr, void* stream) {
  if (!cache || !img_off || !img_h || !img_w || !samples || !tap_min || !tap_size || !tap_w || !lr || !hr) return FSR_ERR_BAD_ARG;
  if (B <= 0 || lr_size <= 0 || scale <= 0 || K <= 0) return FSR_ERR_BAD_SHAPE;
  const size_t hr_size = (size_t)lr_size * scale;
  const size_t smem = (hr_size * hr_size + 15) / 16 * 16 + hr_size * lr_size * sizeof(float);
  if (smem > 200 * 1024) return FSR_ERR_BAD_SHAPE;
  static size_t s_attr = 48 * 1024;
  if (smem > s_attr) {
    FSR_CUDA(cudaFuncSetAttribute(crop_resize_aa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    s_attr = smem;
  }
  CropResizeParams p{cache, (const long long*)img_off, img_h, img_w, samples, tap_min, tap_size, tap_w, lr, hr, B, lr_size, scale, K};
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope scope(FSR_K_NONE - 1, st);
  PdlLaunch(dim3((unsigned)B, 3), 256, smem, st)(crop_resize_aa_kernel, p);
  return cuda_rc(cudaGetLastError());
}

// ------------------------------------------------------------------ contexts
int fsr_ctx_create(void** ctx_out) {
  if (!ctx_out) return FSR_ERR_BAD_ARG;
  *ctx_out = new (std::nothrow) FsrCtxImpl();
  return *ctx_out ? FSR_OK : FSR_ERR_BAD_ARG;
}

int fsr_ctx_destroy(void* ctx) {
  auto* c = reinterpret_cast<FsrCtxImpl*>(ctx);
  if (!c) return FSR_OK;
  if (tl_ctx == c) tl_ctx = nullptr;
  if (c->fork) cudaEventDestroy(c->fork);
  for (int i = 0; i < 4; ++i) {
    if (c->side[i]) cudaStreamDestroy(c->side[i]);
    if (c->join[i]) cudaEventDestroy(c->join[i]);
  }
  delete c;
  return FSR_OK;
}

int fsr_ctx_set(void* ctx, int option, int value) {
  auto* c = reinterpret_cast<FsrCtxImpl*>(ctx);
  if (!c || option < 0 || option >= kOptCount) return FSR_ERR_BAD_ARG;
  c->opt[option] = value < 0 ? -1 : value;
  return FSR_OK;
}

int fsr_ctx_bind(void* ctx) {
  tl_ctx = reinterpret_cast<FsrCtxImpl*>(ctx);
  return FSR_OK;
}

// ------------------------------------------------------------------ precise generator forward (precise.cuh)
int fsr_split_f32(const float* x, void* hi, void* lo, size_t n_elems, void* stream) {
  if (!x || !hi || !lo || n_elems % 8) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope scope(FSR_K_NONE - 1, st);
  PdlLaunch(ew_blocks(n_elems / 8), 256, 0, st)(split_f32_kernel, x, (__half*)hi, (__half*)lo, n_elems / 8);
  return cuda_rc(cudaGetLastError());
}

int fsr_neck_conv3x3_f32(const float* x, const float* w, const float* bias, const float* alpha, float* out, int N, int H, int W,
                         void* stream) {
  if (!x || !w || !alpha || !out || N <= 0 || H <= 0 || W <= 0) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t threads = (size_t)N * H * W * 2;
  LaunchScope scope(FSR_K_NECK, st);
  PdlLaunch((unsigned)((threads + 255) / 256), 256, 0, st)(neck_conv3x3_f32_kernel, x, w, bias, alpha, out, N, H, W);
  return cuda_rc(cudaGetLastError());
}

int fsr_in_stats_f32(const float* x, int64_t* stats, int N, int HW, void* stream) {
  if (!x || !stats || N <= 0 || HW <= 0) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int bpi = (HW + 16 * 32 - 1) / (16 * 32);
  const int cap = (num_sms() * 4 + N - 1) / N;
  if (bpi > cap) bpi = cap;
  if (bpi < 1) bpi = 1;
  LaunchScope scope(FSR_K_NONE - 1, st);
  PdlLaunch(dim3(bpi, N), 256, 0, st)(in_stats_f32_kernel, x, reinterpret_cast<long long*>(stats), HW);
  return cuda_rc(cudaGetLastError());
}

int fsr_in_apply_f32(const float* x, const int64_t* stats, const float* residual, float* out, void* hi, void* lo,
                     const float* alpha, int act, int N, int HW, float eps, void* stream) {
  if (!x || !stats || !out || N <= 0 || HW <= 0 || (hi == nullptr) != (lo == nullptr)) return FSR_ERR_BAD_ARG;
  if (act == FSR_ACT_PRELU && !alpha) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int bpi = (int)(((size_t)HW * 8 + 256 * 4 - 1) / (256 * 4));
  const int cap = (num_sms() * 8 + N - 1) / N;
  if (bpi > cap) bpi = cap;
  if (bpi < 1) bpi = 1;
  LaunchScope scope(FSR_K_IN_APPLY, st);
  PdlLaunch(dim3(bpi, N), 256, 0, st)(in_apply_f32_kernel, x, reinterpret_cast<const long long*>(stats), residual, out, (__half*)hi, (__half*)lo,
                                                     alpha, act, HW, eps);
  return cuda_rc(cudaGetLastError());
}

int fsr_ps_prelu_f32(const float* conv, const float* bias_packed, const float* alpha, float* out, void* hi, void* lo, int N, int H,
                     int W, void* stream) {
  if (!conv || !bias_packed || !alpha || !out || N <= 0 || (hi == nullptr) != (lo == nullptr)) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope scope(FSR_K_NONE - 1, st);
  PdlLaunch(ew_blocks((size_t)N * H * W * 32), 256, 0, st)(ps_prelu_f32_kernel, conv, bias_packed, alpha, out, (__half*)hi, (__half*)lo, N, H, W);
  return cuda_rc(cudaGetLastError());
}

int fsr_tanh_f32(float* pre, uint8_t* out_u8, int N, int HW, void* stream) {
  if (!pre || N <= 0 || HW <= 0) return FSR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)N * 3 * HW;
  LaunchScope scope(FSR_K_NONE - 1, st);
  if (out_u8) PdlLaunch(ew_blocks(n), 256, 0, st)(tanh_u8_kernel, pre, out_u8, N, HW);
  else PdlLaunch(ew_blocks(n), 256, 0, st)(tanh_f32_kernel, pre, n);
  return cuda_rc(cudaGetLastError());
}

// ------------------------------------------------------------------ data-parallel exchange (nccl_comm.cuh)
int fsr_nccl_available(void) { return nccl_api().ok ? 1 : 0; }

int fsr_nccl_version(void) {
  const NcclApi& a = nccl_api();
  if (!a.ok || !a.GetVersion) return FSR_ERR_NO_NCCL;
  int v = 0;
  return a.GetVersion(&v) == 0 ? v : FSR_ERR_NCCL;
}

int fsr_nccl_unique_id(void* id_out_host) {
  const NcclApi& a = nccl_api();
  if (!a.ok) return FSR_ERR_NO_NCCL;
  if (!id_out_host) return FSR_ERR_BAD_ARG;
  return a.GetUniqueId(reinterpret_cast<NcclUniqueId*>(id_out_host)) == 0 ? FSR_OK : FSR_ERR_NCCL;
}

int fsr_nccl_init(const void* id_host, int rank, int world, void** comm_out) {
  const NcclApi& a = nccl_api();
  if (!a.ok) return FSR_ERR_NO_NCCL;
  if (!id_host || !comm_out || world < 1 || rank < 0 || rank >= world) return FSR_ERR_BAD_ARG;
  NcclUniqueId id;
  memcpy(&id, id_host, sizeof(id));
  NcclComm c = nullptr;
  if (a.CommInitRank(&c, world, id, rank) != 0) return FSR_ERR_NCCL;
  *comm_out = c;
  return FSR_OK;
}

int fsr_nccl_allreduce(void* comm, float* buf, size_t n, void* stream) {
  const NcclApi& a = nccl_api();
  if (!a.ok) return FSR_ERR_NO_NCCL;
  if (!comm || !buf) return FSR_ERR_BAD_ARG;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return a.AllReduce(buf, buf, n, kNcclFloat32, kNcclSum, comm, (cudaStream_t)stream) == 0 ? FSR_OK : FSR_ERR_NCCL;
}

int fsr_nccl_broadcast(void* comm, float* buf, size_t n, int root, void* stream) {
  const NcclApi& a = nccl_api();
  if (!a.ok) return FSR_ERR_NO_NCCL;
  if (!comm || !buf) return FSR_ERR_BAD_ARG;
  return a.Broadcast(buf, buf, n, kNcclFloat32, root, comm, (cudaStream_t)stream) == 0 ? FSR_OK : FSR_ERR_NCCL;
}

int fsr_nccl_destroy(void* comm) {
  const NcclApi& a = nccl_api();
  if (!a.ok) return FSR_ERR_NO_NCCL;
  if (!comm) return FSR_OK;
  return a.CommDestroy(comm) == 0 ? FSR_OK : FSR_ERR_NCCL;
}

}  // extern "C"
