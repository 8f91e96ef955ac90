// fsr_common.cuh - sm_100a PTX wrappers shared by the Fast-SRGAN B200 kernels.
// mbarrier / TMA (cp.async.bulk.tensor) / tcgen05 (UMMA + TMEM) primitives, hand-written
// inline PTX (no CUTLASS dependency).  Compile with -gencode arch=compute_100a,code=sm_100a.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define FSR_DEVINL __device__ __forceinline__

namespace fsr {

// ---------------------------------------------------------------- error codes (C-ABI)
enum : int {
  FSR_OK = 0,
  FSR_ERR_BAD_SHAPE = -1,
  FSR_ERR_BAD_ARG = -2,
  FSR_ERR_TENSORMAP = -3,
  FSR_ERR_WORKSPACE = -4,
  FSR_ERR_NO_DRIVER = -5,
  FSR_ERR_NO_NCCL = -6,
  FSR_ERR_NCCL = -7,
  FSR_ERR_CUDA_BASE = -1000,  // -(1000 + cudaError_t)
};

FSR_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute may become
// resident while its predecessor in the stream drains.  griddepcontrol.wait blocks until every prerequisite grid has
// COMPLETED and its memory is visible (a no-op for a normally launched kernel), so everything after it sees exactly
// what plain stream order would show; launch_dependents lets the following grid start launching.  Every kernel of this
// library calls this before its first global-memory access.
FSR_DEVINL void pdl_grid_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

FSR_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
FSR_DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
FSR_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

FSR_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
FSR_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
FSR_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a descriptor/pipeline bug must not hang the GPU box (a hang is a strike).
// ~2^28 polls (seconds) then trap -> the launch fails with an error instead of hanging.
FSR_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 28)) { asm volatile("trap;"); }
  }
}

// Waiters that are NOT on the critical path (TMA producer waiting for a free stage, epilogue warps waiting for an
// accumulator) back off between polls: a blocked try_wait keeps re-reading the barrier through the shared-memory data
// pipe - the pipe the tensor core fetches its operands through (profiles/r02: 684 LSU wavefronts per tile in the 64->64
// conv, of which only 137 are the epilogue's staging stores).
FSR_DEVINL void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (ns) __nanosleep(ns);
    if (++spins > (1u << 26)) { asm volatile("trap;"); }
  }
}

// ---------------------------------------------------------------- explicit shared-space access
FSR_DEVINL void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
FSR_DEVINL uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
FSR_DEVINL uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// ---------------------------------------------------------------- TMA
FSR_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)m) : "memory");
}
FSR_DEVINL void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
FSR_DEVINL void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

FSR_DEVINL void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// TMA store (smem -> global), bulk-group completion; OOB parts of the box are clipped by the hardware
FSR_DEVINL void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"((uint64_t)m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
FSR_DEVINL void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)m), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
FSR_DEVINL void tma_store_5d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"((uint64_t)m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
FSR_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
FSR_DEVINL void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
FSR_DEVINL void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------- tcgen05 / TMEM
FSR_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
FSR_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <uint32_t kCols>
FSR_DEVINL void tmem_alloc(uint32_t* smem_dst) {  // one full warp, .sync.aligned
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
FSR_DEVINL void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16/bf16 in, fp32 accumulate), single CTA.
FSR_DEVINL void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Weight-stationary variants: B is loaded into collector buffer b0 by the `fill` MMA and re-used (no smem read of B)
// by the `lastuse` MMA that follows with the same B operand - the B (weights) tile is fetched once per TWO M tiles.
FSR_DEVINL void umma_f16_ws_fill(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
FSR_DEVINL void umma_f16_ws_lastuse(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
// (implicitly performs tcgen05.fence::before_thread_sync)
FSR_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: warp reads its 32-lane quarter, 32 consecutive fp32 columns per thread.
FSR_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
FSR_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
FSR_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor (PTX ISA "tcgen05 matrix descriptor"):
//  [0,14)  start address >> 4         [16,30) leading-dim byte offset >> 4
//  [32,46) stride-dim byte offset >> 4 [46,48) version = 1 (sm_100)
//  [49,52) base offset                [61,64) swizzle: 0 none, 2 = 128B, 4 = 64B, 6 = 32B
// K-major operand, 128-byte rows (64 fp16), SWIZZLE_128B: 8-row atoms of 1024 B stacked along M/N
// => SBO = 1024 B; LBO is ignored for swizzled K-major layouts (set to 1 like CUTLASS does).
FSR_DEVINL uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)((smem_addr >> 7) & 0x7) << 49;  // base offset: 0 when 1024-B aligned
  d |= (uint64_t)2 << 61;
  return d;
}

// Split form for hot loops: the high word is constant for a 1024-B-aligned tile, the low word is
// (addr >> 4) | LBO; advancing inside a tile = adding (bytes >> 4) to the low word.
constexpr uint32_t kDescHiSw128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
FSR_DEVINL uint32_t desc_lo_sw128(uint32_t smem_addr) { return ((smem_addr & 0x3FFFF) >> 4) | (1u << 16); }
FSR_DEVINL uint64_t desc_join(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; }

// Warp-uniform leader election (same lane every time for the full mask).
FSR_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Instruction descriptor for kind::f16: fp32 accumulate, A/B both K-major.
//  [4,6) c_format (1 = F32)  [7,10) a_format  [10,13) b_format (0 = F16, 1 = BF16)
//  [15] a_major [16] b_major (0 = K)  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, bool bf16) {
  return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- small math helpers
template <typename T>
struct Cvt;
template <>
struct Cvt<__half> {
  FSR_DEVINL static uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  FSR_DEVINL static float2 unpack2(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
  FSR_DEVINL static float to_f(__half v) { return __half2float(v); }
  FSR_DEVINL static __half from_f(float v) { return __float2half_rn(v); }
};
template <>
struct Cvt<__nv_bfloat16> {
  FSR_DEVINL static uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  FSR_DEVINL static float2 unpack2(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u)); }
  FSR_DEVINL static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  FSR_DEVINL static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// InstanceNorm statistics are accumulated as 64-bit FIXED-POINT integers: integer atomics are associative, so the
// result does not depend on the order in which warps / CTAs arrive -> the forward pass is bitwise reproducible
// (fp32 atomics made it differ run to run at 1e-7, which 16-bit rounding + kinked activations amplify chaotically).
constexpr double kStatSumScale = 16777216.0;    // 2^24  (|sum|   < 5.5e11)
constexpr double kStatSqScale = 1048576.0;      // 2^20  (sum sq  < 8.8e12)
// v * 2^k is exact in fp32 (power-of-two scale, no overflow for |v| < 2^100), so fp32 multiply + F2I.S64.F32 gives the
// same integer as the fp64 route it replaces - without touching the FP64 pipe (three DMULs of the old form collected
// 25 % of the res-block conv's stall samples as math-pipe throttle, profiles/r01/ncu_full_resblock_conv_fused_input.md).
FSR_DEVINL long long stat_fix(float v, double scale) { return __float2ll_rn(v * (float)scale); }
FSR_DEVINL void stat_atomic_add(long long* dst, long long v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)v);
}
FSR_DEVINL void stat_mean_rstd(const long long* st, double inv_count, float eps, float& mean, float& rstd) {
  const double m = (double)st[0] / kStatSumScale * inv_count;
  double var = (double)st[1] / kStatSqScale * inv_count - m * m;   // biased variance (InstanceNorm2d), in fp64
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(var + (double)eps));
}

// ---- deterministic cross-block reductions of the backward pass (bias / PReLU-slope / 3-channel weight gradients, loss sums)
// fp32 atomicAdd makes a sum depend on the order in which blocks arrive; these sums go through 64-bit FIXED-POINT integer
// atomics instead (associative -> order-independent) into a zeroed slot of the library's pool (capi.cu det_slot()), and the
// LAST block to arrive converts the totals and adds them to the fp32 outputs, then re-zeroes the slot.  One launch, no
// caller-visible scratch; with the two-stage weight gradients and the in-block InstanceNorm sums this makes the whole
// training step bitwise reproducible run to run.
struct DetRed {
  unsigned long long* acc;   // [kDetSlotLen], all zero between launches
  unsigned int* ticket;      // 0 between launches
};
constexpr int kDetSlotLen = 16384;
constexpr float kDetScale = 68719476736.f;          // 2^36: |sum| < 1.3e8, resolution 1.5e-11 per contribution
FSR_DEVINL unsigned long long det_fix(float v) { return (unsigned long long)__float2ll_rn(v * kDetScale); }
FSR_DEVINL void det_add(const DetRed& r, int i, float v) { atomicAdd(r.acc + i, det_fix(v)); }
// det_arrive: called by EVERY thread of EVERY block exactly once, after the block's det_add()s; true (block-uniform) in the
// last block to arrive, which then det_collect()s the totals (out[i] += total[i0 + i], slot re-zeroed) and det_release()s.
FSR_DEVINL int det_tid() { return threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z); }
FSR_DEVINL bool det_arrive(const DetRed& r) {
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (det_tid() == 0) s_last = (atomicAdd(r.ticket, 1u) == gridDim.x * gridDim.y * gridDim.z - 1u) ? 1u : 0u;
  __syncthreads();
  const bool last = s_last != 0u;
  if (last) __threadfence();
  return last;
}
FSR_DEVINL void det_collect(const DetRed& r, int i0, int n, float* __restrict__ out) {
  const int nthreads = blockDim.x * blockDim.y * blockDim.z;
  for (int i = det_tid(); i < n; i += nthreads) {
    const long long v = (long long)atomicExch(r.acc + i0 + i, 0ull);
    if (v != 0) out[i] += __ll2float_rn(v) * (1.0f / kDetScale);
  }
}
FSR_DEVINL void det_release(const DetRed& r) {
  if (det_tid() == 0) atomicExch(r.ticket, 0u);
}
FSR_DEVINL void det_finish(const DetRed& r, float* __restrict__ out, int nout) {
  if (!det_arrive(r)) return;
  det_collect(r, 0, nout, out);
  det_release(r);
}

FSR_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace fsr
