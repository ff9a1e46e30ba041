// dfcsa-b200: shared device/host helpers for the sm_100a kernels.
// Raw PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld)
// and the UMMA shared-memory / instruction descriptors.  No CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/dfcsa.h"

namespace dfcsa {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing (no exceptions cross the C ABI)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);

#define DFCSA_CHECK_ARG(cond, ...)                                   \
  do { if (!(cond)) { ::dfcsa::set_error(__VA_ARGS__); return DFCSA_ERR_BAD_ARG; } } while (0)
#define DFCSA_CUDA(call)                                             \
  do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return ::dfcsa::cuda_fail(e__, #call); } while (0)
#define DFCSA_LAUNCH_CHECK(name)                                     \
  do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return ::dfcsa::cuda_fail(e__, name); } while (0)

// TMA descriptor encode through the driver entry point (no link-time libcuda dependency, so the
// library still loads on a CPU-only host).
int encode_tensor_map(CUtensorMap* map, int dtype, int rank, const void* base,
                      const uint64_t* dims, const uint64_t* strides_bytes /* rank-1 */,
                      const uint32_t* box, bool swizzle128);

int num_sms();
int wgrad_pick_splits(long long items, long long pix_blocks, int sms, int force_waves, long long* blocks_per_split);

static inline size_t dtype_size(int dt) { return dt == DFCSA_F32 ? 4 : 2; }

// ---------------------------------------------------------------------------------------------
// 16-bit storage helpers.  Activations are fp16 ("act"), gradients bf16 ("grad"); the tensor core
// sees both through kind::f16 with fp32 accumulation.
// ---------------------------------------------------------------------------------------------
// max / min that PROPAGATE NaN (max.NaN.f32): fmaxf / fminf return the non-NaN operand, which would silently turn a
// poisoned activation into 0 (ReLU) or -65504 (fp16 saturation) and hide a diverged batch from the loss; with these a
// NaN reaches the loss and the gradient norm, and the optimizer skips the step like the reference does.
__device__ __forceinline__ float fmax_nan(float a, float b) { float d; asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float fmin_nan(float a, float b) { float d; asm("min.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

template <typename T> struct Cvt;
template <> struct Cvt<float> {
  __device__ __forceinline__ static float to_f(float v) { return v; }
  __device__ __forceinline__ static float from_f(float v) { return v; }
};
template <> struct Cvt<__half> {
  __device__ __forceinline__ static float to_f(__half v) { return __half2float(v); }
  __device__ __forceinline__ static __half from_f(float v) {
    // saturate to the finite fp16 range so an outlier can never poison a BatchNorm with inf
    v = fmin_nan(fmax_nan(v, -65504.f), 65504.f);
    return __float2half_rn(v);
  }
};
template <> struct Cvt<__nv_bfloat16> {
  __device__ __forceinline__ static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ __forceinline__ static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// 8 x 16-bit <-> 8 x fp32 (one 16-byte vector)
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]) {
  uint4 raw = *reinterpret_cast<const uint4*>(p);
  const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = Cvt<T>::to_f(e[i]);
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&v)[8]) {
  uint4 raw;
  T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
  for (int i = 0; i < 8; ++i) e[i] = Cvt<T>::from_f(v[i]);
  *reinterpret_cast<uint4*>(p) = raw;
}
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p)     = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// 16 x 16-bit from 16 fp32: ONE 32-byte store per thread (st.global.v8.b32, sm_100+), i.e. a whole DRAM sector -
// a row-per-thread epilogue that writes 16 bytes at a time makes L2 fill the other half of every sector from DRAM
// (measured with ncu: +25 % DRAM reads on the level-1 GEMMs).  p must be 32-byte aligned.
// two floats -> packed 16-bit pair.  fp16: convert first (an out-of-range value becomes +-inf), then clamp the PAIR to the
// finite range with NaN-propagating half2 min / max - 3 instructions per pair instead of 5 (two float clamps per element);
// the GEMM epilogue is bound by instruction issue on the short-K shapes (ncu: ~440 warp instructions per tile and warp).
template <typename T> __device__ __forceinline__ uint32_t pack2_sat(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2_sat<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  const __half2 hi = __half2half2(__ushort_as_half(static_cast<unsigned short>(0x7BFF)));   //  65504
  const __half2 lo = __half2half2(__ushort_as_half(static_cast<unsigned short>(0xFBFF)));   // -65504
  h = __hmax2_nan(__hmin2_nan(h, hi), lo);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2_sat<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <typename T>
__device__ __forceinline__ void store16_256(T* p, const float (&v)[16]) {
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = pack2_sat<T>(v[2 * i], v[2 * i + 1]);
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// PTX: shared-memory addresses, mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped launch (an error code at the C ABI), never
// as a hung GPU.  ~2 s at 2 GHz.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("dfcsa: mbarrier wait timed out (block %d,%d of %d thread %d of %d, barrier 0x%x parity %u)\n", blockIdx.x, blockIdx.y,
             gridDim.x, threadIdx.x, blockDim.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// PTX: TMA tiled loads (global -> shared, completion on an mbarrier)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
         "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// PTX: tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               :: "r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16 / bf16 operands, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// PTX: CTA pairs (cta_group::2).  Two CTAs of a cluster of 2 (the SMs of one TPC) run ONE tcgen05.mma of M = 256: each CTA
// stages its own 128 rows of A and HALF of the B tile, so per CTA the shared-memory traffic of the B operand (TMA fill +
// UMMA reads) is halved.  Only the leader (cluster rank 0) issues the MMAs; the peer's TMA loads signal the LEADER's
// mbarrier (same shared-memory offset, bit 24 of the shared::cluster address cleared), tcgen05.commit multicasts its
// arrival to both CTAs, and the peer's epilogue warps arrive on the leader's "accumulator drained" barrier remotely.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;     // shared::cluster address -> the same offset in the even CTA of the pair
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem, uint32_t ncols) {  // the same warp of BOTH CTAs, same dst offset
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
               :: "r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs once every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_2cta(void* dst, const CUtensorMap* m, uint64_t* bar,
                                                 int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask),
         "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// arrive on the LEADER's (cluster rank 0) copy of a barrier, from either CTA of the pair: mapa translates the local address
// into the shared::cluster address of the same offset in rank 0.  (Masking bit 24 of the local address, which is what the
// .cta_group::2 TMA form accepts for its mbarrier operand, does NOT redirect a plain mbarrier.arrive: measured - the peer's
// arrivals stayed in the peer and the leader's third tile waited forever.)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(0u));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(remote) : "memory");
}

// UMMA shared-memory matrix descriptor, 128-byte swizzle.
//   K-major  operand: rows (M or N index) are 128 B apart, 8-row groups 1024 B apart (SBO); LBO unused.
//   MN-major operand: 64 contiguous M/N elements per 128 B row, rows = K index; 8-K-row groups SBO apart,
//                     64-element M/N blocks LBO apart.
// (field layout: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout=2 [61,64))
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// The same descriptor `bytes` further into the tile.  Only the 14-bit start-address field (units of 16 B) moves, and a tile
// inside the CTA's shared-memory window cannot carry out of it, so this is ONE 32-bit add on the low word - the issuing
// thread builds the descriptor of a stage once and derives every k step / tap from it by a constant (building each one
// from the address cost ~10 dependent uniform-datapath instructions per tcgen05.mma: at N <= 128 the single issuing
// thread, not the tensor pipe, set the pace).
__device__ __forceinline__ uint64_t umma_desc_add(uint64_t d, uint32_t bytes) {
  const uint32_t lo = static_cast<uint32_t>(d) + (bytes >> 4);
  return (d & 0xFFFFFFFF00000000ull) | static_cast<uint64_t>(lo);
}
// UMMA instruction descriptor for kind::f16: fp32 accumulator, per-operand fp16(0)/bf16(1) format and
// K(0)/MN(1) majorness, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ inline uint32_t umma_idesc_f16(int m, int n, int a_fmt, int b_fmt, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                          // D format: F32
  d |= static_cast<uint32_t>(a_fmt) << 7;
  d |= static_cast<uint32_t>(b_fmt) << 10;
  d |= static_cast<uint32_t>(a_mn_major) << 15;
  d |= static_cast<uint32_t>(b_mn_major) << 16;
  d |= static_cast<uint32_t>(n >> 3) << 17;
  d |= static_cast<uint32_t>(m >> 4) << 24;
  return d;
}
static inline int umma_fmt(int dtype) { return dtype == DFCSA_BF16 ? 1 : 0; }

}  // namespace dfcsa
