// Fused attention forward for large token counts (sm_100a):   O = exp(Q K^T - lse) V   without the [N, N] probabilities
// ever reaching HBM (reference models/unet_dfc_sa_res.py:30-33 / models/unet_dfc_sa_ablation_attention.py:20-24:
// bmm(Q, K), softmax, bmm(V, A^T)).  The row log-sum-exp comes from the statistics pass of dfcsa_bgemm
// (DFCSA_BGEMM_EPI_ROWSTATS + dfcsa_lse_combine), so the probabilities are final the moment they are computed and the
// output accumulator never needs rescaling.
//
// One persistent CTA per SM walks (image, 128-query tile).  Per 128-key tile:
//   warp 0      TMA: K tile [128 keys x Cq] (K-major) and V tile [128 keys x C] (read MN-major) into a stage ring
//   warp 1      tcgen05.mma  S = Q K^T            -> TMEM columns [0,128) / [128,256) (double buffered)
//               tcgen05.mma  O += P V             -> TMEM columns [256, 256 + C)
//   warps 4-11  tcgen05.ld S, p = exp2(s*log2e - lse*log2e), fp16, written into shared memory in the 128-byte-swizzled
//               K-major layout the tensor core reads as the A operand of the second product (one row per thread,
//               64 keys = one 128-byte swizzle row), fence.proxy.async, mbarrier handshake with warp 1
// S of tile j+1 is computed while the softmax warps work on tile j; at the end of a query tile the same eight warps read O
// from TMEM and store it (fp32).
#include "common.cuh"
#include <algorithm>
#include <mutex>

namespace dfcsa {
namespace {

constexpr int kQBytes = 128 * 128;         // 128 rows x 64 k (zero-filled beyond Cq) fp16
constexpr int kKBytes = 128 * 128;
constexpr int kPBytes = 2 * 128 * 128;     // two K blocks of 64 keys
constexpr int kMaxKvStages = 4;
constexpr int kSmemMax = 227 * 1024 - 4096;       // dynamic part; the barriers are static shared memory
constexpr float kLog2e = 1.4426950408889634f;

struct FusedArgs {
  int batch, N, Cq, C, m_tiles, key_tiles, stages;
  const float* lse;
  float* o;                // [batch, N, C]
  uint32_t idesc_s, idesc_o;
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(384, 1)
attn_pv_fused_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const __grid_constant__ FusedArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, q_empty, o_full, o_empty;
  __shared__ __align__(8) uint64_t kv_full[kMaxKvStages], kv_empty[kMaxKvStages];
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2], p_full[2], p_empty[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int v_bytes = 128 * a.C * 2;
  const int stage_bytes = kKBytes + v_bytes;
  uint8_t* q_smem = smem;
  uint8_t* p_smem = smem + kQBytes;                     // 2 buffers
  uint8_t* kv_smem = p_smem + 2 * kPBytes;
  const long long total_tiles = static_cast<long long>(a.batch) * a.m_tiles;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v); }
  if (warp == 1 && lane == 0) {
    mbar_init(&q_full, 1); mbar_init(&q_empty, 1); mbar_init(&o_full, 1); mbar_init(&o_empty, 8);
    for (int i = 0; i < a.stages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); mbar_init(&p_full[i], 4); mbar_init(&p_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_o = tmem_base + 256;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0, qphase = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = static_cast<int>(tile % a.m_tiles);
      const int b = static_cast<int>(tile / a.m_tiles);
      if (elect_one()) {
        mbar_wait(&q_empty, qphase ^ 1);
        mbar_arrive_expect_tx(&q_full, kQBytes);
        tma_load_3d(q_smem, &map_q, &q_full, 0, mt * 128, b);
      }
      qphase ^= 1;
      for (int j = 0; j < a.key_tiles; ++j) {
        if (elect_one()) {
          mbar_wait(&kv_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&kv_full[stage], static_cast<uint32_t>(stage_bytes));
          uint8_t* sk = kv_smem + stage * stage_bytes;
          uint8_t* sv = sk + kKBytes;
          tma_load_3d(sk, &map_k, &kv_full[stage], 0, j * 128, b);
          const int nb = a.C / 64;
          for (int kb = 0; kb < 2; ++kb)
            for (int jn = 0; jn < nb; ++jn)
              tma_load_3d(sv + (kb * nb + jn) * 8192, &map_v, &kv_full[stage], jn * 64, j * 128 + kb * 64, b);
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0; uint32_t phase = 0;            // stage of the next S product
    int pstage = 0;                               // stage of the next P V product
    uint32_t qphase = 0, ophase = 0;
    uint32_t sph[2] = {0, 0}, pph[2] = {0, 0};    // parities of s_empty / p_full per buffer
    const uint32_t q_addr = smem_u32(q_smem);
    const int nb = a.C / 64;
    auto pv = [&](int i) {                        // O (+)= P_i V_i
      const int buf = i & 1;
      mbar_wait(&p_full[buf], pph[buf]); pph[buf] ^= 1;
      if (i == 0) { mbar_wait(&o_empty, ophase ^ 1); ophase ^= 1; }
      tc_fence_after();
      if (elect_one()) {
        const uint32_t p_addr = smem_u32(p_smem + buf * kPBytes);
        const uint32_t v_addr = smem_u32(kv_smem + pstage * stage_bytes + kKBytes);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_desc_add(umma_smem_desc(p_addr + kb * 16384, 16, 1024), k * 32);
            const uint64_t db = umma_desc_add(umma_smem_desc(v_addr + kb * nb * 8192, 8192, 1024), k * 2048);
            umma_f16(tmem_o, da, db, a.idesc_o, (i | kb | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&p_empty[buf]);
        umma_commit(&kv_empty[pstage]);
      }
      __syncwarp();
      if (++pstage == a.stages) pstage = 0;
    };
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&q_full, qphase); qphase ^= 1;
      for (int j = 0; j < a.key_tiles; ++j) {
        const int buf = j & 1;
        mbar_wait(&kv_full[stage], phase);
        mbar_wait(&s_empty[buf], sph[buf] ^ 1); sph[buf] ^= 1;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t k_addr = smem_u32(kv_smem + stage * stage_bytes);
          const int ksteps = (a.Cq + 15) / 16;          // the K block is zero beyond Cq
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t da = umma_desc_add(umma_smem_desc(q_addr, 16, 1024), k * 32);
            const uint64_t db = umma_desc_add(umma_smem_desc(k_addr, 16, 1024), k * 32);
            umma_f16(tmem_base + buf * 128, da, db, a.idesc_s, k != 0 ? 1u : 0u);
          }
          umma_commit(&s_full[buf]);
          if (j + 1 == a.key_tiles) umma_commit(&q_empty);
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
        if (j >= 1) pv(j - 1);
      }
      pv(a.key_tiles - 1);
      if (elect_one()) umma_commit(&o_full);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===================== softmax + epilogue =====================
    // The two groups of four warps own one S / P buffer each and take alternate key tiles, so the two warps that share
    // an SM sub-partition are out of phase: one's barrier / tcgen05.ld / fence latency overlaps the other's exp work.
    const int q4 = (warp - 4) & 3;        // TMEM lane quadrant = rows q4*32 .. +32 of the query tile
    const int half = (warp - 4) >> 2;     // group: key tiles j = half (mod 2), buffer `half`; half of the C output columns
    const int row = q4 * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>(q4 * 32) << 16;
    uint32_t sph = 0, pph = 0, ophase = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = static_cast<int>(tile % a.m_tiles);
      const long long b = tile / a.m_tiles;
      const int m = mt * 128 + row;
      const bool valid = m < a.N;
      const float lse2 = valid ? a.lse[b * a.N + m] * kLog2e : 0.f;
      for (int j = half; j < a.key_tiles; j += 2) {
        const int buf = half;
        mbar_wait(&s_full[buf], sph); sph ^= 1;
        tc_fence_after();
        // 128 keys in four chunks of 32; the tcgen05.ld of chunk c+1 is in flight while chunk c is exponentiated
        uint32_t ra[32], rb[32];
        const uint32_t t_s = tmem_base + buf * 128 + t_lane;
        auto chunk = [&](uint32_t (&r)[32], uint32_t (&nxt)[32], const int c) {
          tmem_ld_wait();
          if (c < 3) tmem_ld_32x32(t_s + (c + 1) * 32, nxt);
          else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[buf]);
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i)
            pk[i] = pack_h2(ex2f(fmaf(__uint_as_float(r[2 * i]), kLog2e, -lse2)), ex2f(fmaf(__uint_as_float(r[2 * i + 1]), kLog2e, -lse2)));
          if (c == 0) { mbar_wait(&p_empty[buf], pph ^ 1); pph ^= 1; }
          // K block c/2 of the P tile: row = query, 8 chunks of 8 keys per 128-byte row, chunk x stored at x ^ (row & 7)
          uint8_t* prow = p_smem + buf * kPBytes + (c >> 1) * 16384 + row * 128;
#pragma unroll
          for (int x = 0; x < 4; ++x)
            *reinterpret_cast<uint4*>(prow + ((((c & 1) * 4 + x) ^ (row & 7)) << 4)) = make_uint4(pk[4 * x], pk[4 * x + 1], pk[4 * x + 2], pk[4 * x + 3]);
        };
        tmem_ld_32x32(t_s, ra);
        chunk(ra, rb, 0); chunk(rb, ra, 1); chunk(ra, rb, 2); chunk(rb, ra, 3);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[buf]);
      }
      // ---- output: O rows of this quadrant, this warp's half of the C columns ----
      mbar_wait(&o_full, ophase); ophase ^= 1;
      tc_fence_after();
      const int ccols = a.C / 2;
      for (int c0 = 0; c0 < ccols; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_o + half * ccols + c0 + t_lane, r);
        tmem_ld_wait();
        if (valid) {
          float* dst = a.o + (b * a.N + m) * a.C + half * ccols + c0;
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<float4*>(dst + g * 4) = make_float4(__uint_as_float(r[g * 4]), __uint_as_float(r[g * 4 + 1]),
                                                                  __uint_as_float(r[g * 4 + 2]), __uint_as_float(r[g * 4 + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

std::once_flag g_attr_once;

}  // namespace
}  // namespace dfcsa

using namespace dfcsa;

extern "C" int dfcsa_attn_pv_fused(const void* qkv, int64_t ld, int32_t batch, int32_t N, int32_t Cq, int32_t C, const float* lse,
                                   float* o, void* stream) {
  DFCSA_CHECK_ARG(qkv && lse && o && batch > 0 && N > 0, "dfcsa_attn_pv_fused: bad args");
  DFCSA_CHECK_ARG(Cq % 8 == 0 && Cq >= 8 && Cq <= 64 && C % 64 == 0 && C >= 64 && C <= 128,
                  "dfcsa_attn_pv_fused: needs 8 <= Cq <= 64 (multiple of 8) and C in {64, 128}");
  DFCSA_CHECK_ARG(ld % 8 == 0 && ld >= 2 * Cq + C && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(o) & 15) == 0, "dfcsa_attn_pv_fused: qkv rows must be 16-byte aligned (q | k | v)");
  FusedArgs a{};
  a.batch = batch; a.N = N; a.Cq = Cq; a.C = C;
  a.m_tiles = (N + 127) / 128; a.key_tiles = (N + 127) / 128;
  a.lse = lse; a.o = o;
  a.idesc_s = umma_idesc_f16(128, 128, 0, 0, 0, 0);
  a.idesc_o = umma_idesc_f16(128, C, 0, 0, 0, 1);
  const int stage_bytes = kKBytes + 128 * C * 2;
  const int budget = kSmemMax - 1024 - kQBytes - 2 * kPBytes;
  a.stages = std::min(kMaxKvStages, budget / stage_bytes);
  DFCSA_CHECK_ARG(a.stages >= 2, "dfcsa_attn_pv_fused: shared memory budget");
  const __half* base = reinterpret_cast<const __half*>(qkv);
  CUtensorMap map_q, map_k, map_v;
  uint64_t dims[3], strides[2];
  uint32_t box[3];
  strides[0] = static_cast<uint64_t>(ld) * 2;
  strides[1] = static_cast<uint64_t>(N) * strides[0];
  dims[1] = static_cast<uint64_t>(N); dims[2] = static_cast<uint64_t>(batch);
  dims[0] = static_cast<uint64_t>(Cq); box[0] = 64; box[1] = 128; box[2] = 1;
  int rc = encode_tensor_map(&map_q, DFCSA_F16, 3, base, dims, strides, box, true);
  if (rc) return rc;
  rc = encode_tensor_map(&map_k, DFCSA_F16, 3, base + Cq, dims, strides, box, true);
  if (rc) return rc;
  dims[0] = static_cast<uint64_t>(C); box[0] = 64; box[1] = 64;
  rc = encode_tensor_map(&map_v, DFCSA_F16, 3, base + 2 * Cq, dims, strides, box, true);
  if (rc) return rc;
  const int smem_bytes = kQBytes + 2 * kPBytes + a.stages * stage_bytes + 1024;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(g_attr_once, [] {
    attr_err = cudaFuncSetAttribute(attn_pv_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(attn_pv_fused_kernel)");
  const long long total = static_cast<long long>(batch) * a.m_tiles;
  const int grid = static_cast<int>(std::min<long long>(total, num_sms()));
  attn_pv_fused_kernel<<<grid, 384, smem_bytes, static_cast<cudaStream_t>(stream)>>>(map_q, map_k, map_v, a);
  DFCSA_LAUNCH_CHECK("attn_pv_fused_kernel");
  return DFCSA_OK;
}
