// Fused attention backward for large token counts (sm_100a).  With P = exp(Q K^T - lse) and D_i = sum_c dO_ic O_ic
// (reference models/unet_dfc_sa_res.py:30-33 and autograd of those lines):
//     dV = P^T dO        dS = P o (dO V^T - D)        dQ = dS K        dK = dS^T Q
// Two kernels, neither of which lets an [N, N] tensor reach HBM - P and dS are rebuilt tile by tile from the operands:
//   attn_dq_fused_kernel    one CTA per 128-QUERY tile, walks 64-key tiles:  S, dP in TMEM -> dS tile in shared memory
//                           (A operand, K-major) -> dQ accumulated in TMEM
//   attn_dkdv_fused_kernel  one CTA per 128-KEY tile, walks 64-query tiles:  S^T, dP^T in TMEM -> P^T and dS^T tiles in
//                           shared memory -> dV, dK accumulated in TMEM (lse / D are per COLUMN here: staged per tile with
//                           a bulk copy)
// The exp is evaluated twice (once per kernel); everything else is tensor-core work on tiles that are already on chip.
// Operand types: the score product uses the fp16 q / k of the forward pass (so that P matches the forward's to rounding),
// everything on the gradient side is bf16.  Operand tiles do double duty: a K-major [rows x 64] tile with the 128-byte
// swizzle is byte-identical to an MN-major [k rows x 64] tile, so the K tile that is the B operand of S is also the B
// operand of dQ = dS K, and dO / Q tiles serve dP^T and dV / dK alike.
//
// Warp roles as in attn_fused.cu: warp 0 TMA, warp 1 tcgen05.mma issue, warp 2 TMEM allocation, warps 4-11 softmax math.
#include "common.cuh"
#include <algorithm>
#include <mutex>

namespace dfcsa {
namespace {

constexpr int kMaxStages = 4;
constexpr int kSmemMax = 227 * 1024 - 4096;
constexpr float kLog2e = 1.4426950408889634f;

struct BwdArgs {
  int batch, N, Cq, C, tiles128, tiles64, stages;
  const float* lse;
  const float* D;
  float* out0;  long long ld0;     // dq kernel: dq;   dk/dv kernel: dk
  float* out1;  long long ld1;     // dk/dv kernel: dv
  uint32_t idesc_s, idesc_dp, idesc_acc0, idesc_acc1;
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// 32 bf16 values of one row -> chunks [c0, c0+4) of a 128-byte swizzled row
__device__ __forceinline__ void store_row_chunks(uint8_t* row_base, int row, int c0, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(row_base + (((c0 + c) ^ (row & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
}

// 16 bf16 values of one row -> chunks c0, c0+1
__device__ __forceinline__ void store_row_2chunks(uint8_t* row_base, int row, int c0, const uint32_t (&pk)[8]) {
#pragma unroll
  for (int c = 0; c < 2; ++c)
    *reinterpret_cast<uint4*>(row_base + (((c0 + c) ^ (row & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
}

// =====================================================================================================================
// dQ:  per 128-query tile, over 64-key tiles
//   TMEM: S [0,128) (2 x 64), dP [128,256) (2 x 64), dQ [256, 320)
//   smem: Q16 16K | dO nb*16K | dS 2 x 16K | stages x (K16 8K | Kb 8K | Vb nb*8K)
// =====================================================================================================================
__global__ void __launch_bounds__(384, 1)
attn_dq_fused_kernel(const __grid_constant__ CUtensorMap map_q16, const __grid_constant__ CUtensorMap map_do,
                     const __grid_constant__ CUtensorMap map_k16, const __grid_constant__ CUtensorMap map_kb,
                     const __grid_constant__ CUtensorMap map_vb, const __grid_constant__ BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full, a_empty, o_full, o_empty;
  __shared__ __align__(8) uint64_t kv_full[kMaxStages], kv_empty[kMaxStages];
  __shared__ __align__(8) uint64_t sd_full[2], sd_empty[2], ds_full[2], ds_empty[2];
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nb = a.C / 64;
  uint8_t* q_smem = smem;
  uint8_t* do_smem = q_smem + 16384;
  uint8_t* ds_smem = do_smem + nb * 16384;
  uint8_t* st_smem = ds_smem + 2 * 16384;
  const int stage_bytes = 8192 + 8192 + nb * 8192;
  const long long total_tiles = static_cast<long long>(a.batch) * a.tiles128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q16); tma_prefetch_desc(&map_do); tma_prefetch_desc(&map_k16); tma_prefetch_desc(&map_kb); tma_prefetch_desc(&map_vb);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&a_full, 1); mbar_init(&a_empty, 1); mbar_init(&o_full, 1); mbar_init(&o_empty, 8);
    for (int i = 0; i < a.stages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sd_full[i], 1); mbar_init(&sd_empty[i], 4); mbar_init(&ds_full[i], 4); mbar_init(&ds_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_dp = tmem_base + 128, tmem_dq = tmem_base + 256;

  if (warp == 0) {
    int stage = 0; uint32_t phase = 0, aphase = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = static_cast<int>(tile % a.tiles128), b = static_cast<int>(tile / a.tiles128);
      if (elect_one()) {
        mbar_wait(&a_empty, aphase ^ 1);
        mbar_arrive_expect_tx(&a_full, static_cast<uint32_t>(16384 + nb * 16384));
        tma_load_3d(q_smem, &map_q16, &a_full, 0, mt * 128, b);
        for (int jn = 0; jn < nb; ++jn) tma_load_3d(do_smem + jn * 16384, &map_do, &a_full, jn * 64, mt * 128, b);
      }
      aphase ^= 1;
      for (int j = 0; j < a.tiles64; ++j) {
        if (elect_one()) {
          mbar_wait(&kv_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&kv_full[stage], static_cast<uint32_t>(stage_bytes));
          uint8_t* s0 = st_smem + stage * stage_bytes;
          tma_load_3d(s0, &map_k16, &kv_full[stage], 0, j * 64, b);
          tma_load_3d(s0 + 8192, &map_kb, &kv_full[stage], 0, j * 64, b);
          for (int jn = 0; jn < nb; ++jn) tma_load_3d(s0 + 16384 + jn * 8192, &map_vb, &kv_full[stage], jn * 64, j * 64, b);
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    int stage = 0; uint32_t phase = 0; int pstage = 0;
    uint32_t aphase = 0, ophase = 0, sph[2] = {0, 0}, pph[2] = {0, 0};
    const uint32_t q_addr = smem_u32(q_smem), do_addr = smem_u32(do_smem);
    const int ksteps = (a.Cq + 15) / 16;
    auto acc = [&](int i) {                       // dQ (+)= dS_i K_i
      const int buf = i & 1;
      mbar_wait(&ds_full[buf], pph[buf]); pph[buf] ^= 1;
      if (i == 0) { mbar_wait(&o_empty, ophase ^ 1); ophase ^= 1; }
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ds_addr = smem_u32(ds_smem + buf * 16384);
        const uint32_t kb_addr = smem_u32(st_smem + pstage * stage_bytes + 8192);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_dq, umma_desc_add(umma_smem_desc(ds_addr, 16, 1024), k * 32), umma_desc_add(umma_smem_desc(kb_addr, 8192, 1024), k * 2048), a.idesc_acc0,
                   (i | k) != 0 ? 1u : 0u);
        umma_commit(&ds_empty[buf]);
        umma_commit(&kv_empty[pstage]);
      }
      __syncwarp();
      if (++pstage == a.stages) pstage = 0;
    };
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&a_full, aphase); aphase ^= 1;
      for (int j = 0; j < a.tiles64; ++j) {
        const int buf = j & 1;
        mbar_wait(&kv_full[stage], phase);
        mbar_wait(&sd_empty[buf], sph[buf] ^ 1); sph[buf] ^= 1;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t k16 = smem_u32(st_smem + stage * stage_bytes);
          const uint32_t vb = k16 + 16384;
          for (int k = 0; k < ksteps; ++k)
            umma_f16(tmem_base + buf * 64, umma_desc_add(umma_smem_desc(q_addr, 16, 1024), k * 32), umma_desc_add(umma_smem_desc(k16, 16, 1024), k * 32), a.idesc_s, k != 0 ? 1u : 0u);
          for (int jn = 0; jn < nb; ++jn)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_dp + buf * 64, umma_desc_add(umma_smem_desc(do_addr + jn * 16384, 16, 1024), k * 32),
                       umma_desc_add(umma_smem_desc(vb + jn * 8192, 16, 1024), k * 32), a.idesc_dp, (jn | k) != 0 ? 1u : 0u);
          umma_commit(&sd_full[buf]);
          if (j + 1 == a.tiles64) umma_commit(&a_empty);
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
        if (j >= 1) acc(j - 1);
      }
      acc(a.tiles64 - 1);
      if (elect_one()) umma_commit(&o_full);
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int q4 = (warp - 4) & 3, half = (warp - 4) >> 2;
    const int row = q4 * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>(q4 * 32) << 16;
    uint32_t sph[2] = {0, 0}, pph[2] = {0, 0}, ophase = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = static_cast<int>(tile % a.tiles128);
      const long long b = tile / a.tiles128;
      const int m = mt * 128 + row;
      const bool valid = m < a.N;
      const float lse2 = valid ? a.lse[b * a.N + m] * kLog2e : 0.f;
      const float Dm = valid ? a.D[b * a.N + m] : 0.f;
      // the two groups of four warps own one buffer each and take alternate key tiles (out of phase on every sub-partition)
      for (int j = half; j < a.tiles64; j += 2) {
        const int buf = half;
        mbar_wait(&sd_full[buf], sph[buf]); sph[buf] ^= 1;
        tc_fence_after();
        // 64 keys in four chunks of 16; the tcgen05.ld pair of chunk c+1 is in flight while chunk c is processed
        uint32_t sa[16], da[16], sb[16], db[16];
        const uint32_t t_s = tmem_base + buf * 64 + t_lane, t_d = tmem_dp + buf * 64 + t_lane;
        auto chunk = [&](uint32_t (&rs)[16], uint32_t (&rp)[16], uint32_t (&ns)[16], uint32_t (&np)[16], const int c) {
          tmem_ld_wait();
          if (c < 3) { tmem_ld_32x16(t_s + (c + 1) * 16, ns); tmem_ld_32x16(t_d + (c + 1) * 16, np); }
          else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sd_empty[buf]);
          }
          uint32_t pk[8];
          const int nvalid = a.N - j * 64 - c * 16;            // key columns of this chunk that exist
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float p0 = ex2f(fmaf(__uint_as_float(rs[2 * i]), kLog2e, -lse2));
            const float p1 = ex2f(fmaf(__uint_as_float(rs[2 * i + 1]), kLog2e, -lse2));
            const float s0 = 2 * i < nvalid ? p0 * (__uint_as_float(rp[2 * i]) - Dm) : 0.f;
            const float s1 = 2 * i + 1 < nvalid ? p1 * (__uint_as_float(rp[2 * i + 1]) - Dm) : 0.f;
            pk[i] = pack_bf2(s0, s1);
          }
          if (c == 0) { mbar_wait(&ds_empty[buf], pph[buf] ^ 1); pph[buf] ^= 1; }
          store_row_2chunks(ds_smem + buf * 16384 + row * 128, row, c * 2, pk);
        };
        tmem_ld_32x16(t_s, sa); tmem_ld_32x16(t_d, da);
        chunk(sa, da, sb, db, 0); chunk(sb, db, sa, da, 1); chunk(sa, da, sb, db, 2); chunk(sb, db, sa, da, 3);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ds_full[buf]);
      }
      mbar_wait(&o_full, ophase); ophase ^= 1;
      tc_fence_after();
      if (half == 0) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_dq + t_lane, r);
        tmem_ld_wait();
        if (valid) {
          float* dst = a.out0 + (b * a.N + m) * a.ld0;
#pragma unroll
          for (int c = 0; c < 32; c += 4)
            if (c < a.Cq)
              *reinterpret_cast<float4*>(dst + c) = make_float4(__uint_as_float(r[c]), __uint_as_float(r[c + 1]), __uint_as_float(r[c + 2]), __uint_as_float(r[c + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// =====================================================================================================================
// dK, dV:  per 128-key tile, over 64-query tiles
//   TMEM: S^T [0,128) (2 x 64), dP^T [128,256) (2 x 64), dV [256, 256+C), dK [384, 448)
//   smem: K16 16K | Vb nb*16K | P^T 2 x 16K | dS^T 2 x 16K | stages x (Q16 8K | Qb 8K | dO nb*8K | lse 256 B | D 256 B | pad)
// =====================================================================================================================
__global__ void __launch_bounds__(384, 1)
attn_dkdv_fused_kernel(const __grid_constant__ CUtensorMap map_k16, const __grid_constant__ CUtensorMap map_vb,
                       const __grid_constant__ CUtensorMap map_q16, const __grid_constant__ CUtensorMap map_qb,
                       const __grid_constant__ CUtensorMap map_do, const __grid_constant__ BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full, a_empty, o_full, o_empty;
  __shared__ __align__(8) uint64_t kv_full[kMaxStages], kv_empty[kMaxStages];
  __shared__ __align__(8) uint64_t sd_full[2], sd_empty[2], ds_full[2], ds_empty[2];
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nb = a.C / 64;
  uint8_t* k_smem = smem;
  uint8_t* v_smem = k_smem + 16384;
  uint8_t* pt_smem = v_smem + nb * 16384;          // 2 buffers
  uint8_t* dst_smem = pt_smem + 2 * 16384;         // 2 buffers
  uint8_t* st_smem = dst_smem + 2 * 16384;
  const int tile_bytes = 8192 + 8192 + nb * 8192;
  const int stage_bytes = tile_bytes + 1024;       // + lse[64] at +tile_bytes, D[64] at +tile_bytes+256
  const long long total_tiles = static_cast<long long>(a.batch) * a.tiles128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_k16); tma_prefetch_desc(&map_vb); tma_prefetch_desc(&map_q16); tma_prefetch_desc(&map_qb); tma_prefetch_desc(&map_do);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&a_full, 1); mbar_init(&a_empty, 1); mbar_init(&o_full, 1); mbar_init(&o_empty, 8);
    for (int i = 0; i < a.stages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sd_full[i], 1); mbar_init(&sd_empty[i], 4); mbar_init(&ds_full[i], 4); mbar_init(&ds_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_dp = tmem_base + 128, tmem_dv = tmem_base + 256, tmem_dk = tmem_base + 384;

  if (warp == 0) {
    int stage = 0; uint32_t phase = 0, aphase = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int kt = static_cast<int>(tile % a.tiles128), b = static_cast<int>(tile / a.tiles128);
      if (elect_one()) {
        mbar_wait(&a_empty, aphase ^ 1);
        mbar_arrive_expect_tx(&a_full, static_cast<uint32_t>(16384 + nb * 16384));
        tma_load_3d(k_smem, &map_k16, &a_full, 0, kt * 128, b);
        for (int jn = 0; jn < nb; ++jn) tma_load_3d(v_smem + jn * 16384, &map_vb, &a_full, jn * 64, kt * 128, b);
      }
      aphase ^= 1;
      for (int i = 0; i < a.tiles64; ++i) {
        if (elect_one()) {
          const int nvalid = min(64, a.N - i * 64);
          mbar_wait(&kv_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&kv_full[stage], static_cast<uint32_t>(tile_bytes + 8 * nvalid));
          uint8_t* s0 = st_smem + stage * stage_bytes;
          tma_load_3d(s0, &map_q16, &kv_full[stage], 0, i * 64, b);
          tma_load_3d(s0 + 8192, &map_qb, &kv_full[stage], 0, i * 64, b);
          for (int jn = 0; jn < nb; ++jn) tma_load_3d(s0 + 16384 + jn * 8192, &map_do, &kv_full[stage], jn * 64, i * 64, b);
          const long long off = static_cast<long long>(b) * a.N + i * 64;
          bulk_load_1d(s0 + tile_bytes, a.lse + off, 4u * nvalid, &kv_full[stage]);
          bulk_load_1d(s0 + tile_bytes + 256, a.D + off, 4u * nvalid, &kv_full[stage]);
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    int stage = 0; uint32_t phase = 0; int pstage = 0;
    uint32_t aphase = 0, ophase = 0, sph[2] = {0, 0}, pph[2] = {0, 0};
    const uint32_t k_addr = smem_u32(k_smem), v_addr = smem_u32(v_smem);
    const int ksteps = (a.Cq + 15) / 16;
    auto acc = [&](int i) {                       // dV (+)= P^T_i dO_i ;  dK (+)= dS^T_i Q_i
      const int buf = i & 1;
      mbar_wait(&ds_full[buf], pph[buf]); pph[buf] ^= 1;
      if (i == 0) { mbar_wait(&o_empty, ophase ^ 1); ophase ^= 1; }
      tc_fence_after();
      if (elect_one()) {
        const uint32_t pt = smem_u32(pt_smem + buf * 16384), dst = smem_u32(dst_smem + buf * 16384);
        const uint32_t s0 = smem_u32(st_smem + pstage * stage_bytes);
        const uint32_t qb = s0 + 8192, dob = s0 + 16384;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_dv, umma_desc_add(umma_smem_desc(pt, 16, 1024), k * 32), umma_desc_add(umma_smem_desc(dob, 8192, 1024), k * 2048), a.idesc_acc1, (i | k) != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_dk, umma_desc_add(umma_smem_desc(dst, 16, 1024), k * 32), umma_desc_add(umma_smem_desc(qb, 8192, 1024), k * 2048), a.idesc_acc0, (i | k) != 0 ? 1u : 0u);
        umma_commit(&ds_empty[buf]);
        umma_commit(&kv_empty[pstage]);
      }
      __syncwarp();
      if (++pstage == a.stages) pstage = 0;
    };
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&a_full, aphase); aphase ^= 1;
      for (int i = 0; i < a.tiles64; ++i) {
        const int buf = i & 1;
        mbar_wait(&kv_full[stage], phase);
        mbar_wait(&sd_empty[buf], sph[buf] ^ 1); sph[buf] ^= 1;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t q16 = smem_u32(st_smem + stage * stage_bytes);
          const uint32_t dob = q16 + 16384;
          for (int k = 0; k < ksteps; ++k)
            umma_f16(tmem_base + buf * 64, umma_desc_add(umma_smem_desc(k_addr, 16, 1024), k * 32), umma_desc_add(umma_smem_desc(q16, 16, 1024), k * 32), a.idesc_s, k != 0 ? 1u : 0u);
          for (int jn = 0; jn < nb; ++jn)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_dp + buf * 64, umma_desc_add(umma_smem_desc(v_addr + jn * 16384, 16, 1024), k * 32),
                       umma_desc_add(umma_smem_desc(dob + jn * 8192, 16, 1024), k * 32), a.idesc_dp, (jn | k) != 0 ? 1u : 0u);
          umma_commit(&sd_full[buf]);
          if (i + 1 == a.tiles64) umma_commit(&a_empty);
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
        if (i >= 1) acc(i - 1);
      }
      acc(a.tiles64 - 1);
      if (elect_one()) umma_commit(&o_full);
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int q4 = (warp - 4) & 3, half = (warp - 4) >> 2;
    const int row = q4 * 32 + lane;                 // key within the tile
    const uint32_t t_lane = static_cast<uint32_t>(q4 * 32) << 16;
    uint32_t sph[2] = {0, 0}, pph[2] = {0, 0}, ophase = 0;
    int stage = 0; uint32_t kvphase = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int kt = static_cast<int>(tile % a.tiles128);
      const long long b = tile / a.tiles128;
      const int key = kt * 128 + row;
      const bool valid = key < a.N;
      // alternate query tiles per group of four warps; `stage` / `kvphase` follow the ring position of tile i
      for (int i = 0; i < a.tiles64; ++i) {
        if ((i & 1) == half) {
          const int buf = half;
          mbar_wait(&kv_full[stage], kvphase);                   // lse / D of this query tile have landed (read below)
          mbar_wait(&sd_full[buf], sph[buf]); sph[buf] ^= 1;
          tc_fence_after();
          const float* lse_t = reinterpret_cast<const float*>(st_smem + stage * stage_bytes + tile_bytes);
          // 64 queries in four chunks of 16; the tcgen05.ld pair of chunk c+1 is in flight while chunk c is processed
          uint32_t sa[16], da[16], sb[16], db[16];
          const uint32_t t_s = tmem_base + buf * 64 + t_lane, t_d = tmem_dp + buf * 64 + t_lane;
          auto chunk = [&](uint32_t (&rs)[16], uint32_t (&rp)[16], uint32_t (&ns)[16], uint32_t (&np)[16], const int c) {
            tmem_ld_wait();
            if (c < 3) { tmem_ld_32x16(t_s + (c + 1) * 16, ns); tmem_ld_32x16(t_d + (c + 1) * 16, np); }
            else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&sd_empty[buf]);
            }
            const float* lse_s = lse_t + c * 16;
            const float* d_s = lse_s + 64;
            const int nvalid = a.N - i * 64 - c * 16;            // query columns of this chunk that exist
            uint32_t pp[8], pd[8];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 l4 = *reinterpret_cast<const float4*>(lse_s + g * 4);
              const float4 d4 = *reinterpret_cast<const float4*>(d_s + g * 4);
              const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
              float p[4], sv[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int col = g * 4 + e;
                const bool ok = col < nvalid;
                p[e] = ok ? ex2f(kLog2e * (__uint_as_float(rs[col]) - lv[e])) : 0.f;
                sv[e] = ok ? p[e] * (__uint_as_float(rp[col]) - dv[e]) : 0.f;
              }
              pp[g * 2] = pack_bf2(p[0], p[1]); pp[g * 2 + 1] = pack_bf2(p[2], p[3]);
              pd[g * 2] = pack_bf2(sv[0], sv[1]); pd[g * 2 + 1] = pack_bf2(sv[2], sv[3]);
            }
            if (c == 0) { mbar_wait(&ds_empty[buf], pph[buf] ^ 1); pph[buf] ^= 1; }
            store_row_2chunks(pt_smem + buf * 16384 + row * 128, row, c * 2, pp);
            store_row_2chunks(dst_smem + buf * 16384 + row * 128, row, c * 2, pd);
          };
          tmem_ld_32x16(t_s, sa); tmem_ld_32x16(t_d, da);
          chunk(sa, da, sb, db, 0); chunk(sb, db, sa, da, 1); chunk(sa, da, sb, db, 2); chunk(sb, db, sa, da, 3);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ds_full[buf]);
        }
        if (++stage == a.stages) { stage = 0; kvphase ^= 1; }
      }
      mbar_wait(&o_full, ophase); ophase ^= 1;
      tc_fence_after();
      const int ccols = a.C / 2;
      for (int c0 = 0; c0 < ccols; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_dv + half * ccols + c0 + t_lane, r);
        tmem_ld_wait();
        if (valid) {
          float* dst = a.out1 + (b * a.N + key) * a.ld1 + half * ccols + c0;
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<float4*>(dst + g * 4) = make_float4(__uint_as_float(r[g * 4]), __uint_as_float(r[g * 4 + 1]),
                                                                  __uint_as_float(r[g * 4 + 2]), __uint_as_float(r[g * 4 + 3]));
        }
      }
      if (half == 0) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_dk + t_lane, r);
        tmem_ld_wait();
        if (valid) {
          float* dst = a.out0 + (b * a.N + key) * a.ld0;
#pragma unroll
          for (int c = 0; c < 32; c += 4)
            if (c < a.Cq)
              *reinterpret_cast<float4*>(dst + c) = make_float4(__uint_as_float(r[c]), __uint_as_float(r[c + 1]), __uint_as_float(r[c + 2]), __uint_as_float(r[c + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

std::once_flag g_attr_once;

int make_map(CUtensorMap* m, int dtype, const void* base, int inner, int N, int batch, long long ld, uint32_t box_rows) {
  uint64_t dims[3] = {static_cast<uint64_t>(inner), static_cast<uint64_t>(N), static_cast<uint64_t>(batch)};
  uint64_t strides[2] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(N) * static_cast<uint64_t>(ld) * 2};
  uint32_t box[3] = {64, box_rows, 1};
  return encode_tensor_map(m, dtype, 3, base, dims, strides, box, true);
}

}  // namespace
}  // namespace dfcsa

using namespace dfcsa;

extern "C" int dfcsa_attn_bwd_fused(const void* qkv16, int64_t ld16, const void* qkvb, int64_t ldb, const void* dO, int64_t ld_do,
                                    int32_t batch, int32_t N, int32_t Cq, int32_t C, const float* lse, const float* D,
                                    float* dqkv, int64_t ld_out, void* stream) {
  DFCSA_CHECK_ARG(qkv16 && qkvb && dO && lse && D && dqkv && batch > 0 && N > 0, "dfcsa_attn_bwd_fused: bad args");
  DFCSA_CHECK_ARG(Cq % 8 == 0 && Cq >= 8 && Cq <= 32 && (C == 64 || C == 128) && N % 8 == 0,
                  "dfcsa_attn_bwd_fused: needs Cq in {8, 16, 24, 32}, C in {64, 128}, N % 8 == 0");
  DFCSA_CHECK_ARG(ld16 % 8 == 0 && ldb % 8 == 0 && ld_do % 8 == 0 && ld_out % 4 == 0 && Cq % 4 == 0 &&
                  ((reinterpret_cast<uintptr_t>(qkv16) | reinterpret_cast<uintptr_t>(qkvb) | reinterpret_cast<uintptr_t>(dO) |
                    reinterpret_cast<uintptr_t>(dqkv) | reinterpret_cast<uintptr_t>(lse) | reinterpret_cast<uintptr_t>(D)) & 15) == 0,
                  "dfcsa_attn_bwd_fused: pitches / alignment");
  const __half* h = reinterpret_cast<const __half*>(qkv16);
  const __nv_bfloat16* bb = reinterpret_cast<const __nv_bfloat16*>(qkvb);
  const int nb = C / 64;
  BwdArgs a{};
  a.batch = batch; a.N = N; a.Cq = Cq; a.C = C;
  a.tiles128 = (N + 127) / 128; a.tiles64 = (N + 63) / 64;
  a.lse = lse; a.D = D;
  a.idesc_s = umma_idesc_f16(128, 64, 0, 0, 0, 0);
  a.idesc_dp = umma_idesc_f16(128, 64, 1, 1, 0, 0);
  a.idesc_acc0 = umma_idesc_f16(128, 64, 1, 1, 0, 1);       // dQ / dK: N padded to 64 (zero columns beyond Cq)
  a.idesc_acc1 = umma_idesc_f16(128, C, 1, 1, 0, 1);        // dV
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(g_attr_once, [] {
    attr_err = cudaFuncSetAttribute(attn_dq_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(attn_dkdv_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(attn bwd fused)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(batch) * a.tiles128;
  const int grid = static_cast<int>(std::min<long long>(total, num_sms()));
  int rc;
  {   // ---- dQ ----
    CUtensorMap mq, mdo, mk16, mkb, mvb;
    if ((rc = make_map(&mq, DFCSA_F16, h, Cq, N, batch, ld16, 128))) return rc;
    if ((rc = make_map(&mdo, DFCSA_BF16, dO, C, N, batch, ld_do, 128))) return rc;
    if ((rc = make_map(&mk16, DFCSA_F16, h + Cq, Cq, N, batch, ld16, 64))) return rc;
    if ((rc = make_map(&mkb, DFCSA_BF16, bb + Cq, Cq, N, batch, ldb, 64))) return rc;
    if ((rc = make_map(&mvb, DFCSA_BF16, bb + 2 * Cq, C, N, batch, ldb, 64))) return rc;
    const int fixed = 16384 + nb * 16384 + 2 * 16384, stage_bytes = 16384 + nb * 8192;
    a.stages = std::min(kMaxStages, (kSmemMax - 1024 - fixed) / stage_bytes);
    DFCSA_CHECK_ARG(a.stages >= 2, "dfcsa_attn_bwd_fused: shared memory budget (dq)");
    a.out0 = dqkv; a.ld0 = ld_out; a.out1 = nullptr; a.ld1 = 0;
    attn_dq_fused_kernel<<<grid, 384, fixed + a.stages * stage_bytes + 1024, st>>>(mq, mdo, mk16, mkb, mvb, a);
    DFCSA_LAUNCH_CHECK("attn_dq_fused_kernel");
  }
  {   // ---- dK, dV ----
    CUtensorMap mk16, mvb, mq16, mqb, mdo;
    if ((rc = make_map(&mk16, DFCSA_F16, h + Cq, Cq, N, batch, ld16, 128))) return rc;
    if ((rc = make_map(&mvb, DFCSA_BF16, bb + 2 * Cq, C, N, batch, ldb, 128))) return rc;
    if ((rc = make_map(&mq16, DFCSA_F16, h, Cq, N, batch, ld16, 64))) return rc;
    if ((rc = make_map(&mqb, DFCSA_BF16, bb, Cq, N, batch, ldb, 64))) return rc;
    if ((rc = make_map(&mdo, DFCSA_BF16, dO, C, N, batch, ld_do, 64))) return rc;
    const int fixed = 16384 + nb * 16384 + 4 * 16384, stage_bytes = 16384 + nb * 8192 + 1024;
    a.stages = std::min(kMaxStages, (kSmemMax - 1024 - fixed) / stage_bytes);
    DFCSA_CHECK_ARG(a.stages >= 2, "dfcsa_attn_bwd_fused: shared memory budget (dk, dv)");
    a.out0 = dqkv + Cq; a.ld0 = ld_out; a.out1 = dqkv + 2 * Cq; a.ld1 = ld_out;
    attn_dkdv_fused_kernel<<<grid, 384, fixed + a.stages * stage_bytes + 1024, st>>>(mk16, mvb, mq16, mqb, mdo, a);
    DFCSA_LAUNCH_CHECK("attn_dkdv_fused_kernel");
  }
  return DFCSA_OK;
}
