// Batched GEMM on tcgen05 (sm_100a) for the attention products:   C[b] = A[b] * B[b]   (16-bit operands, fp32 accumulate)
//
//   S  = Q K^T      A = Q  [N x Cq]  K-major,   B = K  [N x Cq]  K-major        (reference models/unet_dfc_sa_res.py:30)
//   O  = P V        A = P  [N x N]   K-major,   B = V  [N x C]   MN-major       (:33)
//   dV = P^T dO     A = P  [N x N]   MN-major,  B = dO [N x C]   MN-major
//   dP = dO V^T     A = dO [N x C]   K-major,   B = V  [N x C]   K-major
//   dQ = dS K       A = dS [N x N]   K-major,   B = K  [N x Cq]  MN-major
//   dK = dS^T Q     A = dS [N x N]   MN-major,  B = Q  [N x Cq]  MN-major
//
// "K-major" = the reduction index is contiguous in memory (rows of 64 k = one 128-byte swizzle row per m / n);
// "MN-major" = the m / n index is contiguous (a row of the stored matrix is one k): the operand is simply read
// transposed by the UMMA descriptor, no transposed copy is ever made.
//
// One persistent CTA per SM walks (batch, m tile, n tile); warp 0 = TMA producer (3-D tiled loads, out-of-range
// rows / k zero-filled by the hardware), warp 1 = tcgen05.mma issuer (M = 128, N = block_n, K = 16; two accumulators
// in TMEM so the epilogue of a tile overlaps the next main loop), warp 2 = TMEM allocator, warps 4-11 = epilogue
// (tcgen05.ld, one row per thread, 32-byte stores).
#include "common.cuh"
#include <algorithm>
#include <mutex>

namespace dfcsa {
namespace {

constexpr int kMaxStages = 8;
constexpr int kABytes = 128 * 64 * 2;            // 16 KiB: 128 m x 64 k
constexpr int kSmemBudget = 227 * 1024 - 4096;
constexpr int kAccStride = 256;

struct BgemmArgs {
  int batch, M, N, K;
  int m_tiles, n_tiles, block_n, total_kb, stages;
  int a_mn, b_mn;
  void* C; long long c_b, ld_c; int c_dtype;
  int wide;                 // 32-byte stores legal
  uint32_t idesc;
  float* rowstat;           // ROWSTATS: [batch, 2*n_tiles, M, 2] (max, sum of exp) partials, C is not written
  const float* rowvec;      // EXP: [batch, M] log-sum-exp per row, C = exp(acc - lse);  SOFTMAX_BWD: D, C = aux * (acc - D)
  const void* aux;          // SOFTMAX_BWD: the probabilities, laid out like C
  int aux_dtype;
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void load_256(const void* p, uint32_t* w) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}
__device__ __forceinline__ float f16bits_to_float(uint32_t w, int hi, int dtype) {
  const unsigned short h = hi ? static_cast<unsigned short>(w >> 16) : static_cast<unsigned short>(w & 0xffffu);
  if (dtype == DFCSA_F16) return __half2float(__ushort_as_half(h));
  return __uint_as_float(static_cast<uint32_t>(h) << 16);
}

// What the epilogue does with one accumulator row (one thread = one row of the 128-row tile), 32 columns at a time.
template <int EPI>
struct EpiRow {
  float run_max, run_sum;     // ROWSTATS (log2 domain)
  float rowc;                 // EXP: lse * log2(e);  SOFTMAX_BWD: D

  __device__ __forceinline__ void init(const BgemmArgs& a, long long b, int m, bool valid) {
    run_max = -INFINITY; run_sum = 0.f; rowc = 0.f;
    if (EPI == DFCSA_BGEMM_EPI_EXP && valid) rowc = a.rowvec[b * a.M + m] * kLog2e;
    if (EPI == DFCSA_BGEMM_EPI_SOFTMAX_BWD && valid) rowc = a.rowvec[b * a.M + m];
  }

  // SOFTMAX_BWD: the 32 probabilities next to this chunk, 64 bytes per thread
  __device__ __forceinline__ void prefetch(const BgemmArgs& a, long long off, int ncols, bool valid, uint32_t (&x)[16]) {
    if (EPI != DFCSA_BGEMM_EPI_SOFTMAX_BWD) return;
    const unsigned short* src = reinterpret_cast<const unsigned short*>(a.aux) + off;
    if (valid && a.wide && ncols == 32) {
      load_256(src, x);
      load_256(src + 16, x + 8);
    } else {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (valid && g * 8 < ncols) v = __ldg(reinterpret_cast<const uint4*>(src + g * 8));
        x[g * 4] = v.x; x[g * 4 + 1] = v.y; x[g * 4 + 2] = v.z; x[g * 4 + 3] = v.w;
      }
    }
  }

  __device__ __forceinline__ void chunk(const BgemmArgs& a, uint32_t (&raw)[32], const uint32_t (&x)[16], long long off, int ncols,
                                        bool valid) {
    if (EPI == DFCSA_BGEMM_EPI_ROWSTATS) {
      float cmax = -INFINITY;
      if (ncols == 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i) cmax = fmaxf(cmax, __uint_as_float(raw[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) if (i < ncols) cmax = fmaxf(cmax, __uint_as_float(raw[i]));
      }
      const float nmax = fmaxf(run_max, cmax * kLog2e);
      float p0 = 0.f, p1 = 0.f;
      if (ncols == 32) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          p0 += ex2f(fmaf(__uint_as_float(raw[i]), kLog2e, -nmax));
          p1 += ex2f(fmaf(__uint_as_float(raw[i + 1]), kLog2e, -nmax));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) if (i < ncols) p0 += ex2f(fmaf(__uint_as_float(raw[i]), kLog2e, -nmax));
      }
      run_sum = run_sum * (run_max == nmax ? 1.f : ex2f(run_max - nmax)) + (p0 + p1);
      run_max = nmax;
      return;
    }
    if (EPI == DFCSA_BGEMM_EPI_EXP) {
#pragma unroll
      for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(ex2f(fmaf(__uint_as_float(raw[i]), kLog2e, -rowc)));
    }
    if (EPI == DFCSA_BGEMM_EPI_SOFTMAX_BWD) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        raw[i] = __float_as_uint(f16bits_to_float(x[i >> 1], i & 1, a.aux_dtype) * (__uint_as_float(raw[i]) - rowc));
    }
    if (!valid) return;
    if (a.c_dtype == DFCSA_F32) {
      float* dst = reinterpret_cast<float*>(a.C) + off;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        if (g * 4 < ncols)
          *reinterpret_cast<float4*>(dst + g * 4) = make_float4(__uint_as_float(raw[g * 4]), __uint_as_float(raw[g * 4 + 1]),
                                                                __uint_as_float(raw[g * 4 + 2]), __uint_as_float(raw[g * 4 + 3]));
      return;
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if (g * 16 < ncols) {
        float t[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) t[i] = __uint_as_float(raw[g * 16 + i]);
        if (a.wide && ncols - g * 16 >= 16) {
          if (a.c_dtype == DFCSA_F16) store16_256<__half>(reinterpret_cast<__half*>(a.C) + off + g * 16, t);
          else store16_256<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(a.C) + off + g * 16, t);
        } else {
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            if (g * 16 + h8 * 8 < ncols) {
              float u[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) u[i] = t[h8 * 8 + i];
              if (a.c_dtype == DFCSA_F16) store8<__half>(reinterpret_cast<__half*>(a.C) + off + g * 16 + h8 * 8, u);
              else store8<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(a.C) + off + g * 16 + h8 * 8, u);
            }
          }
        }
      }
    }
  }
};

template <int EPI>
__global__ void __launch_bounds__(384, 1)
bgemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ BgemmArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = a.block_n * 128;
  const int stage_bytes = kABytes + b_bytes;
  const long long total_tiles = static_cast<long long>(a.batch) * a.m_tiles * a.n_tiles;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < a.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 8); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0;
    const uint32_t tx = static_cast<uint32_t>(stage_bytes);
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = static_cast<int>(tile % a.n_tiles);
      const long long r = tile / a.n_tiles;
      const int mt = static_cast<int>(r % a.m_tiles);
      const int b = static_cast<int>(r / a.m_tiles);
      const int m0 = mt * 128, n0 = nt * a.block_n;
      for (int kb = 0; kb < a.total_kb; ++kb) {
        if (elect_one()) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], tx);
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + kABytes;
          const int k0 = kb * 64;
          if (a.a_mn) {     // two boxes of (64 m x 64 k): smem row = k, 64 contiguous m
            tma_load_3d(sa, &map_a, &full_bar[stage], m0, k0, b);
            tma_load_3d(sa + 8192, &map_a, &full_bar[stage], m0 + 64, k0, b);
          } else {          // one box of (64 k x 128 m): smem row = m
            tma_load_3d(sa, &map_a, &full_bar[stage], k0, m0, b);
          }
          if (a.b_mn) {
            for (int j = 0; j < a.block_n / 64; ++j) tma_load_3d(sb + j * 8192, &map_b, &full_bar[stage], n0 + j * 64, k0, b);
          } else {
            tma_load_3d(sb, &map_b, &full_bar[stage], k0, n0, b);
          }
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0; uint32_t phase = 0;
    int as = 0; uint32_t aphase = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * kAccStride;
      for (int kb = 0; kb < a.total_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = a.a_mn ? umma_desc_add(umma_smem_desc(a_addr, 8192, 1024), k * 2048) : umma_desc_add(umma_smem_desc(a_addr, 16, 1024), k * 32);
            const uint64_t db = a.b_mn ? umma_desc_add(umma_smem_desc(b_addr, 8192, 1024), k * 2048) : umma_desc_add(umma_smem_desc(b_addr, 16, 1024), k * 32);
            umma_f16(d_tmem, da, db, a.idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb + 1 == a.total_kb) umma_commit(&tmem_full_bar[as]);
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
      as ^= 1; if (as == 0) aphase ^= 1;
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = (warp - 4) & 3;
    const int half = (warp - 4) >> 2;
    int as = 0; uint32_t aphase = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = static_cast<int>(tile % a.n_tiles);
      const long long r = tile / a.n_tiles;
      const int mt = static_cast<int>(r % a.m_tiles);
      const long long b = r / a.m_tiles;
      const int m = mt * 128 + ew * 32 + lane;
      const bool valid = m < a.M;
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + as * kAccStride + (static_cast<uint32_t>(ew * 32) << 16);
      const long long row_off = b * a.c_b + static_cast<long long>(m) * a.ld_c;
      const int n_tile0 = nt * a.block_n;
      const int nchunks = min(a.block_n, a.N - n_tile0 + 31) / 32;       // chunks of this tile that hold real columns
      EpiRow<EPI> row;
      row.init(a, b, m, valid);
      // two register buffers: the tcgen05.ld (and the global loads of the softmax-backward operand) of the next chunk
      // are in flight while the current one is processed
      uint32_t r0[32], r1[32];
      uint32_t x0[16], x1[16];
      int ch = half;
      if (ch < nchunks) {
        tmem_ld_32x32(t_row + ch * 32, r0);
        row.prefetch(a, row_off + n_tile0 + ch * 32, min(32, a.N - n_tile0 - ch * 32), valid, x0);
      }
      while (ch < nchunks) {
        tmem_ld_wait();
        int nx = ch + 2;
        if (nx < nchunks) {
          tmem_ld_32x32(t_row + nx * 32, r1);
          row.prefetch(a, row_off + n_tile0 + nx * 32, min(32, a.N - n_tile0 - nx * 32), valid, x1);
        }
        row.chunk(a, r0, x0, row_off + n_tile0 + ch * 32, min(32, a.N - n_tile0 - ch * 32), valid);
        ch = nx;
        if (ch >= nchunks) break;
        tmem_ld_wait();
        nx = ch + 2;
        if (nx < nchunks) {
          tmem_ld_32x32(t_row + nx * 32, r0);
          row.prefetch(a, row_off + n_tile0 + nx * 32, min(32, a.N - n_tile0 - nx * 32), valid, x0);
        }
        row.chunk(a, r1, x1, row_off + n_tile0 + ch * 32, min(32, a.N - n_tile0 - ch * 32), valid);
        ch = nx;
      }
      if (EPI == DFCSA_BGEMM_EPI_ROWSTATS && valid) {
        // partial (max, sum exp(x - max)) of this thread's columns; (-inf, 0) when it had none.  [batch, part, M, 2]
        float* rs = a.rowstat + ((b * (2LL * a.n_tiles) + 2 * nt + half) * a.M + m) * 2;
        *reinterpret_cast<float2*>(rs) = make_float2(row.run_max * kLn2, row.run_sum);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
      as ^= 1; if (as == 0) aphase ^= 1;
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

int bgemm_tc(const dfcsa_bgemm_params_t* p, cudaStream_t stream) {
  DFCSA_CHECK_ARG(p->batch > 0 && p->M > 0 && p->N > 0 && p->K > 0, "dfcsa_bgemm: empty problem");
  DFCSA_CHECK_ARG(p->ab_dtype == DFCSA_F16 || p->ab_dtype == DFCSA_BF16, "dfcsa_bgemm: operands must be fp16 or bf16");
  DFCSA_CHECK_ARG(p->A && p->B && (p->C || p->epi_mode == DFCSA_BGEMM_EPI_ROWSTATS), "dfcsa_bgemm: null pointer");
  DFCSA_CHECK_ARG(p->epi_mode == DFCSA_BGEMM_EPI_NONE || (p->epi_mode == DFCSA_BGEMM_EPI_ROWSTATS && p->rowstat) ||
                  (p->epi_mode == DFCSA_BGEMM_EPI_EXP && p->rowvec) ||
                  (p->epi_mode == DFCSA_BGEMM_EPI_SOFTMAX_BWD && p->rowvec && p->aux && p->c_dtype != DFCSA_F32 &&
                   (p->aux_dtype == DFCSA_F16 || p->aux_dtype == DFCSA_BF16) && (reinterpret_cast<uintptr_t>(p->aux) & 15) == 0),
                  "dfcsa_bgemm: bad epilogue arguments");
  DFCSA_CHECK_ARG(p->ld_a % 8 == 0 && p->ld_b % 8 == 0 && p->a_b % 8 == 0 && p->b_b % 8 == 0 &&
                  (reinterpret_cast<uintptr_t>(p->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->B) & 15) == 0,
                  "dfcsa_bgemm: operand pitches / batch strides must be multiples of 8 elements and bases 16-byte aligned");
  DFCSA_CHECK_ARG(p->epi_mode == DFCSA_BGEMM_EPI_ROWSTATS ||
                  (p->N % 8 == 0 && p->ld_c % 8 == 0 && p->c_b % 8 == 0 && (reinterpret_cast<uintptr_t>(p->C) & 15) == 0),
                  "dfcsa_bgemm: N, ld_c, c_b must be multiples of 8 and C 16-byte aligned");
  BgemmArgs a{};
  a.batch = p->batch; a.M = p->M; a.N = p->N; a.K = p->K;
  a.a_mn = p->a_mn_major ? 1 : 0; a.b_mn = p->b_mn_major ? 1 : 0;
  a.m_tiles = (p->M + 127) / 128;
  int block_n = p->N <= 256 ? (p->N + 31) / 32 * 32 : 256;
  if (a.b_mn) block_n = (block_n + 63) / 64 * 64;      // MN-major B arrives in 64-wide boxes
  if (p->N > 256) {
    int best_pad = 1 << 30;
    for (int bn = 256; bn >= 128; bn -= 64) {
      const int pad = (p->N + bn - 1) / bn * bn;
      if (pad < best_pad) { best_pad = pad; block_n = bn; }
    }
  }
  a.block_n = block_n;
  a.n_tiles = (p->N + block_n - 1) / block_n;
  a.total_kb = (p->K + 63) / 64;
  const int stage_bytes = kABytes + block_n * 128;
  a.stages = std::min(kMaxStages, (kSmemBudget - 1024) / stage_bytes);
  a.idesc = umma_idesc_f16(128, block_n, umma_fmt(p->ab_dtype), umma_fmt(p->ab_dtype), a.a_mn, a.b_mn);
  a.C = p->C; a.c_b = p->c_b; a.ld_c = p->ld_c; a.c_dtype = p->c_dtype;
  a.rowstat = p->rowstat; a.rowvec = p->rowvec; a.aux = p->aux; a.aux_dtype = p->aux_dtype;
  a.wide = p->c_dtype != DFCSA_F32 && p->ld_c % 16 == 0 && p->c_b % 16 == 0 && (reinterpret_cast<uintptr_t>(p->C) & 31) == 0 &&
           (p->epi_mode != DFCSA_BGEMM_EPI_SOFTMAX_BWD || (reinterpret_cast<uintptr_t>(p->aux) & 31) == 0);

  CUtensorMap map_a, map_b;
  uint64_t dims[3], strides[2];
  uint32_t box[3];
  const uint64_t nb = static_cast<uint64_t>(p->batch);
  {
    const uint64_t rows = a.a_mn ? p->K : p->M, inner = a.a_mn ? p->M : p->K;
    dims[0] = inner; dims[1] = rows; dims[2] = nb;
    strides[0] = static_cast<uint64_t>(p->ld_a) * 2;
    strides[1] = p->batch > 1 ? static_cast<uint64_t>(p->a_b) * 2 : rows * strides[0];
    box[0] = 64; box[1] = a.a_mn ? 64 : 128; box[2] = 1;
    int rc = encode_tensor_map(&map_a, p->ab_dtype, 3, p->A, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    const uint64_t rows = a.b_mn ? p->K : p->N, inner = a.b_mn ? p->N : p->K;
    dims[0] = inner; dims[1] = rows; dims[2] = nb;
    strides[0] = static_cast<uint64_t>(p->ld_b) * 2;
    strides[1] = p->batch > 1 ? static_cast<uint64_t>(p->b_b) * 2 : rows * strides[0];
    box[0] = 64; box[1] = a.b_mn ? 64 : static_cast<uint32_t>(block_n); box[2] = 1;
    int rc = encode_tensor_map(&map_b, p->ab_dtype, 3, p->B, dims, strides, box, true);
    if (rc) return rc;
  }
  const int smem_bytes = a.stages * stage_bytes + 1024;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(g_attr_once, [] {
    attr_err = cudaFuncSetAttribute(bgemm_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(bgemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(bgemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(bgemm_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(bgemm_tc_kernel)");
  const long long total_tiles = static_cast<long long>(p->batch) * a.m_tiles * a.n_tiles;
  const int grid = static_cast<int>(std::min<long long>(total_tiles, num_sms()));
  switch (p->epi_mode) {
    case DFCSA_BGEMM_EPI_ROWSTATS:    bgemm_tc_kernel<1><<<grid, 384, smem_bytes, stream>>>(map_a, map_b, a); break;
    case DFCSA_BGEMM_EPI_EXP:         bgemm_tc_kernel<2><<<grid, 384, smem_bytes, stream>>>(map_a, map_b, a); break;
    case DFCSA_BGEMM_EPI_SOFTMAX_BWD: bgemm_tc_kernel<3><<<grid, 384, smem_bytes, stream>>>(map_a, map_b, a); break;
    default:                          bgemm_tc_kernel<0><<<grid, 384, smem_bytes, stream>>>(map_a, map_b, a); break;
  }
  DFCSA_LAUNCH_CHECK("bgemm_tc_kernel");
  return DFCSA_OK;
}

}  // namespace dfcsa

namespace dfcsa {
namespace {
// lse[r] = log sum_j exp(S[r, j]) from the per-tile partials (max_i, sum_i) written by the ROWSTATS epilogue
// (rowstat is [batch, parts, M, 2]: consecutive threads = consecutive rows read consecutive float2)
__global__ void lse_combine_kernel(const float* __restrict__ rowstat, int parts, int M, long long rows, float* __restrict__ lse) {
  const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const long long b = r / M;
  const float2* p = reinterpret_cast<const float2*>(rowstat) + b * parts * M + (r - b * M);
  float mx = -INFINITY, s = 0.f;                 // online merge, one pass over the partials
  for (int i = 0; i < parts; ++i) {
    const float2 v = __ldg(p + static_cast<long long>(i) * M);
    if (v.y > 0.f) {
      const float nm = fmaxf(mx, v.x);
      s = s * __expf(mx - nm) + v.y * __expf(v.x - nm);
      mx = nm;
    }
  }
  lse[r] = mx + logf(s);
}
}  // namespace
}  // namespace dfcsa

extern "C" int dfcsa_bgemm_rowstat_parts(int32_t N) {
  int block_n = N <= 256 ? (N + 31) / 32 * 32 : 256;
  if (N > 256) {
    int best_pad = 1 << 30;
    for (int bn = 256; bn >= 128; bn -= 64) {
      const int pad = (N + bn - 1) / bn * bn;
      if (pad < best_pad) { best_pad = pad; block_n = bn; }
    }
  }
  return 2 * ((N + block_n - 1) / block_n);
}

extern "C" int dfcsa_lse_combine(const float* rowstat, int32_t parts, int32_t batch, int32_t M, float* lse, void* stream) {
  DFCSA_CHECK_ARG(rowstat && lse && parts > 0 && batch > 0 && M > 0, "dfcsa_lse_combine: bad args");
  const long long rows = static_cast<long long>(batch) * M;
  dfcsa::lse_combine_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(rowstat, parts, M, rows, lse);
  DFCSA_LAUNCH_CHECK("lse_combine_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_bgemm(const dfcsa_bgemm_params_t* p, void* stream) {
  DFCSA_CHECK_ARG(p != nullptr, "dfcsa_bgemm: null params");
  return dfcsa::bgemm_tc(p, static_cast<cudaStream_t>(stream));
}
