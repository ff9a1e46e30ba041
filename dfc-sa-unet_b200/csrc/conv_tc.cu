// Implicit-GEMM convolution on tcgen05 (sm_100a).
//
//   out[m, n] = sum_{seg, tap, c} src_seg[pix(m, tap), c] * w[n, k(seg, tap, c)]
//
// One persistent CTA per SM walks (m-tile, n-tile) pairs.  Warp roles:
//   warp 0   : TMA producer  - one 5-D tiled TMA per K block for the activation patch (the 3x3 taps are the same
//              box shifted by (dh, dw); TMA zero-fills the halo), one 2-D TMA for the packed weights
//   warp 1   : MMA issuer    - lane 0 issues tcgen05.mma (M=128, N=block_n, K=16) into a double-buffered TMEM
//              accumulator; tcgen05.commit releases smem stages / publishes the accumulator
//   warp 2   : TMEM allocator
//   warps 4-11: epilogue     - tcgen05.ld (one pixel row per thread; two warps per TMEM lane quarter split the
//              columns), optional bias / accumulate, per-channel sum / sum^2 (BatchNorm batch statistics) by a
//              shuffle transpose-reduce, 16-byte stores
// Both operands are K-major, 128-byte swizzled: a pixel's 64 channels (128 B) form one swizzle row.
#include "common.cuh"
#include <algorithm>
#include <mutex>

namespace dfcsa {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kMaxStages = 8;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kSmemBudget = 227 * 1024 - 12288;  // dynamic part; ~8 KB of static barriers / stats live beside it
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;                 // TMEM columns between the two accumulators

struct ConvTcArgs {
  int n_seg;
  int seg_kb[3];    // channels / 64
  int seg_mode[3];
  int total_kb;
  int tiles_w, tiles_h, tiles_b, w_t, h_t;
  int H, W;         // output grid as seen by the tile walker (B merged into H when there is no halo)
  int n_tiles_n, block_n, N;
  int stages;
  int kbs;          // K blocks (of 64) per pipeline stage: 2 for narrow tiles so one barrier round-trip feeds 8 MMAs
  int dw3;          // 3x3 segments load ONE halo box (10 x 16 pixels) per kernel row and K block and read the three dw
                    // taps from it through UMMA descriptors shifted by 128 bytes (pixel patch 8 x 16, SBO = 10 rows)
  int stage_bytes;
  int a_box_bytes;  // bytes one activation TMA delivers
  void* out;
  long long ld_out;
  int out_dtype, out_mode, accumulate;
  const float* bias;
  double* stats;
  int convt_co, convt_h, convt_w;
  uint32_t idesc;
  __nv_bfloat16* shadow;
  long long ld_shadow;
  int fixed_nt;     // gridDim.x is a multiple of n_tiles_n: a CTA sees ONE n tile for the whole kernel (nt == blockIdx.x % n_tiles_n)
  int wide_out, wide_shadow;   // 32-byte stores legal (16-bit tensor, pitch % 16 == 0, base 32-byte aligned, no accumulate)
  int act_cols;     // ReLU on the output columns n < act_cols (0: no activation)
  int stats_cols;   // BatchNorm sums only for the output columns n < stats_cols
  int two_cta;      // CTA pairs: M = 256 per tcgen05.mma (cta_group::2), each CTA stages half of the B tile
  int has_bn;       // fold the BatchNorm finalize into this launch (last CTA by ticket)
  dfcsa_bn_fold_t bn;
  int epi_mode;     // DFCSA_EPI_*: fused inference epilogue (conv_tc_kernel<false, false, true> only)
  const __half* epi_p; const __half* epi_q;
  long long ld_epi;
  const float* epi_scale;
};

struct TileCoord { int nt, w0, h0, tb; };

__device__ __forceinline__ TileCoord tile_coord(const ConvTcArgs& a, int tile) {
  TileCoord t;
  if (a.n_tiles_n == 1) {
    t.nt = 0;
    if (a.tiles_h == 1 && a.tiles_b == 1) { t.w0 = tile * a.w_t; t.h0 = 0; t.tb = 0; return t; }   // flattened 1x1 GEMM: no divisions
  }
  t.nt = tile % a.n_tiles_n;
  int mt = tile / a.n_tiles_n;
  t.w0 = (mt % a.tiles_w) * a.w_t;
  int r = mt / a.tiles_w;
  t.h0 = (r % a.tiles_h) * a.h_t;
  t.tb = r / a.tiles_h;
  return t;
}

// CTA pairs: pair tile pt = (pair of adjacent m tiles, n tile); CTA `rank` of the pair owns m tile 2 * mp + rank, which may lie
// past the last one (odd tile count): its TMA boxes are then out of range (zero fill) and its rows are masked (tb >= tiles_b)
__device__ __forceinline__ TileCoord tile_coord_pair(const ConvTcArgs& a, int pt, int rank) {
  TileCoord t;
  t.nt = pt % a.n_tiles_n;
  const int mt = 2 * (pt / a.n_tiles_n) + rank;
  t.w0 = (mt % a.tiles_w) * a.w_t;
  const int r = mt / a.tiles_w;
  t.h0 = (r % a.tiles_h) * a.h_t;
  t.tb = r / a.tiles_h;
  return t;
}

// 32 columns of a 16-bit tensor as two 32-byte (whole-sector) stores; dst is 32-byte aligned
template <typename TOut>
__device__ __forceinline__ void store_chunk_wide(TOut* dst, float (&v)[32]) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = v[g * 16 + i];
    store16_256<TOut>(dst + g * 16, t);
  }
}

template <typename TOut>
__device__ __forceinline__ void store_chunk(TOut* dst, float (&v)[32], int ncols, bool accumulate) {
  // ncols is a multiple of 8; dst is 16-byte aligned
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (g * 8 < ncols) {
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = v[g * 8 + i];
      if (accumulate) {
        float o[8];
        load8<TOut>(dst + g * 8, o);
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] += o[i];
      }
      store8<TOut>(dst + g * 8, t);
    }
  }
}

// lane j ends with sum over the 32 lanes of v[j]
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      float send = upper ? v[i] : v[i + off];
      float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// REG_STATS: the narrow-output variant that keeps the BatchNorm partial sums in registers (64 more registers per
// epilogue thread - kept out of the general instantiation, whose epilogue got measurably slower at 166 registers)
template <bool REG_STATS, bool TWO = false, bool EPI = false>
__global__ void __launch_bounds__(384, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ ConvTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_stats[4][2][256];   // [TMEM lane quarter][sum | sumsq][column]: one writer warp per entry

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = a.stage_bytes;
  // TWO: the two CTAs of a cluster work on one 256-pixel x block_n tile; `cta` / `n_cta` count pairs, `total_tiles` pair tiles
  const int rank = TWO ? static_cast<int>(cluster_ctarank()) : 0;
  const int cta = TWO ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int n_cta = TWO ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int total_tiles = TWO ? ((a.tiles_w * a.tiles_h * a.tiles_b + 1) / 2) * a.n_tiles_n
                              : a.tiles_w * a.tiles_h * a.tiles_b * a.n_tiles_n;

  // No zero fill of the stages: every B box is written in full by TMA (out-of-range rows arrive as zeros), and an
  // activation-tile row a partial box leaves unwritten is a pixel row of the MMA's M axis - it can only reach its own
  // output row, which the epilogue masks (valid == false) before anything is stored or summed.
  // Without BatchNorm statistics the s_stats array is free: it holds the bias (indexed by GEMM column n; for a ConvT
  // the bias repeats every convt_co columns) so the epilogue reads it as shared-memory broadcasts.
  const bool bias_smem = a.bias != nullptr && a.stats == nullptr && a.N <= 4 * 2 * 256 && a.N % 32 == 0;
  for (int i = threadIdx.x; i < 4 * 2 * 256; i += blockDim.x) {
    float b = 0.f;
    if (bias_smem && i < a.N) b = a.bias[a.out_mode == DFCSA_OUT_CONVT2x2 ? i % a.convt_co : i];
    (&s_stats[0][0][0])[i] = b;
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    if (a.n_seg > 1) tma_prefetch_desc(&map_a1);
    if (a.n_seg > 2) tma_prefetch_desc(&map_a2);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < a.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    // the leader's "accumulator drained" barrier collects the epilogue warps of BOTH CTAs of a pair
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], TWO ? 16 : 8); }
    fence_barrier_init();
  }
  if (warp == 2) { if constexpr (TWO) tmem_alloc_2cta(&tmem_base_smem, kTmemCols); else tmem_alloc(&tmem_base_smem, kTmemCols); }
  tc_fence_before();
  if constexpr (TWO) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0;
    const uint32_t kb_tx = static_cast<uint32_t>(a.a_box_bytes + a.block_n * 128);
    for (int tile = cta; tile < total_tiles; tile += n_cta) {
      const TileCoord tc = TWO ? tile_coord_pair(a, tile, rank) : tile_coord(a, tile);
      if (a.dw3) {
        constexpr int kHaloBytes = 10 * 16 * 128;      // 20 KiB: (8 + 2) x 16 pixels x 64 channels
        const int bn_bytes = a.block_n * 128;
        int k_base = 0;                                 // K offset (in 64-blocks) of the current segment in the packed weights
        for (int s = 0; s < a.n_seg; ++s) {
          const CUtensorMap* ma = (s == 0) ? &map_a0 : (s == 1) ? &map_a1 : &map_a2;
          const bool m3 = a.seg_mode[s] == DFCSA_TAP_3x3;
          const int rows = m3 ? 3 : 1;
          for (int dh = 0; dh < rows; ++dh) {
            for (int kb = 0; kb < a.seg_kb[s]; ++kb) {
              if (elect_one()) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * stage_bytes;
                uint8_t* sb = sa + kHaloBytes;
                if (m3) {
                  mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(kHaloBytes + 3 * bn_bytes));
                  tma_load_5d(sa, ma, &full_bar[stage], kb * kBlockK, tc.w0 - 1, tc.h0 + dh - 1, tc.tb, 0);
                  for (int dw = 0; dw < 3; ++dw)
                    tma_load_2d(sb + dw * bn_bytes, &map_b, &full_bar[stage],
                                (k_base + (dh * 3 + dw) * a.seg_kb[s] + kb) * kBlockK, tc.nt * a.block_n);
                } else {
                  mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(kABytes + bn_bytes));
                  tma_load_5d(sa, ma, &full_bar[stage], kb * kBlockK, tc.w0, tc.h0, tc.tb, 0);
                  tma_load_2d(sb, &map_b, &full_bar[stage], (k_base + kb) * kBlockK, tc.nt * a.block_n);
                }
              }
              __syncwarp();
              if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
          }
          k_base += (m3 ? 9 : 1) * a.seg_kb[s];
        }
        continue;
      }
      int kb_global = 0, sub = 0;
      for (int s = 0; s < a.n_seg; ++s) {
        const CUtensorMap* ma = (s == 0) ? &map_a0 : (s == 1) ? &map_a1 : &map_a2;
        const int mode = a.seg_mode[s];
        const int taps = (mode == DFCSA_TAP_1x1) ? 1 : (mode == DFCSA_TAP_3x3) ? 9 : 4;
        for (int t = 0; t < taps; ++t) {
          int c1, c2, c3, c4;
          if (mode == DFCSA_TAP_2x2S2) { c1 = t & 1; c2 = tc.w0; c3 = t >> 1; c4 = tc.h0; }
          else if (mode == DFCSA_TAP_3x3) { c1 = tc.w0 + (t % 3) - 1; c2 = tc.h0 + (t / 3) - 1; c3 = tc.tb; c4 = 0; }
          else { c1 = tc.w0; c2 = tc.h0; c3 = tc.tb; c4 = 0; }
          for (int kb = 0; kb < a.seg_kb[s]; ++kb) {
            if (elect_one()) {
              if constexpr (TWO) {
                // each CTA loads its own 128 pixel rows of A and its HALF of the B tile; both signal the leader's barrier,
                // which expects the bytes of both (the peer's bytes may land before the leader has armed the phase: the
                // transaction count just goes negative for a moment)
                const int bh = a.block_n / 2;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * static_cast<uint32_t>(a.a_box_bytes + bh * 128));
                uint8_t* sa = smem + stage * stage_bytes;
                uint8_t* sb = sa + kABytes;
                tma_load_5d_2cta(sa, ma, &full_bar[stage], kb * kBlockK, c1, c2, c3, c4);
                tma_load_2d_2cta(sb, &map_b, &full_bar[stage], kb_global * kBlockK, tc.nt * a.block_n + rank * bh);
              } else {
                if (sub == 0) {
                  mbar_wait(&empty_bar[stage], phase ^ 1);
                  mbar_arrive_expect_tx(&full_bar[stage], kb_tx * static_cast<uint32_t>(min(a.kbs, a.total_kb - kb_global)));
                }
                uint8_t* sa = smem + stage * stage_bytes + sub * kABytes;
                uint8_t* sb = smem + stage * stage_bytes + a.kbs * kABytes + sub * (a.block_n * 128);
                tma_load_5d(sa, ma, &full_bar[stage], kb * kBlockK, c1, c2, c3, c4);
                tma_load_2d(sb, &map_b, &full_bar[stage], kb_global * kBlockK, tc.nt * a.block_n);
              }
            }
            __syncwarp();
            ++kb_global;
            if (++sub == a.kbs || kb_global == a.total_kb) {
              sub = 0;
              if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1 && (!TWO || rank == 0)) {
    // ===================== MMA issuer (pairs: the leader CTA only) =====================
    int stage = 0; uint32_t phase = 0;
    int as = 0; uint32_t aphase = 0;
    for (int tile = cta; tile < total_tiles; tile += n_cta) {
      mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * kAccStride;
      if (a.dw3) {
        constexpr int kHaloBytes = 10 * 16 * 128;
        const int bn_bytes = a.block_n * 128;
        int items = 0;
        for (int s = 0; s < a.n_seg; ++s) items += (a.seg_mode[s] == DFCSA_TAP_3x3 ? 3 : 1) * a.seg_kb[s];
        int it = 0;
        for (int s = 0; s < a.n_seg; ++s) {
          const bool m3 = a.seg_mode[s] == DFCSA_TAP_3x3;
          const int n_it = (m3 ? 3 : 1) * a.seg_kb[s];
          for (int j = 0; j < n_it; ++j, ++it) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
              const uint32_t b_addr = a_addr + kHaloBytes;
              // one descriptor per operand and stage; every tap / k step is a constant further on (umma_desc_add)
              if (m3) {
                // patch row g = 8 consecutive pixels = one 8-row core group; groups are 10 box rows apart
                const uint64_t da0 = umma_smem_desc(a_addr, 16, 1280);
                uint64_t db0 = umma_smem_desc(b_addr, 16, 1024);
#pragma unroll
                for (int dw = 0; dw < 3; ++dw) {
#pragma unroll
                  for (int k = 0; k < kBlockK / 16; ++k)
                    umma_f16(d_tmem, umma_desc_add(da0, dw * 128 + k * 32), umma_desc_add(db0, k * 32), a.idesc,
                             (dw | k) != 0 ? 1u : (it != 0 ? 1u : 0u));
                  db0 = umma_desc_add(db0, bn_bytes);
                }
              } else {
                const uint64_t da0 = umma_smem_desc(a_addr, 16, 1024), db0 = umma_smem_desc(b_addr, 16, 1024);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma_f16(d_tmem, umma_desc_add(da0, k * 32), umma_desc_add(db0, k * 32), a.idesc, k != 0 ? 1u : (it != 0 ? 1u : 0u));
              }
              umma_commit(&empty_bar[stage]);
              if (it + 1 == items) umma_commit(&tmem_full_bar[as]);
            }
            __syncwarp();
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
          }
        }
        as ^= 1; if (as == 0) aphase ^= 1;
        continue;
      }
      for (int kb = 0; kb < a.total_kb; kb += a.kbs) {
        const int nsub = min(a.kbs, a.total_kb - kb);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint32_t b_addr = a_addr + a.kbs * kABytes;
          uint64_t da0 = umma_smem_desc(a_addr, 16, 1024), db0 = umma_smem_desc(b_addr, 16, 1024);
          if constexpr (TWO) {
            // one M = 256 instruction per 16 channels: rows 0-127 from this CTA's A tile, 128-255 from the peer's (same
            // shared-memory offsets), the N columns split between the two CTAs' B halves; accumulators in both CTAs' TMEM
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_f16_2cta(d_tmem, umma_desc_add(da0, k * 32), umma_desc_add(db0, k * 32), a.idesc, k != 0 ? 1u : (kb != 0 ? 1u : 0u));
            umma_commit_2cta(&empty_bar[stage]);                  // frees the stage in both CTAs
            if (kb + nsub >= a.total_kb) umma_commit_2cta(&tmem_full_bar[as]);
          } else {
            const uint32_t b_sub = static_cast<uint32_t>(a.block_n * 128);
            for (int sub = 0; sub < nsub; ++sub) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma_f16(d_tmem, umma_desc_add(da0, k * 32), umma_desc_add(db0, k * 32), a.idesc, k != 0 ? 1u : ((kb | sub) != 0 ? 1u : 0u));
              da0 = umma_desc_add(da0, kABytes);
              db0 = umma_desc_add(db0, b_sub);
            }
            umma_commit(&empty_bar[stage]);
            if (kb + nsub >= a.total_kb) umma_commit(&tmem_full_bar[as]);
          }
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
      as ^= 1; if (as == 0) aphase ^= 1;
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = (warp - 4) & 3;        // == warp % 4: the TMEM lane quarter this warp may read
    const int half = (warp - 4) >> 2;     // which half of the column chunks this warp takes
    const int r = ew * 32 + lane;         // tile row == TMEM lane
    const int et = threadIdx.x - 128;
    const bool flush_each = a.n_tiles_n > 1 && !a.fixed_nt;
    // Narrow outputs (block_n <= 64, one n tile): every epilogue warp owns ONE fixed 32-column chunk for the whole
    // kernel, so the BatchNorm partial sums stay in registers across tiles and the 2 x 31-shuffle transpose runs once
    // per kernel instead of once per tile (these level-1 GEMMs are bound by the epilogue's instruction issue).
    constexpr bool reg_stats = REG_STATS;
    float rs[REG_STATS ? 32 : 1], rq[REG_STATS ? 32 : 1];
#pragma unroll
    for (int i = 0; i < (REG_STATS ? 32 : 1); ++i) { rs[i] = 0.f; rq[i] = 0.f; }
    int as = 0; uint32_t aphase = 0;
    int last_nt = 0;
    bool have_stats = false;
    const int dw = r % a.w_t, dh = r / a.w_t;       // this thread's pixel inside every tile
    for (int tile = cta; tile < total_tiles; tile += n_cta) {
      const TileCoord tc = TWO ? tile_coord_pair(a, tile, rank) : tile_coord(a, tile);
      const int w = tc.w0 + dw, h = tc.h0 + dh;
      const bool valid = (r < a.w_t * a.h_t) && (w < a.W) && (h < a.H) && (!TWO || tc.tb < a.tiles_b);
      long long row_off;
      if (a.out_mode == DFCSA_OUT_CONVT2x2) {
        // flattened input pixel m -> (b, i, j); quadrant added per chunk below
        // (the host checks that the pixel count fits 31 bits: 32-bit divisions, ~10x cheaper than the 64-bit ones)
        const unsigned m = (static_cast<unsigned>(tc.tb) * a.H + h) * a.W + w;
        const unsigned j = m % static_cast<unsigned>(a.convt_w);
        const unsigned t2 = m / static_cast<unsigned>(a.convt_w);
        const unsigned i = t2 % static_cast<unsigned>(a.convt_h);
        const long long b = t2 / static_cast<unsigned>(a.convt_h);
        row_off = ((b * 2 * a.convt_h + 2 * i) * (2LL * a.convt_w) + 2 * j);  // pixel index of quadrant (0,0)
      } else {
        row_off = (static_cast<long long>(tc.tb) * a.H + h) * a.W + w;
      }
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const int nchunks = a.block_n / 32;
      for (int ch = half; ch < nchunks; ch += 2) {
        const int n0 = tc.nt * a.block_n + ch * 32;
        if (n0 >= a.N) break;
        const int ncols = min(32, a.N - n0);
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + as * kAccStride + ch * 32 + (static_cast<uint32_t>(ew * 32) << 16), raw);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
        long long pix = row_off;
        int cn = n0;
        if (a.out_mode == DFCSA_OUT_CONVT2x2) {
          const int q = n0 / a.convt_co;
          cn = n0 - q * a.convt_co;
          pix += (q >> 1) * (2LL * a.convt_w) + (q & 1);
        }
        if (bias_smem) {          // whole 32-column chunk, staged once per CTA: 8 LDS.128 broadcasts instead of 32 loads
          const float4* bs = reinterpret_cast<const float4*>(&s_stats[0][0][0] + n0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 t = bs[i];
            v[i * 4] += t.x; v[i * 4 + 1] += t.y; v[i * 4 + 2] += t.z; v[i * 4 + 3] += t.w;
          }
        } else if (a.bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i) if (i < ncols) v[i] += __ldg(a.bias + cn + i);
        }
        if (n0 < a.act_cols) {      // folded conv + BatchNorm + ReLU of the inference path
#pragma unroll
          for (int i = 0; i < 32; ++i) if (n0 + i < a.act_cols) v[i] = fmax_nan(v[i], 0.f);
        }
        if constexpr (EPI) {
          // fused inference epilogues (own instantiation: the training epilogue keeps its register budget).  The second
          // operands are the rows this tile's TMA loads brought through L2 a moment ago (gate mix) or a tensor the
          // separate pass would have read anyway (residual).
          if (valid && ncols == 32) {
            const uint4* pp = reinterpret_cast<const uint4*>(a.epi_p + pix * a.ld_epi + cn);
            uint4 pr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) pr[i] = __ldg(pp + i);
            const __half2* ph = reinterpret_cast<const __half2*>(pr);
            if (a.epi_mode == DFCSA_EPI_GATE_MIX) {
              const uint4* qp = reinterpret_cast<const uint4*>(a.epi_q + pix * a.ld_epi + cn);
              uint4 qr[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) qr[i] = __ldg(qp + i);
              const __half2* qh = reinterpret_cast<const __half2*>(qr);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float2 l = __half22float2(ph[i]), q = __half22float2(qh[i]);
                const float g0 = __fdividef(1.f, 1.f + __expf(-v[2 * i])), g1 = __fdividef(1.f, 1.f + __expf(-v[2 * i + 1]));
                v[2 * i] = fmaf(g0, l.x - q.x, q.x);
                v[2 * i + 1] = fmaf(g1, l.y - q.y, q.y);
              }
            } else {
              const float sc = __ldg(a.epi_scale);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float2 l = __half22float2(ph[i]);
                v[2 * i] = fmaf(sc, l.x, v[2 * i]);
                v[2 * i + 1] = fmaf(sc, l.y, v[2 * i + 1]);
              }
            }
          }
        }
        if (valid) {
          if (a.wide_out && ncols == 32) {
            if (a.out_dtype == DFCSA_F16) store_chunk_wide<__half>(reinterpret_cast<__half*>(a.out) + pix * a.ld_out + cn, v);
            else store_chunk_wide<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(a.out) + pix * a.ld_out + cn, v);
          } else if (a.out_dtype == DFCSA_F16)
            store_chunk<__half>(reinterpret_cast<__half*>(a.out) + pix * a.ld_out + cn, v, ncols, a.accumulate != 0);
          else if (a.out_dtype == DFCSA_BF16)
            store_chunk<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(a.out) + pix * a.ld_out + cn, v, ncols, a.accumulate != 0);
          else
            store_chunk<float>(reinterpret_cast<float*>(a.out) + pix * a.ld_out + cn, v, ncols, a.accumulate != 0);
          if (a.shadow != nullptr) {
            if (a.wide_shadow && ncols == 32) store_chunk_wide<__nv_bfloat16>(a.shadow + pix * a.ld_shadow + cn, v);
            else store_chunk<__nv_bfloat16>(a.shadow + pix * a.ld_shadow + cn, v, ncols, false);
          }
        }
        if (n0 >= a.stats_cols) continue;       // e.g. the residual half of the [W2 ; W5] GEMM: no BatchNorm behind it
        if constexpr (REG_STATS) {
          if (valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) { rs[i] += v[i]; rq[i] = fmaf(v[i], v[i], rq[i]); }
          }
        } else if (a.stats != nullptr) {
          float sq[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) { v[i] = valid ? v[i] : 0.f; sq[i] = v[i] * v[i]; }
          const float s1 = transpose_reduce32(v, lane);
          const float s2 = transpose_reduce32(sq, lane);
          s_stats[ew][0][ch * 32 + lane] += s1;     // this warp is the only writer of (ew, column)
          s_stats[ew][1][ch * 32 + lane] += s2;
          have_stats = true;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if constexpr (TWO) mbar_arrive_leader(&tmem_empty_bar[as]); else mbar_arrive(&tmem_empty_bar[as]); }
      as ^= 1; if (as == 0) aphase ^= 1;
      last_nt = tc.nt;
      if (a.stats != nullptr && flush_each) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int c = et; c < a.block_n; c += 256) {
          const int n = last_nt * a.block_n + c;
          if (n < a.N) {
            atomicAdd(a.stats + n, static_cast<double>(s_stats[0][0][c] + s_stats[1][0][c] + s_stats[2][0][c] + s_stats[3][0][c]));
            atomicAdd(a.stats + a.N + n, static_cast<double>(s_stats[0][1][c] + s_stats[1][1][c] + s_stats[2][1][c] + s_stats[3][1][c]));
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) { s_stats[q][0][c] = 0.f; s_stats[q][1][c] = 0.f; }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
    (void)have_stats;
    if constexpr (REG_STATS) {
      if (half < a.block_n / 32) {       // this warp's chunk is ch == half
        const float s1 = transpose_reduce32(rs, lane);
        const float s2 = transpose_reduce32(rq, lane);
        s_stats[ew][0][half * 32 + lane] += s1;
        s_stats[ew][1][half * 32 + lane] += s2;
      }
    }
    (void)reg_stats;
  }
  // final per-CTA flush of the BatchNorm partial sums (single n-tile case)
  tc_fence_before();
  // pairs: neither CTA may free its TMEM / leave while the other still runs MMAs on its shared memory or reads the accumulators
  if constexpr (TWO) cluster_sync_all(); else __syncthreads();
  if (a.stats != nullptr && (a.n_tiles_n == 1 || a.fixed_nt) && cta < total_tiles) {
    const int n_base = (blockIdx.x % a.n_tiles_n) * a.block_n;     // 0 for a single n tile
    for (int c = threadIdx.x; c < a.block_n; c += blockDim.x) {
      const int n = n_base + c;
      if (n < a.N) {
        atomicAdd(a.stats + n, static_cast<double>(s_stats[0][0][c] + s_stats[1][0][c] + s_stats[2][0][c] + s_stats[3][0][c]));
        atomicAdd(a.stats + a.N + n, static_cast<double>(s_stats[0][1][c] + s_stats[1][1][c] + s_stats[2][1][c] + s_stats[3][1][c]));
      }
    }
  }
  if (warp == 2) {
    tc_fence_after();
    if constexpr (TWO) tmem_dealloc_2cta(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
  if (a.has_bn) {
    // BatchNorm finalize by the last CTA: every CTA's double atomics above are ordered before its ticket by the fence; the
    // CTA that draws the last ticket therefore sees the complete sums (read with ld.global.cg, i.e. from L2 where the
    // atomics were performed).  Nobody waits, so CTAs need not be co-resident.
    __shared__ unsigned s_ticket;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(a.bn.ticket, 1u);
    __syncthreads();
    if (s_ticket == gridDim.x - 1) {
      __threadfence();
      const double n = static_cast<double>(a.bn.count);
      for (int c = threadIdx.x; c < a.bn.channels; c += blockDim.x) {
        const double mean = __ldcg(a.stats + c) / n;
        double var = __ldcg(a.stats + a.N + c) / n - mean * mean;
        if (var < 0.0) var = 0.0;
        const double invstd = rsqrt(var + static_cast<double>(a.bn.eps));
        const float g = a.bn.gamma[c];
        a.bn.scale[c] = static_cast<float>(g * invstd);
        a.bn.shift[c] = static_cast<float>(a.bn.beta[c] - mean * g * invstd);
        a.bn.mean[c] = static_cast<float>(mean);
        a.bn.invstd[c] = static_cast<float>(invstd);
        if (a.bn.running_mean != nullptr) {
          const double b = a.bn.conv_bias ? a.bn.conv_bias[c] : 0.0;
          const double mom = a.bn.momentum;
          a.bn.running_mean[c] = static_cast<float>((1.0 - mom) * a.bn.running_mean[c] + mom * (mean + b));
          const double unb = a.bn.count > 1 ? var * n / (n - 1.0) : var;
          a.bn.running_var[c] = static_cast<float>((1.0 - mom) * a.bn.running_var[c] + mom * unb);
        }
      }
    }
  }
}

std::once_flag g_attr_once;

void pick_spatial_tile(int H, int W, int& w_t, int& h_t) {
  double best = -1.0;
  w_t = 1; h_t = 1;
  for (int wt = 1; wt <= std::min(W, kBlockM); ++wt) {
    int ht = std::min(H, kBlockM / wt);
    if (ht < 1) continue;
    long long tiles = static_cast<long long>((W + wt - 1) / wt) * ((H + ht - 1) / ht);
    double eff = static_cast<double>(H) * W / (static_cast<double>(tiles) * kBlockM);
    if (eff > best + 1e-9 || (eff > best - 1e-9 && wt > w_t)) { best = eff; w_t = wt; h_t = ht; }
  }
}

}  // namespace

int conv_gemm_tc(const dfcsa_conv_params_t* p, cudaStream_t stream) {
  DFCSA_CHECK_ARG(p->n_seg >= 1 && p->n_seg <= 3, "conv_gemm_tc: n_seg must be 1..3");
  DFCSA_CHECK_ARG(p->src_dtype != DFCSA_F32 && p->w_dtype != DFCSA_F32, "conv_gemm_tc: 16-bit operands required");
  // measured on B200: a kind::f16 tcgen05.mma with A fp16 and B bf16 (or the reverse) raises an illegal instruction
  DFCSA_CHECK_ARG(p->src_dtype == p->w_dtype, "conv_gemm_tc: activations and weights must share one 16-bit format");
  DFCSA_CHECK_ARG(p->N % 8 == 0 && p->ld_out % 8 == 0, "conv_gemm_tc: N and ld_out must be multiples of 8");
  DFCSA_CHECK_ARG((reinterpret_cast<uintptr_t>(p->out) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->w) & 15) == 0,
                  "conv_gemm_tc: out / w must be 16-byte aligned");
  const long long Mtot = static_cast<long long>(p->B) * p->H * p->W;
  DFCSA_CHECK_ARG(Mtot > 0, "conv_gemm_tc: empty problem");
  bool any3 = false, any2 = false;
  int ktot = 0;
  for (int s = 0; s < p->n_seg; ++s) {
    const dfcsa_seg_t& sg = p->seg[s];
    DFCSA_CHECK_ARG(sg.channels > 0 && sg.channels % 64 == 0, "conv_gemm_tc: segment channels must be a multiple of 64 (got %d)", sg.channels);
    DFCSA_CHECK_ARG(sg.ld % 8 == 0 && (reinterpret_cast<uintptr_t>(sg.ptr) & 15) == 0, "conv_gemm_tc: segment pitch/alignment");
    any3 |= sg.tap_mode == DFCSA_TAP_3x3;
    any2 |= sg.tap_mode == DFCSA_TAP_2x2S2;
    ktot += sg.channels * (sg.tap_mode == DFCSA_TAP_1x1 ? 1 : sg.tap_mode == DFCSA_TAP_3x3 ? 9 : 4);
  }
  DFCSA_CHECK_ARG(!(any2 && (any3 || p->n_seg != 1)), "conv_gemm_tc: a 2x2s2 segment must be the only segment");
  if (p->out_mode == DFCSA_OUT_CONVT2x2) {
    DFCSA_CHECK_ARG(!any3 && !any2 && p->N % 4 == 0 && (p->N / 4) % 32 == 0 && !p->accumulate,
                    "conv_gemm_tc: ConvT output needs 1x1 taps and Co %% 32 == 0");
  }

  ConvTcArgs a{};
  a.n_seg = p->n_seg;
  a.total_kb = ktot / 64;
  a.N = p->N;
  // ---- n tiling (needed first: the halo-box variant of the 3x3 path needs block_n <= 128) ----
  const int sms = num_sms();
  int block_n;
  if (p->N <= 256) block_n = (p->N + 31) / 32 * 32;
  else {
    int best_pad = 1 << 30; block_n = 256;
    for (int bn = 256; bn >= 128; bn -= 32) {
      int pad = (p->N + bn - 1) / bn * bn;
      if (pad < best_pad) { best_pad = pad; block_n = bn; }
    }
  }
  a.dw3 = (any3 && !any2 && block_n <= 128) ? 1 : 0;
  // ---- tile geometry ----
  if (any3) {
    a.H = p->H; a.W = p->W; a.tiles_b = p->B;
    if (a.dw3) { a.w_t = 8; a.h_t = 16; }
    else pick_spatial_tile(p->H, p->W, a.w_t, a.h_t);
  } else if (any2) {
    a.H = p->B * p->H; a.W = p->W; a.tiles_b = 1;
    pick_spatial_tile(a.H, a.W, a.w_t, a.h_t);
  } else {
    DFCSA_CHECK_ARG(Mtot < (1LL << 31), "conv_gemm_tc: too many pixels");
    a.H = 1; a.W = static_cast<int>(Mtot); a.tiles_b = 1; a.w_t = kBlockM; a.h_t = 1;
  }
  a.tiles_w = (a.W + a.w_t - 1) / a.w_t;
  a.tiles_h = (a.H + a.h_t - 1) / a.h_t;
  a.a_box_bytes = a.w_t * a.h_t * 128;
  const long long m_tiles = static_cast<long long>(a.tiles_w) * a.tiles_h * a.tiles_b;

  while (!a.dw3 && block_n > 64 && block_n % 64 == 0 && m_tiles * ((p->N + block_n - 1) / block_n) < sms) block_n /= 2;
  // BatchNorm statistics on a short-K 1x1 GEMM: the kernel is bound by the epilogue's instruction issue (the shuffle
  // transpose-reduce of the general path costs ~4x the store path), not by the tensor pipe.  Use 64-wide n tiles with
  // a grid that is a multiple of the n-tile count, so every CTA keeps ONE n tile and the register-resident statistics
  // variant applies; the n tiles of one pixel tile run side by side on neighbouring CTAs and share its A box through L2.
  static const int fix_nt_max_kb = [] { const char* e = getenv("DFCSA_CONV_FIXNT_MAXKB"); return e ? atoi(e) : 2; }();
  if (p->stats != nullptr && !any3 && !any2 && p->out_mode == DFCSA_OUT_DIRECT && p->N % 64 == 0 && p->N > 64 && p->N / 64 <= sms &&
      a.total_kb <= fix_nt_max_kb) {
    block_n = 64;
    a.fixed_nt = 1;
  }
  a.block_n = block_n;
  a.n_tiles_n = (p->N + block_n - 1) / block_n;
  // CTA pairs for the wide, deep GEMMs (levels 3-5): with a 128 x 256 tile per CTA the shared-memory port carries 96 B/clk
  // of UMMA operand reads plus 94 B/clk of TMA fill - above its 128 B/clk, which is why ncu shows the tensor pipe at 56-60 %
  // on these shapes.  A pair shares the B tile (each CTA stages and reads half of it): 64 + 62 B/clk.
  static const bool pairs_ok = [] { const char* e = getenv("DFCSA_CONV_2CTA"); return !e || atoi(e) != 0; }();
  // Measured (profiles/gemm_shapes_r02_l_*.json): K >= 1024 gains 5-8 % (e.g. 28^2 N=512 K=9216: 947 -> 1019 TFLOP/s, 56^2 N=256
  // K=4608: 994 -> 1070), K <= 512 loses up to 25 % (the pair handshake per tile outweighs the saved operand traffic), hence
  // the K threshold (DFCSA_CONV_2CTA_MINKB overrides it for experiments).
  static const int pairs_min_kb = [] { const char* e = getenv("DFCSA_CONV_2CTA_MINKB"); return e ? atoi(e) : 16; }();
  a.two_cta = (pairs_ok && !a.dw3 && block_n == 256 && p->N % 256 == 0 && p->out_mode == DFCSA_OUT_DIRECT && m_tiles >= 2 &&
               a.total_kb >= pairs_min_kb && sms >= 2 && !(p->epi != nullptr && p->epi->mode != DFCSA_EPI_NONE)) ? 1 : 0;
  if (a.dw3) {
    a.kbs = 1;
    a.stage_bytes = 10 * 16 * 128 + 3 * block_n * 128;
  } else {
    a.kbs = (block_n <= 128 && a.total_kb >= 2) ? 2 : 1;
    a.stage_bytes = a.kbs * (kABytes + (a.two_cta ? block_n / 2 : block_n) * 128);
  }
  const int stage_bytes = a.stage_bytes;
  a.stages = std::min(kMaxStages, (kSmemBudget - 1024) / stage_bytes);
  a.idesc = umma_idesc_f16(a.two_cta ? 2 * kBlockM : kBlockM, block_n, umma_fmt(p->src_dtype), umma_fmt(p->w_dtype), 0, 0);

  // ---- tensor maps ----
  CUtensorMap maps[3];
  for (int s = 0; s < 3; ++s) {
    const dfcsa_seg_t& sg = p->seg[s < p->n_seg ? s : 0];
    a.seg_kb[s] = sg.channels / 64;
    a.seg_mode[s] = sg.tap_mode;
    const uint64_t ldb = static_cast<uint64_t>(sg.ld) * 2;
    uint64_t dims[5], strides[4];
    uint32_t box[5];
    if (sg.tap_mode == DFCSA_TAP_2x2S2) {
      // source grid (B, 2H, 2W): (c, dj, j, di, b*H+i)
      dims[0] = sg.channels; dims[1] = 2; dims[2] = p->W; dims[3] = 2; dims[4] = static_cast<uint64_t>(p->B) * p->H;
      strides[0] = ldb; strides[1] = 2 * ldb; strides[2] = 2ull * p->W * ldb; strides[3] = 4ull * p->W * ldb;
      box[0] = 64; box[1] = 1; box[2] = a.w_t; box[3] = 1; box[4] = a.h_t;
    } else if (any3) {
      dims[0] = sg.channels; dims[1] = p->W; dims[2] = p->H; dims[3] = p->B; dims[4] = 1;
      strides[0] = ldb; strides[1] = static_cast<uint64_t>(p->W) * ldb; strides[2] = static_cast<uint64_t>(p->H) * p->W * ldb;
      strides[3] = static_cast<uint64_t>(p->B) * p->H * p->W * ldb;
      box[0] = 64; box[1] = (a.dw3 && sg.tap_mode == DFCSA_TAP_3x3) ? a.w_t + 2 : a.w_t; box[2] = a.h_t; box[3] = 1; box[4] = 1;
    } else {
      dims[0] = sg.channels; dims[1] = static_cast<uint64_t>(Mtot); dims[2] = 1; dims[3] = 1; dims[4] = 1;
      strides[0] = ldb; strides[1] = static_cast<uint64_t>(Mtot) * ldb; strides[2] = strides[1]; strides[3] = strides[1];
      box[0] = 64; box[1] = kBlockM; box[2] = 1; box[3] = 1; box[4] = 1;
    }
    int rc = encode_tensor_map(&maps[s], p->src_dtype, 5, sg.ptr, dims, strides, box, true);
    if (rc) return rc;
  }
  CUtensorMap map_b;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(ktot), static_cast<uint64_t>(p->N)};
    uint64_t strides[1] = {static_cast<uint64_t>(ktot) * 2};
    uint32_t box[2] = {64, static_cast<uint32_t>(a.two_cta ? block_n / 2 : block_n)};
    int rc = encode_tensor_map(&map_b, p->w_dtype, 2, p->w, dims, strides, box, true);
    if (rc) return rc;
  }

  a.out = p->out; a.ld_out = p->ld_out; a.out_dtype = p->out_dtype; a.out_mode = p->out_mode;
  a.accumulate = p->accumulate; a.bias = p->bias; a.stats = p->stats;
  a.shadow = reinterpret_cast<__nv_bfloat16*>(p->shadow); a.ld_shadow = p->ld_shadow;
  DFCSA_CHECK_ARG(p->shadow == nullptr || (!p->accumulate && p->ld_shadow % 8 == 0 && (reinterpret_cast<uintptr_t>(p->shadow) & 15) == 0),
                  "conv_gemm_tc: bad shadow output");
  if (p->out_mode == DFCSA_OUT_CONVT2x2) { a.convt_co = p->N / 4; a.convt_h = p->H; a.convt_w = p->W; }
  DFCSA_CHECK_ARG(p->act == 0 || (p->act == 1 && p->out_mode == DFCSA_OUT_DIRECT && !p->accumulate && p->stats == nullptr),
                  "conv_gemm_tc: the ReLU epilogue needs a direct, non-accumulating output without statistics");
  a.act_cols = p->act ? (p->act_cols > 0 ? p->act_cols : p->N) : 0;
  a.stats_cols = p->stats != nullptr ? ((p->stats_cols > 0 && p->stats_cols % 32 == 0) ? p->stats_cols : p->N) : 0;
  if (p->epi != nullptr && p->epi->mode != DFCSA_EPI_NONE) {
    const dfcsa_conv_epi_t& e = *p->epi;
    DFCSA_CHECK_ARG((e.mode == DFCSA_EPI_GATE_MIX && e.p && e.q) || (e.mode == DFCSA_EPI_RESIDUAL && e.p && e.scale),
                    "conv_gemm_tc: bad fused epilogue (gate mix needs p and q, residual needs p and scale)");
    DFCSA_CHECK_ARG(p->out_mode == DFCSA_OUT_DIRECT && !p->accumulate && p->stats == nullptr && p->shadow == nullptr &&
                    p->out_dtype == DFCSA_F16 && p->N % 32 == 0 && e.ld % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(e.p) & 15) == 0 && (reinterpret_cast<uintptr_t>(e.q) & 15) == 0,
                    "conv_gemm_tc: the fused epilogue needs a direct fp16 output with N % 32 == 0, no statistics, 16-byte aligned operand rows");
    a.epi_mode = e.mode;
    a.epi_p = reinterpret_cast<const __half*>(e.p); a.epi_q = reinterpret_cast<const __half*>(e.q);
    a.ld_epi = e.ld; a.epi_scale = e.scale;
  }
  if (p->bn != nullptr) {
    const dfcsa_bn_fold_t& b = *p->bn;
    DFCSA_CHECK_ARG(p->stats != nullptr && b.gamma && b.beta && b.scale && b.shift && b.mean && b.invstd && b.ticket && b.count > 0 &&
                    b.channels > 0 && b.channels <= a.stats_cols, "conv_gemm_tc: bad BatchNorm fold (needs stats, outputs, a ticket)");
    a.has_bn = 1;
    a.bn = b;
  }
  a.wide_out = p->out_dtype != DFCSA_F32 && !p->accumulate && p->ld_out % 16 == 0 && (reinterpret_cast<uintptr_t>(p->out) & 31) == 0;
  a.wide_shadow = p->shadow != nullptr && p->ld_shadow % 16 == 0 && (reinterpret_cast<uintptr_t>(p->shadow) & 31) == 0;

  const int smem_bytes = a.stages * stage_bytes + 1024;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(g_attr_once, [] {
    attr_err = cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(conv_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(conv_tc_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(conv_tc_kernel)");
  const long long total_tiles = m_tiles * a.n_tiles_n;
  int grid = static_cast<int>(std::min<long long>(total_tiles, sms));
  if (a.fixed_nt) grid = static_cast<int>(std::min<long long>(total_tiles, static_cast<long long>(sms / a.n_tiles_n) * a.n_tiles_n));
  if (a.two_cta) {
    // clusters of two CTAs (the SMs of one TPC); one pair per pair tile, at most sms / 2 pairs
    const long long pair_tiles = (m_tiles + 1) / 2 * a.n_tiles_n;
    const int pairs = static_cast<int>(std::min<long long>(pair_tiles, sms / 2));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<false, true>, maps[0], maps[1], maps[2], map_b, a);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(conv_tc_kernel<pairs>)");
  } else if (a.epi_mode != DFCSA_EPI_NONE)
    conv_tc_kernel<false, false, true><<<grid, 384, smem_bytes, stream>>>(maps[0], maps[1], maps[2], map_b, a);
  else if (p->stats != nullptr && (a.n_tiles_n == 1 || a.fixed_nt) && a.block_n <= 64)
    conv_tc_kernel<true><<<grid, 384, smem_bytes, stream>>>(maps[0], maps[1], maps[2], map_b, a);
  else
    conv_tc_kernel<false><<<grid, 384, smem_bytes, stream>>>(maps[0], maps[1], maps[2], map_b, a);
  DFCSA_LAUNCH_CHECK("conv_tc_kernel");
  return DFCSA_OK;
}

}  // namespace dfcsa
