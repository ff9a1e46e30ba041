// Convolution weight gradient on tcgen05 (sm_100a).
//
//   dw[n, t*C + c] += alpha * sum_m dy[pix_dy(m, t), n] * x[pix_x(m, t), c]
//
// The reduction runs over pixels, which are the *rows* of the NHWC tensors, so both operands are MN-major:
// a TMA box of (64 channels x <=64 pixels) lands as <=64 swizzled 128-byte rows (row = pixel = UMMA K index,
// 64 contiguous channels = UMMA M/N index).  The 3x3 taps shift the x box by (dh, dw); TMA zero-fills the halo.
// 3x3: a CTA owns one kernel ROW (dh) and computes its three taps (dw = -1, 0, +1) from ONE x box loaded with a
// one-pixel halo in W: the tap shift is a 128-byte shift of the UMMA descriptor start address inside that box (the
// hardware applies the 128-byte swizzle to absolute shared-memory addresses, so any row offset is legal - measured,
// tools/tc_probe.cu probe 10).  dy is loaded once for the three taps: ~3x less L2->SMEM traffic than a box per tap.
// One CTA per (tap row, n tile, c tile, pixel split); fp32 partials leave through atomics.  The n tile is 256 wide (two
// M=128 accumulators that share every x box: 128 FLOP per byte staged from L2 instead of 85) when N >= 256, else 128.
//   warps 0-3: epilogue (TMEM -> atomicAdd),  warp 4: TMA producer,  warp 5: TMEM alloc + MMA issuer
// x may be fp16 (a forward activation) while dy is bf16 (a gradient): kind::f16 cannot mix the two formats, so warps
// 0-3 - idle until the epilogue - rewrite every landed x box to bf16 in place in shared memory before the MMA warp
// reads it.  That replaces the bf16 "shadow" copy of every activation the forward pass used to write to HBM.
// The kernel is bound by the 128 B/clk shared-memory port (profiles/ncu_wgrad_tc_r01_final.csv): operand re-reads of every
// UMMA + that rewrite + the TMA writes.  Measured dead end (round 2, profiles/bench_r02_a_xreg.json): moving x through
// registers instead (four warps: 16-byte global loads two pixel blocks ahead -> convert -> one swizzled store) was correct
// but 4x SLOWER (32.3 ms against 7.7 ms per step over the 56 launches) - 128 threads cannot keep enough bytes in flight to
// replace a TMA box; the variant was removed.  A timing-only run with the rewrite switched off (wrong numbers on purpose,
// profiles/bench_r02_k_nocvt.json) puts its whole cost at 0.35 ms of the 7.8 ms per step: the rewrite overlaps the MMAs, so
// fp16 gradients with a loss scale (x and dy in one format) would buy ~1 %, not the third its wavefront share suggests.
#include "common.cuh"
#include <algorithm>
#include <mutex>

namespace dfcsa {
namespace {

constexpr int kMaxStages = 6;
constexpr int kBoxBytes = 64 * 128;          // one (64 ch x 64 px) box
constexpr int kBMaxBytes = 4 * kBoxBytes;    // up to 256 c

struct WgradTcArgs {
  int taps, n_tiles, c_tiles, splits;
  int block_n;             // 128 or 256 (two accumulators)
  int stages, a_bytes, stage_bytes;
  int block_c;             // multiple of 64, <= 256
  int N, C;
  int tiles_w, tiles_h, tiles_b, w_t, h_t;   // pixel-block geometry
  long long pix_blocks, blocks_per_split;
  int x_mode, dy_mode;
  int dw3;                 // 3x3 mode: three dw taps per CTA from one halo box (patch 16 x 4 pixels, x box 18 x 4)
  int x_box_bytes;         // shared-memory pitch of one x box (also the LBO between its 64-channel blocks)
  int x_tx_bytes;          // bytes one x TMA box delivers
  int box_bytes;           // bytes one dy TMA box delivers (w_t*h_t*128)
  float* dw; long long ld_dw;
  const float* alpha;
  uint32_t idesc;
  uint32_t tmem_cols;
  int zero_fill;           // shared-memory stages must be zeroed first (partial pixel boxes)
  int cvt_x;               // x arrives as fp16 and is converted to bf16 in shared memory (dy is bf16)
  int vec_red;             // dw rows are 16-byte aligned: red.global.add.v4.f32 (4x fewer L2 atomic operations)
  // second gradient over the same x (rows N1 .. N1+N2 of the virtual dy = [dy | dy2]; N1 % 64 == 0)
  int N1;                  // rows of the first gradient (== N when there is no second one)
  int c_begin2;
  float* dw2; long long ld_dw2;
  const float* alpha2;
  int vec_red2;
};

__global__ void __launch_bounds__(192, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                const __grid_constant__ CUtensorMap map_dy2, const __grid_constant__ WgradTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t cvt_bar[kMaxStages];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  // work item
  // tap fastest: the CTAs resident together are the taps / tiles of the SAME pixel range, so the shifted re-reads of
  // x and the repeated reads of dy hit L2 instead of HBM (x alone is 822 MB at level 1, far beyond the 126 MB L2)
  int id = blockIdx.x;
  const int tap = id % a.taps;     id /= a.taps;
  const int nt = id % a.n_tiles;   id /= a.n_tiles;
  const int ct = id % a.c_tiles;   id /= a.c_tiles;
  const int split = id;
  const int n0 = nt * a.block_n;
  const int c0 = ct * a.block_c;
  const int kStages = a.stages, kStageBytes = a.stage_bytes, kABytes = a.a_bytes;
  const long long pb_beg = split * a.blocks_per_split;
  // a tile of the second gradient whose x channels all lie below c_begin2 has nothing to compute
  const bool dead = n0 >= a.N1 && c0 + a.block_c <= a.c_begin2;
  const long long pb_end = dead ? pb_beg : min(a.pix_blocks, pb_beg + a.blocks_per_split);
  const int n_boxes_a = min((a.N - n0 + 63) / 64, a.block_n / 64);
  const int n_halves = (n_boxes_a + 1) / 2;
  const int n_boxes_b = min(a.block_c, a.C - c0) / 64;

  if (a.zero_fill) {
    // only when a pixel box is partial (fewer than 64 pixels per box): the rows TMA never writes lie on the reduction
    // axis of BOTH operands and must read as zero.  (An unwritten 64-channel block of dy only feeds output rows >= N,
    // which are never stored.)
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4* p = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < a.stages * a.stage_bytes / 16; i += blockDim.x) p[i] = z;
    fence_proxy_async();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_x);
    if (a.N1 < a.N) tma_prefetch_desc(&map_dy2);
    for (int i = 0; i < a.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); mbar_init(&cvt_bar[i], 128); }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(&tmem_base_smem, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 4) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0;
    const uint32_t tx_bytes = static_cast<uint32_t>(n_boxes_a * a.box_bytes + n_boxes_b * a.x_tx_bytes);
    for (long long pb = pb_beg; pb < pb_end; ++pb) {
      const int tw = static_cast<int>(pb % a.tiles_w);
      const long long r = pb / a.tiles_w;
      const int th = static_cast<int>(r % a.tiles_h);
      const int tb = static_cast<int>(r / a.tiles_h);
      const int w0 = tw * a.w_t, h0 = th * a.h_t;
      if (elect_one()) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
        uint8_t* sa = smem + stage * kStageBytes;
        uint8_t* sb = sa + kABytes;
        for (int j = 0; j < n_boxes_a; ++j) {
          if (a.dy_mode == DFCSA_TAP_2x2S2)
            tma_load_5d(sa + j * kBoxBytes, &map_dy, &full_bar[stage], n0 + j * 64, tap & 1, w0, tap >> 1, h0);
          else if (n0 + j * 64 >= a.N1)
            tma_load_5d(sa + j * kBoxBytes, &map_dy2, &full_bar[stage], n0 + j * 64 - a.N1, w0, h0, tb, 0);
          else
            tma_load_5d(sa + j * kBoxBytes, &map_dy, &full_bar[stage], n0 + j * 64, w0, h0, tb, 0);
        }
        for (int j = 0; j < n_boxes_b; ++j) {
          if (a.dw3)   // tap == dh: rows h0+dh-1 .., columns w0-1 .. w0+16 (halo of one pixel on each side)
            tma_load_5d(sb + j * a.x_box_bytes, &map_x, &full_bar[stage], c0 + j * 64, w0 - 1, h0 + tap - 1, tb, 0);
          else
            tma_load_5d(sb + j * kBoxBytes, &map_x, &full_bar[stage], c0 + j * 64, w0, h0, tb, 0);
        }
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    int stage = 0; uint32_t phase = 0;
    bool first = true;
    for (long long pb = pb_beg; pb < pb_end; ++pb) {
      mbar_wait(a.cvt_x ? &cvt_bar[stage] : &full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
        const uint32_t b_addr = a_addr + kABytes;
        // one descriptor per operand and stage; every k step / tap is a constant further on (umma_desc_add)
        const uint32_t acc0 = first ? 0u : 1u;
        if (a.dw3) {
          const uint64_t da0 = umma_smem_desc(a_addr, kBoxBytes, 1024);
          const uint64_t db0 = umma_smem_desc(b_addr, a.x_box_bytes, 1024);
          // k outer, tap inner: consecutive instructions accumulate into different TMEM tiles and share the dy operand
          // (measured 1 % faster than tap outer once the issue path was short, profiles/bench_r02_u_*.json)
#pragma unroll
          for (int k = 0; k < 4; ++k) {     // patch row k: 16 dy pixels against x pixels shifted by dw inside the 18-wide row
#pragma unroll
            for (int dw = 0; dw < 3; ++dw)
              umma_f16(tmem_base + dw * 128, umma_desc_add(da0, k * 2048), umma_desc_add(db0, (k * 18 + dw) * 128), a.idesc,
                       k == 0 ? acc0 : 1u);
          }
        } else {
          uint64_t da0 = umma_smem_desc(a_addr, kBoxBytes, 1024);
          const uint64_t db0 = umma_smem_desc(b_addr, kBoxBytes, 1024);
          for (int h = 0; h < n_halves; ++h) {
#pragma unroll
            for (int k = 0; k < 4; ++k)     // 16 pixels (rows) per instruction = 2 KiB
              umma_f16(tmem_base + h * 256, umma_desc_add(da0, k * 2048), umma_desc_add(db0, k * 2048), a.idesc, k == 0 ? acc0 : 1u);
            da0 = umma_desc_add(da0, 2 * kBoxBytes);
          }
        }
        umma_commit(&empty_bar[stage]);
      }
      first = false;
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(&done_bar);
    __syncwarp();
  } else {
    // ===================== x: fp16 -> bf16 in place (only when the operand formats differ) =====================
    if (a.cvt_x) {
      int stage = 0; uint32_t phase = 0;
      const int n16 = n_boxes_b * a.x_box_bytes / 16;
      for (long long pb = pb_beg; pb < pb_end; ++pb) {
        mbar_wait(&full_bar[stage], phase);
        uint4* xs = reinterpret_cast<uint4*>(smem + stage * kStageBytes + kABytes);
        for (int i = threadIdx.x; i < n16; i += 128) {
          uint4 v = xs[i];
          uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
            const __nv_bfloat162 bb = __floats2bfloat162_rn(f.x, f.y);
            w[j] = *reinterpret_cast<const uint32_t*>(&bb);
          }
          xs[i] = v;
        }
        fence_proxy_async();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
        mbar_arrive(&cvt_bar[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
    // ===================== epilogue =====================
    if (pb_end > pb_beg) {
      mbar_wait(&done_bar, 0);
      tc_fence_after();
      const float alpha1 = a.alpha ? __ldg(a.alpha) : 1.f;
      const float alpha2 = a.alpha2 ? __ldg(a.alpha2) : 1.f;
      const int ccols = min(a.block_c, a.C - c0);
      const int n_acc = a.dw3 ? 3 : n_halves;           // accumulators: the three dw taps, or the two 128-row halves
      for (int h = 0; h < n_acc; ++h) {
        const int n = n0 + (a.dw3 ? 0 : h * 128) + warp * 32 + lane;
        const int tap_out = a.dw3 ? tap * 3 + h : tap;
        const uint32_t acc_col = a.dw3 ? h * 128 : h * 256;
        for (int ch = 0; ch * 32 < ccols; ++ch) {
          uint32_t raw[32];
          tmem_ld_32x32(tmem_base + acc_col + ch * 32 + (static_cast<uint32_t>(warp * 32) << 16), raw);
          tmem_ld_wait();
          const bool seg2 = n >= a.N1;
          const int ccol = c0 + ch * 32;       // first x channel of this chunk
          if (n < a.N && !(seg2 && ccol < a.c_begin2)) {
            float* dst = seg2 ? a.dw2 + static_cast<long long>(n - a.N1) * a.ld_dw2 + (ccol - a.c_begin2)
                              : a.dw + static_cast<long long>(n) * a.ld_dw + static_cast<long long>(tap_out) * a.C + ccol;
            const float alpha = seg2 ? alpha2 : alpha1;
            if (seg2 ? a.vec_red2 : a.vec_red) {      // ccols is a multiple of 64, so a 32-column chunk is always complete
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                             :: "l"(dst + i), "f"(alpha * __uint_as_float(raw[i])), "f"(alpha * __uint_as_float(raw[i + 1])),
                                "f"(alpha * __uint_as_float(raw[i + 2])), "f"(alpha * __uint_as_float(raw[i + 3])) : "memory");
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (ch * 32 + i < ccols) atomicAdd(dst + i, alpha * __uint_as_float(raw[i]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

std::once_flag g_attr_once;

void pick_patch(int H, int W, int& w_t, int& h_t) {
  double best = -1.0;
  w_t = 1; h_t = 1;
  for (int wt = 1; wt <= std::min(W, 64); ++wt) {
    int ht = std::min(H, 64 / wt);
    if (ht < 1) continue;
    long long tiles = static_cast<long long>((W + wt - 1) / wt) * ((H + ht - 1) / ht);
    double eff = static_cast<double>(H) * W / (static_cast<double>(tiles) * 64);
    if (eff > best + 1e-9 || (eff > best - 1e-9 && wt > w_t)) { best = eff; w_t = wt; h_t = ht; }
  }
}

}  // namespace

// Pixel splits: the kernel needs a whole SM per CTA (shared memory + 512 TMEM columns), so a grid of 2*SMs + 1 CTAs
// costs three waves, not two (measured: 297 CTAs of the level-1 3x3 gradient kept every SM idle for a third of the
// launch).  Pick the split count whose grid fills 1..4 whole waves at the lowest cost
//   waves * (pixel blocks per CTA + fixed cost of a CTA: prologue, pipeline fill, atomic epilogue ~ 8 blocks).
// items = independent (tap row, n tile, c tile) work items; returns the split count, *blocks_per_split = pixel blocks
// of one CTA.  force_waves > 0 restricts the choice to that wave count (experiments).
int wgrad_pick_splits(long long items, long long pix_blocks, int sms, int force_waves, long long* blocks_per_split) {
  long long best_cost = -1, best_s = 1, best_bps = pix_blocks;
  for (int k = 1; k <= 4; ++k) {
    if (force_waves > 0 && k != force_waves) continue;
    long long s = std::max<long long>(1, static_cast<long long>(k) * sms / items);
    s = std::min(s, std::max<long long>(1, pix_blocks / 4));
    const long long bps = (pix_blocks + s - 1) / s;
    s = (pix_blocks + bps - 1) / bps;
    const long long waves = (items * s + sms - 1) / sms;
    const long long cost = waves * (bps + 8);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_s = s; best_bps = bps; }
  }
  *blocks_per_split = best_bps;
  return static_cast<int>(best_s);
}

int conv_wgrad_tc(const dfcsa_wgrad_params_t* p, cudaStream_t stream) {
  DFCSA_CHECK_ARG(p->x_dtype != DFCSA_F32 && p->dy_dtype != DFCSA_F32, "conv_wgrad_tc: 16-bit operands required");
  const bool cvt_x = p->x_dtype == DFCSA_F16 && p->dy_dtype == DFCSA_BF16;
  DFCSA_CHECK_ARG(p->x_dtype == p->dy_dtype || cvt_x,
                  "conv_wgrad_tc: x and dy must share one 16-bit format, or x fp16 with dy bf16 (converted in shared memory)");
  DFCSA_CHECK_ARG(p->C % 64 == 0 && p->N % 8 == 0, "conv_wgrad_tc: C must be a multiple of 64 and N of 8 (C=%d N=%d)", p->C, p->N);
  DFCSA_CHECK_ARG(p->ld_x % 8 == 0 && p->ld_dy % 8 == 0, "conv_wgrad_tc: pitches must be multiples of 8");
  DFCSA_CHECK_ARG((reinterpret_cast<uintptr_t>(p->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->dy) & 15) == 0,
                  "conv_wgrad_tc: operands must be 16-byte aligned");
  const long long Mtot = static_cast<long long>(p->B) * p->H * p->W;
  DFCSA_CHECK_ARG(Mtot > 0 && Mtot < (1LL << 31), "conv_wgrad_tc: bad pixel count");

  WgradTcArgs a{};
  a.N = p->N; a.C = p->C;
  a.N1 = p->N;
  if (p->dy2 != nullptr) {
    DFCSA_CHECK_ARG(p->N % 64 == 0 && p->N2 % 8 == 0 && p->c_begin2 % 32 == 0 && p->dy_dtype != DFCSA_F32 && p->ld_dy2 % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(p->dy2) & 15) == 0,
                    "conv_wgrad_tc: second gradient needs N %% 64 == 0, N2 %% 8 == 0, c_begin2 %% 32 == 0 and 16-byte aligned dy2");
    a.N = p->N + p->N2;        // rows of the virtual concatenation [dy | dy2]
    a.c_begin2 = p->c_begin2; a.dw2 = p->dw2; a.ld_dw2 = p->ld_dw2; a.alpha2 = p->alpha2;
    a.vec_red2 = ((reinterpret_cast<uintptr_t>(p->dw2) & 15) == 0 && p->ld_dw2 % 4 == 0) ? 1 : 0;
  }
  a.x_mode = p->x_tap_mode; a.dy_mode = p->dy_tap_mode;
  a.taps = p->x_tap_mode == DFCSA_TAP_3x3 ? 9 : (p->dy_tap_mode == DFCSA_TAP_2x2S2 ? 4 : 1);
  a.block_n = a.N >= 256 ? 256 : 128;
  a.n_tiles = (a.N + a.block_n - 1) / a.block_n;
  a.a_bytes = (a.block_n / 64) * kBoxBytes;
  a.stage_bytes = a.a_bytes + kBMaxBytes;
  a.stages = a.block_n == 256 ? 3 : 4;
  if (p->C <= 256) a.block_c = p->C;
  else {
    int best_pad = 1 << 30; a.block_c = 256;
    for (int bc = 256; bc >= 128; bc -= 64) {
      int pad = (p->C + bc - 1) / bc * bc;
      if (pad < best_pad) { best_pad = pad; a.block_c = bc; }
    }
  }
  a.c_tiles = (p->C + a.block_c - 1) / a.block_c;

  // ---- pixel-block geometry and tensor maps ----
  CUtensorMap map_dy, map_x, map_dy2;
  const uint64_t ldx = static_cast<uint64_t>(p->ld_x) * 2, ldy = static_cast<uint64_t>(p->ld_dy) * 2;
  uint64_t dims[5], strides[4];
  uint32_t box[5];
  a.x_box_bytes = kBoxBytes;
  if (p->x_tap_mode == DFCSA_TAP_3x3) {
    // one kernel row per CTA; pixel patch 16 wide x 4 high so that every UMMA K step (16 pixels) is one patch row and
    // the dw shift never crosses a row boundary; the x box carries one halo pixel on each side: 18 x 4 pixels
    a.dw3 = 1; a.taps = 3;
    a.w_t = 16; a.h_t = 4;
    a.block_n = 128; a.n_tiles = (p->N + 127) / 128;
    a.block_c = std::min(p->C, 128); a.c_tiles = (p->C + a.block_c - 1) / a.block_c;
    a.x_box_bytes = 18 * 4 * 128;                       // 9216
    a.a_bytes = 2 * kBoxBytes;
    a.stage_bytes = a.a_bytes + 2 * a.x_box_bytes;      // 34816
    a.stages = 5;
    a.tiles_w = (p->W + a.w_t - 1) / a.w_t; a.tiles_h = (p->H + a.h_t - 1) / a.h_t; a.tiles_b = p->B;
    box[0] = 64; box[1] = 18; box[2] = 4; box[3] = 1; box[4] = 1;
    dims[0] = p->C; dims[1] = p->W; dims[2] = p->H; dims[3] = p->B; dims[4] = 1;
    strides[0] = ldx; strides[1] = p->W * ldx; strides[2] = static_cast<uint64_t>(p->H) * p->W * ldx; strides[3] = static_cast<uint64_t>(Mtot) * ldx;
    int rc = encode_tensor_map(&map_x, p->x_dtype, 5, p->x, dims, strides, box, true);
    if (rc) return rc;
    box[1] = 16;
    dims[0] = p->N;
    strides[0] = ldy; strides[1] = p->W * ldy; strides[2] = static_cast<uint64_t>(p->H) * p->W * ldy; strides[3] = static_cast<uint64_t>(Mtot) * ldy;
    rc = encode_tensor_map(&map_dy, p->dy_dtype, 5, p->dy, dims, strides, box, true);
    if (rc) return rc;
  } else if (p->dy_tap_mode == DFCSA_TAP_2x2S2) {
    const int Hm = p->B * p->H;
    pick_patch(Hm, p->W, a.w_t, a.h_t);
    a.tiles_w = (p->W + a.w_t - 1) / a.w_t; a.tiles_h = (Hm + a.h_t - 1) / a.h_t; a.tiles_b = 1;
    box[0] = 64; box[1] = a.w_t; box[2] = a.h_t; box[3] = 1; box[4] = 1;
    dims[0] = p->C; dims[1] = p->W; dims[2] = Hm; dims[3] = 1; dims[4] = 1;
    strides[0] = ldx; strides[1] = p->W * ldx; strides[2] = static_cast<uint64_t>(Mtot) * ldx; strides[3] = strides[2];
    int rc = encode_tensor_map(&map_x, p->x_dtype, 5, p->x, dims, strides, box, true);
    if (rc) return rc;
    // dy on the (B, 2H, 2W) grid: (n, dj, j, di, b*H+i)
    dims[0] = p->N; dims[1] = 2; dims[2] = p->W; dims[3] = 2; dims[4] = Hm;
    strides[0] = ldy; strides[1] = 2 * ldy; strides[2] = 2ull * p->W * ldy; strides[3] = 4ull * p->W * ldy;
    box[0] = 64; box[1] = 1; box[2] = a.w_t; box[3] = 1; box[4] = a.h_t;
    rc = encode_tensor_map(&map_dy, p->dy_dtype, 5, p->dy, dims, strides, box, true);
    if (rc) return rc;
  } else {
    a.w_t = 64; a.h_t = 1;
    a.tiles_w = static_cast<int>((Mtot + 63) / 64); a.tiles_h = 1; a.tiles_b = 1;
    box[0] = 64; box[1] = 64; box[2] = 1; box[3] = 1; box[4] = 1;
    dims[0] = p->C; dims[1] = Mtot; dims[2] = 1; dims[3] = 1; dims[4] = 1;
    strides[0] = ldx; strides[1] = static_cast<uint64_t>(Mtot) * ldx; strides[2] = strides[1]; strides[3] = strides[1];
    int rc = encode_tensor_map(&map_x, p->x_dtype, 5, p->x, dims, strides, box, true);
    if (rc) return rc;
    dims[0] = p->N;
    strides[0] = ldy; strides[1] = static_cast<uint64_t>(Mtot) * ldy; strides[2] = strides[1]; strides[3] = strides[1];
    rc = encode_tensor_map(&map_dy, p->dy_dtype, 5, p->dy, dims, strides, box, true);
    if (rc) return rc;
    if (p->dy2 != nullptr) {
      const uint64_t ldy2 = static_cast<uint64_t>(p->ld_dy2) * 2;
      dims[0] = p->N2;
      strides[0] = ldy2; strides[1] = static_cast<uint64_t>(Mtot) * ldy2; strides[2] = strides[1]; strides[3] = strides[1];
      rc = encode_tensor_map(&map_dy2, p->dy_dtype, 5, p->dy2, dims, strides, box, true);
      if (rc) return rc;
    }
  }
  if (p->dy2 == nullptr) map_dy2 = map_dy;
  a.box_bytes = a.w_t * a.h_t * 128;
  a.x_tx_bytes = a.dw3 ? a.x_box_bytes : a.box_bytes;
  a.zero_fill = (a.box_bytes < kBoxBytes) ? 1 : 0;
  a.pix_blocks = static_cast<long long>(a.tiles_w) * a.tiles_h * a.tiles_b;

  const long long items = static_cast<long long>(a.taps) * a.n_tiles * a.c_tiles;
  static const int force_waves = [] { const char* e = getenv("DFCSA_WGRAD_WAVES"); return e ? atoi(e) : 0; }();
  long long bps = 0;
  a.splits = wgrad_pick_splits(items, a.pix_blocks, num_sms(), force_waves, &bps);
  a.blocks_per_split = bps;
  a.dw = p->dw; a.ld_dw = p->ld_dw; a.alpha = p->alpha;
  a.vec_red = ((reinterpret_cast<uintptr_t>(p->dw) & 15) == 0 && p->ld_dw % 4 == 0) ? 1 : 0;
  a.cvt_x = cvt_x ? 1 : 0;
  a.idesc = umma_idesc_f16(128, a.block_c, umma_fmt(p->dy_dtype), umma_fmt(cvt_x ? DFCSA_BF16 : p->x_dtype), 1, 1);
  a.tmem_cols = (a.block_n == 256 || a.dw3) ? 512 : (a.block_c <= 64 ? 64 : a.block_c <= 128 ? 128 : 256);

  const int smem_bytes = a.stages * a.stage_bytes + 1024;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(g_attr_once, [] {
    attr_err = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * (4 * kBoxBytes + kBMaxBytes) + 1024);  // == 4 stages of the 128-wide tile
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(wgrad_tc_kernel)");
  const long long grid = items * a.splits;
  DFCSA_CHECK_ARG(grid < (1LL << 31), "conv_wgrad_tc: grid too large");
  wgrad_tc_kernel<<<static_cast<unsigned>(grid), 192, smem_bytes, stream>>>(map_dy, map_x, map_dy2, a);
  DFCSA_LAUNCH_CHECK("wgrad_tc_kernel");
  return DFCSA_OK;
}

}  // namespace dfcsa
