// Bandwidth-bound kernels of the DFC-SA block (forward and backward): BatchNorm finalize / apply, ReLU, the
// adaptive average pool and bilinear up-sample (and their transposes), gating, residual + max-pool, layout moves.
// Activations fp16, gradients bf16, all arithmetic fp32, per-channel reductions flushed in double.
// Every kernel walks NHWC so that consecutive threads touch consecutive 16-byte channel vectors of a pixel.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>

namespace dfcsa {
namespace {

typedef __half act_t;
typedef __nv_bfloat16 grad_t;

template <int VEC, typename T> __device__ __forceinline__ void ldv(const T* p, float (&v)[VEC]) {
  if constexpr (VEC == 8) load8<T>(p, v);
  else v[0] = Cvt<T>::to_f(p[0]);
}
template <int VEC, typename T> __device__ __forceinline__ void stv(T* p, const float (&v)[VEC]) {
  if constexpr (VEC == 8) store8<T>(p, v);
  else p[0] = Cvt<T>::from_f(v[0]);
}
// raw 16-bit channel vectors: several loads in flight cost 4 registers each until they are converted at their use
template <int VEC, typename T> struct RawV { uint4 r; };
template <typename T> struct RawV<1, T> { T r; };
template <int VEC, typename T> __device__ __forceinline__ RawV<VEC, T> ldraw(const T* p) {
  RawV<VEC, T> w;
  if constexpr (VEC == 8) w.r = *reinterpret_cast<const uint4*>(p);
  else w.r = p[0];
  return w;
}
template <int VEC, typename T> __device__ __forceinline__ void cvtraw(const RawV<VEC, T>& w, float (&v)[VEC]) {
  if constexpr (VEC == 8) {
    const T* e = reinterpret_cast<const T*>(&w.r);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = Cvt<T>::to_f(e[i]);
  } else v[0] = Cvt<T>::to_f(w.r);
}
template <int VEC> __device__ __forceinline__ void ldf(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 8) load8<float>(p, v);
  else if constexpr (VEC == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else v[0] = p[0];
}
template <int VEC> __device__ __forceinline__ void stf(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) p[i] = v[i];
  }
}

static inline bool vec8_ok(int C, std::initializer_list<long long> lds, std::initializer_list<const void*> ptrs) {
  if (C % 8) return false;
  for (long long l : lds) if (l % 8) return false;
  for (const void* p : ptrs) if (p && (reinterpret_cast<uintptr_t>(p) & 15)) return false;
  return true;
}

constexpr int kGatePix = 2;   // pixels per grid-stride step in the gate kernels

// sigmoid with the fast reciprocal (the IEEE division costs ~10 issue slots per element; the gate kernels are bound by
// instruction issue + memory latency together, profiles/ncu_block1_sections_r02.csv)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// adaptive_avg_pool2d window of output index i: [lo, hi)
__device__ __forceinline__ void pool_win(int i, int s, int P, int& lo, int& hi) {
  lo = (i * s) / P;
  hi = ((i + 1) * s + P - 1) / P;
}
// windows containing source index y: [ilo, ihi]
__device__ __forceinline__ void pool_win_of(int y, int s, int P, int& ilo, int& ihi) {
  ilo = (y * P) / s;
  ihi = ((y + 1) * P - 1) / s;
  if (ihi > P - 1) ihi = P - 1;
}
// bilinear, align_corners=False: source taps of destination index d (ATen area_pixel_compute_source_index)
__device__ __forceinline__ void bilerp_taps(int d, int P, int s, int& i0, int& i1, float& l1) {
  const float scale = static_cast<float>(P) / static_cast<float>(s);
  float src = scale * (static_cast<float>(d) + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = min(static_cast<int>(src), P - 1);
  i1 = min(i0 + 1, P - 1);
  l1 = src - static_cast<float>(i0);
}

// flat item -> (cv, j, y, b) with extents (CV, P, H, *): 32-bit divisions whenever the item count allows (the 64-bit
// ones cost more than the few loads of an item when the pooled grid is fine)
__device__ __forceinline__ void decode4(long long i, bool small, int CV, int P, int H, int& cv, int& j, int& y, long long& b) {
  if (small) {
    unsigned u = static_cast<unsigned>(i);
    unsigned q = u / static_cast<unsigned>(CV); cv = static_cast<int>(u - q * CV); u = q;
    q = u / static_cast<unsigned>(P); j = static_cast<int>(u - q * P); u = q;
    q = u / static_cast<unsigned>(H); y = static_cast<int>(u - q * H);
    b = q;
  } else {
    cv = static_cast<int>(i % CV);
    long long r = i / CV;
    j = static_cast<int>(r % P); r /= P;
    y = static_cast<int>(r % H);
    b = r / H;
  }
}

// (b, y, x) of pixel m for a grid-stride walk, advanced without divisions
struct PixIter {
  unsigned x, y, b, sx, sy, sb;
  __device__ __forceinline__ void init(unsigned m0, unsigned stride, unsigned H, unsigned W) {
    x = m0 % W; unsigned r = m0 / W; y = r % H; b = r / H;
    sx = stride % W; r = stride / W; sy = r % H; sb = r / H;
  }
  __device__ __forceinline__ void next(unsigned H, unsigned W) {
    x += sx; if (x >= W) { x -= W; ++y; }
    y += sy; if (y >= H) { y -= H; ++b; }
    b += sb;
  }
};

// adaptive_avg_pool^T lookup tables in shared memory: for every source row / column the range of windows that contain
// it, and 1/|window| per window (integer divisions done once per block instead of ~10 per 16-byte vector)
struct PoolTabs { const int* yt; const int* xt; const float* wy; const float* wx; };
__device__ __forceinline__ PoolTabs build_pool_tabs(unsigned char* smem, int H, int W, int P) {
  int* yt = reinterpret_cast<int*>(smem);
  int* xt = yt + H;
  float* wy = reinterpret_cast<float*>(xt + W);
  float* wx = wy + P;
  for (int t = threadIdx.x; t < H; t += blockDim.x) { int lo, hi; pool_win_of(t, H, P, lo, hi); yt[t] = lo | (hi << 16); }
  for (int t = threadIdx.x; t < W; t += blockDim.x) { int lo, hi; pool_win_of(t, W, P, lo, hi); xt[t] = lo | (hi << 16); }
  for (int t = threadIdx.x; t < P; t += blockDim.x) {
    int a, e; pool_win(t, H, P, a, e); wy[t] = 1.f / static_cast<float>(e - a);
    pool_win(t, W, P, a, e); wx[t] = 1.f / static_cast<float>(e - a);
  }
  __syncthreads();
  PoolTabs t; t.yt = yt; t.xt = xt; t.wy = wy; t.wx = wx;
  return t;
}
static inline size_t pool_tabs_bytes(int H, int W, int P) { return static_cast<size_t>(H + W + 2 * P) * 4; }

// ---- per-channel block reduction: thread (cl, pl) holds NRED x VEC partials for channel vector cl ----
template <int VEC, int NRED>
__device__ __forceinline__ void flush_channel_partials(float (&acc)[NRED][VEC], int cl, int pl, int CL, int PL,
                                                       int c_base, int C, double* const (&outs)[NRED], float* s_red) {
  // s_red: [NRED][PL][CL*VEC]
  const int row = CL * VEC;
  if (cl < CL && pl < PL) {
#pragma unroll
    for (int r = 0; r < NRED; ++r)
#pragma unroll
      for (int v = 0; v < VEC; ++v) s_red[(r * PL + pl) * row + cl * VEC + v] = acc[r][v];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < row; idx += blockDim.x) {
    const int c = c_base + idx;
    if (c < C) {
#pragma unroll
      for (int r = 0; r < NRED; ++r) {
        float s = 0.f;
        for (int p = 0; p < PL; ++p) s += s_red[(r * PL + p) * row + idx];
        atomicAdd(outs[r] + c, static_cast<double>(s));
      }
    }
  }
}

__device__ __forceinline__ void block_scalar_reduce_add(float v, double* out) {
  __shared__ float s_w[32];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_w[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < (blockDim.x + 31) / 32 ? s_w[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) atomicAdd(out, static_cast<double>(t));
  }
}

// Resident CTAs per SM each heavy streaming kernel is compiled for (register cap 65536 / (256 * occ)).  Measured on
// B200 per kernel (profiles/README.md): the kernels with many live vectors (branch_bwd_reduce1 / reduce2 / apply) are
// fastest at 2 (128 registers, no spills), the gate kernels at 3, block_out_bwd_reduce at 4.  DFCSA_EW_OCC overrides
// all of them for experiments.
static int ew_occ(int dflt) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("DFCSA_EW_OCC");
    forced = e ? atoi(e) : 0;
    if (forced < 2 || forced > 4) forced = 0;
  }
  return forced ? forced : dflt;
}
#define OCC_DISPATCH(dflt, ...)                                                  \
  do { const int occ__ = ew_occ(dflt);                                           \
       if (occ__ == 2) { constexpr int OCC = 2; __VA_ARGS__; }                   \
       else if (occ__ == 4) { constexpr int OCC = 4; __VA_ARGS__; }              \
       else { constexpr int OCC = 3; __VA_ARGS__; } } while (0)

struct RedGeom { int CL, PL, chunks; };
static RedGeom red_geom(int C, int VEC) {
  RedGeom g;
  const int cv = (C + VEC - 1) / VEC;
  g.CL = std::min(cv, VEC == 8 ? 32 : 64);
  g.PL = 256 / g.CL;
  g.chunks = (cv + g.CL - 1) / g.CL;
  return g;
}
// blocks along x so that the whole grid (x * other_dims) is ONE full wave of `occ` resident blocks per SM: the kernels
// are grid-stride loops, and 2.67 waves of short blocks cost a ~12 % tail
static int red_blocks(long long items, int PL, int other_dims = 1, int occ = 4) {
  long long b = (items + PL - 1) / PL;
  const long long wave = std::max<long long>(1, static_cast<long long>(num_sms()) * occ / other_dims);
  return static_cast<int>(std::max<long long>(1, std::min<long long>(b, wave)));
}

// =============================================================================================
// BatchNorm finalize / eval affine / param grads
// =============================================================================================
__global__ void bn_finalize_kernel(const double* sum, const double* sumsq, long long count, int C, const float* gamma, const float* beta,
                                   const float* conv_bias, float* rmean, float* rvar, float momentum, float eps,
                                   float* scale, float* shift, float* mean_out, float* invstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double n = static_cast<double>(count);
  const double mean = sum[c] / n;
  double var = sumsq[c] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const double invstd = rsqrt(var + static_cast<double>(eps));
  const float sc = static_cast<float>(gamma[c] * invstd);
  scale[c] = sc;
  shift[c] = static_cast<float>(beta[c] - mean * gamma[c] * invstd);
  mean_out[c] = static_cast<float>(mean);
  invstd_out[c] = static_cast<float>(invstd);
  if (rmean != nullptr) {
    const double b = conv_bias ? conv_bias[c] : 0.0;
    rmean[c] = static_cast<float>((1.0 - momentum) * rmean[c] + momentum * (mean + b));
    const double unb = count > 1 ? var * n / (n - 1.0) : var;
    rvar[c] = static_cast<float>((1.0 - momentum) * rvar[c] + momentum * unb);
  }
}
__global__ void bn_eval_affine_kernel(int C, const float* gamma, const float* beta, const float* conv_bias,
                                      const float* rmean, const float* rvar, float eps, float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] * rsqrtf(rvar[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] + ((conv_bias ? conv_bias[c] : 0.f) - rmean[c]) * sc;
}
__global__ void bn_param_grads_kernel(const double* red, int C, float* dgamma, float* dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  dbeta[c] = static_cast<float>(red[c]);
  dgamma[c] = static_cast<float>(red[C + c]);
}

struct BlockGradPtrs { float* g[4]; float* b[4]; float* drs; float* dgam; };
__global__ void block_param_grads_kernel(const double* red, int C, BlockGradPtrs o) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (o.b[i] != nullptr) o.b[i][c] = static_cast<float>(red[2 * i * C + c]);
      if (o.g[i] != nullptr) o.g[i][c] = static_cast<float>(red[(2 * i + 1) * C + c]);
    }
  }
  if (c == 0) {
    if (o.drs != nullptr) *o.drs = static_cast<float>(red[8 * C]);
    if (o.dgam != nullptr) *o.dgam = static_cast<float>(red[8 * C + 1]);
  }
}

// =============================================================================================
// forward: bn+relu -> adaptive pool (separable)
// =============================================================================================
// MASKS: two more planes behind tmp ([3][B, H, P, C]): the row means of m = [bn(a0) > 0] and of m * a0.  Pooled like the
// activation itself they give, per pooling window w, mean_w(m) and mean_w(m a0), from which the backward gets the
// pool^T(dpooled) part of the BatchNorm-2 reductions without a gather pass over the feature map:
//   sum_pix poolT(dp)[pix] m[pix]      = sum_w dp[w] mean_w(m)
//   sum_pix poolT(dp)[pix] m[pix] a0   = sum_w dp[w] mean_w(m a0)          (dfcsa_pool_window_terms)
template <int VEC, bool MASKS>
__global__ void pool_rows_kernel(const act_t* a0, long long ld, int B, int H, int W, int C, const float* scale,
                                 const float* shift, int P, float* tmp) {
  const int CV = C / VEC;
  const long long total = static_cast<long long>(B) * H * P * CV;
  const long long plane = static_cast<long long>(B) * H * P * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int cv, j, y; long long b;
    decode4(i, total < (1LL << 31), CV, P, H, cv, j, y, b);
    int lo, hi; pool_win(j, W, P, lo, hi);
    float sc[VEC], sh[VEC], acc[VEC], am[MASKS ? VEC : 1], ax[MASKS ? VEC : 1];
    ldf<VEC>(scale + cv * VEC, sc); ldf<VEC>(shift + cv * VEC, sh);
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
#pragma unroll
    for (int v = 0; v < (MASKS ? VEC : 1); ++v) { am[v] = 0.f; ax[v] = 0.f; }
    const act_t* row = a0 + ((b * H + y) * W) * ld + cv * VEC;
    // four pixels of the window per step with their loads issued together: one dependent 16-byte load per thread at a time
    // left this pass at 0.4 - 0.6 of copy bandwidth (ncu: 34 % occupancy, long-scoreboard stalls)
    // (eight in flight: 0.1274 -> 0.1259 ms at level 1, not kept)
    for (int x = lo; x < hi; x += 4) {
      RawV<VEC, act_t> raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) raw[u] = ldraw<VEC>(row + min(x + u, hi - 1) * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (x + u < hi) {
          float t[VEC]; cvtraw<VEC>(raw[u], t);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const float bnv = fmaf(t[v], sc[v], sh[v]);
            if constexpr (MASKS) {
              // sum relu(bn(a0)) = scale * sum m a0 + shift * sum m: with the two mask sums the activation sum is free
              // (this pass is bound by instruction issue: ~60 instructions per 16-byte load left 0.5 of copy bandwidth)
              // (a NaN in A0 does not reach the pooled value this way; it still reaches the loss through A = ... + relu(bn2 A0),
              // whose max propagates NaN, so the optimizer's non-finite check sees it)
              if (bnv > 0.f) { am[v] += 1.f; ax[v] += t[v]; }
            } else {
              acc[v] += fmax_nan(bnv, 0.f);
            }
          }
        }
      }
    }
    const float inv = 1.f / static_cast<float>(hi - lo);
    float* o = tmp + ((b * H + y) * P + j) * C + cv * VEC;
#pragma unroll
    for (int v = 0; v < VEC; ++v) o[v] = (MASKS ? fmaf(sc[v], ax[v], sh[v] * am[v]) : acc[v]) * inv;
    if constexpr (MASKS) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) { o[plane + v] = am[v] * inv; o[2 * plane + v] = ax[v] * inv; }
    }
  }
}
// out[b, i, j, c] = scale_out * sum_y wgt(i, y) * tmp[b, y, j, c]; mode 0: pool windows over y, mode 1: bilinear^T.
// V channels per thread (4 with 16-byte accesses when C % 4 == 0): at fine pooled grids (P = 16 / 32, where the pooled
// map of the deep levels is larger than the feature map) this pass is instruction bound, not bandwidth bound, and the
// index decode + tap weights are per thread, not per channel.
template <int V>
__global__ void __launch_bounds__(256)
cols_reduce_kernel(const float* __restrict__ tmp, int B, int H, int P, int C, int mode, const float* mul, float* __restrict__ out,
                   const float* __restrict__ dot_with, double* dot_out) {
  const int CV = C / V;
  const long long total = static_cast<long long>(B) * P * P * CV;
  const float m = mul ? *mul : 1.f;
  float dot = 0.f;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    int c, j, i; long long b;
    decode4(idx, total < (1LL << 31), CV, P, P, c, j, i, b);
    c *= V;
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
    const float* col = tmp + (b * H * P + j) * C + c;          // + y * P * C
    const long long ystride = static_cast<long long>(P) * C;
    if (mode == 0) {
      int lo, hi; pool_win(i, H, P, lo, hi);
      for (int y = lo; y < hi; ++y) {
        float t[V]; ldf<V>(col + y * ystride, t);
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] += t[v];
      }
      const float den = static_cast<float>(hi - lo);
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] /= den;
    } else {
      const float ratio = static_cast<float>(H) / static_cast<float>(P);
      int lo = static_cast<int>(floorf((i - 0.5f) * ratio - 0.5f)) - 1;
      int hi = static_cast<int>(ceilf((i + 1.5f) * ratio - 0.5f)) + 1;
      lo = max(lo, 0); hi = min(hi, H - 1);
      for (int y = lo; y <= hi; ++y) {
        int i0, i1; float l1; bilerp_taps(y, P, H, i0, i1, l1);
        float wgt = 0.f;
        if (i0 == i) wgt += 1.f - l1;
        if (i1 == i) wgt += l1;
        if (wgt != 0.f) {
          float t[V]; ldf<V>(col + y * ystride, t);
#pragma unroll
          for (int v = 0; v < V; ++v) acc[v] += wgt * t[v];
        }
      }
    }
    const long long o = idx * V;
    float r[V];
#pragma unroll
    for (int v = 0; v < V; ++v) r[v] = acc[v] * m;
    stf<V>(out + o, r);
    if (dot_with != nullptr) {
      float w[V]; ldf<V>(dot_with + o, w);
#pragma unroll
      for (int v = 0; v < V; ++v) dot = fmaf(acc[v], w[v], dot);
    }
  }
  if (dot_out != nullptr) block_scalar_reduce_add(dot, dot_out);   // <unscaled result, dot_with>
}

// The same reduction with ONE CTA per pooled cell (b, i, j): CV = C / V channel vectors x 256 / CV row parts, the parts meet
// in shared memory.  For coarse pooled maps (P = 4 at 224^2: 1024 cells x 16 channel vectors, each walking 56 - 116 rows
// with one dependent load in flight) the thread-per-output kernel above is a latency chain on 64 CTAs: 45 us for the
// 14.7 MB of tmp at level 1 (ncu: 12 % of the warp slots active).
template <int V>
__global__ void __launch_bounds__(256)
cols_reduce_par_kernel(const float* __restrict__ tmp, int B, int H, int P, int C, int mode, const float* mul,
                       float* __restrict__ out, const float* __restrict__ dot_with, double* dot_out) {
  __shared__ __align__(16) float s_part[256 * V];
  const int CV = C / V;                       // divides 256 (host)
  const int parts = 256 / CV;
  const int cv = threadIdx.x % CV, part = threadIdx.x / CV;
  const float m = mul ? *mul : 1.f;
  const long long cells = static_cast<long long>(B) * P * P;
  const long long ystride = static_cast<long long>(P) * C;
  float dot = 0.f;
  for (long long cell = blockIdx.x; cell < cells; cell += gridDim.x) {
    const int j = static_cast<int>(cell % P);
    const int i = static_cast<int>((cell / P) % P);
    const long long b = cell / (static_cast<long long>(P) * P);
    const float* col = tmp + (b * H * P + j) * C + cv * V;          // + y * P * C
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
    float den = 1.f;
    if (mode == 0) {
      int lo, hi; pool_win(i, H, P, lo, hi);
      for (int y = lo + part; y < hi; y += parts) {
        float t[V]; ldf<V>(col + y * ystride, t);
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] += t[v];
      }
      den = static_cast<float>(hi - lo);
    } else {
      const float ratio = static_cast<float>(H) / static_cast<float>(P);
      int lo = static_cast<int>(floorf((i - 0.5f) * ratio - 0.5f)) - 1;
      int hi = static_cast<int>(ceilf((i + 1.5f) * ratio - 0.5f)) + 1;
      lo = max(lo, 0); hi = min(hi, H - 1);
      for (int y = lo + part; y <= hi; y += parts) {
        int i0, i1; float l1; bilerp_taps(y, P, H, i0, i1, l1);
        float wgt = 0.f;
        if (i0 == i) wgt += 1.f - l1;
        if (i1 == i) wgt += l1;
        if (wgt != 0.f) {
          float t[V]; ldf<V>(col + y * ystride, t);
#pragma unroll
          for (int v = 0; v < V; ++v) acc[v] += wgt * t[v];
        }
      }
    }
    stf<V>(s_part + threadIdx.x * V, acc);
    __syncthreads();
    if (part == 0) {
      for (int p2 = 1; p2 < parts; ++p2) {
        float t[V]; ldf<V>(s_part + (p2 * CV + cv) * V, t);
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] += t[v];
      }
      float r[V];
#pragma unroll
      for (int v = 0; v < V; ++v) { if (mode == 0) acc[v] /= den; r[v] = acc[v] * m; }
      const long long o = cell * C + cv * V;
      stf<V>(out + o, r);
      if (dot_with != nullptr) {
        float w[V]; ldf<V>(dot_with + o, w);
#pragma unroll
        for (int v = 0; v < V; ++v) dot = fmaf(acc[v], w[v], dot);
      }
    }
    __syncthreads();
  }
  if (dot_out != nullptr) block_scalar_reduce_add(dot, dot_out);   // <unscaled result, dot_with>
}

// =============================================================================================
// forward: L / A into the concat buffer
// =============================================================================================
template <int VEC>
__device__ __forceinline__ void bilerp_gather(const float* o, long long b, int y, int x, int H, int W, int P, int C,
                                              int c, float (&u)[VEC]) {
  int y0, y1, x0, x1; float ly, lx;
  bilerp_taps(y, P, H, y0, y1, ly);
  bilerp_taps(x, P, W, x0, x1, lx);
  const float* base = o + b * P * P * C + c;
  float a00[VEC], a01[VEC], a10[VEC], a11[VEC];
  ldf<VEC>(base + (y0 * P + x0) * C, a00); ldf<VEC>(base + (y0 * P + x1) * C, a01);
  ldf<VEC>(base + (y1 * P + x0) * C, a10); ldf<VEC>(base + (y1 * P + x1) * C, a11);
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
#pragma unroll
  for (int v = 0; v < VEC; ++v) u[v] = w00 * a00[v] + w01 * a01[v] + w10 * a10[v] + w11 * a11[v];
}

template <int VEC, int OCC>
__global__ void __launch_bounds__(256, OCC)
branch_act_fwd_kernel(const act_t* __restrict__ l0, long long ld_l0, const act_t* __restrict__ a0, long long ld_a0, int B, int H,
                      int W, int C, const float* s1, const float* t1, const float* s2, const float* t2, const float* o, int P,
                      const float* gamma, act_t* __restrict__ z, long long ld_z, grad_t* __restrict__ zb, long long ld_zb, int CL,
                      int PL) {
  const int cl = threadIdx.x % CL, pl = threadIdx.x / CL;
  const int c = (blockIdx.y * CL + cl) * VEC;
  if (pl >= PL || c >= C) return;
  const unsigned M = static_cast<unsigned>(B) * H * W;
  const float gm = *gamma;
  const bool has_l = l0 != nullptr;       // inference path: L is written by the folded conv + ReLU, only A is computed here
  float sc1[VEC], sh1[VEC], sc2[VEC], sh2[VEC];
  if (has_l) { ldf<VEC>(s1 + c, sc1); ldf<VEC>(t1 + c, sh1); }
  ldf<VEC>(s2 + c, sc2); ldf<VEC>(t2 + c, sh2);
  // kActPix pixels per iteration with all their loads issued before the first store: one pixel at a time (L0, store,
  // then A0 - the store to z kept the second load from moving up) left 16 bytes in flight per thread, ~12 KB per SM,
  // and the kernel at 0.61 of copy bandwidth.  The raw vectors cost 8 registers per pixel, so the kernel is built for
  // two resident blocks per SM (118 registers): 512 threads x 96 bytes = 48 KB on the way per SM.
  constexpr int kActPix = 3;
  const unsigned start = blockIdx.x * PL + pl, stride = gridDim.x * PL;
  PixIter it[kActPix];
#pragma unroll
  for (int u = 0; u < kActPix; ++u) it[u].init((start + u * stride) % M, kActPix * stride, H, W);
  for (unsigned m32 = start; m32 < M; m32 += kActPix * stride) {
    RawV<VEC, act_t> lraw[kActPix], araw[kActPix];
    bool ok[kActPix];
#pragma unroll
    for (int u = 0; u < kActPix; ++u) {
      const long long m = static_cast<long long>(m32) + static_cast<long long>(u) * stride;
      ok[u] = m < M;
      if (ok[u]) { if (has_l) lraw[u] = ldraw<VEC>(l0 + m * ld_l0 + c); araw[u] = ldraw<VEC>(a0 + m * ld_a0 + c); }
    }
#pragma unroll
    for (int u = 0; u < kActPix; ++u) {
      if (ok[u]) {
        const long long m = static_cast<long long>(m32) + static_cast<long long>(u) * stride;
        float outv[VEC], xin[VEC];
        if (has_l) {
          cvtraw<VEC>(lraw[u], xin);
#pragma unroll
          for (int v = 0; v < VEC; ++v) outv[v] = fmax_nan(fmaf(xin[v], sc1[v], sh1[v]), 0.f);
          stv<VEC>(z + m * ld_z + C + c, outv);
          if (zb != nullptr) stv<VEC>(zb + m * ld_zb + C + c, outv);
        }
        float up[VEC];
        bilerp_gather<VEC>(o, it[u].b, it[u].y, it[u].x, H, W, P, C, c, up);
        cvtraw<VEC>(araw[u], xin);
#pragma unroll
        for (int v = 0; v < VEC; ++v) outv[v] = gm * up[v] + fmax_nan(fmaf(xin[v], sc2[v], sh2[v]), 0.f);
        stv<VEC>(z + m * ld_z + 2 * C + c, outv);
        if (zb != nullptr) stv<VEC>(zb + m * ld_zb + 2 * C + c, outv);
      }
      it[u].next(H, W);
    }
  }
}

// Row-walk variant of branch_act_fwd for coarse pooled maps (W / P >= 8, 16-byte channel vectors): a thread owns channel
// vector cv of one row segment (b, y, k) and walks its pixels along x.  The bilinear taps of the pooled attention output,
// already interpolated along y and scaled by gamma, live in two register vectors that are reloaded only when the x cell
// changes (about P times per row) instead of four 32-byte gathers per pixel - those held the per-pixel kernel at 0.77 of
// copy bandwidth in training and at 0.35 in the inference path, where only the A half is computed (ncu: L1 throughput 76 %).
template <int OCC>
__global__ void __launch_bounds__(256, OCC)
branch_act_rows_kernel(const act_t* __restrict__ l0, long long ld_l0, const act_t* __restrict__ a0, long long ld_a0, int B, int H,
                       int W, int C, const float* s1, const float* t1, const float* s2, const float* t2, const float* __restrict__ o,
                       int P, const float* gamma, act_t* __restrict__ z, long long ld_z, int nseg) {
  constexpr int VEC = 8;
  const int CV = C / VEC;
  const long long total = static_cast<long long>(B) * H * nseg * CV;
  const bool has_l = l0 != nullptr;
  const float gm = *gamma;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int cv, k, y; long long b;
    decode4(i, total < (1LL << 31), CV, nseg, H, cv, k, y, b);
    const int c = cv * VEC;
    const int xlo = static_cast<int>(static_cast<long long>(k) * W / nseg), xhi = static_cast<int>(static_cast<long long>(k + 1) * W / nseg);
    float sc1[VEC], sh1[VEC], sc2[VEC], sh2[VEC];
    if (has_l) { ldf<VEC>(s1 + c, sc1); ldf<VEC>(t1 + c, sh1); }
    ldf<VEC>(s2 + c, sc2); ldf<VEC>(t2 + c, sh2);
    int y0, y1; float ly;
    bilerp_taps(y, P, H, y0, y1, ly);
    const float* ob = o + b * P * P * C + c;
    const long long mrow = (b * H + y) * W;
    float c0v[VEC], c1v[VEC];
    int cur_x0 = -1;
    for (int x = xlo; x < xhi; x += 4) {
      RawV<VEC, act_t> lraw[4], araw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long m = mrow + min(x + u, xhi - 1);
        if (has_l) lraw[u] = ldraw<VEC>(l0 + m * ld_l0 + c);
        araw[u] = ldraw<VEC>(a0 + m * ld_a0 + c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int xx = x + u;
        if (xx < xhi) {
          int x0, x1; float lx;
          bilerp_taps(xx, P, W, x0, x1, lx);
          if (x0 != cur_x0) {          // new x cell: y-interpolate the two tap columns once
            float a00[VEC], a01[VEC], a10[VEC], a11[VEC];
            ldf<VEC>(ob + (y0 * P + x0) * C, a00); ldf<VEC>(ob + (y0 * P + x1) * C, a01);
            ldf<VEC>(ob + (y1 * P + x0) * C, a10); ldf<VEC>(ob + (y1 * P + x1) * C, a11);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              c0v[v] = gm * fmaf(ly, a10[v] - a00[v], a00[v]);
              c1v[v] = gm * fmaf(ly, a11[v] - a01[v], a01[v]);
            }
            cur_x0 = x0;
          }
          const long long m = mrow + xx;
          float outv[VEC], xin[VEC];
          if (has_l) {
            cvtraw<VEC>(lraw[u], xin);
#pragma unroll
            for (int v = 0; v < VEC; ++v) outv[v] = fmax_nan(fmaf(xin[v], sc1[v], sh1[v]), 0.f);
            stv<VEC>(z + m * ld_z + C + c, outv);
          }
          cvtraw<VEC>(araw[u], xin);
#pragma unroll
          for (int v = 0; v < VEC; ++v)
            outv[v] = fmaf(lx, c1v[v] - c0v[v], c0v[v]) + fmax_nan(fmaf(xin[v], sc2[v], sh2[v]), 0.f);
          stv<VEC>(z + m * ld_z + 2 * C + c, outv);
        }
      }
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256)
gate_mix_fwd_kernel(const act_t* g0, long long ld_g0, long long M, int C, const float* s3, const float* t3, act_t* z,
                    long long ld_z, grad_t* zb, long long ld_zb, int CL, int PL) {
  const int cl = threadIdx.x % CL, pl = threadIdx.x / CL;
  const int c = (blockIdx.y * CL + cl) * VEC;
  if (pl >= PL || c >= C) return;
  float sc[VEC], sh[VEC];
  ldf<VEC>(s3 + c, sc); ldf<VEC>(t3 + c, sh);
  // kGatePix pixels per step, every load issued before the first use (see branch_act_fwd_kernel): the sigmoid makes the
  // compute phase long enough that one pixel at a time leaves the memory pipe idle a third of the time
  const long long stride = static_cast<long long>(gridDim.x) * PL;
  for (long long m0 = static_cast<long long>(blockIdx.x) * PL + pl; m0 < M; m0 += kGatePix * stride) {
    RawV<VEC, act_t> graw[kGatePix], lraw[kGatePix], araw[kGatePix];
#pragma unroll
    for (int u = 0; u < kGatePix; ++u) {
      const long long m = min(m0 + u * stride, M - 1);
      graw[u] = ldraw<VEC>(g0 + m * ld_g0 + c);
      lraw[u] = ldraw<VEC>(z + m * ld_z + C + c); araw[u] = ldraw<VEC>(z + m * ld_z + 2 * C + c);
    }
#pragma unroll
    for (int u = 0; u < kGatePix; ++u) {
      const long long m = m0 + u * stride;
      if (m < M) {
        float g[VEC], l[VEC], a[VEC], f[VEC];
        cvtraw<VEC>(graw[u], g); cvtraw<VEC>(lraw[u], l); cvtraw<VEC>(araw[u], a);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const float gg = fast_sigmoid(fmaf(g[v], sc[v], sh[v]));
          f[v] = gg * l[v] + (1.f - gg) * a[v];
        }
        stv<VEC>(z + m * ld_z + c, f);
        if (zb != nullptr) stv<VEC>(zb + m * ld_zb + c, f);
      }
    }
  }
}

// y = relu(bn4(F0)) + rs*R, optional 2x2 max pool; one thread per 2x2 window and channel vector
// PLAIN (the addition / attention-only ablation blocks): y = F0 [+ F1] + rs*R, no BatchNorm / ReLU
// PRE: the eight (twelve) loads of a window are issued before its first store - a store to y orders the next pixel's loads
// behind it (same element type, may alias), which makes a window four dependent round trips.
template <int VEC, bool PLAIN = false, bool PRE = false>
__global__ void __launch_bounds__(256, PRE ? 3 : 4)
block_out_fwd_kernel(const act_t* f0, long long ld_f0, const act_t* r, long long ld_r, int B, int H, int W, int C,
                     const float* s4, const float* t4, const float* res_scale, act_t* y, long long ld_y, act_t* yp,
                     long long ld_yp, grad_t* yb, long long ld_yb, grad_t* ypb, long long ld_ypb, int CL, int PL,
                     const act_t* f1 = nullptr, long long ld_f1 = 0) {
  const int cl = threadIdx.x % CL, pl = threadIdx.x / CL;
  const int c = (blockIdx.y * CL + cl) * VEC;
  if (pl >= PL || c >= C) return;
  const unsigned Hw = (H + 1) / 2, Ww = (W + 1) / 2;
  const unsigned Hp = H / 2, Wp = W / 2;
  const unsigned nwin = static_cast<unsigned>(B) * Hw * Ww;
  const float rs = *res_scale;
  float sc[VEC], sh[VEC];
  if (!PLAIN) { ldf<VEC>(s4 + c, sc); ldf<VEC>(t4 + c, sh); }
  for (unsigned wi = blockIdx.x * PL + pl; wi < nwin; wi += gridDim.x * PL) {
    const unsigned xo = wi % Ww, q = wi / Ww, yo = q % Hw, b = q / Hw;
    float mx[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) mx[v] = -INFINITY;
    RawV<VEC, act_t> fraw[PRE ? 4 : 1], rraw[PRE ? 4 : 1], graw[PRE && PLAIN ? 4 : 1];
    if constexpr (PRE) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // edge windows of odd maps: the clamped duplicate is loaded and dropped
        const unsigned yy = min(2 * yo + (k >> 1), static_cast<unsigned>(H) - 1), xx = min(2 * xo + (k & 1), static_cast<unsigned>(W) - 1);
        const long long m = (static_cast<long long>(b) * H + yy) * W + xx;
        fraw[k] = ldraw<VEC>(f0 + m * ld_f0 + c); rraw[k] = ldraw<VEC>(r + m * ld_r + c);
        if constexpr (PLAIN) { if (f1 != nullptr) graw[k] = ldraw<VEC>(f1 + m * ld_f1 + c); }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned yy = 2 * yo + (k >> 1), xx = 2 * xo + (k & 1);
      if (yy >= static_cast<unsigned>(H) || xx >= static_cast<unsigned>(W)) continue;
      const long long m = (static_cast<long long>(b) * H + yy) * W + xx;
      float fv[VEC], rv[VEC], ov[VEC];
      if constexpr (PRE) { cvtraw<VEC>(fraw[k], fv); cvtraw<VEC>(rraw[k], rv); }
      else { ldv<VEC>(f0 + m * ld_f0 + c, fv); ldv<VEC>(r + m * ld_r + c, rv); }
      if (PLAIN) {
        if (f1 != nullptr) {
          float gv[VEC];
          if constexpr (PRE && PLAIN) cvtraw<VEC>(graw[k], gv);
          else ldv<VEC>(f1 + m * ld_f1 + c, gv);
#pragma unroll
          for (int v = 0; v < VEC; ++v) fv[v] += gv[v];
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) ov[v] = fv[v] + rs * rv[v];
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) ov[v] = fmax_nan(fmaf(fv[v], sc[v], sh[v]), 0.f) + rs * rv[v];
      }
      stv<VEC>(y + m * ld_y + c, ov);
      if (yb != nullptr) stv<VEC>(yb + m * ld_yb + c, ov);
      // pool over the values as stored (fp16), so backward can recompute the argmax from y
#pragma unroll
      for (int v = 0; v < VEC; ++v) mx[v] = fmaxf(mx[v], Cvt<act_t>::to_f(Cvt<act_t>::from_f(ov[v])));
    }
    if (yp != nullptr && yo < Hp && xo < Wp) {
      const long long mp = (static_cast<long long>(b) * Hp + yo) * Wp + xo;
      stv<VEC>(yp + mp * ld_yp + c, mx);
      if (ypb != nullptr) stv<VEC>(ypb + mp * ld_ypb + c, mx);
    }
  }
}

// =============================================================================================
// backward
// Register budget: these kernels are pure streaming, so what matters is resident threads per SM.  Per-channel
// constants are kept to (scale, shift) for the activation mask plus two folded BN-backward coefficients
//   dx = scale*(d - k1 - (x - mean)*k2)  =  scale*d + p*x + q,   p = -scale*k2,  q = scale*(k2*mean - k1)
// and the second BatchNorm reduction is accumulated as sum(d*x) and turned into sum(d*xhat) = invstd*(sum(d*x) -
// mean*sum(d)) in double when a block flushes, so mean / invstd never live in registers.
// =============================================================================================
template <int VEC>
__device__ __forceinline__ void flush_bn_partials(float (&acc)[2][VEC], int cl, int pl, int CL, int PL, int c_base, int C,
                                                  const float* mean, const float* invstd, double* red, float* s_red) {
  // s_red: [2][PL][CL*VEC]
  const int row = CL * VEC;
  if (cl < CL && pl < PL) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int v = 0; v < VEC; ++v) s_red[(r * PL + pl) * row + cl * VEC + v] = acc[r][v];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < row; idx += blockDim.x) {
    const int c = c_base + idx;
    if (c < C) {
      float s1 = 0.f, s2 = 0.f;
      for (int p = 0; p < PL; ++p) { s1 += s_red[p * row + idx]; s2 += s_red[(PL + p) * row + idx]; }
      atomicAdd(red + c, static_cast<double>(s1));
      atomicAdd(red + C + c, static_cast<double>(invstd[c]) * (static_cast<double>(s2) - static_cast<double>(mean[c]) * static_cast<double>(s1)));
    }
  }
}

// folded BN-backward coefficients of channel vector c (see the section comment)
template <int VEC>
__device__ __forceinline__ void bn_bwd_coeffs(const float* scale, const float* mean, const float* invstd, const double* red,
                                              int C, int c, double invn, float (&sc)[VEC], float (&p)[VEC], float (&q)[VEC]) {
  float mu[VEC], is[VEC];
  ldf<VEC>(scale + c, sc); ldf<VEC>(mean + c, mu); ldf<VEC>(invstd + c, is);
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const float k1 = static_cast<float>(red[c + v] * invn);
    const float k2 = static_cast<float>(red[C + c + v] * invn) * is[v];
    p[v] = -sc[v] * k2;
    q[v] = sc[v] * (k2 * mu[v] - k1);
  }
}

template <int VEC, int OCC>
__global__ void __launch_bounds__(256, OCC)
block_out_bwd_reduce_kernel(const grad_t* dskip, long long ld_dskip, const grad_t* dyp, long long ld_dyp,
                            const act_t* y, long long ld_y, const act_t* f0, long long ld_f0, const act_t* r,
                            long long ld_r, int B, int H, int W, int C, const float* s4, const float* t4,
                            const float* mean4, const float* invstd4, grad_t* dy_out, long long ld_dy, double* red4,
                            double* drs, int CL, int PL) {
  __shared__ float s_red[2 * 256 * VEC];
  const int cl = threadIdx.x % CL, pl = threadIdx.x / CL;
  const int c_base = blockIdx.y * CL * VEC;
  const int c = c_base + cl * VEC;
  const bool active = pl < PL && c < C;
  const int Hw = (H + 1) / 2, Ww = (W + 1) / 2;
  const int Hp = H / 2, Wp = W / 2;
  const unsigned nwin = static_cast<unsigned>(B) * Hw * Ww;
  float acc[2][VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { acc[0][v] = 0.f; acc[1][v] = 0.f; }
  float drs_acc = 0.f;
  if (active) {
    float sc[VEC], sh[VEC];
    ldf<VEC>(s4 + c, sc); ldf<VEC>(t4 + c, sh);
    PixIter it; it.init(blockIdx.x * PL + pl, gridDim.x * PL, Hw, Ww);
    for (unsigned wi = blockIdx.x * PL + pl; wi < nwin; wi += gridDim.x * PL, it.next(Hw, Ww)) {
      const int xo = it.x, yo = it.y;
      const long long b = it.b;
      // argmax of the stored y over the window (first maximum in scan order, like ATen's max_pool2d)
      int arg[VEC];
      float gp[VEC];
      const bool pooled = dyp != nullptr && yo < Hp && xo < Wp;
      if (pooled) {
        float best[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { best[v] = -INFINITY; arg[v] = 0; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const long long m = (b * H + 2 * yo + (k >> 1)) * W + 2 * xo + (k & 1);
          float yv[VEC]; ldv<VEC>(y + m * ld_y + c, yv);
#pragma unroll
          for (int v = 0; v < VEC; ++v) if (yv[v] > best[v]) { best[v] = yv[v]; arg[v] = k; }
        }
        ldv<VEC>(dyp + ((b * Hp + yo) * Wp + xo) * ld_dyp + c, gp);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int yy = 2 * yo + (k >> 1), xx = 2 * xo + (k & 1);
        if (yy >= H || xx >= W) continue;
        const long long m = (b * H + yy) * W + xx;
        float d[VEC];
        if (dskip != nullptr) ldv<VEC>(dskip + m * ld_dskip + c, d);
        else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) d[v] = 0.f;
        }
        if (pooled) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) if (arg[v] == k) d[v] += gp[v];
        }
        if (dy_out != nullptr && (dy_out != dskip || pooled)) {
          // round to the stored precision first, so the sums see exactly what the later passes read
#pragma unroll
          for (int v = 0; v < VEC; ++v) d[v] = Cvt<grad_t>::to_f(Cvt<grad_t>::from_f(d[v]));
          stv<VEC>(dy_out + m * ld_dy + c, d);
        }
        float fv[VEC], rv[VEC];
        ldv<VEC>(f0 + m * ld_f0 + c, fv); ldv<VEC>(r + m * ld_r + c, rv);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          drs_acc = fmaf(d[v], rv[v], drs_acc);
          const float d4 = fmaf(fv[v], sc[v], sh[v]) > 0.f ? d[v] : 0.f;
          acc[0][v] += d4;
          acc[1][v] = fmaf(d4, fv[v], acc[1][v]);
        }
      }
    }
  }
  flush_bn_partials<VEC>(acc, cl, pl, CL, PL, c_base, C, mean4, invstd4, red4, s_red);
  block_scalar_reduce_add(drs_acc, drs);
}

// The same pass for the blocks with a max pool behind them, every load of a 2x2 window issued before the first use.  Above,
// the in-place store of dy (dy_out == dskip) orders the loads of the next pixel behind it, so a window is four dependent
// round trips after the argmax loads: 0.70 of copy bandwidth on the encoder blocks against 0.92 on the decoder blocks.
// 17 sixteen-byte loads in flight per thread: two resident CTAs per SM.
template <int VEC>
__global__ void __launch_bounds__(256, 2)
block_out_bwd_reduce_pre_kernel(const grad_t* dskip, long long ld_dskip, const grad_t* dyp, long long ld_dyp,
                                const act_t* y, long long ld_y, const act_t* f0, long long ld_f0, const act_t* r,
                                long long ld_r, int B, int H, int W, int C, const float* s4, const float* t4,
                                const float* mean4, const float* invstd4, grad_t* dy_out, long long ld_dy, double* red4,
                                double* drs, int CL, int PL) {
  __shared__ float s_red[2 * 256 * VEC];
  __shared__ __align__(16) float s_aff[2 * 32 * VEC];     // (scale, shift) of the CTA's channels: read at use, not held in registers
  const int cl = threadIdx.x % CL, pl = threadIdx.x / CL;
  const int c_base = blockIdx.y * CL * VEC;
  const int c = c_base + cl * VEC;
  const bool active = pl < PL && c < C;
  const int Hw = (H + 1) / 2, Ww = (W + 1) / 2;
  const int Hp = H / 2, Wp = W / 2;
  const unsigned nwin = static_cast<unsigned>(B) * Hw * Ww;
  float acc[2][VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { acc[0][v] = 0.f; acc[1][v] = 0.f; }
  float drs_acc = 0.f;
  if (threadIdx.x < CL * VEC && c_base + threadIdx.x < C) {
    s_aff[threadIdx.x] = s4[c_base + threadIdx.x];
    s_aff[32 * VEC + threadIdx.x] = t4[c_base + threadIdx.x];
  }
  __syncthreads();
  if (active) {
    PixIter it; it.init(blockIdx.x * PL + pl, gridDim.x * PL, Hw, Ww);
    for (unsigned wi = blockIdx.x * PL + pl; wi < nwin; wi += gridDim.x * PL, it.next(Hw, Ww)) {
      const int xo = it.x, yo = it.y;
      const long long b = it.b;
      const bool pooled = yo < Hp && xo < Wp;         // dyp is given (host); edge windows of odd maps have no pooled cell
      RawV<VEC, act_t> yraw[4], fraw[4], rraw[4];
      RawV<VEC, grad_t> draw[4], gpraw;
      long long mk[4];
      bool ok[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int yy = 2 * yo + (k >> 1), xx = 2 * xo + (k & 1);
        ok[k] = yy < H && xx < W;
        mk[k] = (b * H + min(yy, H - 1)) * W + min(xx, W - 1);   // clamped duplicates are loaded and dropped, never stored
        yraw[k] = ldraw<VEC>(y + mk[k] * ld_y + c);
        if (dskip != nullptr) draw[k] = ldraw<VEC>(dskip + mk[k] * ld_dskip + c);
        fraw[k] = ldraw<VEC>(f0 + mk[k] * ld_f0 + c);
        rraw[k] = ldraw<VEC>(r + mk[k] * ld_r + c);
      }
      gpraw = ldraw<VEC>(dyp + ((b * Hp + min(yo, Hp - 1)) * Wp + min(xo, Wp - 1)) * ld_dyp + c);
      // argmax of the stored y over the window (first maximum in scan order, like ATen's max_pool2d)
      int arg[VEC];
      float gp[VEC], best[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) { best[v] = -INFINITY; arg[v] = 0; }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float yv[VEC]; cvtraw<VEC>(yraw[k], yv);
#pragma unroll
        for (int v = 0; v < VEC; ++v) if (yv[v] > best[v]) { best[v] = yv[v]; arg[v] = k; }
      }
      cvtraw<VEC>(gpraw, gp);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!ok[k]) continue;
        float d[VEC];
        if (dskip != nullptr) cvtraw<VEC>(draw[k], d);
        else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) d[v] = 0.f;
        }
        if (pooled) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) if (arg[v] == k) d[v] += gp[v];
        }
        if (dy_out != dskip || pooled) {
          // round to the stored precision first, so the sums see exactly what the later passes read
#pragma unroll
          for (int v = 0; v < VEC; ++v) d[v] = Cvt<grad_t>::to_f(Cvt<grad_t>::from_f(d[v]));
          stv<VEC>(dy_out + mk[k] * ld_dy + c, d);
        }
        float fv[VEC], rv[VEC], sc[VEC], sh[VEC];
        cvtraw<VEC>(fraw[k], fv); cvtraw<VEC>(rraw[k], rv);
        ldf<VEC>(s_aff + cl * VEC, sc); ldf<VEC>(s_aff + 32 * VEC + cl * VEC, sh);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          drs_acc = fmaf(d[v], rv[v], drs_acc);
          const float d4 = fmaf(fv[v], sc[v], sh[v]) > 0.f ? d[v] : 0.f;
          acc[0][v] += d4;
          acc[1][v] = fmaf(d4, fv[v], acc[1][v]);
        }
      }
    }
  }
  flush_bn_partials<VEC>(acc, cl, pl, CL, PL, c_base, C, mean4, invstd4, red4, s_red);
  block_scalar_reduce_add(drs_acc, drs);
}

// dx = scale * (d - k1 - xhat*k2), d = dy * [relu mask] (act_mode 0) or dy (act_mode 2)
template <int VEC>
__global__ void __launch_bounds__(256, 4)
bn_bwd_apply_kernel(const grad_t* dy, long long ld_dy, const act_t* x, long long ld_x, long long M, int C,
                    const float* scale, const float* shift, const float* mean, const float* invstd, const double* red,
                    int act_mode, grad_t* dx, long long ld_dx, int CL, int PL) {
  const int cl = threadIdx.x % CL, pl = threadIdx.x / CL;
  const int c = (blockIdx.y * CL + cl) * VEC;
  if (pl >= PL || c >= C) return;
  float sc[VEC], sh[VEC], p[VEC], q[VEC];
  bn_bwd_coeffs<VEC>(scale, mean, invstd, red, C, c, 1.0 / static_cast<double>(M), sc, p, q);
  ldf<VEC>(shift + c, sh);
  for (long long m = static_cast<long long>(blockIdx.x) * PL + pl; m < M; m += static_cast<long long>(gridDim.x) * PL) {
    float d[VEC], xv[VEC], o[VEC];
    ldv<VEC>(dy + m * ld_dy + c, d); ldv<VEC>(x + m * ld_x + c, xv);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float dd = d[v];
      if (act_mode == 0 && !(fmaf(xv[v], sc[v], sh[v]) > 0.f)) dd = 0.f;
      o[v] = fmaf(sc[v], dd, fmaf(p[v], xv[v], q[v]));
    }
    stv<VEC>(dx + m * ld_dx + c, o);
  }
}

template <int VEC, int OCC>
__global__ void __launch_bounds__(256, OCC)
gate_mix_bwd_reduce_kernel(const grad_t* dz, long long ld_dz, const act_t* z, long long ld_z, const act_t* g0,
                           long long ld_g0, long long M, int C, const float* s3, const float* t3, const float* mean3,
                           const float* invstd3, double* red3, int CL, int PL) {
  __shared__ float s_red[2 * 256 * VEC];
  const int cl = threadIdx.x % CL, pl = threadIdx.x / CL;
  const int c_base = blockIdx.y * CL * VEC;
  const int c = c_base + cl * VEC;
  float acc[2][VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { acc[0][v] = 0.f; acc[1][v] = 0.f; }
  if (pl < PL && c < C) {
    float sc[VEC], sh[VEC];
    ldf<VEC>(s3 + c, sc); ldf<VEC>(t3 + c, sh);
    const long long stride = static_cast<long long>(gridDim.x) * PL;
    for (long long m0 = static_cast<long long>(blockIdx.x) * PL + pl; m0 < M; m0 += kGatePix * stride) {
      RawV<VEC, grad_t> draw[kGatePix];
      RawV<VEC, act_t> lraw[kGatePix], araw[kGatePix], graw[kGatePix];
#pragma unroll
      for (int u = 0; u < kGatePix; ++u) {
        const long long m = min(m0 + u * stride, M - 1);
        draw[u] = ldraw<VEC>(dz + m * ld_dz + c);
        lraw[u] = ldraw<VEC>(z + m * ld_z + C + c); araw[u] = ldraw<VEC>(z + m * ld_z + 2 * C + c);
        graw[u] = ldraw<VEC>(g0 + m * ld_g0 + c);
      }
#pragma unroll
      for (int u = 0; u < kGatePix; ++u) {
        if (m0 + u * stride < M) {
          float df[VEC], l[VEC], a[VEC], g[VEC];
          cvtraw<VEC>(draw[u], df); cvtraw<VEC>(lraw[u], l); cvtraw<VEC>(araw[u], a); cvtraw<VEC>(graw[u], g);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const float gg = fast_sigmoid(fmaf(g[v], sc[v], sh[v]));
            const float ds = df[v] * (l[v] - a[v]) * gg * (1.f - gg);
            acc[0][v] += ds;
            acc[1][v] = fmaf(ds, g[v], acc[1][v]);
          }
        }
      }
    }
  }
  flush_bn_partials<VEC>(acc, cl, pl, CL, PL, c_base, C, mean3, invstd3, red3, s_red);
}

template <int VEC, int OCC>
__global__ void __launch_bounds__(256, OCC)
gate_mix_bwd_apply_kernel(const grad_t* dz, long long ld_dz, const act_t* z, long long ld_z, const act_t* g0, long long ld_g0,
                          long long M, int C, const float* s3, const float* t3, const float* mean3, const float* invstd3,
                          const double* red3, grad_t* dg0, long long ld_dg0, int CL, int PL) {
  const int cl = threadIdx.x % CL, pl = threadIdx.x / CL;
  const int c = (blockIdx.y * CL + cl) * VEC;
  if (pl >= PL || c >= C) return;
  float sc[VEC], sh[VEC], p[VEC], q[VEC];
  bn_bwd_coeffs<VEC>(s3, mean3, invstd3, red3, C, c, 1.0 / static_cast<double>(M), sc, p, q);
  ldf<VEC>(t3 + c, sh);
  const long long stride = static_cast<long long>(gridDim.x) * PL;
  for (long long m0 = static_cast<long long>(blockIdx.x) * PL + pl; m0 < M; m0 += kGatePix * stride) {
    RawV<VEC, grad_t> draw[kGatePix];
    RawV<VEC, act_t> lraw[kGatePix], araw[kGatePix], graw[kGatePix];
#pragma unroll
    for (int u = 0; u < kGatePix; ++u) {
      const long long m = min(m0 + u * stride, M - 1);
      draw[u] = ldraw<VEC>(dz + m * ld_dz + c);
      lraw[u] = ldraw<VEC>(z + m * ld_z + C + c); araw[u] = ldraw<VEC>(z + m * ld_z + 2 * C + c);
      graw[u] = ldraw<VEC>(g0 + m * ld_g0 + c);
    }
#pragma unroll
    for (int u = 0; u < kGatePix; ++u) {
      const long long m = m0 + u * stride;
      if (m < M) {
        float df[VEC], l[VEC], a[VEC], g[VEC], o[VEC];
        cvtraw<VEC>(draw[u], df); cvtraw<VEC>(lraw[u], l); cvtraw<VEC>(araw[u], a); cvtraw<VEC>(graw[u], g);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const float gg = fast_sigmoid(fmaf(g[v], sc[v], sh[v]));
          const float ds = df[v] * (l[v] - a[v]) * gg * (1.f - gg);
          o[v] = fmaf(sc[v], ds, fmaf(p[v], g[v], q[v]));
        }
        stv<VEC>(dg0 + m * ld_dg0 + c, o);
      }
    }
  }
}

// pass 1 of the branch backward: gate-mix terms into dL / dA (in place) and the BN1 reductions.
// PIX pixels per grid-stride step with every load issued before the first use, like the gate kernels: five 16-byte loads
// per pixel and a sigmoid per element leave the memory pipe idle between steps when one pixel is walked at a time.
template <int VEC, int OCC, int PIX>
__global__ void __launch_bounds__(256, OCC)
branch_bwd_reduce1_kernel(grad_t* dz, long long ld_dz, const act_t* l0, long long ld_l0, const act_t* g0, long long ld_g0,
                          long long M, int C, const float* s1, const float* t1, const float* mean1,
                          const float* invstd1, const float* s3, const float* t3, double* red1, int CL, int PL) {
  __shared__ float s_red[2 * 256 * VEC];
  const int cl = threadIdx.x % CL, pl = threadIdx.x / CL;
  const int c_base = blockIdx.y * CL * VEC;
  const int c = c_base + cl * VEC;
  float acc[2][VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { acc[0][v] = 0.f; acc[1][v] = 0.f; }
  if (pl < PL && c < C) {
    const bool gated = g0 != nullptr;       // the ablation blocks without a gate: dL / dA are final as they arrive
    float sc[VEC], sh[VEC], sc3[VEC], sh3[VEC];
    ldf<VEC>(s1 + c, sc); ldf<VEC>(t1 + c, sh);
    if (gated) { ldf<VEC>(s3 + c, sc3); ldf<VEC>(t3 + c, sh3); }
    const long long stride = static_cast<long long>(gridDim.x) * PL;
    for (long long m0 = static_cast<long long>(blockIdx.x) * PL + pl; m0 < M; m0 += PIX * stride) {
      RawV<VEC, grad_t> dlraw[PIX], daraw[PIX], dfraw[PIX];
      RawV<VEC, act_t> lraw[PIX], graw[PIX];
#pragma unroll
      for (int u = 0; u < PIX; ++u) {
        const long long m = min(m0 + u * stride, M - 1);   // a clamped duplicate is loaded and dropped, never stored
        dlraw[u] = ldraw<VEC>(dz + m * ld_dz + C + c);
        lraw[u] = ldraw<VEC>(l0 + m * ld_l0 + c);
        if (gated) {
          daraw[u] = ldraw<VEC>(dz + m * ld_dz + 2 * C + c);
          dfraw[u] = ldraw<VEC>(dz + m * ld_dz + c);
          graw[u] = ldraw<VEC>(g0 + m * ld_g0 + c);
        }
      }
#pragma unroll
      for (int u = 0; u < PIX; ++u) {
        const long long m = m0 + u * stride;
        if (m < M) {
          float dl[VEC], lv[VEC];
          cvtraw<VEC>(dlraw[u], dl); cvtraw<VEC>(lraw[u], lv);
          if (gated) {
            float da[VEC], df[VEC], gv[VEC];
            cvtraw<VEC>(daraw[u], da); cvtraw<VEC>(dfraw[u], df); cvtraw<VEC>(graw[u], gv);
            // fused = g*L + (1-g)*A  ->  dL += df*g, dA += df*(1-g); rounded to the stored precision so that the sums see
            // exactly what the later passes read back
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              const float gg = fast_sigmoid(fmaf(gv[v], sc3[v], sh3[v]));
              dl[v] = Cvt<grad_t>::to_f(Cvt<grad_t>::from_f(fmaf(df[v], gg, dl[v])));
              da[v] = fmaf(df[v], 1.f - gg, da[v]);
            }
            stv<VEC>(dz + m * ld_dz + C + c, dl); stv<VEC>(dz + m * ld_dz + 2 * C + c, da);
          }
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const float d1 = fmaf(lv[v], sc[v], sh[v]) > 0.f ? dl[v] : 0.f;
            acc[0][v] += d1;
            acc[1][v] = fmaf(d1, lv[v], acc[1][v]);
          }
        }
      }
    }
  }
  flush_bn_partials<VEC>(acc, cl, pl, CL, PL, c_base, C, mean1, invstd1, red1, s_red);
}

// bilinear^T along x: tmp[b, y, px, c] = sum_x wx(x, px) * dA[b, y, x, c]
// `parts` lanes share one output: lane `part` takes every parts-th pixel of the ~2*W/P wide window with four
// independent loads in flight, and the partial sums meet through shuffles (the lanes of one output are `group` = CV *
// parts consecutive threads, a power of two <= 32).  One thread per output walked the 116-pixel window of level 1 with
// a single dependent load in flight: 226 us for 411 MB (ncu), 0.28 of copy bandwidth.
template <int VEC>
__global__ void __launch_bounds__(256)
bilerpT_rows_kernel(const grad_t* dz, long long ld_dz, int B, int H, int W, int C, int P, float* tmp, int parts) {
  const int CV = C / VEC;
  const int group = CV * parts;
  const long long total = static_cast<long long>(B) * H * P * group;
  const bool small = total < (1LL << 31);
  const float ratio = static_cast<float>(W) / static_cast<float>(P);
  for (long long base = static_cast<long long>(blockIdx.x) * blockDim.x; base < total;
       base += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = base + threadIdx.x;
    const bool active = i < total;           // groups are aligned and total % group == 0: a group is active as a whole
    int lg, px, y; long long b;
    decode4(active ? i : total - 1, small, group, P, H, lg, px, y, b);
    const int part = lg / CV;
    const int c = (lg - part * CV) * VEC;
    int lo = static_cast<int>(floorf((px - 0.5f) * ratio - 0.5f)) - 1;
    int hi = static_cast<int>(ceilf((px + 1.5f) * ratio - 0.5f)) + 1;
    lo = max(lo, 0); hi = min(hi, W - 1);
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    const grad_t* row = dz + ((b * H + y) * W) * ld_dz + 2 * C + c;
    for (int x = lo + part; x <= hi; x += 4 * parts) {
      float d[4][VEC], wg[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int xx = x + u * parts;
        const bool ok = xx <= hi;
        const int xc = ok ? xx : hi;
        int x0, x1; float lx; bilerp_taps(xc, P, W, x0, x1, lx);
        float wgt = 0.f;
        if (x0 == px) wgt += 1.f - lx;
        if (x1 == px) wgt += lx;
        wg[u] = ok ? wgt : 0.f;
        ldv<VEC>(row + xc * ld_dz, d[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = fmaf(wg[u], d[u][v], acc[v]);
    }
    for (int off = CV; off < group; off <<= 1) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], off);
    }
    if (active && part == 0) {
      float* o = tmp + ((b * H + y) * P + px) * C + c;
#pragma unroll
      for (int v = 0; v < VEC; ++v) o[v] = acc[v];
    }
  }
}

// Measured and not kept (session AI / AJ, profiles/experiments/): a one-cell variant - each lane group walks only the pixels
// whose LEFT tap is its cell, keeps (1 - lx) d and lx d as two sums, neighbours meet in shared memory - loads every pixel
// once and issues 25 % fewer instructions, yet takes the same 127 us at level 1 (ncu: issue 54 %, DRAM 3.2 TB/s).
// lanes per output of bilerpT_rows_kernel: CV * parts must be a power of two <= 32, and a part should keep >= 8 pixels
static int bilerpT_parts(int CV, int W, int P) {
  if (CV <= 0 || (CV & (CV - 1)) != 0 || CV >= 32) return 1;
  const int window = 2 * ((W + P - 1) / P) + 4;
  int parts = 1;
  while (parts * 2 * CV <= 32 && parts * 2 * 8 <= window) parts *= 2;
  return parts;
}

constexpr int kPixInFlight = 4;   // pixels per grid-stride step in the A-branch backward kernels

// adaptive_avg_pool^T gather of dpooled at pixel (y, x)
template <int VEC>
__device__ __forceinline__ void poolT_gather(const float* dp, unsigned b, int y, int x, int P, int C, int c,
                                             const PoolTabs& tb, float (&g)[VEC]) {
#pragma unroll
  for (int v = 0; v < VEC; ++v) g[v] = 0.f;
  const int ty = tb.yt[y], tx = tb.xt[x];
  const int ilo = ty & 0xffff, ihi = ty >> 16, jlo = tx & 0xffff, jhi = tx >> 16;
  const float* base = dp + static_cast<long long>(b) * P * P * C + c;
  for (int i = ilo; i <= ihi; ++i) {
    const float wi = tb.wy[i];
    for (int j = jlo; j <= jhi; ++j) {
      const float wgt = wi * tb.wx[j];
      float t[VEC]; ldf<VEC>(base + (i * P + j) * C, t);
#pragma unroll
      for (int v = 0; v < VEC; ++v) g[v] = fmaf(wgt, t[v], g[v]);
    }
  }
}

template <int VEC, int OCC>
__global__ void __launch_bounds__(256, OCC)
branch_bwd_reduce2_kernel(const grad_t* dz, long long ld_dz, const act_t* a0, long long ld_a0, int B, int H, int W,
                          int C, const float* s2, const float* t2, const float* mean2, const float* invstd2,
                          const float* dpooled, int P, double* red2, int CL, int PL) {
  __shared__ float s_red[2 * 256 * VEC];
  extern __shared__ unsigned char s_dyn[];
  const PoolTabs tabs = build_pool_tabs(s_dyn, H, W, P);
  const int cl = threadIdx.x % CL, pl = threadIdx.x / CL;
  const int c_base = blockIdx.y * CL * VEC;
  const int c = c_base + cl * VEC;
  const unsigned M = static_cast<unsigned>(B) * H * W;
  float acc[2][VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { acc[0][v] = 0.f; acc[1][v] = 0.f; }
  if (pl < PL && c < C) {
    float sc[VEC], sh[VEC];
    ldf<VEC>(s2 + c, sc); ldf<VEC>(t2 + c, sh);
    // kPixInFlight pixels per iteration, every load issued before the first use: with 2 loads per pixel and 512 resident
    // threads per SM one pixel in flight leaves ~16 KB outstanding per SM, two 32 KB - the HBM latency-bandwidth
    // product needs ~45 KB (two pixels: 0.46 of copy bandwidth at level 1, ncu)
    const unsigned start = blockIdx.x * PL + pl, stride = gridDim.x * PL;
    PixIter it[kPixInFlight];
#pragma unroll
    for (int u = 0; u < kPixInFlight; ++u) it[u].init((start + u * stride) % M, kPixInFlight * stride, H, W);
    for (unsigned m32 = start; m32 < M; m32 += kPixInFlight * stride) {
      RawV<VEC, grad_t> draw[kPixInFlight];
      RawV<VEC, act_t> araw[kPixInFlight];
      bool ok[kPixInFlight];
#pragma unroll
      for (int u = 0; u < kPixInFlight; ++u) {
        const long long m = static_cast<long long>(m32) + static_cast<long long>(u) * stride;
        ok[u] = m < M;
        if (ok[u]) { draw[u] = ldraw<VEC>(dz + m * ld_dz + 2 * C + c); araw[u] = ldraw<VEC>(a0 + m * ld_a0 + c); }
      }
#pragma unroll
      for (int u = 0; u < kPixInFlight; ++u) {
        if (ok[u]) {
          float gp[VEC], da[VEC], av[VEC];
          poolT_gather<VEC>(dpooled, it[u].b, it[u].y, it[u].x, P, C, c, tabs, gp);
          cvtraw<VEC>(draw[u], da); cvtraw<VEC>(araw[u], av);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const float d2 = fmaf(av[v], sc[v], sh[v]) > 0.f ? da[v] + gp[v] : 0.f;
            acc[0][v] += d2;
            acc[1][v] = fmaf(d2, av[v], acc[1][v]);
          }
        }
        it[u].next(H, W);
      }
    }
  }
  flush_bn_partials<VEC>(acc, cl, pl, CL, PL, c_base, C, mean2, invstd2, red2, s_red);
}

// pass 3, A branch: dA0 from (dA + pool^T(dpooled), A0, BN2).  (The L branch, dL0 from (dL, L0, BN1), is a plain BatchNorm
// + ReLU backward and runs on bn_bwd_apply_kernel.)
template <int VEC, int OCC>
__global__ void __launch_bounds__(256, OCC)
branch_bwd_apply_kernel(const grad_t* __restrict__ dz, long long ld_dz, const act_t* __restrict__ a0, long long ld_a0, int B, int H,
                        int W, int C, const float* s2, const float* t2, const float* mean2, const float* invstd2,
                        const double* red2, const float* dpooled, int P, grad_t* __restrict__ da0, long long ld_da0, int CL,
                        int PL) {
  extern __shared__ unsigned char s_dyn[];
  const PoolTabs tabs = build_pool_tabs(s_dyn, H, W, P);
  const int cl = threadIdx.x % CL, pl = threadIdx.x / CL;
  const int c = (blockIdx.y * CL + cl) * VEC;
  if (pl >= PL || c >= C) return;
  const unsigned M = static_cast<unsigned>(B) * H * W;
  float sc[VEC], sh[VEC], p[VEC], q[VEC];
  bn_bwd_coeffs<VEC>(s2, mean2, invstd2, red2, C, c, 1.0 / static_cast<double>(M), sc, p, q);
  ldf<VEC>(t2 + c, sh);
  const grad_t* dsrc = dz + 2 * C + c;
  const unsigned start = blockIdx.x * PL + pl, stride = gridDim.x * PL;   // kPixInFlight pixels per iteration (see reduce2)
  PixIter it[kPixInFlight];
#pragma unroll
  for (int u = 0; u < kPixInFlight; ++u) it[u].init((start + u * stride) % M, kPixInFlight * stride, H, W);
  for (unsigned m32 = start; m32 < M; m32 += kPixInFlight * stride) {
    RawV<VEC, grad_t> draw[kPixInFlight];
    RawV<VEC, act_t> xraw[kPixInFlight];
    bool ok[kPixInFlight];
#pragma unroll
    for (int u = 0; u < kPixInFlight; ++u) {
      const long long m = static_cast<long long>(m32) + static_cast<long long>(u) * stride;
      ok[u] = m < M;
      if (ok[u]) { draw[u] = ldraw<VEC>(dsrc + m * ld_dz); xraw[u] = ldraw<VEC>(a0 + m * ld_a0 + c); }
    }
#pragma unroll
    for (int u = 0; u < kPixInFlight; ++u) {
      if (ok[u]) {
        const long long m = static_cast<long long>(m32) + static_cast<long long>(u) * stride;
        float gp[VEC], o[VEC], d[VEC], xv[VEC];
        poolT_gather<VEC>(dpooled, it[u].b, it[u].y, it[u].x, P, C, c, tabs, gp);
        cvtraw<VEC>(draw[u], d); cvtraw<VEC>(xraw[u], xv);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const float dd = fmaf(xv[v], sc[v], sh[v]) > 0.f ? d[v] + gp[v] : 0.f;
          o[v] = fmaf(sc[v], dd, fmaf(p[v], xv[v], q[v]));
        }
        stv<VEC>(da0 + m * ld_da0 + c, o);
      }
      it[u].next(H, W);
    }
  }
}

// Row-walk variant of the A-branch apply pass (see branch_act_rows_kernel): the pool^T(dpooled) vector of a pixel depends only
// on the pooling windows that contain it, so a thread walking a row segment re-gathers it only when the window set changes
// (about P times per row) instead of for every pixel.
template <int OCC>
__global__ void __launch_bounds__(256, OCC)
branch_bwd_apply_rows_kernel(const grad_t* __restrict__ dz, long long ld_dz, const act_t* __restrict__ a0, long long ld_a0, int B, int H,
                             int W, int C, const float* s2, const float* t2, const float* mean2, const float* invstd2,
                             const double* red2, const float* __restrict__ dpooled, int P, grad_t* __restrict__ da0, long long ld_da0,
                             int nseg) {
  constexpr int VEC = 8;
  extern __shared__ unsigned char s_dyn[];
  const PoolTabs tabs = build_pool_tabs(s_dyn, H, W, P);
  const int CV = C / VEC;
  const long long total = static_cast<long long>(B) * H * nseg * CV;
  const double invn = 1.0 / (static_cast<double>(B) * H * W);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int cv, k, y; long long b;
    decode4(i, total < (1LL << 31), CV, nseg, H, cv, k, y, b);
    const int c = cv * VEC;
    const int xlo = static_cast<int>(static_cast<long long>(k) * W / nseg), xhi = static_cast<int>(static_cast<long long>(k + 1) * W / nseg);
    float sc[VEC], sh[VEC], p[VEC], q[VEC];
    bn_bwd_coeffs<VEC>(s2, mean2, invstd2, red2, C, c, invn, sc, p, q);
    ldf<VEC>(t2 + c, sh);
    const int ty = tabs.yt[y];
    const int ilo = ty & 0xffff, ihi = ty >> 16;
    const float* dpb = dpooled + b * P * P * C + c;
    const long long mrow = (b * H + y) * W;
    const grad_t* dsrc = dz + 2 * C + c;
    float gp[VEC];
    int cur_key = -1;
    for (int x = xlo; x < xhi; x += 4) {
      RawV<VEC, grad_t> draw[4];
      RawV<VEC, act_t> xraw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long m = mrow + min(x + u, xhi - 1);
        draw[u] = ldraw<VEC>(dsrc + m * ld_dz); xraw[u] = ldraw<VEC>(a0 + m * ld_a0 + c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int xx = x + u;
        if (xx < xhi) {
          const int key = tabs.xt[xx];
          if (key != cur_key) {        // the set of pooling windows over this pixel changed: gather their gradients once
            const int jlo = key & 0xffff, jhi = key >> 16;
#pragma unroll
            for (int v = 0; v < VEC; ++v) gp[v] = 0.f;
            for (int wi = ilo; wi <= ihi; ++wi)
              for (int wj = jlo; wj <= jhi; ++wj) {
                const float wgt = tabs.wy[wi] * tabs.wx[wj];
                float t[VEC]; ldf<VEC>(dpb + (wi * P + wj) * C, t);
#pragma unroll
                for (int v = 0; v < VEC; ++v) gp[v] = fmaf(wgt, t[v], gp[v]);
              }
            cur_key = key;
          }
          float d[VEC], xv[VEC], ov[VEC];
          cvtraw<VEC>(draw[u], d); cvtraw<VEC>(xraw[u], xv);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const float dd = fmaf(xv[v], sc[v], sh[v]) > 0.f ? d[v] + gp[v] : 0.f;
            ov[v] = fmaf(sc[v], dd, fmaf(p[v], xv[v], q[v]));
          }
          stv<VEC>(da0 + (mrow + xx) * ld_da0 + c, ov);
        }
      }
    }
  }
}

// red += (sum d m, sum d m xhat) with m = [bn(x) > 0]: the BatchNorm + ReLU backward reduction over a plain (dy, x) pair -
// what is left of branch_bwd_reduce2 once the pool^T(dpooled) part comes from the window means (no gather per pixel)
template <int VEC, int OCC>
__global__ void __launch_bounds__(256, OCC)
bn_bwd_reduce_kernel(const grad_t* __restrict__ dy, long long ld_dy, const act_t* __restrict__ x, long long ld_x, long long M, int C,
                     const float* scale, const float* shift, const float* mean, const float* invstd, double* red, int CL, int PL) {
  __shared__ float s_red[2 * 256 * VEC];
  const int cl = threadIdx.x % CL, pl = threadIdx.x / CL;
  const int c_base = blockIdx.y * CL * VEC;
  const int c = c_base + cl * VEC;
  float acc[2][VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { acc[0][v] = 0.f; acc[1][v] = 0.f; }
  if (pl < PL && c < C) {
    float sc[VEC], sh[VEC];
    ldf<VEC>(scale + c, sc); ldf<VEC>(shift + c, sh);
    const long long stride = static_cast<long long>(gridDim.x) * PL;
    for (long long m0 = static_cast<long long>(blockIdx.x) * PL + pl; m0 < M; m0 += kPixInFlight * stride) {
      RawV<VEC, grad_t> draw[kPixInFlight];
      RawV<VEC, act_t> xraw[kPixInFlight];
      bool ok[kPixInFlight];
#pragma unroll
      for (int u = 0; u < kPixInFlight; ++u) {
        const long long m = m0 + u * stride;
        ok[u] = m < M;
        if (ok[u]) { draw[u] = ldraw<VEC>(dy + m * ld_dy + c); xraw[u] = ldraw<VEC>(x + m * ld_x + c); }
      }
#pragma unroll
      for (int u = 0; u < kPixInFlight; ++u) {
        if (ok[u]) {
          float d[VEC], xv[VEC];
          cvtraw<VEC>(draw[u], d); cvtraw<VEC>(xraw[u], xv);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const float dm = fmaf(xv[v], sc[v], sh[v]) > 0.f ? d[v] : 0.f;
            acc[0][v] += dm;
            acc[1][v] = fmaf(dm, xv[v], acc[1][v]);
          }
        }
      }
    }
  }
  flush_bn_partials<VEC>(acc, cl, pl, CL, PL, c_base, C, mean, invstd, red, s_red);
}

// red[c] += sum_{b,w} dp m_mean;  red[C + c] += invstd (sum dp ma_mean - mean sum dp m_mean)   over the [B, P, P] windows
__global__ void __launch_bounds__(256)
pool_window_terms_kernel(const float* __restrict__ dp, const float* __restrict__ means, long long BW, int C, const float* mean,
                         const float* invstd, double* red) {
  // block: 32 channels x 8 window lanes
  const int cl = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + cl;
  float s0 = 0.f, s1 = 0.f;
  if (c < C) {
    const float* mm = means;                       // mean_w(m)
    const float* ma = means + BW * C;              // mean_w(m a0)
    for (long long w = static_cast<long long>(blockIdx.x) * 8 + wl; w < BW; w += static_cast<long long>(gridDim.x) * 8) {
      const float d = dp[w * C + c];
      s0 = fmaf(d, mm[w * C + c], s0);
      s1 = fmaf(d, ma[w * C + c], s1);
    }
  }
  __shared__ float s[2][8][33];
  s[0][wl][cl] = s0; s[1][wl][cl] = s1;
  __syncthreads();
  if (wl == 0 && c < C) {
    float t0 = 0.f, t1 = 0.f;
    for (int p = 0; p < 8; ++p) { t0 += s[0][p][cl]; t1 += s[1][p][cl]; }
    atomicAdd(red + c, static_cast<double>(t0));
    atomicAdd(red + C + c, static_cast<double>(invstd[c]) * (static_cast<double>(t1) - static_cast<double>(mean[c]) * static_cast<double>(t0)));
  }
}

// =============================================================================================
// layout / misc
// =============================================================================================
__global__ void nchw_to_nhwc_kernel(const float* src, void* dst, int ddt, long long ld, int B, int C, int H, int W) {
  const long long total = static_cast<long long>(B) * H * W * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long m = i / C;
    const long long hw = m % (static_cast<long long>(H) * W);
    const long long b = m / (static_cast<long long>(H) * W);
    const float v = src[(b * C + c) * H * W + hw];
    if (ddt == DFCSA_F32) reinterpret_cast<float*>(dst)[m * ld + c] = v;
    else if (ddt == DFCSA_F16) reinterpret_cast<__half*>(dst)[m * ld + c] = Cvt<__half>::from_f(v);
    else reinterpret_cast<__nv_bfloat16*>(dst)[m * ld + c] = __float2bfloat16_rn(v);
  }
}
__global__ void nhwc_to_nchw_kernel(const void* src, int sdt, long long ld, float* dst, int B, int C, int H, int W) {
  const long long total = static_cast<long long>(B) * H * W * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long hw = i % (static_cast<long long>(H) * W);
    const long long r = i / (static_cast<long long>(H) * W);
    const int c = static_cast<int>(r % C);
    const long long b = r / C;
    const long long m = b * H * W + hw;
    float v;
    if (sdt == DFCSA_F32) v = reinterpret_cast<const float*>(src)[m * ld + c];
    else if (sdt == DFCSA_F16) v = __half2float(reinterpret_cast<const __half*>(src)[m * ld + c]);
    else v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[m * ld + c]);
    dst[i] = v;
  }
}
__global__ void __launch_bounds__(256)
colsum_kernel(const void* x, int dt, long long ld, long long M, int C, float* out) {
  // block: 32 channels x 8 pixel lanes
  const int cl = threadIdx.x % 32, pl = threadIdx.x / 32;
  const int c = blockIdx.y * 32 + cl;
  float acc = 0.f;
  if (c < C) {
    for (long long m = static_cast<long long>(blockIdx.x) * 8 + pl; m < M; m += static_cast<long long>(gridDim.x) * 8) {
      if (dt == DFCSA_F32) acc += reinterpret_cast<const float*>(x)[m * ld + c];
      else if (dt == DFCSA_F16) acc += __half2float(reinterpret_cast<const __half*>(x)[m * ld + c]);
      else acc += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[m * ld + c]);
    }
  }
  __shared__ float s[8][33];
  s[pl][cl] = acc;
  __syncthreads();
  if (pl == 0 && c < C) {
    float t = 0.f;
    for (int p = 0; p < 8; ++p) t += s[p][cl];
    atomicAdd(out + c, t);
  }
}
// 16-bit inputs with C % 8 == 0: 16-byte loads, the (channel-lane, pixel-lane) layout of the other streaming kernels
template <typename T>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T* x, long long ld, long long M, int C, float* out, int CL, int PL) {
  __shared__ float s_red[256 * 8];
  const int cl = threadIdx.x % CL, pl = threadIdx.x / CL;
  const int c_base = blockIdx.y * CL * 8;
  const int c = c_base + cl * 8;
  float acc[8];
#pragma unroll
  for (int v = 0; v < 8; ++v) acc[v] = 0.f;
  if (pl < PL && c < C) {
    for (long long m = static_cast<long long>(blockIdx.x) * PL + pl; m < M; m += static_cast<long long>(gridDim.x) * PL) {
      float t[8];
      load8<T>(x + m * ld + c, t);
#pragma unroll
      for (int v = 0; v < 8; ++v) acc[v] += t[v];
    }
  }
  const int row = CL * 8;
  if (cl < CL && pl < PL) {
#pragma unroll
    for (int v = 0; v < 8; ++v) s_red[pl * row + cl * 8 + v] = acc[v];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < row; idx += blockDim.x) {
    if (c_base + idx < C) {
      float t = 0.f;
      for (int p = 0; p < PL; ++p) t += s_red[p * row + idx];
      atomicAdd(out + c_base + idx, t);
    }
  }
}
__global__ void cast2d_kernel(const void* x, int xdt, long long ld_x, void* y, int ydt, long long ld_y, long long M, int C) {
  const long long total = M * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long m = i / C;
    float v;
    if (xdt == DFCSA_F32) v = reinterpret_cast<const float*>(x)[m * ld_x + c];
    else if (xdt == DFCSA_F16) v = __half2float(reinterpret_cast<const __half*>(x)[m * ld_x + c]);
    else v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[m * ld_x + c]);
    if (ydt == DFCSA_F32) reinterpret_cast<float*>(y)[m * ld_y + c] = v;
    else if (ydt == DFCSA_F16) reinterpret_cast<__half*>(y)[m * ld_y + c] = Cvt<__half>::from_f(v);
    else reinterpret_cast<__nv_bfloat16*>(y)[m * ld_y + c] = __float2bfloat16_rn(v);
  }
}

// 8 elements per thread (16-byte accesses on the 16-bit side); C % 8 == 0, pitches % 8 == 0, 16-byte aligned bases
template <typename TX, typename TY>
__global__ void __launch_bounds__(256) cast2d_vec_kernel(const TX* __restrict__ x, long long ld_x, TY* __restrict__ y, long long ld_y,
                                                         long long M, int C8) {
  const long long total = M * C8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long m; int c;
    if (total < (1LL << 31)) {
      const unsigned ii = static_cast<unsigned>(i);
      m = ii / static_cast<unsigned>(C8); c = static_cast<int>(ii - static_cast<unsigned>(m) * C8) * 8;
    } else {
      m = i / C8; c = static_cast<int>(i - m * C8) * 8;
    }
    float t[8];
    load8<TX>(x + m * ld_x + c, t);
    store8<TY>(y + m * ld_y + c, t);
  }
}

static int ew_blocks(long long total) {
  return static_cast<int>(std::max<long long>(1, std::min<long long>((total + 255) / 256, 148LL * 32)));
}
static bool bout_fwd_pre() {
  static const bool on = [] { const char* e = getenv("DFCSA_BOUT_FWD_PRE"); return !e || atoi(e) != 0; }();
  return on;
}
// cols_reduce_par_kernel: C / 4 channel vectors must divide the CTA, and the rows of a window must be worth splitting
static bool cols_par_ok(int C, int H, int P) {
  static const bool on = [] { const char* e = getenv("DFCSA_COLS_PAR"); return !e || atoi(e) != 0; }();
  const int cv = C / 4;
  return on && C % 4 == 0 && cv <= 256 && 256 % cv == 0 && H / P >= 8;
}
static int cols_par_blocks(int B, int P) {
  return static_cast<int>(std::min<long long>(static_cast<long long>(B) * P * P, 148LL * 8));
}

}  // namespace
}  // namespace dfcsa

using namespace dfcsa;
#define ST static_cast<cudaStream_t>(stream)
#define A_(p) reinterpret_cast<const act_t*>(p)
#define AM_(p) reinterpret_cast<act_t*>(p)
#define G_(p) reinterpret_cast<const grad_t*>(p)
#define GM_(p) reinterpret_cast<grad_t*>(p)
// dispatch on the vector width
#define VEC_DISPATCH(ok, ...)            \
  do { if (ok) { constexpr int VEC = 8; __VA_ARGS__; } else { constexpr int VEC = 1; __VA_ARGS__; } } while (0)

extern "C" int dfcsa_bn_finalize(const double* sum, const double* sumsq, int64_t count, int32_t C, const float* gamma, const float* beta,
                                 const float* conv_bias, float* running_mean, float* running_var, float momentum,
                                 float eps, float* scale, float* shift, float* mean, float* invstd, void* stream) {
  DFCSA_CHECK_ARG(sum && sumsq && gamma && beta && scale && shift && mean && invstd && C > 0 && count > 0, "dfcsa_bn_finalize: bad args");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, ST>>>(sum, sumsq, count, C, gamma, beta, conv_bias, running_mean, running_var,
                                                      momentum, eps, scale, shift, mean, invstd);
  DFCSA_LAUNCH_CHECK("bn_finalize_kernel");
  return DFCSA_OK;
}
extern "C" int dfcsa_bn_eval_affine(int32_t C, const float* gamma, const float* beta, const float* conv_bias,
                                    const float* running_mean, const float* running_var, float eps, float* scale,
                                    float* shift, void* stream) {
  DFCSA_CHECK_ARG(gamma && beta && running_mean && running_var && scale && shift && C > 0, "dfcsa_bn_eval_affine: bad args");
  bn_eval_affine_kernel<<<(C + 127) / 128, 128, 0, ST>>>(C, gamma, beta, conv_bias, running_mean, running_var, eps, scale, shift);
  DFCSA_LAUNCH_CHECK("bn_eval_affine_kernel");
  return DFCSA_OK;
}
extern "C" int dfcsa_bn_param_grads(const double* red, int32_t C, float* dgamma, float* dbeta, void* stream) {
  DFCSA_CHECK_ARG(red && dgamma && dbeta && C > 0, "dfcsa_bn_param_grads: bad args");
  bn_param_grads_kernel<<<(C + 127) / 128, 128, 0, ST>>>(red, C, dgamma, dbeta);
  DFCSA_LAUNCH_CHECK("bn_param_grads_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_block_param_grads(const double* red, int32_t C, float* dg1, float* db1, float* dg2, float* db2, float* dg3,
                                       float* db3, float* dg4, float* db4, float* drs, float* dgam, void* stream) {
  DFCSA_CHECK_ARG(red && C > 0, "dfcsa_block_param_grads: bad args");
  BlockGradPtrs o;
  o.g[0] = dg1; o.b[0] = db1; o.g[1] = dg2; o.b[1] = db2; o.g[2] = dg3; o.b[2] = db3; o.g[3] = dg4; o.b[3] = db4;
  o.drs = drs; o.dgam = dgam;
  block_param_grads_kernel<<<(C + 127) / 128, 128, 0, ST>>>(red, C, o);
  DFCSA_LAUNCH_CHECK("block_param_grads_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_bnrelu_pool_fwd(const void* a0, int64_t ld, int32_t B, int32_t H, int32_t W, int32_t C,
                                     const float* scale, const float* shift, int32_t P, float* tmp, float* pooled,
                                     int32_t with_masks, void* stream) {
  DFCSA_CHECK_ARG(a0 && scale && shift && tmp && pooled && B > 0 && H > 0 && W > 0 && C > 0 && P > 0, "dfcsa_bnrelu_pool_fwd: bad args");
  const bool v8 = vec8_ok(C, {ld}, {a0, scale, shift, tmp});
  const long long total = static_cast<long long>(B) * H * P * (v8 ? C / 8 : C);
  if (with_masks)
    VEC_DISPATCH(v8, (pool_rows_kernel<VEC, true><<<ew_blocks(total), 256, 0, ST>>>(A_(a0), ld, B, H, W, C, scale, shift, P, tmp)));
  else
    VEC_DISPATCH(v8, (pool_rows_kernel<VEC, false><<<ew_blocks(total), 256, 0, ST>>>(A_(a0), ld, B, H, W, C, scale, shift, P, tmp)));
  DFCSA_LAUNCH_CHECK("pool_rows_kernel");
  // with_masks: tmp / pooled hold three planes ([3][B, ...]); the column pass sees them as 3 B images
  const int Bc = with_masks ? 3 * B : B;
  if (cols_par_ok(C, H, P) && ((reinterpret_cast<uintptr_t>(tmp) | reinterpret_cast<uintptr_t>(pooled)) & 15) == 0)
    cols_reduce_par_kernel<4><<<cols_par_blocks(Bc, P), 256, 0, ST>>>(tmp, Bc, H, P, C, 0, nullptr, pooled, nullptr, nullptr);
  else if (C % 4 == 0 && ((reinterpret_cast<uintptr_t>(tmp) | reinterpret_cast<uintptr_t>(pooled)) & 15) == 0)
    cols_reduce_kernel<4><<<ew_blocks(static_cast<long long>(Bc) * P * P * (C / 4)), 256, 0, ST>>>(tmp, Bc, H, P, C, 0, nullptr, pooled, nullptr, nullptr);
  else
    cols_reduce_kernel<1><<<ew_blocks(static_cast<long long>(Bc) * P * P * C), 256, 0, ST>>>(tmp, Bc, H, P, C, 0, nullptr, pooled, nullptr, nullptr);
  DFCSA_LAUNCH_CHECK("cols_reduce_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_bn_bwd_reduce(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, int64_t M, int32_t C,
                                   const float* scale, const float* shift, const float* mean, const float* invstd, double* red,
                                   void* stream) {
  DFCSA_CHECK_ARG(dy && x && scale && shift && mean && invstd && red && M > 0 && C > 0, "dfcsa_bn_bwd_reduce: bad args");
  const bool v8 = vec8_ok(C, {ld_dy, ld_x}, {dy, x, scale, shift, mean, invstd});
  const RedGeom g = red_geom(C, v8 ? 8 : 1);
  dim3 grid(red_blocks(M, g.PL, g.chunks, ew_occ(3)), g.chunks);
  OCC_DISPATCH(3, VEC_DISPATCH(v8, (bn_bwd_reduce_kernel<VEC, OCC><<<grid, 256, 0, ST>>>(G_(dy), ld_dy, A_(x), ld_x, M, C, scale, shift, mean, invstd,
                                                                                        red, g.CL, g.PL))));
  DFCSA_LAUNCH_CHECK("bn_bwd_reduce_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_pool_window_terms(const float* dpooled, const float* means, int32_t B, int32_t P, int32_t C, const float* mean,
                                       const float* invstd, double* red, void* stream) {
  DFCSA_CHECK_ARG(dpooled && means && mean && invstd && red && B > 0 && P > 0 && C > 0, "dfcsa_pool_window_terms: bad args");
  const long long BW = static_cast<long long>(B) * P * P;
  dim3 grid(static_cast<unsigned>(std::max<long long>(1, std::min<long long>((BW + 63) / 64, 148 * 2))), (C + 31) / 32);
  pool_window_terms_kernel<<<grid, 256, 0, ST>>>(dpooled, means, BW, C, mean, invstd, red);
  DFCSA_LAUNCH_CHECK("pool_window_terms_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_branch_act_fwd(const void* l0, int64_t ld_l0, const void* a0, int64_t ld_a0, int32_t B, int32_t H,
                                    int32_t W, int32_t C, const float* scale1, const float* shift1, const float* scale2,
                                    const float* shift2, const float* o, int32_t P, const float* gamma, void* z,
                                    int64_t ld_z, void* zb, int64_t ld_zb, void* stream) {
  DFCSA_CHECK_ARG(a0 && (l0 == nullptr || (scale1 && shift1)) && scale2 && shift2 && o && gamma && z, "dfcsa_branch_act_fwd: null pointer");
  const bool v8 = vec8_ok(C, {l0 ? ld_l0 : 0, ld_a0, ld_z, zb ? ld_zb : 0}, {l0, a0, z, zb, o, scale1, shift1, scale2, shift2});
  DFCSA_CHECK_ARG(static_cast<long long>(B) * H * W < (1LL << 31), "dfcsa_branch_act_fwd: too many pixels");
  static const int act_occ = [] { const char* e = getenv("DFCSA_ACT_OCC"); return e ? atoi(e) : 2; }();
  static const bool rows_ok = [] { const char* e = getenv("DFCSA_ACT_ROWS"); return !e || atoi(e) != 0; }();
  if (rows_ok && v8 && zb == nullptr && W / P >= 8) {
    // row segments: at least P per row (one per pooled cell), more when the batch is small, never shorter than 8 pixels
    const long long per_seg = static_cast<long long>(B) * H * (C / 8);
    int nseg = static_cast<int>(std::min<long long>(W / 8, std::max<long long>(P, (150000 + per_seg - 1) / per_seg)));
    const long long total = per_seg * nseg;
    const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>((total + 255) / 256, static_cast<long long>(num_sms()) * 2 * 8)));
    branch_act_rows_kernel<2><<<blocks, 256, 0, ST>>>(A_(l0), ld_l0, A_(a0), ld_a0, B, H, W, C, scale1, shift1, scale2, shift2, o, P, gamma,
                                                      AM_(z), ld_z, nseg);
    DFCSA_LAUNCH_CHECK("branch_act_rows_kernel");
    return DFCSA_OK;
  }
  const RedGeom g = red_geom(C, v8 ? 8 : 1);
  dim3 grid(red_blocks(static_cast<long long>(B) * H * W, g.PL, g.chunks, ew_occ(act_occ)), g.chunks);
  OCC_DISPATCH(act_occ, VEC_DISPATCH(v8, (branch_act_fwd_kernel<VEC, OCC><<<grid, 256, 0, ST>>>(A_(l0), ld_l0, A_(a0), ld_a0, B, H, W, C, scale1, shift1, scale2,
                                                                     shift2, o, P, gamma, AM_(z), ld_z, GM_(zb), ld_zb, g.CL, g.PL))));
  DFCSA_LAUNCH_CHECK("branch_act_fwd_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_gate_mix_fwd(const void* g0, int64_t ld_g0, int64_t M, int32_t C, const float* scale3,
                                  const float* shift3, void* z, int64_t ld_z, void* zb, int64_t ld_zb, void* stream) {
  DFCSA_CHECK_ARG(g0 && scale3 && shift3 && z && M > 0 && C > 0, "dfcsa_gate_mix_fwd: bad args");
  const bool v8 = vec8_ok(C, {ld_g0, ld_z, zb ? ld_zb : 0}, {g0, z, zb, scale3, shift3});
  const RedGeom g = red_geom(C, v8 ? 8 : 1);
  dim3 grid(red_blocks(M, g.PL, g.chunks, 4), g.chunks);
  VEC_DISPATCH(v8, (gate_mix_fwd_kernel<VEC><<<grid, 256, 0, ST>>>(A_(g0), ld_g0, M, C, scale3, shift3, AM_(z), ld_z, GM_(zb), ld_zb,
                                                                   g.CL, g.PL)));
  DFCSA_LAUNCH_CHECK("gate_mix_fwd_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_sum_out_fwd(const void* a, int64_t ld_a, const void* b, int64_t ld_b, const void* r, int64_t ld_r, int32_t B,
                                 int32_t H, int32_t W, int32_t C, const float* res_scale, void* y, int64_t ld_y, void* yp,
                                 int64_t ld_yp, void* stream) {
  DFCSA_CHECK_ARG(a && r && res_scale && y, "dfcsa_sum_out_fwd: null pointer");
  const bool v8 = vec8_ok(C, {ld_a, b ? ld_b : 0, ld_r, ld_y, yp ? ld_yp : 0}, {a, b, r, y, yp});
  const long long nwin = static_cast<long long>(B) * ((H + 1) / 2) * ((W + 1) / 2);
  DFCSA_CHECK_ARG(nwin < (1LL << 31), "dfcsa_sum_out_fwd: too many pixels");
  const RedGeom g = red_geom(C, v8 ? 8 : 1);
  if (bout_fwd_pre() && v8) {
    dim3 grid3(red_blocks(nwin, g.PL, g.chunks, 3), g.chunks);
    block_out_fwd_kernel<8, true, true><<<grid3, 256, 0, ST>>>(A_(a), ld_a, A_(r), ld_r, B, H, W, C, nullptr, nullptr, res_scale,
                                                               AM_(y), ld_y, AM_(yp), ld_yp, nullptr, 0, nullptr, 0,
                                                               g.CL, g.PL, A_(b), ld_b);
    DFCSA_LAUNCH_CHECK("block_out_fwd_kernel<plain, pre>");
    return DFCSA_OK;
  }
  dim3 grid(red_blocks(nwin, g.PL, g.chunks, 4), g.chunks);
  VEC_DISPATCH(v8, (block_out_fwd_kernel<VEC, true><<<grid, 256, 0, ST>>>(A_(a), ld_a, A_(r), ld_r, B, H, W, C, nullptr, nullptr, res_scale,
                                                                          AM_(y), ld_y, AM_(yp), ld_yp, nullptr, 0, nullptr, 0,
                                                                          g.CL, g.PL, A_(b), ld_b)));
  DFCSA_LAUNCH_CHECK("block_out_fwd_kernel<plain>");
  return DFCSA_OK;
}

extern "C" int dfcsa_block_out_fwd(const void* f0, int64_t ld_f0, const void* r, int64_t ld_r, int32_t B, int32_t H,
                                   int32_t W, int32_t C, const float* scale4, const float* shift4, const float* res_scale,
                                   void* y, int64_t ld_y, void* yp, int64_t ld_yp, void* yb, int64_t ld_yb, void* ypb,
                                   int64_t ld_ypb, void* stream) {
  DFCSA_CHECK_ARG(f0 && r && scale4 && shift4 && res_scale && y, "dfcsa_block_out_fwd: null pointer");
  const bool v8 = vec8_ok(C, {ld_f0, ld_r, ld_y, yp ? ld_yp : 0, yb ? ld_yb : 0, ypb ? ld_ypb : 0},
                          {f0, r, y, yp, yb, ypb, scale4, shift4});
  const long long nwin = static_cast<long long>(B) * ((H + 1) / 2) * ((W + 1) / 2);
  DFCSA_CHECK_ARG(nwin < (1LL << 31), "dfcsa_block_out_fwd: too many pixels");
  const RedGeom g = red_geom(C, v8 ? 8 : 1);
  if (bout_fwd_pre() && v8) {
    dim3 grid3(red_blocks(nwin, g.PL, g.chunks, 3), g.chunks);
    block_out_fwd_kernel<8, false, true><<<grid3, 256, 0, ST>>>(A_(f0), ld_f0, A_(r), ld_r, B, H, W, C, scale4, shift4, res_scale,
                                                                AM_(y), ld_y, AM_(yp), ld_yp, GM_(yb), ld_yb, GM_(ypb), ld_ypb, g.CL, g.PL);
    DFCSA_LAUNCH_CHECK("block_out_fwd_kernel<pre>");
    return DFCSA_OK;
  }
  dim3 grid(red_blocks(nwin, g.PL, g.chunks, 4), g.chunks);
  VEC_DISPATCH(v8, (block_out_fwd_kernel<VEC><<<grid, 256, 0, ST>>>(A_(f0), ld_f0, A_(r), ld_r, B, H, W, C, scale4, shift4, res_scale,
                                                                    AM_(y), ld_y, AM_(yp), ld_yp, GM_(yb), ld_yb, GM_(ypb), ld_ypb,
                                                                    g.CL, g.PL)));
  DFCSA_LAUNCH_CHECK("block_out_fwd_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_block_out_bwd_reduce(const void* dskip, int64_t ld_dskip, const void* dyp, int64_t ld_dyp,
                                          const void* y, int64_t ld_y, const void* f0, int64_t ld_f0, const void* r,
                                          int64_t ld_r, int32_t B, int32_t H, int32_t W, int32_t C, const float* scale4,
                                          const float* shift4, const float* mean4, const float* invstd4, void* dy_out,
                                          int64_t ld_dy, double* red4, double* drs, void* stream) {
  DFCSA_CHECK_ARG((dskip || dyp) && y && f0 && r && scale4 && shift4 && mean4 && invstd4 && red4 && drs, "dfcsa_block_out_bwd_reduce: null pointer");
  DFCSA_CHECK_ARG(dy_out != nullptr || dyp == nullptr, "dfcsa_block_out_bwd_reduce: dy_out required when dyp is given");
  const bool v8 = vec8_ok(C, {dskip ? ld_dskip : 0, dyp ? ld_dyp : 0, ld_y, ld_f0, ld_r, dy_out ? ld_dy : 0},
                          {dskip, dyp, y, f0, r, dy_out, scale4, shift4, mean4, invstd4});
  const RedGeom g = red_geom(C, v8 ? 8 : 1);
  const long long nwin = static_cast<long long>(B) * ((H + 1) / 2) * ((W + 1) / 2);
  static const bool pre_ok = [] { const char* e = getenv("DFCSA_BOUT_PRE"); return !e || atoi(e) != 0; }();
  if (pre_ok && v8 && dyp != nullptr && H >= 2 && W >= 2) {
    dim3 grid2(red_blocks(nwin, g.PL, g.chunks, 2), g.chunks);
    block_out_bwd_reduce_pre_kernel<8><<<grid2, 256, 0, ST>>>(G_(dskip), ld_dskip, G_(dyp), ld_dyp, A_(y), ld_y, A_(f0), ld_f0, A_(r), ld_r,
                                                              B, H, W, C, scale4, shift4, mean4, invstd4, GM_(dy_out), ld_dy, red4, drs,
                                                              g.CL, g.PL);
    DFCSA_LAUNCH_CHECK("block_out_bwd_reduce_pre_kernel");
    return DFCSA_OK;
  }
  dim3 grid(red_blocks(nwin, g.PL, g.chunks, ew_occ(4)), g.chunks);
  OCC_DISPATCH(4, VEC_DISPATCH(v8, (block_out_bwd_reduce_kernel<VEC, OCC><<<grid, 256, 0, ST>>>(G_(dskip), ld_dskip, G_(dyp), ld_dyp, A_(y), ld_y, A_(f0), ld_f0,
                                                                           A_(r), ld_r, B, H, W, C, scale4, shift4, mean4, invstd4,
                                                                           GM_(dy_out), ld_dy, red4, drs, g.CL, g.PL))));
  DFCSA_LAUNCH_CHECK("block_out_bwd_reduce_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_bn_bwd_apply(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, int64_t M, int32_t C,
                                  const float* scale, const float* shift, const float* mean, const float* invstd,
                                  const float* gamma, const double* red, int32_t act_mode, void* dx, int64_t ld_dx,
                                  void* stream) {
  (void)gamma;
  DFCSA_CHECK_ARG(dy && x && scale && shift && mean && invstd && red && dx && M > 0, "dfcsa_bn_bwd_apply: bad args");
  const bool v8 = vec8_ok(C, {ld_dy, ld_x, ld_dx}, {dy, x, dx, scale, shift, mean, invstd});
  const RedGeom g = red_geom(C, v8 ? 8 : 1);
  dim3 grid(red_blocks(M, g.PL, g.chunks, 4), g.chunks);
  VEC_DISPATCH(v8, (bn_bwd_apply_kernel<VEC><<<grid, 256, 0, ST>>>(G_(dy), ld_dy, A_(x), ld_x, M, C, scale, shift, mean, invstd, red,
                                                                   act_mode, GM_(dx), ld_dx, g.CL, g.PL)));
  DFCSA_LAUNCH_CHECK("bn_bwd_apply_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_gate_mix_bwd_reduce(const void* dz, int64_t ld_dz, const void* z, int64_t ld_z, const void* g0,
                                         int64_t ld_g0, int64_t M, int32_t C, const float* scale3, const float* shift3,
                                         const float* mean3, const float* invstd3, double* red3, void* stream) {
  DFCSA_CHECK_ARG(dz && z && g0 && scale3 && shift3 && mean3 && invstd3 && red3 && M > 0, "dfcsa_gate_mix_bwd_reduce: bad args");
  const bool v8 = vec8_ok(C, {ld_dz, ld_z, ld_g0}, {dz, z, g0, scale3, shift3, mean3, invstd3});
  const RedGeom g = red_geom(C, v8 ? 8 : 1);
  dim3 grid(red_blocks(M, g.PL, g.chunks, ew_occ(2)), g.chunks);      // two pixels in flight: 106 registers, no spills at 2 blocks / SM
  OCC_DISPATCH(2, VEC_DISPATCH(v8, (gate_mix_bwd_reduce_kernel<VEC, OCC><<<grid, 256, 0, ST>>>(G_(dz), ld_dz, A_(z), ld_z, A_(g0), ld_g0, M, C, scale3, shift3,
                                                                          mean3, invstd3, red3, g.CL, g.PL))));
  DFCSA_LAUNCH_CHECK("gate_mix_bwd_reduce_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_gate_mix_bwd_apply(const void* dz, int64_t ld_dz, const void* z, int64_t ld_z, const void* g0,
                                        int64_t ld_g0, int64_t M, int32_t C, const float* scale3, const float* shift3,
                                        const float* mean3, const float* invstd3, const float* gamma3, const double* red3,
                                        void* dg0, int64_t ld_dg0, void* stream) {
  (void)gamma3;
  DFCSA_CHECK_ARG(dz && z && g0 && scale3 && shift3 && mean3 && invstd3 && red3 && dg0 && M > 0, "dfcsa_gate_mix_bwd_apply: bad args");
  const bool v8 = vec8_ok(C, {ld_dz, ld_z, ld_g0, ld_dg0}, {dz, z, g0, dg0, scale3, shift3, mean3, invstd3});
  const RedGeom g = red_geom(C, v8 ? 8 : 1);
  dim3 grid(red_blocks(M, g.PL, g.chunks, ew_occ(2)), g.chunks);      // measured: 0.98 of copy bandwidth at 2 blocks / SM, 0.91 at 3
  OCC_DISPATCH(2, VEC_DISPATCH(v8, (gate_mix_bwd_apply_kernel<VEC, OCC><<<grid, 256, 0, ST>>>(G_(dz), ld_dz, A_(z), ld_z, A_(g0), ld_g0, M, C, scale3, shift3,
                                                                         mean3, invstd3, red3, GM_(dg0), ld_dg0, g.CL, g.PL))));
  DFCSA_LAUNCH_CHECK("gate_mix_bwd_apply_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_branch_bwd_reduce1(void* dz, int64_t ld_dz, const void* l0, int64_t ld_l0, const void* g0, int64_t ld_g0,
                                        int32_t B, int32_t H, int32_t W, int32_t C, const float* scale1, const float* shift1,
                                        const float* mean1, const float* invstd1, const float* scale3, const float* shift3,
                                        const float* o, int32_t P, const float* gamma, double* red1, double* dgamma, float* tmp,
                                        float* d_o, void* stream) {
  DFCSA_CHECK_ARG(dz && l0 && scale1 && shift1 && mean1 && invstd1 && o && gamma && red1 && dgamma && tmp && d_o &&
                  (g0 == nullptr || (scale3 && shift3)), "dfcsa_branch_bwd_reduce1: null pointer");
  DFCSA_CHECK_ARG(static_cast<long long>(B) * H * W < (1LL << 31), "dfcsa_branch_bwd_reduce1: too many pixels");
  const bool v8 = vec8_ok(C, {ld_dz, ld_l0, ld_g0}, {dz, l0, g0, o, tmp, scale1, shift1, mean1, invstd1, scale3, shift3});
  const RedGeom g = red_geom(C, v8 ? 8 : 1);
  const long long M = static_cast<long long>(B) * H * W;
  dim3 grid(red_blocks(M, g.PL, g.chunks, ew_occ(2)), g.chunks);
  static const bool two_pix = [] { const char* e = getenv("DFCSA_RED1_PIX"); return !e || atoi(e) != 1; }();
  if (two_pix)
    OCC_DISPATCH(2, VEC_DISPATCH(v8, (branch_bwd_reduce1_kernel<VEC, OCC, 2><<<grid, 256, 0, ST>>>(GM_(dz), ld_dz, A_(l0), ld_l0, A_(g0), ld_g0, M, C, scale1,
                                                                              shift1, mean1, invstd1, scale3, shift3, red1, g.CL, g.PL))));
  else
    OCC_DISPATCH(2, VEC_DISPATCH(v8, (branch_bwd_reduce1_kernel<VEC, OCC, 1><<<grid, 256, 0, ST>>>(GM_(dz), ld_dz, A_(l0), ld_l0, A_(g0), ld_g0, M, C, scale1,
                                                                              shift1, mean1, invstd1, scale3, shift3, red1, g.CL, g.PL))));
  DFCSA_LAUNCH_CHECK("branch_bwd_reduce1_kernel");
  const int parts = bilerpT_parts(v8 ? C / 8 : C, W, P);
  const long long total = static_cast<long long>(B) * H * P * (v8 ? C / 8 : C) * parts;
  VEC_DISPATCH(v8, (bilerpT_rows_kernel<VEC><<<ew_blocks(total), 256, 0, ST>>>(G_(dz), ld_dz, B, H, W, C, P, tmp, parts)));
  DFCSA_LAUNCH_CHECK("bilerpT_rows_kernel");
  // dgamma = sum dA*U = <bilinear_up^T(dA), o>: a dot product over the small pooled map instead of a gather per pixel
  if (cols_par_ok(C, H, P) && ((reinterpret_cast<uintptr_t>(tmp) | reinterpret_cast<uintptr_t>(d_o) | reinterpret_cast<uintptr_t>(o)) & 15) == 0)
    cols_reduce_par_kernel<4><<<cols_par_blocks(B, P), 256, 0, ST>>>(tmp, B, H, P, C, 1, gamma, d_o, o, dgamma);
  else if (C % 4 == 0 && ((reinterpret_cast<uintptr_t>(tmp) | reinterpret_cast<uintptr_t>(d_o) | reinterpret_cast<uintptr_t>(o)) & 15) == 0)
    cols_reduce_kernel<4><<<ew_blocks(static_cast<long long>(B) * P * P * (C / 4)), 256, 0, ST>>>(tmp, B, H, P, C, 1, gamma, d_o, o, dgamma);
  else
    cols_reduce_kernel<1><<<ew_blocks(static_cast<long long>(B) * P * P * C), 256, 0, ST>>>(tmp, B, H, P, C, 1, gamma, d_o, o, dgamma);
  DFCSA_LAUNCH_CHECK("cols_reduce_kernel(bilerpT)");
  return DFCSA_OK;
}

extern "C" int dfcsa_branch_bwd_reduce2(const void* dz, int64_t ld_dz, const void* a0, int64_t ld_a0, int32_t B, int32_t H,
                                        int32_t W, int32_t C, const float* scale2, const float* shift2, const float* mean2,
                                        const float* invstd2, const float* dpooled, int32_t P, double* red2, void* stream) {
  DFCSA_CHECK_ARG(dz && a0 && scale2 && shift2 && mean2 && invstd2 && dpooled && red2, "dfcsa_branch_bwd_reduce2: null pointer");
  DFCSA_CHECK_ARG(static_cast<long long>(B) * H * W < (1LL << 31) && H < 32768 && W < 32768, "dfcsa_branch_bwd_reduce2: too many pixels");
  const bool v8 = vec8_ok(C, {ld_dz, ld_a0}, {dz, a0, dpooled, scale2, shift2, mean2, invstd2});
  const RedGeom g = red_geom(C, v8 ? 8 : 1);
  const long long M = static_cast<long long>(B) * H * W;
  dim3 grid(red_blocks(M, g.PL, g.chunks, ew_occ(2)), g.chunks);
  OCC_DISPATCH(2, VEC_DISPATCH(v8, (branch_bwd_reduce2_kernel<VEC, OCC><<<grid, 256, pool_tabs_bytes(H, W, P), ST>>>(G_(dz), ld_dz, A_(a0), ld_a0, B, H, W, C, scale2, shift2, mean2,
                                                                         invstd2, dpooled, P, red2, g.CL, g.PL))));
  DFCSA_LAUNCH_CHECK("branch_bwd_reduce2_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_branch_bwd_apply(const void* dz, int64_t ld_dz, const void* l0, int64_t ld_l0, const void* a0,
                                      int64_t ld_a0, int32_t B, int32_t H, int32_t W, int32_t C, const float* scale1,
                                      const float* shift1, const float* mean1, const float* invstd1, const float* gamma1,
                                      const double* red1, const float* scale2, const float* shift2, const float* mean2,
                                      const float* invstd2, const float* gamma2, const double* red2, const float* dpooled,
                                      int32_t P, void* dl0, int64_t ld_dl0, void* da0, int64_t ld_da0, void* stream) {
  (void)gamma1; (void)gamma2;
  DFCSA_CHECK_ARG(dz && l0 && a0 && red1 && red2 && dpooled && dl0 && da0, "dfcsa_branch_bwd_apply: null pointer");
  const bool v8 = vec8_ok(C, {ld_dz, ld_l0, ld_a0, ld_dl0, ld_da0},
                          {dz, l0, a0, dl0, da0, dpooled, scale1, shift1, mean1, invstd1, scale2, shift2, mean2, invstd2});
  DFCSA_CHECK_ARG(static_cast<long long>(B) * H * W < (1LL << 31) && H < 32768 && W < 32768, "dfcsa_branch_bwd_apply: too many pixels");
  const RedGeom g = red_geom(C, v8 ? 8 : 1);
  // The L branch is a plain BatchNorm + ReLU backward: it runs on the light bn_bwd_apply_kernel (4 resident blocks per
  // SM, 0.97 of copy bandwidth at level 1) instead of sharing the register budget of the pool^T gather of the A branch
  // (measured together: 0.63); the coefficients and the arithmetic are the same as in the fused version.
  const long long M = static_cast<long long>(B) * H * W;
  dim3 gl(red_blocks(M, g.PL, g.chunks, 4), g.chunks);
  VEC_DISPATCH(v8, (bn_bwd_apply_kernel<VEC><<<gl, 256, 0, ST>>>(G_(dz) + C, ld_dz, A_(l0), ld_l0, M, C, scale1, shift1, mean1, invstd1,
                                                                 red1, 0, GM_(dl0), ld_dl0, g.CL, g.PL)));
  DFCSA_LAUNCH_CHECK("bn_bwd_apply_kernel");
  static const bool rows_ok = [] { const char* e = getenv("DFCSA_ACT_ROWS"); return !e || atoi(e) != 0; }();
  if (rows_ok && v8 && W / P >= 8) {
    const long long per_seg = static_cast<long long>(B) * H * (C / 8);
    const int nseg = static_cast<int>(std::min<long long>(W / 8, std::max<long long>(P, (150000 + per_seg - 1) / per_seg)));
    const long long total = per_seg * nseg;
    const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>((total + 255) / 256, static_cast<long long>(num_sms()) * 2 * 8)));
    branch_bwd_apply_rows_kernel<2><<<blocks, 256, pool_tabs_bytes(H, W, P), ST>>>(G_(dz), ld_dz, A_(a0), ld_a0, B, H, W, C, scale2, shift2,
                                                                                   mean2, invstd2, red2, dpooled, P, GM_(da0), ld_da0, nseg);
    DFCSA_LAUNCH_CHECK("branch_bwd_apply_rows_kernel");
    return DFCSA_OK;
  }
  dim3 grid(red_blocks(M, g.PL, g.chunks, ew_occ(2)), g.chunks);
  OCC_DISPATCH(2, VEC_DISPATCH(v8, (branch_bwd_apply_kernel<VEC, OCC><<<grid, 256, pool_tabs_bytes(H, W, P), ST>>>(
                                        G_(dz), ld_dz, A_(a0), ld_a0, B, H, W, C, scale2, shift2, mean2, invstd2, red2, dpooled, P,
                                        GM_(da0), ld_da0, g.CL, g.PL))));
  DFCSA_LAUNCH_CHECK("branch_bwd_apply_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_nchw_to_nhwc(const float* src, void* dst, int dst_dtype, int64_t ld, int32_t B, int32_t C, int32_t H,
                                  int32_t W, void* stream) {
  DFCSA_CHECK_ARG(src && dst && B > 0 && C > 0 && H > 0 && W > 0, "dfcsa_nchw_to_nhwc: bad args");
  nchw_to_nhwc_kernel<<<ew_blocks(static_cast<long long>(B) * C * H * W), 256, 0, ST>>>(src, dst, dst_dtype, ld, B, C, H, W);
  DFCSA_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return DFCSA_OK;
}
extern "C" int dfcsa_nhwc_to_nchw(const void* src, int src_dtype, int64_t ld, float* dst, int32_t B, int32_t C, int32_t H,
                                  int32_t W, void* stream) {
  DFCSA_CHECK_ARG(src && dst && B > 0 && C > 0 && H > 0 && W > 0, "dfcsa_nhwc_to_nchw: bad args");
  nhwc_to_nchw_kernel<<<ew_blocks(static_cast<long long>(B) * C * H * W), 256, 0, ST>>>(src, src_dtype, ld, dst, B, C, H, W);
  DFCSA_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return DFCSA_OK;
}
extern "C" int dfcsa_colsum(const void* x, int x_dtype, int64_t ld, int64_t M, int32_t C, float* out, void* stream) {
  DFCSA_CHECK_ARG(x && out && M > 0 && C > 0, "dfcsa_colsum: bad args");
  if (x_dtype != DFCSA_F32 && vec8_ok(C, {ld}, {x})) {
    const RedGeom g = red_geom(C, 8);
    dim3 grid(red_blocks(M, g.PL, g.chunks, 4), g.chunks);
    if (x_dtype == DFCSA_F16) colsum_vec_kernel<__half><<<grid, 256, 0, ST>>>(reinterpret_cast<const __half*>(x), ld, M, C, out, g.CL, g.PL);
    else colsum_vec_kernel<__nv_bfloat16><<<grid, 256, 0, ST>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, M, C, out, g.CL, g.PL);
    DFCSA_LAUNCH_CHECK("colsum_vec_kernel");
    return DFCSA_OK;
  }
  dim3 grid(static_cast<unsigned>(std::max<long long>(1, std::min<long long>((M + 63) / 64, 148 * 4))), (C + 31) / 32);
  colsum_kernel<<<grid, 256, 0, ST>>>(x, x_dtype, ld, M, C, out);
  DFCSA_LAUNCH_CHECK("colsum_kernel");
  return DFCSA_OK;
}
extern "C" int dfcsa_cast2d(const void* x, int x_dtype, int64_t ld_x, void* y, int y_dtype, int64_t ld_y, int64_t M,
                            int32_t C, void* stream) {
  DFCSA_CHECK_ARG(x && y && M > 0 && C > 0, "dfcsa_cast2d: bad args");
  if (C % 8 == 0 && ld_x % 8 == 0 && ld_y % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    const int C8 = C / 8;
    const int blocks = ew_blocks(M * C8);
#define DFCSA_CAST(TX, TY) cast2d_vec_kernel<TX, TY><<<blocks, 256, 0, ST>>>(reinterpret_cast<const TX*>(x), ld_x, reinterpret_cast<TY*>(y), ld_y, M, C8)
    if (x_dtype == DFCSA_F32 && y_dtype == DFCSA_F16) DFCSA_CAST(float, __half);
    else if (x_dtype == DFCSA_F32 && y_dtype == DFCSA_BF16) DFCSA_CAST(float, __nv_bfloat16);
    else if (x_dtype == DFCSA_F16 && y_dtype == DFCSA_F32) DFCSA_CAST(__half, float);
    else if (x_dtype == DFCSA_BF16 && y_dtype == DFCSA_F32) DFCSA_CAST(__nv_bfloat16, float);
    else if (x_dtype == DFCSA_F16 && y_dtype == DFCSA_BF16) DFCSA_CAST(__half, __nv_bfloat16);
    else if (x_dtype == DFCSA_BF16 && y_dtype == DFCSA_F16) DFCSA_CAST(__nv_bfloat16, __half);
    else if (x_dtype == DFCSA_F32) DFCSA_CAST(float, float);                 // same type: a pitched copy
    else if (x_dtype == DFCSA_F16) DFCSA_CAST(__half, __half);
    else DFCSA_CAST(__nv_bfloat16, __nv_bfloat16);
#undef DFCSA_CAST
    DFCSA_LAUNCH_CHECK("cast2d_vec_kernel");
    return DFCSA_OK;
  }
  cast2d_kernel<<<ew_blocks(M * C), 256, 0, ST>>>(x, x_dtype, ld_x, y, y_dtype, ld_y, M, C);
  DFCSA_LAUNCH_CHECK("cast2d_kernel");
  return DFCSA_OK;
}
