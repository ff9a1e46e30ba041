// Bandwidth-bound convolutions with a tiny channel count on one side: the first layer (Ci = 3: 3x3 conv, 1x1 attn /
// residual projections and their weight gradients over B*H*W = millions of pixels) and the final 1x1 conv (Co = 1).
// They are HBM-bound (K <= 27 or N <= 4), so they run as coalesced fp32 FMA kernels: a warp owns 32 consecutive
// pixels, a lane owns two adjacent "wide" channels (one 128-byte row per pixel per warp).
#include "common.cuh"
#include <algorithm>

namespace dfcsa {
namespace {

template <typename T> __device__ __forceinline__ float ld1(const void* p, long long i) {
  return Cvt<T>::to_f(reinterpret_cast<const T*>(p)[i]);
}
template <typename T> __device__ __forceinline__ float2 ld2(const void* p, long long i);   // two adjacent elements
template <> __device__ __forceinline__ float2 ld2<__half>(const void* p, long long i) {
  return __half22float2(*reinterpret_cast<const __half2*>(reinterpret_cast<const __half*>(p) + i));
}
template <> __device__ __forceinline__ float2 ld2<__nv_bfloat16>(const void* p, long long i) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(p) + i));
}
template <typename T> __device__ __forceinline__ void st2(void* p, long long i, float a, float b);
template <> __device__ __forceinline__ void st2<__half>(void* p, long long i, float a, float b) {
  a = fmin_nan(fmax_nan(a, -65504.f), 65504.f); b = fmin_nan(fmax_nan(b, -65504.f), 65504.f);
  *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(p) + i) = __floats2half2_rn(a, b);
}
template <> __device__ __forceinline__ void st2<__nv_bfloat16>(void* p, long long i, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(p) + i) = __floats2bfloat162_rn(a, b);
}

// gather the K = TAPS*CI patch values of pixel m into dst[0..K)
template <int TAPS, int CI, typename TX>
__device__ __forceinline__ void load_patch(const void* x, long long ld, long long m, long long M, int H, int W, float* dst) {
  if (m >= M) {
#pragma unroll
    for (int k = 0; k < TAPS * CI; ++k) dst[k] = 0.f;
    return;
  }
  if (TAPS == 1) {
#pragma unroll
    for (int c = 0; c < CI; ++c) dst[c] = ld1<TX>(x, m * ld + c);
  } else {
    const int w = static_cast<int>(m % W);
    const long long r = m / W;
    const int h = static_cast<int>(r % H);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
      const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
      const long long pix = m + (t / 3 - 1) * W + (t % 3 - 1);
#pragma unroll
      for (int c = 0; c < CI; ++c) dst[t * CI + c] = ok ? ld1<TX>(x, pix * ld + c) : 0.f;
    }
  }
}

constexpr int kWarps = 4;

// out[m, n] = sum_k patch(m)[k] * w[n, k] (+ bias[n]); per-channel sum / sum^2 in double
template <int TAPS, int CI, typename TX, typename TO>
__global__ void __launch_bounds__(kWarps * 32)
conv_smallk_kernel(const void* x, long long ld_x, long long M, int H, int W, const float* wgt, int N, void* out,
                   long long ld_out, const float* bias, double* stats, int act_cols) {
  constexpr int K = TAPS * CI;
  constexpr int KP = (K + 4) / 4 * 4;      // row pitch: 16-byte multiple so a pixel's patch is read with LDS.128 broadcasts
  __shared__ __align__(16) float s_patch[kWarps][32][KP];
  __shared__ float s_red[kWarps][2][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.y * 64 + lane * 2;
  float2 wr[K];
#pragma unroll
  for (int k = 0; k < K; ++k) wr[k] = make_float2(wgt[static_cast<long long>(n) * K + k], wgt[static_cast<long long>(n + 1) * K + k]);
  const float b0 = bias ? bias[n] : 0.f, b1 = bias ? bias[n + 1] : 0.f;
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  const long long groups = (M + 31) / 32;
  for (long long g = static_cast<long long>(blockIdx.x) * kWarps + warp; g < groups; g += static_cast<long long>(gridDim.x) * kWarps) {
    const long long m0 = g * 32;
    load_patch<TAPS, CI, TX>(x, ld_x, m0 + lane, M, H, W, s_patch[warp][lane]);
    __syncwarp();
    const int cnt = static_cast<int>(min(32LL, M - m0));
    for (int p = 0; p < cnt; ++p) {
      float a0 = b0, a1 = b1;
      float xr[KP];
#pragma unroll
      for (int k4 = 0; k4 < KP / 4; ++k4) {
        const float4 t = *reinterpret_cast<const float4*>(&s_patch[warp][p][k4 * 4]);
        xr[k4 * 4] = t.x; xr[k4 * 4 + 1] = t.y; xr[k4 * 4 + 2] = t.z; xr[k4 * 4 + 3] = t.w;
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        a0 = fmaf(xr[k], wr[k].x, a0);
        a1 = fmaf(xr[k], wr[k].y, a1);
      }
      if (n < act_cols) a0 = fmax_nan(a0, 0.f);          // folded conv + BatchNorm + ReLU (inference path)
      if (n + 1 < act_cols) a1 = fmax_nan(a1, 0.f);
      st2<TO>(out, (m0 + p) * ld_out + n, a0, a1);
      s0 += a0; s1 += a1; q0 += a0 * a0; q1 += a1 * a1;
    }
    __syncwarp();
  }
  if (stats != nullptr) {
    s_red[warp][0][lane * 2] = s0; s_red[warp][0][lane * 2 + 1] = s1;
    s_red[warp][1][lane * 2] = q0; s_red[warp][1][lane * 2 + 1] = q1;
    __syncthreads();
    if (threadIdx.x < 128) {
      const int r = threadIdx.x >> 6, c = threadIdx.x & 63;
      float v = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < kWarps; ++w2) v += s_red[w2][r][c];
      atomicAdd(stats + r * N + blockIdx.y * 64 + c, static_cast<double>(v));
    }
  }
}

// dw[wch * osw + k * osk] += alpha * sum_m wide[m, wch] * patch(m)[k]
template <int TAPS, int CI, typename TX, typename TW>
__global__ void __launch_bounds__(kWarps * 32)
wgrad_small_kernel(const void* x, long long ld_x, long long M, int H, int W, const void* wide, long long ld_w, float* dw,
                   long long osw, long long osk, const float* alpha) {
  constexpr int K = TAPS * CI;
  constexpr int KP = (K + 4) / 4 * 4;
  __shared__ __align__(16) float s_patch[kWarps][32][KP];
  __shared__ float s_acc[kWarps][K][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.y * 64 + lane * 2;
  float2 acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = make_float2(0.f, 0.f);
  const long long groups = (M + 31) / 32;
  for (long long g = static_cast<long long>(blockIdx.x) * kWarps + warp; g < groups; g += static_cast<long long>(gridDim.x) * kWarps) {
    const long long m0 = g * 32;
    load_patch<TAPS, CI, TX>(x, ld_x, m0 + lane, M, H, W, s_patch[warp][lane]);
    __syncwarp();
    const int cnt = static_cast<int>(min(32LL, M - m0));
    // four pixels per step: their wide-side loads are issued together (a dependent 500 ns load per pixel made this
    // loop latency bound), and each pixel's patch comes from shared memory as LDS.128 broadcasts
    for (int p0 = 0; p0 < 32; p0 += 4) {
      float2 d[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) d[u] = (p0 + u < cnt) ? ld2<TW>(wide, (m0 + p0 + u) * ld_w + n) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float xr[KP];
#pragma unroll
        for (int k4 = 0; k4 < KP / 4; ++k4) {
          const float4 t = *reinterpret_cast<const float4*>(&s_patch[warp][p0 + u][k4 * 4]);
          xr[k4 * 4] = t.x; xr[k4 * 4 + 1] = t.y; xr[k4 * 4 + 2] = t.z; xr[k4 * 4 + 3] = t.w;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
          acc[k].x = fmaf(xr[k], d[u].x, acc[k].x);
          acc[k].y = fmaf(xr[k], d[u].y, acc[k].y);
        }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int k = 0; k < K; ++k) { s_acc[warp][k][lane * 2] = acc[k].x; s_acc[warp][k][lane * 2 + 1] = acc[k].y; }
  __syncthreads();
  const float al = alpha ? *alpha : 1.f;
  for (int i = threadIdx.x; i < K * 64; i += blockDim.x) {
    const int k = i / 64, c = i % 64;
    float v = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < kWarps; ++w2) v += s_acc[w2][k][c];
    atomicAdd(dw + (static_cast<long long>(blockIdx.y) * 64 + c) * osw + static_cast<long long>(k) * osk, al * v);
  }
}

// out[m, n] = sum_c x[m, c] * w[n, c] + bias[n], N <= 4, C % 64 == 0: 8 lanes per pixel, 16-byte loads
template <typename TX>
__global__ void __launch_bounds__(256)
conv_smalln_kernel(const TX* x, long long ld_x, long long M, int C, const float* wgt, int N, float* out, long long ld_out,
                   const float* bias) {
  extern __shared__ float s_w[];   // [N][C]
  for (int i = threadIdx.x; i < N * C; i += blockDim.x) s_w[i] = wgt[i];
  __syncthreads();
  const int sub = threadIdx.x & 7;
  const long long pix_per_iter = static_cast<long long>(gridDim.x) * (blockDim.x / 8);
  for (long long m = static_cast<long long>(blockIdx.x) * (blockDim.x / 8) + (threadIdx.x >> 3); m < (M + 3) / 4 * 4; m += pix_per_iter) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < M) {
      for (int c0 = sub * 8; c0 < C; c0 += 64) {
        float v[8];
        load8<TX>(x + m * ld_x + c0, v);
#pragma unroll
        for (int nn = 0; nn < 4; ++nn) {
          if (nn < N) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[nn] = fmaf(v[j], s_w[nn * C + c0 + j], acc[nn]);
          }
        }
      }
    }
#pragma unroll
    for (int nn = 0; nn < 4; ++nn) {
      acc[nn] += __shfl_xor_sync(0xffffffffu, acc[nn], 1);
      acc[nn] += __shfl_xor_sync(0xffffffffu, acc[nn], 2);
      acc[nn] += __shfl_xor_sync(0xffffffffu, acc[nn], 4);
    }
    if (m < M && sub < N) {
      const float r = sub == 0 ? acc[0] : sub == 1 ? acc[1] : sub == 2 ? acc[2] : acc[3];
      out[m * ld_out + sub] = r + (bias ? bias[sub] : 0.f);
    }
  }
}

// dw[(n0 + c) * osw + k * osk] += alpha * sum_m wide[m, n0 + c] * narrow[m, k]   (1x1 taps, K <= 3 narrow channels)
// The wide rows are read as 16-byte vectors: 8 lanes per pixel, 32 pixels per block and step, UNR steps in flight per
// thread (64 bytes), so that a resident set of 4 blocks keeps ~64 KB per SM on the way - the 2-channel-per-lane kernel
// above had 4 x 4 bytes per lane in flight and ran at a fifth of copy bandwidth on the first / last layer gradients.
template <int K, typename TN, typename TW>
__global__ void __launch_bounds__(256, K == 1 ? 4 : 3)
wgrad_narrow_kernel(const TN* __restrict__ narrow, long long ld_n, long long M, const TW* __restrict__ wide, long long ld_w,
                    float* dw, long long osw, long long osk, const float* alpha) {
  constexpr int UNR = 4;
  __shared__ float s_red[8][K][64];
  const int sub = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n0 = static_cast<long long>(blockIdx.y) * 64;
  const TW* wbase = wide + n0 + sub * 8;
  float acc[K][8];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * 32;
  for (long long mb = static_cast<long long>(blockIdx.x) * 32; mb < M; mb += stride * UNR) {
    uint4 raw[UNR];
    float nv[UNR][K];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long m = mb + u * stride + pl;
      if (m < M) {
        raw[u] = *reinterpret_cast<const uint4*>(wbase + m * ld_w);
#pragma unroll
        for (int k = 0; k < K; ++k) nv[u][k] = Cvt<TN>::to_f(narrow[m * ld_n + k]);
      } else {
        raw[u] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < K; ++k) nv[u][k] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const TW* e = reinterpret_cast<const TW*>(&raw[u]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float w = Cvt<TW>::to_f(e[j]);
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k][j] = fmaf(nv[u][k], w, acc[k][j]);
      }
    }
  }
  // the four pixel lanes of a warp, then the eight warps
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[k][j];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 8) s_red[warp][k][sub * 8 + j] = v;
    }
  __syncthreads();
  const float al = alpha ? *alpha : 1.f;
  for (int i = threadIdx.x; i < K * 64; i += blockDim.x) {
    const int k = i / 64, c = i % 64;
    float v = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < 8; ++w2) v += s_red[w2][k][c];
    atomicAdd(dw + (n0 + c) * osw + static_cast<long long>(k) * osk, al * v);
  }
}

// one wave of `per_sm` resident 256-thread blocks per SM, split over the 64-channel groups of blockIdx.y
int narrow_blocks(long long M, int groups_y, int per_sm) {
  const long long want = std::max<long long>(1, static_cast<long long>(per_sm) * num_sms() / std::max(1, groups_y));
  return static_cast<int>(std::max<long long>(1, std::min<long long>((M + 127) / 128, want)));
}

int small_blocks(long long M) {
  const long long groups = (M + 31) / 32;
  return static_cast<int>(std::max<long long>(1, std::min<long long>((groups + kWarps - 1) / kWarps, 148LL * 8)));
}

}  // namespace

// returns 1 if handled, 0 if the shape is not one of the specialised ones, < 0 on error
int conv_gemm_small(const dfcsa_conv_params_t* p, cudaStream_t stream, int* rc_out) {
  *rc_out = DFCSA_OK;
  if (p->n_seg != 1 || p->out_mode != DFCSA_OUT_DIRECT || p->accumulate || p->shadow != nullptr) return 0;
  if (p->act != 0 && (p->act != 1 || p->stats != nullptr)) return 0;
  const int act_cols = p->act ? (p->act_cols > 0 ? p->act_cols : p->N) : 0;
  const dfcsa_seg_t& sg = p->seg[0];
  const long long M = static_cast<long long>(p->B) * p->H * p->W;
  // ---- tiny N (final 1x1 conv) ----
  if (act_cols == 0 && sg.tap_mode == DFCSA_TAP_1x1 && p->N <= 4 && sg.channels % 64 == 0 && sg.channels <= 1024 && p->w_dtype == DFCSA_F32 &&
      p->out_dtype == DFCSA_F32 && p->src_dtype != DFCSA_F32 && p->stats == nullptr && sg.ld % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(sg.ptr) & 15) == 0) {
    const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>((M + 31) / 32, 148LL * 8)));
    const size_t smem = static_cast<size_t>(p->N) * sg.channels * sizeof(float);
    if (p->src_dtype == DFCSA_F16)
      conv_smalln_kernel<__half><<<blocks, 256, smem, stream>>>(reinterpret_cast<const __half*>(sg.ptr), sg.ld, M, sg.channels,
                                                                reinterpret_cast<const float*>(p->w), p->N,
                                                                reinterpret_cast<float*>(p->out), p->ld_out, p->bias);
    else
      conv_smalln_kernel<__nv_bfloat16><<<blocks, 256, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(sg.ptr), sg.ld, M,
                                                                       sg.channels, reinterpret_cast<const float*>(p->w), p->N,
                                                                       reinterpret_cast<float*>(p->out), p->ld_out, p->bias);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) *rc_out = cuda_fail(e, "conv_smalln_kernel");
    return 1;
  }
  // ---- tiny K ----
  if (p->w_dtype != DFCSA_F32 || p->N % 64 != 0 || p->out_dtype == DFCSA_F32 || p->ld_out % 2 != 0 ||
      (reinterpret_cast<uintptr_t>(p->out) & 3) != 0)
    return 0;
  dim3 grid(small_blocks(M), p->N / 64);
  const float* w = reinterpret_cast<const float*>(p->w);
#define LAUNCH_SMALLK(TAPS, CI, TX, TO)                                                                                   \
  conv_smallk_kernel<TAPS, CI, TX, TO><<<grid, kWarps * 32, 0, stream>>>(sg.ptr, sg.ld, M, p->H, p->W, w, p->N, p->out, p->ld_out, \
                                                                         p->bias, p->stats, act_cols)
  bool done = false;
  if (sg.channels == 3 && p->src_dtype == DFCSA_F32 && sg.tap_mode == DFCSA_TAP_3x3) {
    if (p->out_dtype == DFCSA_F16) LAUNCH_SMALLK(9, 3, float, __half); else LAUNCH_SMALLK(9, 3, float, __nv_bfloat16);
    done = true;
  } else if (sg.channels == 3 && p->src_dtype == DFCSA_F32 && sg.tap_mode == DFCSA_TAP_1x1) {
    if (p->out_dtype == DFCSA_F16) LAUNCH_SMALLK(1, 3, float, __half); else LAUNCH_SMALLK(1, 3, float, __nv_bfloat16);
    done = true;
  } else if (sg.channels == 1 && sg.tap_mode == DFCSA_TAP_1x1 && p->src_dtype == DFCSA_BF16) {
    if (p->out_dtype == DFCSA_F16) LAUNCH_SMALLK(1, 1, __nv_bfloat16, __half); else LAUNCH_SMALLK(1, 1, __nv_bfloat16, __nv_bfloat16);
    done = true;
  }
#undef LAUNCH_SMALLK
  if (!done) return 0;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) *rc_out = cuda_fail(e, "conv_smallk_kernel");
  return 1;
}

int conv_wgrad_small(const dfcsa_wgrad_params_t* p, cudaStream_t stream, int* rc_out) {
  *rc_out = DFCSA_OK;
  if (p->dy_tap_mode != DFCSA_TAP_1x1) return 0;
  const long long M = static_cast<long long>(p->B) * p->H * p->W;
  bool done = false;
  // narrow x (C = 3, fp32), wide dy: dw[n, t*3 + c]
  if (p->C == 3 && p->x_dtype == DFCSA_F32 && p->dy_dtype == DFCSA_BF16 && p->N % 64 == 0 && p->ld_dy % 2 == 0 &&
      (reinterpret_cast<uintptr_t>(p->dy) & 3) == 0) {
    dim3 grid(small_blocks(M), p->N / 64);
    const bool vec_dy = p->ld_dy % 8 == 0 && (reinterpret_cast<uintptr_t>(p->dy) & 15) == 0;
    if (p->x_tap_mode == DFCSA_TAP_3x3)
      wgrad_small_kernel<9, 3, float, __nv_bfloat16><<<grid, kWarps * 32, 0, stream>>>(p->x, p->ld_x, M, p->H, p->W, p->dy, p->ld_dy,
                                                                                       p->dw, p->ld_dw, 1, p->alpha);
    else if (vec_dy)
      wgrad_narrow_kernel<3, float, __nv_bfloat16><<<dim3(narrow_blocks(M, p->N / 64, 3), p->N / 64), 256, 0, stream>>>(
          reinterpret_cast<const float*>(p->x), p->ld_x, M, reinterpret_cast<const __nv_bfloat16*>(p->dy), p->ld_dy, p->dw, p->ld_dw, 1,
          p->alpha);
    else
      wgrad_small_kernel<1, 3, float, __nv_bfloat16><<<grid, kWarps * 32, 0, stream>>>(p->x, p->ld_x, M, p->H, p->W, p->dy, p->ld_dy,
                                                                                       p->dw, p->ld_dw, 1, p->alpha);
    done = true;
  } else if (p->N == 1 && p->x_tap_mode == DFCSA_TAP_1x1 && p->dy_dtype == DFCSA_BF16 &&
             (p->x_dtype == DFCSA_BF16 || p->x_dtype == DFCSA_F16) &&
             p->C % 64 == 0 && p->ld_x % 2 == 0 && (reinterpret_cast<uintptr_t>(p->x) & 3) == 0) {
    // narrow dy (1 channel), wide x: dw[0, c] = sum_m dy[m] * x[m, c]
    dim3 grid(small_blocks(M), p->C / 64);
    const bool vec_x = p->ld_x % 8 == 0 && (reinterpret_cast<uintptr_t>(p->x) & 15) == 0;
    const dim3 ngrid(narrow_blocks(M, p->C / 64, 4), p->C / 64);
    const __nv_bfloat16* dy1 = reinterpret_cast<const __nv_bfloat16*>(p->dy);
    if (vec_x && p->x_dtype == DFCSA_BF16)
      wgrad_narrow_kernel<1, __nv_bfloat16, __nv_bfloat16><<<ngrid, 256, 0, stream>>>(
          dy1, p->ld_dy, M, reinterpret_cast<const __nv_bfloat16*>(p->x), p->ld_x, p->dw, 1, 0, p->alpha);
    else if (vec_x)
      wgrad_narrow_kernel<1, __nv_bfloat16, __half><<<ngrid, 256, 0, stream>>>(
          dy1, p->ld_dy, M, reinterpret_cast<const __half*>(p->x), p->ld_x, p->dw, 1, 0, p->alpha);
    else if (p->x_dtype == DFCSA_BF16)
      wgrad_small_kernel<1, 1, __nv_bfloat16, __nv_bfloat16><<<grid, kWarps * 32, 0, stream>>>(p->dy, p->ld_dy, M, p->H, p->W, p->x,
                                                                                             p->ld_x, p->dw, 1, 0, p->alpha);
    else
      wgrad_small_kernel<1, 1, __nv_bfloat16, __half><<<grid, kWarps * 32, 0, stream>>>(p->dy, p->ld_dy, M, p->H, p->W, p->x,
                                                                                      p->ld_x, p->dw, 1, 0, p->alpha);
    done = true;
  }
  if (!done) return 0;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) *rc_out = cuda_fail(e, "wgrad_small_kernel");
  return 1;
}

}  // namespace dfcsa
