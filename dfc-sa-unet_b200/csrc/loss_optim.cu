// bce_dice loss (forward sums, finalize, backward) and the fused clip-by-global-norm + SGD-momentum step.
#include "common.cuh"
#include <algorithm>

namespace dfcsa {
namespace {

__device__ __forceinline__ float sigmoidf_exact(float x) { return 1.f / (1.f + expf(-x)); }

// sums: 0 bce_sum, 1 sum p*t, 2 sum p, 3 sum t, 4 sum [p>.5]*t, 5 sum [p>.5]
__global__ void __launch_bounds__(256)
bce_dice_sums_kernel(const float* __restrict__ x, const float* __restrict__ t, long long n, int from_logits, double* sums) {
  // blockIdx.y = sample of the batched entry point (n elements and 8 sums per sample); 0 for the whole-batch one
  x += static_cast<long long>(blockIdx.y) * n; t += static_cast<long long>(blockIdx.y) * n; sums += blockIdx.y * 8;
  float a[6] = {0, 0, 0, 0, 0, 0};
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float xv = x[i], tv = t[i];
    const float p = from_logits ? sigmoidf_exact(xv) : xv;
    // nn.BCELoss: log terms clamped at -100 (ATen binary_cross_entropy)
    const float lp = fmax_nan(logf(p), -100.f);          // NaN-propagating, like ATen's std::max in BCELoss
    const float l1p = fmax_nan(logf(1.f - p), -100.f);
    a[0] -= tv * lp + (1.f - tv) * l1p;
    a[1] += p * tv; a[2] += p; a[3] += tv;
    const float hard = p > 0.5f ? 1.f : 0.f;
    a[4] += hard * tv; a[5] += hard;
  }
  __shared__ float s[6][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float v = warp_sum(a[k]);
    if (lane == 0) s[k][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += s[threadIdx.x][i];
    atomicAdd(sums + threadIdx.x, static_cast<double>(v));
  }
}

__global__ void bce_dice_finalize_kernel(const double* sums, long long n, float w_bce, float w_dice, float smooth, float* out,
                                         int n_samples = 1) {
  const int smp = blockIdx.x * blockDim.x + threadIdx.x;
  if (smp >= n_samples) return;
  sums += smp * 8; out += smp * 5;
  const double bce = sums[0] / static_cast<double>(n);
  const double dice_l = 1.0 - (2.0 * sums[1] + smooth) / (sums[2] + sums[3] + smooth);
  out[0] = static_cast<float>(w_bce * bce + w_dice * dice_l);
  out[1] = static_cast<float>(bce);
  out[2] = static_cast<float>(dice_l);
  const double inter = sums[4];
  const double uni = sums[5] + sums[3] - inter;           // reference utils/metrics.py:232
  out[3] = static_cast<float>(inter / (uni + 1e-7));      // iou   :233
  out[4] = static_cast<float>(2.0 * inter / (sums[5] + sums[3] + 1e-7));  // dice :236
}

template <typename TOut>
__global__ void bce_dice_bwd_kernel(const float* __restrict__ x, const float* __restrict__ t, long long n, int from_logits,
                                    const double* sums, float w_bce, float w_dice, float smooth, const float* gout, TOut* dx) {
  const float go = gout ? *gout : 1.f;
  const double D = sums[2] + sums[3] + smooth;
  const float num = static_cast<float>(2.0 * sums[1] + smooth);
  const float invD = static_cast<float>(1.0 / D), invD2 = static_cast<float>(1.0 / (D * D));
  const float invn = 1.f / static_cast<float>(n);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float xv = x[i], tv = t[i];
    const float p = from_logits ? sigmoidf_exact(xv) : xv;
    // ATen binary_cross_entropy_backward: (p - t) / max(p*(1-p), 1e-12) / n
    const float dbce = (p - tv) / fmaxf(p * (1.f - p), 1e-12f) * invn;
    // d/dp [1 - (2I+s)/D] = -(2t*D - (2I+s)) / D^2
    const float ddice = -(2.f * tv * invD - num * invD2);
    float g = go * (w_bce * dbce + w_dice * ddice);
    if (from_logits) g *= p * (1.f - p);
    dx[i] = Cvt<TOut>::from_f(g);
  }
}

// ---------------------------------------------------------------------------------------------
// optimizer: chunked multi-tensor kernels.  chunk c covers elements [off, off+len) of tensor tix.
// ---------------------------------------------------------------------------------------------
constexpr int kChunk = 8192;

__global__ void __launch_bounds__(256)
grad_sumsq_kernel(const dfcsa_param_t* table, int n_tensors, long long max_n, double* sumsq) {
  // grid.y = tensor, grid.x = chunk of the tensor
  const dfcsa_param_t d = table[blockIdx.y];
  const long long beg = static_cast<long long>(blockIdx.x) * kChunk;
  if (beg >= d.n) return;
  const long long end = min(static_cast<long long>(d.n), beg + kChunk);
  float acc = 0.f;
  for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) { const float g = d.g[i]; acc += g * g; }
  __shared__ float s[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += s[i];
    atomicAdd(sumsq, static_cast<double>(v));
  }
}

__global__ void __launch_bounds__(256)
sgd_step_kernel(const dfcsa_param_t* table, int n_tensors, long long max_n, const double* sumsq, float gscale,
                float max_norm, float lr, float momentum, float wd, int first_step) {
  const dfcsa_param_t d = table[blockIdx.y];
  const long long beg = static_cast<long long>(blockIdx.x) * kChunk;
  if (beg >= d.n) return;
  const long long end = min(static_cast<long long>(d.n), beg + kChunk);
  // A non-finite gradient norm (NaN loss, overflow) skips the whole update, weights and momentum untouched: the
  // device-side, rank-consistent form of the reference's "NaN loss -> skip this batch" (utils/trainer.py:134-139) - the
  // norm is computed from the all-reduced gradients, so every rank takes the same decision without a host sync.
  const double ss = *sumsq;
  if (!(ss == ss) || ss > 1.7e308) return;
  // torch.nn.utils.clip_grad_norm_: coef = clamp(max_norm / (total_norm + 1e-6), max=1)
  const float total = sqrtf(static_cast<float>(ss)) * fabsf(gscale);
  const float coef = max_norm > 0.f ? fminf(max_norm / (total + 1e-6f), 1.f) * gscale : gscale;
  for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const float w = d.w[i];
    const float g = coef * d.g[i] + wd * w;
    const float m = first_step ? g : momentum * d.m[i] + g;
    d.m[i] = m;
    d.w[i] = w - lr * m;
  }
}

// dst += src over a flat fp32 buffer (gradient accumulation across micro-batches)
__global__ void __launch_bounds__(256) accumulate_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n4, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 d = reinterpret_cast<float4*>(dst)[i];
    const float4 s = reinterpret_cast<const float4*>(src)[i];
    d.x += s.x; d.y += s.y; d.z += s.z; d.w += s.w;
    reinterpret_cast<float4*>(dst)[i] = d;
  }
  for (long long i = n4 * 4 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] += src[i];
}

}  // namespace
}  // namespace dfcsa

using namespace dfcsa;
#define ST static_cast<cudaStream_t>(stream)

static int loss_blocks(long long n) {
  return static_cast<int>(std::max<long long>(1, std::min<long long>((n + 1023) / 1024, 148 * 8)));
}

extern "C" int dfcsa_bce_dice_sums(const float* x, const float* t, int64_t n, int from_logits, double* sums, void* stream) {
  DFCSA_CHECK_ARG(x && t && sums && n > 0, "dfcsa_bce_dice_sums: bad args");
  bce_dice_sums_kernel<<<loss_blocks(n), 256, 0, ST>>>(x, t, n, from_logits, sums);
  DFCSA_LAUNCH_CHECK("bce_dice_sums_kernel");
  return DFCSA_OK;
}
extern "C" int dfcsa_bce_dice_finalize(const double* sums, int64_t n, float w_bce, float w_dice, float smooth, float* out,
                                       void* stream) {
  DFCSA_CHECK_ARG(sums && out && n > 0, "dfcsa_bce_dice_finalize: bad args");
  bce_dice_finalize_kernel<<<1, 1, 0, ST>>>(sums, n, w_bce, w_dice, smooth, out);
  DFCSA_LAUNCH_CHECK("bce_dice_finalize_kernel");
  return DFCSA_OK;
}
extern "C" int dfcsa_bce_dice_sums_batched(const float* x, const float* t, int64_t n_per_sample, int32_t n_samples, int from_logits,
                                           double* sums, void* stream) {
  DFCSA_CHECK_ARG(x && t && sums && n_per_sample > 0 && n_samples > 0 && n_samples <= 65535, "dfcsa_bce_dice_sums_batched: bad args");
  dim3 grid(std::max(1, loss_blocks(n_per_sample) / std::max(1, n_samples / 4)), n_samples);
  bce_dice_sums_kernel<<<grid, 256, 0, ST>>>(x, t, n_per_sample, from_logits, sums);
  DFCSA_LAUNCH_CHECK("bce_dice_sums_kernel(batched)");
  return DFCSA_OK;
}
extern "C" int dfcsa_bce_dice_finalize_batched(const double* sums, int64_t n_per_sample, int32_t n_samples, float w_bce, float w_dice,
                                               float smooth, float* out, void* stream) {
  DFCSA_CHECK_ARG(sums && out && n_per_sample > 0 && n_samples > 0, "dfcsa_bce_dice_finalize_batched: bad args");
  bce_dice_finalize_kernel<<<(n_samples + 127) / 128, 128, 0, ST>>>(sums, n_per_sample, w_bce, w_dice, smooth, out, n_samples);
  DFCSA_LAUNCH_CHECK("bce_dice_finalize_kernel(batched)");
  return DFCSA_OK;
}
extern "C" int dfcsa_bce_dice_bwd(const float* x, const float* t, int64_t n, int from_logits, const double* sums, float w_bce,
                                  float w_dice, float smooth, const float* gout, void* dx, int dx_dtype, void* stream) {
  DFCSA_CHECK_ARG(x && t && sums && dx && n > 0, "dfcsa_bce_dice_bwd: bad args");
  const int blocks = loss_blocks(n);
  if (dx_dtype == DFCSA_F32)
    bce_dice_bwd_kernel<float><<<blocks, 256, 0, ST>>>(x, t, n, from_logits, sums, w_bce, w_dice, smooth, gout, reinterpret_cast<float*>(dx));
  else if (dx_dtype == DFCSA_BF16)
    bce_dice_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, ST>>>(x, t, n, from_logits, sums, w_bce, w_dice, smooth, gout, reinterpret_cast<__nv_bfloat16*>(dx));
  else
    bce_dice_bwd_kernel<__half><<<blocks, 256, 0, ST>>>(x, t, n, from_logits, sums, w_bce, w_dice, smooth, gout, reinterpret_cast<__half*>(dx));
  DFCSA_LAUNCH_CHECK("bce_dice_bwd_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_grad_sumsq(const dfcsa_param_t* table_dev, int32_t n_tensors, int64_t max_n, double* sumsq, void* stream) {
  DFCSA_CHECK_ARG(table_dev && sumsq && n_tensors > 0 && n_tensors <= 65535 && max_n > 0, "dfcsa_grad_sumsq: bad args");
  dim3 grid(static_cast<unsigned>((max_n + kChunk - 1) / kChunk), n_tensors);
  grad_sumsq_kernel<<<grid, 256, 0, ST>>>(table_dev, n_tensors, max_n, sumsq);
  DFCSA_LAUNCH_CHECK("grad_sumsq_kernel");
  return DFCSA_OK;
}
extern "C" int dfcsa_sgd_step(const dfcsa_param_t* table_dev, int32_t n_tensors, int64_t max_n, const double* sumsq, float gscale,
                              float max_norm, float lr, float momentum, float weight_decay, int first_step, void* stream) {
  DFCSA_CHECK_ARG(table_dev && sumsq && n_tensors > 0 && n_tensors <= 65535 && max_n > 0, "dfcsa_sgd_step: bad args");
  dim3 grid(static_cast<unsigned>((max_n + kChunk - 1) / kChunk), n_tensors);
  sgd_step_kernel<<<grid, 256, 0, ST>>>(table_dev, n_tensors, max_n, sumsq, gscale, max_norm, lr, momentum, weight_decay, first_step);
  DFCSA_LAUNCH_CHECK("sgd_step_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_accumulate(float* dst, const float* src, int64_t n, void* stream) {
  DFCSA_CHECK_ARG(dst && src && n > 0, "dfcsa_accumulate: bad args");
  const bool v4 = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0;
  const long long n4 = v4 ? n / 4 : 0;
  accumulate_kernel<<<loss_blocks(n / 4 + 1), 256, 0, ST>>>(dst, src, n4, n);
  DFCSA_LAUNCH_CHECK("accumulate_kernel");
  return DFCSA_OK;
}
