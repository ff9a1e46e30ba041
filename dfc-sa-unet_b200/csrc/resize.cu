// General bilinear re-size and adaptive average pool on NHWC tensors (forward and transposed), any storage dtype.
//
// Two users, both OFF the 224 / 512 / 1024 hot path (plain one-thread-per-output gather kernels, no tuning):
//   * the reference's post-ConvTranspose re-size for inputs whose side is not a multiple of 16
//     (models/unet_dfc_sa_res.py:180-181: F.interpolate(x, size=skip.shape[2:], mode="bilinear", align_corners=False));
//   * LightSelfAttention / FullResolutionAttention called on their own (models/unet_dfc_sa_res.py:20-39): pool without the
//     BatchNorm + ReLU the block kernels fuse in front of it, and gamma * up(o) + x.
#include "common.cuh"
#include <algorithm>

namespace dfcsa {
namespace {

__device__ __forceinline__ float ld_any(const void* p, int dt, long long i) {
  if (dt == DFCSA_F32) return reinterpret_cast<const float*>(p)[i];
  if (dt == DFCSA_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void st_any(void* p, int dt, long long i, float v) {
  if (dt == DFCSA_F32) reinterpret_cast<float*>(p)[i] = v;
  else if (dt == DFCSA_F16) reinterpret_cast<__half*>(p)[i] = Cvt<__half>::from_f(v);
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

// ATen area_pixel_compute_source_index, align_corners=False: source taps of destination index d (n_src -> n_dst)
__device__ __forceinline__ void taps(int d, int n_src, int n_dst, int& i0, int& i1, float& l1) {
  const float scale = static_cast<float>(n_src) / static_cast<float>(n_dst);
  float src = scale * (static_cast<float>(d) + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = min(static_cast<int>(src), n_src - 1);
  i1 = min(i0 + 1, n_src - 1);
  l1 = src - static_cast<float>(i0);
}
// destination indices whose taps can touch source index i: a conservative [lo, hi] range
__device__ __forceinline__ void dst_range(int i, int n_src, int n_dst, int& lo, int& hi) {
  const float ratio = static_cast<float>(n_dst) / static_cast<float>(n_src);
  lo = max(static_cast<int>(floorf((i - 0.5f) * ratio - 0.5f)) - 1, 0);
  hi = min(static_cast<int>(ceilf((i + 1.5f) * ratio - 0.5f)) + 1, n_dst - 1);
}

// dst[b, y, x, c] = alpha * bilinear(src)[b, y, x, c] + add[b, y, x, c]
__global__ void __launch_bounds__(256)
resize_fwd_kernel(const void* src, int sdt, long long ld_s, int B, int Hi, int Wi, int C, void* dst, int ddt, long long ld_d, int Ho,
                  int Wo, const float* alpha, const void* add, int adt, long long ld_a) {
  const long long total = static_cast<long long>(B) * Ho * Wo * C;
  const float al = alpha ? *alpha : 1.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long r = i / C;
    const int x = static_cast<int>(r % Wo); r /= Wo;
    const int y = static_cast<int>(r % Ho);
    const long long b = r / Ho;
    int y0, y1, x0, x1; float ly, lx;
    taps(y, Hi, Ho, y0, y1, ly);
    taps(x, Wi, Wo, x0, x1, lx);
    const long long base = b * Hi * Wi;
    const float v00 = ld_any(src, sdt, (base + y0 * Wi + x0) * ld_s + c), v01 = ld_any(src, sdt, (base + y0 * Wi + x1) * ld_s + c);
    const float v10 = ld_any(src, sdt, (base + y1 * Wi + x0) * ld_s + c), v11 = ld_any(src, sdt, (base + y1 * Wi + x1) * ld_s + c);
    float v = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
    v *= al;
    const long long m = (b * Ho + y) * Wo + x;
    if (add != nullptr) v += ld_any(add, adt, m * ld_a + c);
    st_any(dst, ddt, m * ld_d + c, v);
  }
}

// dsrc[b, yi, xi, c] = alpha * sum over destination pixels of their tap weight on (yi, xi) * ddst   (the transpose)
__global__ void __launch_bounds__(256)
resize_bwd_kernel(const void* ddst, int ddt, long long ld_d, int B, int Hi, int Wi, int C, void* dsrc, int sdt, long long ld_s, int Ho,
                  int Wo, const float* alpha) {
  const long long total = static_cast<long long>(B) * Hi * Wi * C;
  const float al = alpha ? *alpha : 1.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long r = i / C;
    const int xi = static_cast<int>(r % Wi); r /= Wi;
    const int yi = static_cast<int>(r % Hi);
    const long long b = r / Hi;
    int ylo, yhi, xlo, xhi;
    dst_range(yi, Hi, Ho, ylo, yhi);
    dst_range(xi, Wi, Wo, xlo, xhi);
    float acc = 0.f;
    for (int y = ylo; y <= yhi; ++y) {
      int y0, y1; float ly; taps(y, Hi, Ho, y0, y1, ly);
      float wy = 0.f;
      if (y0 == yi) wy += 1.f - ly;
      if (y1 == yi) wy += ly;
      if (wy == 0.f) continue;
      float row = 0.f;
      for (int x = xlo; x <= xhi; ++x) {
        int x0, x1; float lx; taps(x, Wi, Wo, x0, x1, lx);
        float wx = 0.f;
        if (x0 == xi) wx += 1.f - lx;
        if (x1 == xi) wx += lx;
        if (wx != 0.f) row = fmaf(wx, ld_any(ddst, ddt, ((b * Ho + y) * Wo + x) * ld_d + c), row);
      }
      acc = fmaf(wy, row, acc);
    }
    st_any(dsrc, sdt, ((b * Hi + yi) * Wi + xi) * ld_s + c, al * acc);
  }
}

// pooled[b, i, j, c] = mean of src over window (i, j) of adaptive_avg_pool2d(., P)
__global__ void __launch_bounds__(256)
pool_fwd_kernel(const void* src, int sdt, long long ld_s, int B, int H, int W, int C, int P, float* pooled) {
  const long long total = static_cast<long long>(B) * P * P * C;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    long long r = idx / C;
    const int j = static_cast<int>(r % P); r /= P;
    const int i = static_cast<int>(r % P);
    const long long b = r / P;
    const int ylo = (i * H) / P, yhi = ((i + 1) * H + P - 1) / P, xlo = (j * W) / P, xhi = ((j + 1) * W + P - 1) / P;
    float acc = 0.f;
    for (int y = ylo; y < yhi; ++y)
      for (int x = xlo; x < xhi; ++x) acc += ld_any(src, sdt, ((b * H + y) * W + x) * ld_s + c);
    pooled[idx] = acc / static_cast<float>((yhi - ylo) * (xhi - xlo));
  }
}

// dst[b, y, x, c] = add[b, y, x, c] + sum over windows (i, j) containing (y, x) of dpooled[b, i, j, c] / |window|
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const float* dpooled, int B, int H, int W, int C, int P, const void* add, int adt, long long ld_a, void* dst, int ddt,
                long long ld_d) {
  const long long total = static_cast<long long>(B) * H * W * C;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    long long r = idx / C;
    const int x = static_cast<int>(r % W); r /= W;
    const int y = static_cast<int>(r % H);
    const long long b = r / H;
    const int ilo = (y * P) / H, ihi = min(((y + 1) * P - 1) / H, P - 1);
    const int jlo = (x * P) / W, jhi = min(((x + 1) * P - 1) / W, P - 1);
    float acc = 0.f;
    for (int i = ilo; i <= ihi; ++i) {
      const int ya = (i * H) / P, yb = ((i + 1) * H + P - 1) / P;
      for (int j = jlo; j <= jhi; ++j) {
        const int xa = (j * W) / P, xb = ((j + 1) * W + P - 1) / P;
        acc += dpooled[((b * P + i) * P + j) * C + c] / static_cast<float>((yb - ya) * (xb - xa));
      }
    }
    const long long m = (b * H + y) * W + x;
    if (add != nullptr) acc += ld_any(add, adt, m * ld_a + c);
    st_any(dst, ddt, m * ld_d + c, acc);
  }
}

int blocks_for(long long total) { return static_cast<int>(std::max<long long>(1, std::min<long long>((total + 255) / 256, 148LL * 16))); }
bool dt_ok(int dt) { return dt == DFCSA_F32 || dt == DFCSA_F16 || dt == DFCSA_BF16; }

}  // namespace
}  // namespace dfcsa

using namespace dfcsa;
#define ST static_cast<cudaStream_t>(stream)

extern "C" int dfcsa_resize_bilinear(const void* src, int src_dtype, int64_t ld_src, int32_t B, int32_t Hi, int32_t Wi, int32_t C,
                                     void* dst, int dst_dtype, int64_t ld_dst, int32_t Ho, int32_t Wo, const float* alpha,
                                     const void* add, int add_dtype, int64_t ld_add, void* stream) {
  DFCSA_CHECK_ARG(src && dst && B > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C > 0 && dt_ok(src_dtype) && dt_ok(dst_dtype) &&
                  (add == nullptr || dt_ok(add_dtype)), "dfcsa_resize_bilinear: bad args");
  const long long total = static_cast<long long>(B) * Ho * Wo * C;
  resize_fwd_kernel<<<blocks_for(total), 256, 0, ST>>>(src, src_dtype, ld_src, B, Hi, Wi, C, dst, dst_dtype, ld_dst, Ho, Wo, alpha, add,
                                                       add_dtype, ld_add);
  DFCSA_LAUNCH_CHECK("resize_fwd_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_resize_bilinear_bwd(const void* ddst, int ddst_dtype, int64_t ld_ddst, int32_t B, int32_t Hi, int32_t Wi, int32_t C,
                                         void* dsrc, int dsrc_dtype, int64_t ld_dsrc, int32_t Ho, int32_t Wo, const float* alpha,
                                         void* stream) {
  DFCSA_CHECK_ARG(ddst && dsrc && B > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C > 0 && dt_ok(ddst_dtype) && dt_ok(dsrc_dtype),
                  "dfcsa_resize_bilinear_bwd: bad args");
  const long long total = static_cast<long long>(B) * Hi * Wi * C;
  resize_bwd_kernel<<<blocks_for(total), 256, 0, ST>>>(ddst, ddst_dtype, ld_ddst, B, Hi, Wi, C, dsrc, dsrc_dtype, ld_dsrc, Ho, Wo, alpha);
  DFCSA_LAUNCH_CHECK("resize_bwd_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_adaptive_pool(const void* src, int src_dtype, int64_t ld_src, int32_t B, int32_t H, int32_t W, int32_t C, int32_t P,
                                   float* pooled, void* stream) {
  DFCSA_CHECK_ARG(src && pooled && B > 0 && H > 0 && W > 0 && C > 0 && P > 0 && dt_ok(src_dtype), "dfcsa_adaptive_pool: bad args");
  pool_fwd_kernel<<<blocks_for(static_cast<long long>(B) * P * P * C), 256, 0, ST>>>(src, src_dtype, ld_src, B, H, W, C, P, pooled);
  DFCSA_LAUNCH_CHECK("pool_fwd_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_adaptive_pool_bwd(const float* dpooled, int32_t B, int32_t H, int32_t W, int32_t C, int32_t P, const void* add,
                                       int add_dtype, int64_t ld_add, void* dst, int dst_dtype, int64_t ld_dst, void* stream) {
  DFCSA_CHECK_ARG(dpooled && dst && B > 0 && H > 0 && W > 0 && C > 0 && P > 0 && dt_ok(dst_dtype) && (add == nullptr || dt_ok(add_dtype)),
                  "dfcsa_adaptive_pool_bwd: bad args");
  pool_bwd_kernel<<<blocks_for(static_cast<long long>(B) * H * W * C), 256, 0, ST>>>(dpooled, B, H, W, C, P, add, add_dtype, ld_add, dst,
                                                                                      dst_dtype, ld_dst);
  DFCSA_LAUNCH_CHECK("pool_bwd_kernel");
  return DFCSA_OK;
}
