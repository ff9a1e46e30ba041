// GPU-side data path: the reference's per-sample transform chain (utils/data_loader.py:25-74, run there by PIL +
// torchvision inside DataLoader workers) for a ragged batch of decoded uint8 images, bit-exact with Pillow's byte
// arithmetic:
//
//   resize BILINEAR (image) / NEAREST (mask)  ->  rotate BILINEAR / NEAREST  ->  horizontal flip
//   ->  ToTensor (/255)  ->  Normalize ((x - mean) / std)  |  mask -> (m / 255 > 0.5)
//
// Pillow's resampler is integer work on bytes (22-bit fixed-point taps, int32 accumulation, byte intermediates); its
// rotation is a double-precision lerp truncated to a byte, the mask's a 16.16 fixed-point walk.  Every double operation
// below is an explicit round-to-nearest intrinsic in Pillow's evaluation order (no FMA contraction), so the tap tables
// and coordinates are the same bits the C library computes on the host.
//
//   pp_tables_kernel    per sample: tap tables of both passes + the nearest-neighbour index tables of the mask
//   pp_resize_h_kernel  horizontal pass  src [h, w, 3] -> tmp [h, out_w, 3]   (bytes)
//   pp_resize_v_kernel  vertical pass    tmp -> resized [out_h, out_w, 3]      (bytes)
//   pp_final_kernel     rotate + flip + tensor conversion, one thread per output pixel, NCHW fp32 stores coalesced in x
#include "common.cuh"

namespace dfcsa {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

struct PpGeom {
  int n, max_h, max_w, out_h, out_w, kx, ky;      // kx / ky: tap-table widths (max over the batch)
  // workspace sections (byte offsets), each indexed by sample
  long long off_xmin, off_xcnt, off_xk, off_ymin, off_ycnt, off_yk, off_xt, off_yt, off_tmp, off_res;
};

__device__ __forceinline__ int ksize_of(double support) { return static_cast<int>(ceil(support)) * 2 + 1; }

// taps of one output index of one pass (Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc, triangle filter)
__device__ void resample_taps(int in_size, int out_size, int xx, int kmax, int* xmin_out, int* cnt_out, int* k) {
  const double scale = __ddiv_rn(static_cast<double>(in_size), static_cast<double>(out_size));
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;                       // filter support 1.0 * filterscale
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dmul_rn(static_cast<double>(xx) + 0.5, scale);
  int xmin = static_cast<int>(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) {
    double t = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
    if (t < 0.0) t = -t;
    const double w = t < 1.0 ? __dsub_rn(1.0, t) : 0.0;
    ww = __dadd_rn(ww, w);
  }
  for (int x = 0; x < kmax; ++x) {
    int q = 0;
    if (x < xmax) {
      double t = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
      if (t < 0.0) t = -t;
      double w = t < 1.0 ? __dsub_rn(1.0, t) : 0.0;
      if (ww != 0.0) w = __ddiv_rn(w, ww);
      q = w < 0.0 ? static_cast<int>(__dadd_rn(-0.5, __dmul_rn(w, static_cast<double>(1 << kPrecisionBits))))
                  : static_cast<int>(__dadd_rn(0.5, __dmul_rn(w, static_cast<double>(1 << kPrecisionBits))));
    }
    k[x] = q;
  }
  *xmin_out = xmin;
  *cnt_out = xmax;
}

// source index per output index of the NEAREST resize (Geometry.c ImagingScaleAffine): the coordinate is advanced by
// repeated addition, so the table is a serial walk (a few hundred double adds per axis)
__device__ void nearest_table(int in_size, int out_size, int* tab) {
  const double a = __ddiv_rn(static_cast<double>(in_size), static_cast<double>(out_size));
  double xo = __dadd_rn(0.0, __dmul_rn(a, 0.5));
  for (int x = 0; x < out_size; ++x) {
    const int xin = xo < 0.0 ? -1 : static_cast<int>(xo);
    tab[x] = (xin >= 0 && xin < in_size) ? xin : -1;
    xo = __dadd_rn(xo, a);
  }
}

template <typename T>
__device__ __forceinline__ T* ws_ptr(void* ws, long long off) { return reinterpret_cast<T*>(reinterpret_cast<char*>(ws) + off); }

__global__ void __launch_bounds__(256) pp_tables_kernel(const dfcsa_sample_t* __restrict__ smp, PpGeom g, void* ws) {
  const int s = blockIdx.x;
  const int h = smp[s].h, w = smp[s].w;
  int* xmin = ws_ptr<int>(ws, g.off_xmin) + static_cast<long long>(s) * g.out_w;
  int* xcnt = ws_ptr<int>(ws, g.off_xcnt) + static_cast<long long>(s) * g.out_w;
  int* xk = ws_ptr<int>(ws, g.off_xk) + static_cast<long long>(s) * g.out_w * g.kx;
  int* ymin = ws_ptr<int>(ws, g.off_ymin) + static_cast<long long>(s) * g.out_h;
  int* ycnt = ws_ptr<int>(ws, g.off_ycnt) + static_cast<long long>(s) * g.out_h;
  int* yk = ws_ptr<int>(ws, g.off_yk) + static_cast<long long>(s) * g.out_h * g.ky;
  for (int i = threadIdx.x; i < g.out_w + g.out_h; i += blockDim.x) {
    if (i < g.out_w) resample_taps(w, g.out_w, i, g.kx, xmin + i, xcnt + i, xk + static_cast<long long>(i) * g.kx);
    else { const int j = i - g.out_w; resample_taps(h, g.out_h, j, g.ky, ymin + j, ycnt + j, yk + static_cast<long long>(j) * g.ky); }
  }
  // the two serial walks on the last two warps so that they do not delay the tap tables
  if (threadIdx.x == 192) nearest_table(w, g.out_w, ws_ptr<int>(ws, g.off_xt) + static_cast<long long>(s) * g.out_w);
  if (threadIdx.x == 224) nearest_table(h, g.out_h, ws_ptr<int>(ws, g.off_yt) + static_cast<long long>(s) * g.out_h);
}

__device__ __forceinline__ unsigned char clip8(int v) {
  v >>= kPrecisionBits;
  return static_cast<unsigned char>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// tmp[s][y][xx][c] = clip8(2^21 + sum_t src[y][xmin+t][c] * k[t])
__global__ void __launch_bounds__(256) pp_resize_h_kernel(const dfcsa_sample_t* __restrict__ smp, PpGeom g, void* ws) {
  const int s = blockIdx.y;
  const int h = smp[s].h, w = smp[s].w;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(h) * g.out_w) return;
  const int xx = static_cast<int>(idx % g.out_w), y = static_cast<int>(idx / g.out_w);
  const int xmin = ws_ptr<int>(ws, g.off_xmin)[static_cast<long long>(s) * g.out_w + xx];
  const int cnt = ws_ptr<int>(ws, g.off_xcnt)[static_cast<long long>(s) * g.out_w + xx];
  const int* k = ws_ptr<int>(ws, g.off_xk) + (static_cast<long long>(s) * g.out_w + xx) * g.kx;
  const unsigned char* row = smp[s].img + (static_cast<long long>(y) * w + xmin) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int t = 0; t < cnt; ++t) {
    const int kt = k[t];
    s0 += row[t * 3 + 0] * kt; s1 += row[t * 3 + 1] * kt; s2 += row[t * 3 + 2] * kt;
  }
  unsigned char* o = ws_ptr<unsigned char>(ws, g.off_tmp) + (static_cast<long long>(s) * g.max_h * g.out_w + idx) * 3;
  o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
}

// res[s][yy][xx][c] = clip8(2^21 + sum_t tmp[ymin+t][xx][c] * k[t])
__global__ void __launch_bounds__(256) pp_resize_v_kernel(const dfcsa_sample_t* __restrict__ smp, PpGeom g, void* ws) {
  const int s = blockIdx.y;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;     // (yy, xx, c)
  if (idx >= static_cast<long long>(g.out_h) * g.out_w * 3) return;
  const int yy = static_cast<int>(idx / (g.out_w * 3));
  const int xc = static_cast<int>(idx - static_cast<long long>(yy) * g.out_w * 3);
  const int ymin = ws_ptr<int>(ws, g.off_ymin)[static_cast<long long>(s) * g.out_h + yy];
  const int cnt = ws_ptr<int>(ws, g.off_ycnt)[static_cast<long long>(s) * g.out_h + yy];
  const int* k = ws_ptr<int>(ws, g.off_yk) + (static_cast<long long>(s) * g.out_h + yy) * g.ky;
  const unsigned char* col = ws_ptr<unsigned char>(ws, g.off_tmp) + (static_cast<long long>(s) * g.max_h + ymin) * g.out_w * 3 + xc;
  int acc = 1 << (kPrecisionBits - 1);
  for (int t = 0; t < cnt; ++t) acc += col[static_cast<long long>(t) * g.out_w * 3] * k[t];
  ws_ptr<unsigned char>(ws, g.off_res)[static_cast<long long>(s) * g.out_h * g.out_w * 3 + idx] = clip8(acc);
}

__device__ __forceinline__ int floor_i(double v) { return v < 0.0 ? static_cast<int>(floor(v)) : static_cast<int>(v); }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ double lerp_rn(double a, double b, double d) { return __dadd_rn(a, __dmul_rn(__dsub_rn(b, a), d)); }

struct Norm { float mean[3], std[3]; };

// one thread per output pixel (s, y, x): undo the flip, then the rotation, sample the resized image / mask
__global__ void __launch_bounds__(256) pp_final_kernel(const dfcsa_sample_t* __restrict__ smp, PpGeom g, const void* ws, Norm nm,
                                                       float* __restrict__ img_out, float* __restrict__ mask_out) {
  const int s = blockIdx.y;
  const int H = g.out_h, W = g.out_w;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * W) return;
  const int y = idx / W;
  int x = idx - y * W;
  const dfcsa_sample_t sm = smp[s];
  if (sm.flip) x = W - 1 - x;
  const unsigned char* res = ws_ptr<const unsigned char>(const_cast<void*>(ws), g.off_res) + static_cast<long long>(s) * H * W * 3;
  const int* xt = ws_ptr<const int>(const_cast<void*>(ws), g.off_xt) + static_cast<long long>(s) * W;
  const int* yt = ws_ptr<const int>(const_cast<void*>(ws), g.off_yt) + static_cast<long long>(s) * H;
  auto mask_at = [&](int my, int mx) -> unsigned char {       // resized (NEAREST) mask, never materialised
    if (sm.mask == nullptr) return 0;
    const int sy = yt[my], sx = xt[mx];
    return (sy < 0 || sx < 0) ? static_cast<unsigned char>(0) : sm.mask[static_cast<long long>(sy) * sm.w + sx];
  };
  unsigned char px[3] = {0, 0, 0}, mk = 0;
  if (sm.rot_mode == DFCSA_ROT_AFFINE) {
    // image: Geometry.c affine_transform + bilinear_filter32RGB
    const double xs = static_cast<double>(x) + 0.5, ys = static_cast<double>(y) + 0.5;
    double xin = __dadd_rn(__dadd_rn(__dmul_rn(sm.a[0], xs), __dmul_rn(sm.a[1], ys)), sm.a[2]);
    double yin = __dadd_rn(__dadd_rn(__dmul_rn(sm.a[3], xs), __dmul_rn(sm.a[4], ys)), sm.a[5]);
    if (!(xin < 0.0 || xin >= static_cast<double>(W) || yin < 0.0 || yin >= static_cast<double>(H))) {
      xin = __dsub_rn(xin, 0.5); yin = __dsub_rn(yin, 0.5);
      const int xi = floor_i(xin), yi = floor_i(yin);
      const double dx = __dsub_rn(xin, static_cast<double>(xi)), dy = __dsub_rn(yin, static_cast<double>(yi));
      const int x0 = clampi(xi, 0, W - 1), x1 = clampi(xi + 1, 0, W - 1), yc = clampi(yi, 0, H - 1);
      const bool y1ok = yi + 1 >= 0 && yi + 1 < H;
      const unsigned char* r0 = res + static_cast<long long>(yc) * W * 3;
      const unsigned char* r1 = res + static_cast<long long>(y1ok ? yi + 1 : yc) * W * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double v1 = lerp_rn(static_cast<double>(r0[x0 * 3 + c]), static_cast<double>(r0[x1 * 3 + c]), dx);
        const double v2 = y1ok ? lerp_rn(static_cast<double>(r1[x0 * 3 + c]), static_cast<double>(r1[x1 * 3 + c]), dx) : v1;
        px[c] = static_cast<unsigned char>(__double2int_rz(lerp_rn(v1, v2, dy)));
      }
    }
    // mask: affine_fixed, 16.16
    auto fix = [](double v) { return static_cast<long long>(floor(__dadd_rn(__dmul_rn(v, 65536.0), 0.5))); };
    const long long a0 = fix(sm.a[0]), a1 = fix(sm.a[1]), a3 = fix(sm.a[3]), a4 = fix(sm.a[4]);
    const long long a2 = fix(__dadd_rn(__dadd_rn(sm.a[2], __dmul_rn(sm.a[0], 0.5)), __dmul_rn(sm.a[1], 0.5)));
    const long long a5 = fix(__dadd_rn(__dadd_rn(sm.a[5], __dmul_rn(sm.a[3], 0.5)), __dmul_rn(sm.a[4], 0.5)));
    const long long mx = (a2 + a1 * y + a0 * x) >> 16, my = (a5 + a4 * y + a3 * x) >> 16;
    if (mx >= 0 && mx < W && my >= 0 && my < H) mk = mask_at(static_cast<int>(my), static_cast<int>(mx));
  } else {
    int sy = y, sx = x;
    if (sm.rot_mode == DFCSA_ROT_90) { sy = x; sx = W - 1 - y; }             // counter-clockwise quarter turn (square)
    else if (sm.rot_mode == DFCSA_ROT_180) { sy = H - 1 - y; sx = W - 1 - x; }
    else if (sm.rot_mode == DFCSA_ROT_270) { sy = H - 1 - x; sx = y; }
    const unsigned char* r = res + (static_cast<long long>(sy) * W + sx) * 3;
    px[0] = r[0]; px[1] = r[1]; px[2] = r[2];
    mk = mask_at(sy, sx);
  }
  const int xo = idx - y * W;       // the output column (x was mirrored above)
  const long long plane = static_cast<long long>(H) * W;
  float* io = img_out + static_cast<long long>(s) * 3 * plane + static_cast<long long>(y) * W + xo;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float t = __fdiv_rn(static_cast<float>(px[c]), 255.0f);
    io[c * plane] = __fdiv_rn(__fsub_rn(t, nm.mean[c]), nm.std[c]);
  }
  if (mask_out != nullptr)
    mask_out[static_cast<long long>(s) * plane + static_cast<long long>(y) * W + xo] = __fdiv_rn(static_cast<float>(mk), 255.0f) > 0.5f ? 1.f : 0.f;
}

long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

bool make_geom(int n, int max_h, int max_w, int out_h, int out_w, PpGeom* g, long long* total) {
  if (n <= 0 || max_h <= 0 || max_w <= 0 || out_h <= 0 || out_w <= 0) return false;
  g->n = n; g->max_h = max_h; g->max_w = max_w; g->out_h = out_h; g->out_w = out_w;
  auto ks = [](int in, int out) {
    const double scale = static_cast<double>(in) / out;
    const double support = scale < 1.0 ? 1.0 : scale;
    return static_cast<int>(std::ceil(support)) * 2 + 1;
  };
  g->kx = ks(max_w, out_w); g->ky = ks(max_h, out_h);
  long long off = 0;
  auto take = [&](long long bytes) { const long long o = off; off = align_up(off + bytes, 256); return o; };
  g->off_xmin = take(4LL * n * out_w); g->off_xcnt = take(4LL * n * out_w); g->off_xk = take(4LL * n * out_w * g->kx);
  g->off_ymin = take(4LL * n * out_h); g->off_ycnt = take(4LL * n * out_h); g->off_yk = take(4LL * n * out_h * g->ky);
  g->off_xt = take(4LL * n * out_w); g->off_yt = take(4LL * n * out_h);
  g->off_tmp = take(3LL * n * max_h * out_w);
  g->off_res = take(3LL * n * out_h * out_w);
  *total = off;
  return true;
}

}  // namespace
}  // namespace dfcsa

using namespace dfcsa;

extern "C" int64_t dfcsa_preprocess_workspace_bytes(int32_t n, int32_t max_src_h, int32_t max_src_w, int32_t out_h, int32_t out_w) {
  PpGeom g; long long total = 0;
  if (!make_geom(n, max_src_h, max_src_w, out_h, out_w, &g, &total)) return -1;
  return total;
}

extern "C" int dfcsa_preprocess(const dfcsa_sample_t* samples_dev, int32_t n, int32_t max_src_h, int32_t max_src_w, int32_t out_h,
                                int32_t out_w, const float* mean3, const float* std3, float* img_out, float* mask_out,
                                void* workspace, int64_t workspace_bytes, void* stream) {
  DFCSA_CHECK_ARG(samples_dev && mean3 && std3 && img_out && workspace, "dfcsa_preprocess: null pointer");
  PpGeom g; long long total = 0;
  DFCSA_CHECK_ARG(make_geom(n, max_src_h, max_src_w, out_h, out_w, &g, &total), "dfcsa_preprocess: empty batch or image");
  DFCSA_CHECK_ARG(workspace_bytes >= total, "dfcsa_preprocess: workspace too small (dfcsa_preprocess_workspace_bytes)");
  DFCSA_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "dfcsa_preprocess: workspace must be 256-byte aligned");
  DFCSA_CHECK_ARG(max_src_h < 32768 && max_src_w < 32768 && out_h < 32768 && out_w < 32768 && n <= 65535,
                  "dfcsa_preprocess: sizes beyond the 16.16 fixed-point range of the rotation / the grid limits");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Norm nm;
  for (int c = 0; c < 3; ++c) { nm.mean[c] = mean3[c]; nm.std[c] = std3[c]; }
  pp_tables_kernel<<<n, 256, 0, st>>>(samples_dev, g, workspace);
  DFCSA_LAUNCH_CHECK("pp_tables_kernel");
  {
    const long long items = static_cast<long long>(max_src_h) * out_w;
    dim3 grid(static_cast<unsigned>((items + 255) / 256), n);
    pp_resize_h_kernel<<<grid, 256, 0, st>>>(samples_dev, g, workspace);
    DFCSA_LAUNCH_CHECK("pp_resize_h_kernel");
  }
  {
    const long long items = static_cast<long long>(out_h) * out_w * 3;
    dim3 grid(static_cast<unsigned>((items + 255) / 256), n);
    pp_resize_v_kernel<<<grid, 256, 0, st>>>(samples_dev, g, workspace);
    DFCSA_LAUNCH_CHECK("pp_resize_v_kernel");
  }
  {
    dim3 grid(static_cast<unsigned>((static_cast<long long>(out_h) * out_w + 255) / 256), n);
    pp_final_kernel<<<grid, 256, 0, st>>>(samples_dev, g, workspace, nm, img_out, mask_out);
    DFCSA_LAUNCH_CHECK("pp_final_kernel");
  }
  return DFCSA_OK;
}
