// fp32 SIMT kernels: the implicit-GEMM convolution and weight gradient for channel counts the tensor-core
// path does not take (the Ci=3 first layer, small test networks), the strided batched GEMM behind the pooled
// attention, softmax and the weight re-layout.  Same parameter structs and layouts as the tcgen05 kernels.
#include "common.cuh"
#include <algorithm>

namespace dfcsa {
namespace {

template <typename T> __device__ __forceinline__ float ldf(const void* p, long long i) {
  return Cvt<T>::to_f(reinterpret_cast<const T*>(p)[i]);
}
__device__ __forceinline__ float ld_any(const void* p, long long i, int dt) {
  if (dt == DFCSA_F32) return reinterpret_cast<const float*>(p)[i];
  if (dt == DFCSA_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void st_any(void* p, long long i, int dt, float v) {
  if (dt == DFCSA_F32) reinterpret_cast<float*>(p)[i] = v;
  else if (dt == DFCSA_F16) reinterpret_cast<__half*>(p)[i] = Cvt<__half>::from_f(v);
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

// ---------------------------------------------------------------------------------------------
// conv forward / dgrad, 64 pixels x 64 channels per CTA, 4x4 per thread
// ---------------------------------------------------------------------------------------------
struct SimtConvArgs {
  dfcsa_conv_params_t p;
  long long M;
  int ktot;
  int convt_co;
};

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256)
conv_simt_kernel(const SimtConvArgs a) {
  __shared__ float sA[TK][TM + 1];
  __shared__ float sB[TK][TN + 1];
  __shared__ float sRed[2][TN];
  const dfcsa_conv_params_t& p = a.p;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;   // tx -> 4 output channels, ty -> 4 pixels
  const long long m0 = static_cast<long long>(blockIdx.x) * TM;
  const int n0 = blockIdx.y * TN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // this thread loads A element (row lr, k lk) x4 rows and B element (n ln, k lk)
  const int lr = tid / TK;      // 0..15
  const int lk = tid % TK;
  int koff = 0;
  for (int s = 0; s < p.n_seg; ++s) {
    const dfcsa_seg_t sg = p.seg[s];
    const int taps = sg.tap_mode == DFCSA_TAP_1x1 ? 1 : sg.tap_mode == DFCSA_TAP_3x3 ? 9 : 4;
    for (int t = 0; t < taps; ++t) {
      for (int c0 = 0; c0 < sg.channels; c0 += TK) {
        // ---- load A tile ----
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int r = lr + rr * 16;
          const long long m = m0 + r;
          float v = 0.f;
          const int c = c0 + lk;
          if (m < a.M && c < sg.channels) {
            const int w = static_cast<int>(m % p.W);
            const long long t2 = m / p.W;
            const int h = static_cast<int>(t2 % p.H);
            const long long b = t2 / p.H;
            long long pix = -1;
            if (sg.tap_mode == DFCSA_TAP_1x1) pix = m;
            else if (sg.tap_mode == DFCSA_TAP_3x3) {
              const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
              if (hh >= 0 && hh < p.H && ww >= 0 && ww < p.W) pix = (b * p.H + hh) * p.W + ww;
            } else {
              pix = (b * 2 * p.H + 2 * h + (t >> 1)) * (2LL * p.W) + 2 * w + (t & 1);
            }
            if (pix >= 0) v = ld_any(sg.ptr, pix * sg.ld + c, p.src_dtype);
          }
          sA[lk][r] = v;
        }
        // ---- load B tile ----
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int nn = lr + rr * 16;
          const int n = n0 + nn;
          const int c = c0 + lk;
          float v = 0.f;
          if (n < p.N && c < sg.channels)
            v = ld_any(p.w, static_cast<long long>(n) * a.ktot + koff + t * sg.channels + c, p.w_dtype);
          sB[lk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
          float av[4], bv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) av[i] = sA[k][ty * 4 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) bv[j] = sB[k][tx * 4 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
    koff += taps * sg.channels;
  }

  // ---- epilogue ----
  if (p.stats != nullptr) {
    if (tid < TN) { sRed[0][tid] = 0.f; sRed[1][tid] = 0.f; }
    __syncthreads();
  }
  float cs[4] = {0, 0, 0, 0}, cq[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      long long pix = m;
      int cn = n;
      if (p.out_mode == DFCSA_OUT_CONVT2x2) {
        const int q = n / a.convt_co;
        cn = n - q * a.convt_co;
        const int w = static_cast<int>(m % p.W);
        const long long t2 = m / p.W;
        const int h = static_cast<int>(t2 % p.H);
        const long long b = t2 / p.H;
        pix = (b * 2 * p.H + 2 * h + (q >> 1)) * (2LL * p.W) + 2 * w + (q & 1);
      }
      if (p.bias != nullptr) v += p.bias[cn];
      if (p.act == 1 && n < (p.act_cols > 0 ? p.act_cols : p.N)) v = fmax_nan(v, 0.f);
      const long long o = pix * p.ld_out + cn;
      if (p.accumulate) v += ld_any(p.out, o, p.out_dtype);
      st_any(p.out, o, p.out_dtype, v);
      if (p.shadow != nullptr) reinterpret_cast<__nv_bfloat16*>(p.shadow)[pix * p.ld_shadow + cn] = __float2bfloat16_rn(v);
      cs[j] += v; cq[j] += v * v;
    }
  }
  if (p.stats != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&sRed[0][tx * 4 + j], cs[j]);
      atomicAdd(&sRed[1][tx * 4 + j], cq[j]);
    }
    __syncthreads();
    if (tid < TN && n0 + tid < p.N) {
      atomicAdd(p.stats + n0 + tid, static_cast<double>(sRed[0][tid]));
      atomicAdd(p.stats + p.N + n0 + tid, static_cast<double>(sRed[1][tid]));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient: dw[n, t*C + c] += alpha * sum_m dy[pdy(m,t), n] * x[px(m,t), c]
// grid (n tiles, c tiles * taps, pixel splits); 64 x 64 per CTA, K = pixels
// ---------------------------------------------------------------------------------------------
struct SimtWgradArgs {
  dfcsa_wgrad_params_t p;
  long long M;
  int taps;
  int c_tiles;
  long long pix_per_split;
};

__global__ void __launch_bounds__(256)
wgrad_simt_kernel(const SimtWgradArgs a) {
  __shared__ float sA[TK][TM + 1];  // dy  [k pixel][n]
  __shared__ float sB[TK][TN + 1];  // x   [k pixel][c]
  const dfcsa_wgrad_params_t& p = a.p;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int n0 = blockIdx.x * TM;
  const int t = blockIdx.y / a.c_tiles;
  const int c0 = (blockIdx.y % a.c_tiles) * TN;
  const long long mbeg = blockIdx.z * a.pix_per_split;
  const long long mend = min(a.M, mbeg + a.pix_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int lk = tid / 64;   // 0..3 (pixel within group of 4), x4 -> 16
  const int lc = tid % 64;   // channel within tile (coalesced along channels)
  for (long long mk = mbeg; mk < mend; mk += TK) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = lk + kk * 4;
      const long long m = mk + k;
      float va = 0.f, vb = 0.f;
      if (m < mend) {
        const int w = static_cast<int>(m % p.W);
        const long long t2 = m / p.W;
        const int h = static_cast<int>(t2 % p.H);
        const long long b = t2 / p.H;
        long long pdy = m, px = m;
        if (p.dy_tap_mode == DFCSA_TAP_2x2S2) pdy = (b * 2 * p.H + 2 * h + (t >> 1)) * (2LL * p.W) + 2 * w + (t & 1);
        if (p.x_tap_mode == DFCSA_TAP_3x3) {
          const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
          px = (hh >= 0 && hh < p.H && ww >= 0 && ww < p.W) ? (b * p.H + hh) * p.W + ww : -1;
        }
        if (n0 + lc < p.N) va = ld_any(p.dy, pdy * p.ld_dy + n0 + lc, p.dy_dtype);
        if (px >= 0 && c0 + lc < p.C) vb = ld_any(p.x, px * p.ld_x + c0 + lc, p.x_dtype);
      }
      sA[k][lc] = va;
      sB[k][lc] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = sA[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = sB[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float alpha = p.alpha ? *p.alpha : 1.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= p.N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      if (c >= p.C) continue;
      atomicAdd(p.dw + static_cast<long long>(n) * p.ld_dw + static_cast<long long>(t) * p.C + c, alpha * acc[i][j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// strided batched sgemm
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sgemm_kernel(const dfcsa_sgemm_params_t p) {
  __shared__ float sA[TK][TM + 1];
  __shared__ float sB[TK][TN + 1];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int b = blockIdx.z;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const float* A = p.A + b * p.a_b;
  const float* B = p.B + b * p.b_b;
  float* C = p.C + b * p.c_b;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // choose the load mapping so the unit-stride axis is walked by consecutive threads
  const bool a_k_fast = p.a_k == 1;
  const bool b_k_fast = p.b_k == 1;
  for (int k0 = 0; k0 < p.K; k0 += TK) {
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      int r, k;
      if (a_k_fast) { k = tid % TK; r = tid / TK + rr * 16; } else { r = tid % 64; k = tid / 64 + rr * 4; }
      float v = 0.f;
      if (m0 + r < p.M && k0 + k < p.K) v = A[(m0 + r) * p.a_m + (k0 + k) * p.a_k];
      sA[k][r] = v;
      int n, kb;
      if (b_k_fast) { kb = tid % TK; n = tid / TK + rr * 16; } else { n = tid % 64; kb = tid / 64 + rr * 4; }
      float u = 0.f;
      if (n0 + n < p.N && k0 + kb < p.K) u = B[(k0 + kb) * p.b_k + (n0 + n) * p.b_n];
      sB[kb][n] = u;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = sA[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = sB[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = p.alpha * acc[i][j];
      if (p.bias_n) v += p.bias_n[n];
      if (p.bias_m) v += p.bias_m[m];
      float* c = C + m * p.c_m + n * p.c_n;
      if (p.beta != 0.f) v += p.beta * (*c);
      *c = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// softmax rows (one warp per row) and backward
// ---------------------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(const float* __restrict__ x, void* __restrict__ y, int ydt, long long rows, int cols) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= rows) return;
  const float* xr = x + row * cols;
  // one pass for (max, sum) with the online rescaling, one pass to write: the row is read twice, not three times
  float mx = -INFINITY, sum = 0.f;
  for (int c = lane; c < cols; c += 32) {
    const float v = xr[c];
    const float m2 = fmaxf(mx, v);
    sum = sum * __expf(mx - m2) + __expf(v - m2);
    mx = m2;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float mo = __shfl_xor_sync(0xffffffffu, mx, o), so = __shfl_xor_sync(0xffffffffu, sum, o);
    const float m2 = fmaxf(mx, mo);     // lanes without elements carry (-inf, 0): avoid exp(-inf - -inf)
    sum = sum * (mx == m2 ? 1.f : __expf(mx - m2)) + so * (mo == m2 ? 1.f : __expf(mo - m2));
    mx = m2;
  }
  const float inv = 1.f / sum;
  for (int c = lane; c < cols; c += 32) st_any(y, row * cols + c, ydt, __expf(xr[c] - mx) * inv);
}
__global__ void softmax_rows_bwd_kernel(const void* __restrict__ y, int ydt, const float* __restrict__ dy,
                                        void* __restrict__ dx, int dxdt, long long rows, int cols) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= rows) return;
  const float* dr = dy + row * cols;
  float dot = 0.f;
  for (int c = lane; c < cols; c += 32) dot += ld_any(y, row * cols + c, ydt) * dr[c];
  dot = warp_sum(dot);
  for (int c = lane; c < cols; c += 32) st_any(dx, row * cols + c, dxdt, ld_any(y, row * cols + c, ydt) * (dr[c] - dot));
}

// dS = P * (dP - D) with the row constant D_i = sum_j dP_ij P_ij supplied by the caller as sum_c dO_ic O_ic (the same
// number, since O = P V): ONE pass over the [rows, cols] matrices instead of two.  16 bytes per thread.
__global__ void __launch_bounds__(256)
softmax_bwd_d_kernel(const void* __restrict__ y, int ydt, const void* __restrict__ dy, int dydt, const float* __restrict__ D,
                     void* __restrict__ dx, int dxdt, long long rows, int cols) {
  const long long total = rows * cols;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / cols;
    st_any(dx, i, dxdt, ld_any(y, i, ydt) * (ld_any(dy, i, dydt) - D[row]));
  }
}
// out[r] = sum_c a[r, c] * b[r, c]  (one warp per row)
__global__ void rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b, long long rows, int cols, float* __restrict__ out) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s = fmaf(a[row * cols + c], b[row * cols + c], s);
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}

// ---------------------------------------------------------------------------------------------
// permute3 (weight re-layout with cast)
// ---------------------------------------------------------------------------------------------
__global__ void permute3_kernel(const void* src, int sdt, void* dst, int ddt, long long D0, long long D1, long long D2,
                                long long s0, long long s1, long long s2, int flip1, const float* scale, long long ld_dst) {
  const long long total = D0 * D1 * D2;
  const float sc = scale ? *scale : 1.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i2 = i % D2;
    const long long r = i / D2;
    long long i1 = r % D1;
    const long long i0 = r / D1;
    if (flip1) i1 = D1 - 1 - i1;
    const long long o = i0 * ld_dst + (r % D1) * D2 + i2;
    st_any(dst, o, ddt, sc * ld_any(src, i0 * s0 + i1 * s1 + i2 * s2, sdt));
  }
}

// many permute3 jobs in one launch: chunk_prefix[j] = first 1024-element chunk of job j (n_jobs + 1 entries)
__global__ void __launch_bounds__(256)
pack_jobs_kernel(const dfcsa_pack_job_t* jobs, int n_jobs, const long long* chunk_prefix, long long total_chunks) {
  for (long long ch = blockIdx.x; ch < total_chunks; ch += gridDim.x) {
    int lo = 0, hi = n_jobs - 1;               // last job whose first chunk is <= ch
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (chunk_prefix[mid] <= ch) lo = mid; else hi = mid - 1;
    }
    const dfcsa_pack_job_t j = jobs[lo];
    const long long total = j.D0 * j.D1 * j.D2;
    const long long base = (ch - chunk_prefix[lo]) * 1024;
    const float sc0 = j.scale ? *j.scale : 1.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = base + u * 256 + threadIdx.x;
      if (i < total) {
        long long i0, i1, i2;
        if (total < (1LL << 31)) {        // every weight tensor of the network: two 32-bit divisions instead of four 64-bit ones
          const unsigned iu = static_cast<unsigned>(i), d2 = static_cast<unsigned>(j.D2), d1 = static_cast<unsigned>(j.D1);
          const unsigned r = iu / d2, q = r / d1;
          i2 = iu - r * d2; i1 = r - q * d1; i0 = q;
        } else {
          i2 = i % j.D2;
          const long long r = i / j.D2;
          i1 = r % j.D1;
          i0 = r / j.D1;
        }
        const long long i1s = j.flip1 ? j.D1 - 1 - i1 : i1;
        const float sc = j.row_scale ? sc0 * j.row_scale[i0] : sc0;
        st_any(j.dst, i0 * j.ld_dst + i1 * j.D2 + i2, j.dst_dtype, sc * ld_any(j.src, i0 * j.s0 + i1s * j.s1 + i2 * j.s2, j.src_dtype));
      }
    }
  }
}

}  // namespace

int conv_gemm_simt(const dfcsa_conv_params_t* p, cudaStream_t stream) {
  DFCSA_CHECK_ARG(p->n_seg >= 1 && p->n_seg <= 3, "conv_gemm_simt: n_seg must be 1..3");
  SimtConvArgs a{};
  a.p = *p;
  a.M = static_cast<long long>(p->B) * p->H * p->W;
  DFCSA_CHECK_ARG(a.M > 0 && p->N > 0, "conv_gemm_simt: empty problem");
  a.ktot = 0;
  for (int s = 0; s < p->n_seg; ++s)
    a.ktot += p->seg[s].channels * (p->seg[s].tap_mode == DFCSA_TAP_1x1 ? 1 : p->seg[s].tap_mode == DFCSA_TAP_3x3 ? 9 : 4);
  a.convt_co = p->out_mode == DFCSA_OUT_CONVT2x2 ? p->N / 4 : 0;
  dim3 grid(static_cast<unsigned>((a.M + TM - 1) / TM), (p->N + TN - 1) / TN);
  conv_simt_kernel<<<grid, 256, 0, stream>>>(a);
  DFCSA_LAUNCH_CHECK("conv_simt_kernel");
  return DFCSA_OK;
}

int conv_wgrad_simt(const dfcsa_wgrad_params_t* p, cudaStream_t stream) {
  SimtWgradArgs a{};
  a.p = *p;
  a.M = static_cast<long long>(p->B) * p->H * p->W;
  DFCSA_CHECK_ARG(a.M > 0 && p->N > 0 && p->C > 0, "conv_wgrad_simt: empty problem");
  a.taps = p->x_tap_mode == DFCSA_TAP_3x3 ? 9 : (p->dy_tap_mode == DFCSA_TAP_2x2S2 ? 4 : 1);
  a.c_tiles = (p->C + TN - 1) / TN;
  const int n_tiles = (p->N + TM - 1) / TM;
  const long long base = static_cast<long long>(n_tiles) * a.c_tiles * a.taps;
  long long splits = std::max<long long>(1, (4LL * num_sms() + base - 1) / base);
  splits = std::min<long long>(splits, (a.M + 255) / 256);
  splits = std::max<long long>(1, std::min<long long>(splits, 65535));
  a.pix_per_split = ((a.M + splits - 1) / splits + TK - 1) / TK * TK;
  splits = (a.M + a.pix_per_split - 1) / a.pix_per_split;
  dim3 grid(n_tiles, a.c_tiles * a.taps, static_cast<unsigned>(splits));
  wgrad_simt_kernel<<<grid, 256, 0, stream>>>(a);
  DFCSA_LAUNCH_CHECK("wgrad_simt_kernel");
  return DFCSA_OK;
}

}  // namespace dfcsa

using namespace dfcsa;

extern "C" int dfcsa_sgemm(const dfcsa_sgemm_params_t* p, void* stream) {
  DFCSA_CHECK_ARG(p && p->batch > 0 && p->M > 0 && p->N > 0 && p->K > 0, "dfcsa_sgemm: empty problem");
  DFCSA_CHECK_ARG(p->batch <= 65535, "dfcsa_sgemm: batch too large");
  dim3 grid((p->M + TM - 1) / TM, (p->N + TN - 1) / TN, p->batch);
  sgemm_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*p);
  DFCSA_LAUNCH_CHECK("sgemm_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_softmax_rows(const float* x, void* y, int y_dtype, int64_t rows, int32_t cols, void* stream) {
  DFCSA_CHECK_ARG(x && y && rows > 0 && cols > 0 && x != y, "dfcsa_softmax_rows: bad args");
  const long long blocks = (rows + 7) / 8;
  DFCSA_CHECK_ARG(blocks < (1LL << 31), "dfcsa_softmax_rows: too many rows");
  softmax_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, y_dtype, rows, cols);
  DFCSA_LAUNCH_CHECK("softmax_rows_kernel");
  return DFCSA_OK;
}
extern "C" int dfcsa_softmax_rows_bwd(const void* y, int y_dtype, const float* dy, void* dx, int dx_dtype, int64_t rows,
                                      int32_t cols, void* stream) {
  DFCSA_CHECK_ARG(y && dy && dx && rows > 0 && cols > 0, "dfcsa_softmax_rows_bwd: bad args");
  const long long blocks = (rows + 7) / 8;
  DFCSA_CHECK_ARG(blocks < (1LL << 31), "dfcsa_softmax_rows_bwd: too many rows");
  softmax_rows_bwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, y_dtype, dy, dx, dx_dtype, rows, cols);
  DFCSA_LAUNCH_CHECK("softmax_rows_bwd_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_softmax_rows_bwd_d(const void* y, int y_dtype, const void* dy, int dy_dtype, const float* D, void* dx,
                                        int dx_dtype, int64_t rows, int32_t cols, void* stream) {
  DFCSA_CHECK_ARG(y && dy && D && dx && rows > 0 && cols > 0, "dfcsa_softmax_rows_bwd_d: bad args");
  const long long total = rows * cols;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 32));
  softmax_bwd_d_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(y, y_dtype, dy, dy_dtype, D, dx, dx_dtype, rows, cols);
  DFCSA_LAUNCH_CHECK("softmax_bwd_d_kernel");
  return DFCSA_OK;
}
extern "C" int dfcsa_rowdot(const float* a, const float* b, int64_t rows, int32_t cols, float* out, void* stream) {
  DFCSA_CHECK_ARG(a && b && out && rows > 0 && cols > 0, "dfcsa_rowdot: bad args");
  const long long blocks = (rows + 7) / 8;
  DFCSA_CHECK_ARG(blocks < (1LL << 31), "dfcsa_rowdot: too many rows");
  rowdot_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, rows, cols, out);
  DFCSA_LAUNCH_CHECK("rowdot_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_pack_jobs(const dfcsa_pack_job_t* jobs_dev, int32_t n_jobs, const int64_t* chunk_prefix_dev,
                               int64_t total_chunks, void* stream) {
  DFCSA_CHECK_ARG(jobs_dev && chunk_prefix_dev && n_jobs > 0 && total_chunks > 0, "dfcsa_pack_jobs: bad args");
  const int blocks = static_cast<int>(std::min<long long>(total_chunks, 148 * 16));
  pack_jobs_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(jobs_dev, n_jobs, reinterpret_cast<const long long*>(chunk_prefix_dev),
                                                                          total_chunks);
  DFCSA_LAUNCH_CHECK("pack_jobs_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_permute3(const void* src, int src_dtype, void* dst, int dst_dtype,
                              int64_t D0, int64_t D1, int64_t D2, int64_t s0, int64_t s1, int64_t s2,
                              int flip1, const float* scale, int64_t ld_dst, void* stream) {
  const long long total = D0 * D1 * D2;
  DFCSA_CHECK_ARG(total > 0, "dfcsa_permute3: empty");
  if (ld_dst <= 0) ld_dst = D1 * D2;
  DFCSA_CHECK_ARG(ld_dst >= D1 * D2, "dfcsa_permute3: ld_dst smaller than a destination row");
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
  permute3_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, src_dtype, dst, dst_dtype, D0, D1, D2, s0, s1, s2, flip1, scale, ld_dst);
  DFCSA_LAUNCH_CHECK("permute3_kernel");
  return DFCSA_OK;
}
