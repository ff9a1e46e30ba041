// Pooled self-attention core for small maps (N = P*P <= 32, i.e. pool_size <= 5; the DFC-SA-Res-Block configs use
// P = 4): softmax(q k^T) v and its backward, one CTA per image, everything in fp32.  At N = 16 the three products are
// a few KFLOP per image - tensor cores are pointless - but as separate GEMM / softmax launches they were 8 kernels
// per block and direction; here they are one.  Reference: models/unet_dfc_sa_res.py:28-34.
//   qkv  : [B*N, ld] fp32 rows (q[0:Cq] | k[Cq:2Cq] | v[2Cq:2Cq+C])
//   attn : [B, N, N] fp32 (saved for backward),  o / d_o : [B*N, C] fp32,  dqkv : [B*N, ld] fp32 (same column layout)
#include "common.cuh"

namespace dfcsa {
namespace {

constexpr int kMaxN = 32;

__global__ void __launch_bounds__(256)
attn_small_fwd_kernel(const float* __restrict__ qkv, long long ld, int N, int Cq, int C, float* __restrict__ attn,
                      float* __restrict__ o) {
  __shared__ float sS[kMaxN][kMaxN + 1];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* base = qkv + static_cast<long long>(b) * N * ld;
  for (int idx = tid; idx < N * N; idx += blockDim.x) {
    const int i = idx / N, j = idx % N;
    const float* q = base + static_cast<long long>(i) * ld;
    const float* k = base + static_cast<long long>(j) * ld + Cq;
    float s = 0.f;
    for (int c = 0; c < Cq; ++c) s = fmaf(q[c], k[c], s);
    sS[i][j] = s;
  }
  __syncthreads();
  if (tid < N) {
    float mx = -INFINITY;
    for (int j = 0; j < N; ++j) mx = fmaxf(mx, sS[tid][j]);
    float sum = 0.f;
    for (int j = 0; j < N; ++j) { const float e = expf(sS[tid][j] - mx); sS[tid][j] = e; sum += e; }
    const float inv = 1.f / sum;
    for (int j = 0; j < N; ++j) {
      const float p = sS[tid][j] * inv;
      sS[tid][j] = p;
      attn[(static_cast<long long>(b) * N + tid) * N + j] = p;
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += blockDim.x) {
    float acc[kMaxN];
#pragma unroll
    for (int i = 0; i < kMaxN; ++i) acc[i] = 0.f;
    for (int j = 0; j < N; ++j) {
      const float v = base[static_cast<long long>(j) * ld + 2 * Cq + c];
#pragma unroll
      for (int i = 0; i < kMaxN; ++i) if (i < N) acc[i] = fmaf(sS[i][j], v, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < kMaxN; ++i) if (i < N) o[(static_cast<long long>(b) * N + i) * C + c] = acc[i];
  }
}

__global__ void __launch_bounds__(256)
attn_small_bwd_kernel(const float* __restrict__ qkv, long long ld, const float* __restrict__ attn,
                      const float* __restrict__ d_o, int N, int Cq, int C, float* __restrict__ dqkv, float* dbq, float* dbk, float* dbv) {
  __shared__ float sA[kMaxN][kMaxN + 1];    // probabilities
  __shared__ float sD[kMaxN][kMaxN + 1];    // d attn, then dS
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* base = qkv + static_cast<long long>(b) * N * ld;
  const float* dob = d_o + static_cast<long long>(b) * N * C;
  float* dbase = dqkv + static_cast<long long>(b) * N * ld;
  for (int idx = tid; idx < N * N; idx += blockDim.x) sA[idx / N][idx % N] = attn[static_cast<long long>(b) * N * N + idx];
  // d attn[i][j] = sum_c do[i][c] v[j][c]: one warp per (i, j) pair, lanes over c
  for (int idx = warp; idx < N * N; idx += blockDim.x / 32) {
    const int i = idx / N, j = idx % N;
    const float* dr = dob + static_cast<long long>(i) * C;
    const float* vr = base + static_cast<long long>(j) * ld + 2 * Cq;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(dr[c], vr[c], s);
    s = warp_sum(s);
    if (lane == 0) sD[i][j] = s;
  }
  __syncthreads();
  if (tid < N) {   // dS = A * (dA - rowsum(dA * A))
    float dot = 0.f;
    for (int j = 0; j < N; ++j) dot = fmaf(sD[tid][j], sA[tid][j], dot);
    for (int j = 0; j < N; ++j) sD[tid][j] = sA[tid][j] * (sD[tid][j] - dot);
  }
  // dv[j][c] = sum_i A[i][j] do[i][c]   (does not need dS: overlaps with the row pass above for other threads)
  for (int c = tid; c < C; c += blockDim.x) {
    float acc[kMaxN];
#pragma unroll
    for (int j = 0; j < kMaxN; ++j) acc[j] = 0.f;
    for (int i = 0; i < N; ++i) {
      const float d = dob[static_cast<long long>(i) * C + c];
#pragma unroll
      for (int j = 0; j < kMaxN; ++j) if (j < N) acc[j] = fmaf(sA[i][j], d, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < kMaxN; ++j) if (j < N) dbase[static_cast<long long>(j) * ld + 2 * Cq + c] = acc[j];
    if (dbv != nullptr) {      // bias gradient of the value conv: column sum of dv over tokens (and, by atomics, images)
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxN; ++j) if (j < N) t += acc[j];
      atomicAdd(dbv + c, t);
    }
  }
  __syncthreads();
  // dq[i][c] = sum_j dS[i][j] k[j][c];  dk[j][c] = sum_i dS[i][j] q[i][c]
  for (int idx = tid; idx < N * Cq; idx += blockDim.x) {
    const int r = idx / Cq, c = idx % Cq;
    float dq = 0.f, dk = 0.f;
    for (int t = 0; t < N; ++t) {
      dq = fmaf(sD[r][t], base[static_cast<long long>(t) * ld + Cq + c], dq);
      dk = fmaf(sD[t][r], base[static_cast<long long>(t) * ld + c], dk);
    }
    dbase[static_cast<long long>(r) * ld + c] = dq;
    dbase[static_cast<long long>(r) * ld + Cq + c] = dk;
    if (dbq != nullptr) { atomicAdd(dbq + c, dq); atomicAdd(dbk + c, dk); }
  }
}

}  // namespace
}  // namespace dfcsa

using namespace dfcsa;

extern "C" int dfcsa_attn_small_fwd(const float* qkv, int64_t ld, int32_t B, int32_t N, int32_t Cq, int32_t C, float* attn,
                                    float* o, void* stream) {
  DFCSA_CHECK_ARG(qkv && attn && o && B > 0 && N > 0 && N <= kMaxN && Cq > 0 && C > 0 && ld >= 2 * Cq + C,
                  "dfcsa_attn_small_fwd: bad args (N must be <= 32)");
  attn_small_fwd_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(qkv, ld, N, Cq, C, attn, o);
  DFCSA_LAUNCH_CHECK("attn_small_fwd_kernel");
  return DFCSA_OK;
}

extern "C" int dfcsa_attn_small_bwd(const float* qkv, int64_t ld, const float* attn, const float* d_o, int32_t B, int32_t N,
                                    int32_t Cq, int32_t C, float* dqkv, float* dbq, float* dbk, float* dbv, void* stream) {
  DFCSA_CHECK_ARG(qkv && attn && d_o && dqkv && B > 0 && N > 0 && N <= kMaxN && Cq > 0 && C > 0 && ld >= 2 * Cq + C,
                  "dfcsa_attn_small_bwd: bad args (N must be <= 32)");
  DFCSA_CHECK_ARG((dbq == nullptr) == (dbk == nullptr), "dfcsa_attn_small_bwd: dbq and dbk go together");
  attn_small_bwd_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(qkv, ld, attn, d_o, N, Cq, C, dqkv, dbq, dbk, dbv);
  DFCSA_LAUNCH_CHECK("attn_small_bwd_kernel");
  return DFCSA_OK;
}
