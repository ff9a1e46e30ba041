// Bring-up ladder for the tcgen05 / TMA building blocks in csrc/common.cuh.  Each probe runs in its own process
// (a CUDA fault kills the context):   build/tc_probe <id>
#include "../dfc-sa-unet_b200/csrc/common.cuh"
#include <vector>
#include <cstdlib>
#include <cmath>
#include <cstdarg>

namespace dfcsa {
static char g_e[512];
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_e, sizeof g_e, fmt, ap); va_end(ap); }
int cuda_fail(cudaError_t e, const char* what) { set_error("%s: %s", what, cudaGetErrorString(e)); return 2; }
}
using namespace dfcsa;
#include <cudaTypedefs.h>

static int encode(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                  const uint64_t* strides, const uint32_t* box, bool sw) {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, dt, rank, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
  return r;
}

__global__ void k_trap(int x) { if (x == 1) __trap(); }
__global__ void k_tmem() {
  __shared__ uint32_t base;
  if (threadIdx.x < 32) tmem_alloc(&base, 64);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x < 32) tmem_dealloc(base, 64);
}
__global__ void k_mbar(int* out) {
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) mbar_arrive(&bar);
  mbar_wait(&bar, 0);
  if (threadIdx.x == 0) *out = 7;
}
__global__ void k_tma2d(const __grid_constant__ CUtensorMap m, __half* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* s = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_arrive_expect_tx(&bar, 64 * 32 * 2); tma_load_2d(s, &m, &bar, 0, 0); }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) out[i] = reinterpret_cast<__half*>(s)[i];
}
__global__ void k_tma5d(const __grid_constant__ CUtensorMap m, __half* out, int c1, int c2) {
  extern __shared__ uint8_t raw[];
  uint8_t* s = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_arrive_expect_tx(&bar, 128 * 128); tma_load_5d(s, &m, &bar, 0, c1, c2, 0, 0); }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) out[i] = reinterpret_cast<__half*>(s)[i];
}
// A [128 x 64] K-major, B [N x 64] K-major written by threads in the swizzled layout; D -> out [128 x N]
__global__ void k_mma_manual(const __half* A, const __half* B, float* out, int N, uint32_t idesc, int use_ld) {
  extern __shared__ uint8_t raw[];
  uint8_t* s = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tbase;
  uint8_t* sa = s; uint8_t* sb = s + 16384;
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    int r = i / 64, k = i % 64;
    int off = r * 128 + ((((k * 2) / 16) ^ (r % 8)) * 16) + (k * 2) % 16;
    *reinterpret_cast<__half*>(sa + off) = A[i];
  }
  for (int i = threadIdx.x; i < N * 64; i += blockDim.x) {
    int r = i / 64, k = i % 64;
    int off = r * 128 + ((((k * 2) / 16) ^ (r % 8)) * 16) + (k * 2) % 16;
    *reinterpret_cast<__half*>(sb + off) = B[i];
  }
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tbase, 256);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tbase;
  if (threadIdx.x == 0) {
    for (int k = 0; k < 4; ++k) {
      uint64_t da = umma_smem_desc(smem_u32(sa) + k * 32, 16, 1024);
      uint64_t db = umma_smem_desc(smem_u32(sb) + k * 32, 16, 1024);
      umma_f16(tb, da, db, idesc, k != 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (use_ld) {
    const int w = threadIdx.x / 32, lane = threadIdx.x % 32;
    for (int ch = 0; ch < N / 32; ++ch) {
      uint32_t v[32];
      tmem_ld_32x32(tb + ch * 32 + (uint32_t(w * 32) << 16), v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[(w * 32 + lane) * N + ch * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tb, 256); }
}

// descriptor with an explicit matrix base offset (bits [49,52)): needed (or not) when the start address is not
// aligned to the 1024-byte swizzle repeat - probes 9 / 10 measure which
__device__ __forceinline__ uint64_t desc_bo(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t bo) {
  return umma_smem_desc(addr, lbo, sbo) | (static_cast<uint64_t>(bo & 7) << 49);
}
// probe 9: K-major A tile of 144 rows in the TMA SW128 layout; the MMA reads rows [s, s+128) through a descriptor
// whose start address is shifted by s*128 bytes (a 3x3 tap shift inside one halo tile).  out[m,n] = A[m+s,:].B[n,:]
__global__ void k_mma_kshift(const __half* A, const __half* B, float* out, int N, uint32_t idesc, int s, int bo_mode, int sbo = 1024) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tbase;
  uint8_t* sa = sm; uint8_t* sb = sm + 176 * 128;
  for (int i = threadIdx.x; i < 176 * 64; i += blockDim.x) {
    int r = i / 64, k = i % 64;
    int off = r * 128 + ((((k * 2) / 16) ^ (r % 8)) * 16) + (k * 2) % 16;
    *reinterpret_cast<__half*>(sa + off) = A[i];
  }
  for (int i = threadIdx.x; i < N * 64; i += blockDim.x) {
    int r = i / 64, k = i % 64;
    int off = r * 128 + ((((k * 2) / 16) ^ (r % 8)) * 16) + (k * 2) % 16;
    *reinterpret_cast<__half*>(sb + off) = B[i];
  }
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tbase, 256);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tbase;
  if (threadIdx.x == 0) {
    for (int k = 0; k < 4; ++k) {
      const uint32_t a_addr = smem_u32(sa) + s * 128 + k * 32;
      uint64_t da = desc_bo(a_addr, 16, sbo, bo_mode ? (a_addr >> 7) & 7 : 0);
      uint64_t db = umma_smem_desc(smem_u32(sb) + k * 32, 16, 1024);
      umma_f16(tb, da, db, idesc, k != 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int w = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int ch = 0; ch < N / 32; ++ch) {
    uint32_t v[32];
    tmem_ld_32x32(tb + ch * 32 + (uint32_t(w * 32) << 16), v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(w * 32 + lane) * N + ch * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tb, 256); }
}
// probe 10: MN-major operands (rows = K index = pixel, 64 contiguous M/N elements per 128-byte row, SW128).
// A: 80 K-rows x 128 M (two 64-wide blocks LBO apart), read from K-row s on;  B: 64 K-rows x 64 N unshifted.
// out[m,n] = sum_{k<64} A[k+s, m] * B[k, n]
__global__ void k_mma_mnshift(const __half* A, const __half* B, float* out, uint32_t idesc, int s, int bo_mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tbase;
  const int RA = 80;
  uint8_t* sa = sm; uint8_t* sb = sm + 2 * RA * 128;
  for (int i = threadIdx.x; i < RA * 128; i += blockDim.x) {       // A[k][m]
    int k = i / 128, m = i % 128;
    int blk = m / 64, mm = m % 64;
    int off = blk * RA * 128 + k * 128 + ((((mm * 2) / 16) ^ (k % 8)) * 16) + (mm * 2) % 16;
    *reinterpret_cast<__half*>(sa + off) = A[i];
  }
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {        // B[k][n]
    int k = i / 64, n = i % 64;
    int off = k * 128 + ((((n * 2) / 16) ^ (k % 8)) * 16) + (n * 2) % 16;
    *reinterpret_cast<__half*>(sb + off) = B[i];
  }
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tbase, 64);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tbase;
  if (threadIdx.x == 0) {
    for (int k = 0; k < 4; ++k) {
      const uint32_t a_addr = smem_u32(sa) + s * 128 + k * 2048;
      uint64_t da = desc_bo(a_addr, RA * 128, 1024, bo_mode ? (a_addr >> 7) & 7 : 0);
      uint64_t db = umma_smem_desc(smem_u32(sb) + k * 2048, 64 * 128, 1024);
      umma_f16(tb, da, db, idesc, k != 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int w = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int ch = 0; ch < 2; ++ch) {
    uint32_t v[32];
    tmem_ld_32x32(tb + ch * 32 + (uint32_t(w * 32) << 16), v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(w * 32 + lane) * 64 + ch * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tb, 64); }
}

static void report(const char* name) {
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s: %s\n", name, e == cudaSuccess ? "ran OK" : cudaGetErrorString(e));
}

int main(int argc, char** argv) {
  int id = argc > 1 ? atoi(argv[1]) : 0;
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  printf("device %s cc %d.%d, probe %d\n", pr.name, pr.major, pr.minor, id);
  if (id == 0) { k_trap<<<1, 32>>>(1); report("trap"); return 0; }
  if (id == 1) { k_tmem<<<1, 128>>>(); report("tmem alloc/dealloc"); return 0; }
  if (id == 2) {
    int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
    k_mbar<<<1, 64>>>(d); report("mbarrier");
    int h = 0; cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost); printf("  value %d (expect 7)\n", h); return 0;
  }
  if (id == 3 || id == 4) {
    // source: [rows=300][cols=64] half, value = row*64+col (mod 2048 to stay exact)
    const int R = 300, Cc = 64;
    std::vector<__half> h(R * Cc);
    for (int i = 0; i < R * Cc; ++i) h[i] = __float2half((float)(i % 2048));
    __half* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    __half* o; cudaMalloc(&o, 128 * 64 * 2); cudaMemset(o, 0, 128 * 64 * 2);
    CUtensorMap m;
    if (id == 3) {
      uint64_t dims[2] = {64, (uint64_t)R}; uint64_t str[1] = {128}; uint32_t box[2] = {32, 64};
      if (encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, str, box, false)) return 1;
      cudaFuncSetAttribute(k_tma2d, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
      k_tma2d<<<1, 128, 40000>>>(m, o); report("tma 2d noswizzle");
      std::vector<__half> r(64 * 32); cudaMemcpy(r.data(), o, r.size() * 2, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int rr = 0; rr < 64; ++rr) for (int c = 0; c < 32; ++c)
        if (__half2float(r[rr * 32 + c]) != (float)((rr * 64 + c) % 2048)) ++bad;
      printf("  mismatches %d\n", bad);
    } else {
      // view as (C=64, W=10, H=30, B=1, 1): box (64, 8, 16, 1, 1) at w=-1,h=-1 -> 128 rows, halo zero-filled
      uint64_t dims[5] = {64, 10, 30, 1, 1}; uint64_t str[4] = {128, 1280, 38400, 38400}; uint32_t box[5] = {64, 8, 16, 1, 1};
      if (encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, d, dims, str, box, true)) return 1;
      cudaFuncSetAttribute(k_tma5d, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
      k_tma5d<<<1, 128, 40000>>>(m, o, -1, -1); report("tma 5d swizzle128");
      std::vector<__half> r(128 * 64); cudaMemcpy(r.data(), o, r.size() * 2, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int row = 0; row < 128; ++row) for (int k = 0; k < 64; ++k) {
        int w = row % 8 - 1, hh = row / 8 - 1;
        float exp = (w < 0 || hh < 0 || w >= 10 || hh >= 30) ? 0.f : (float)(((hh * 10 + w) * 64 + k) % 2048);
        int off = row * 128 + ((((k * 2) / 16) ^ (row % 8)) * 16) + (k * 2) % 16;
        if (__half2float(r[off / 2]) != exp) { if (bad < 5) printf("  row %d k %d got %f exp %f\n", row, k, __half2float(r[off / 2]), exp); ++bad; }
      }
      printf("  mismatches %d\n", bad);
    }
    return 0;
  }
  if (id >= 5 && id <= 8) {
    const int N = (id == 6) ? 64 : 128;
    const int use_ld = id != 8;
    const int afmt = 0, bfmt = (id == 7) ? 1 : 0;
    std::vector<__half> A(128 * 64), B(N * 64);
    std::vector<float> Af(128 * 64), Bf(N * 64);
    srand(1);
    for (size_t i = 0; i < A.size(); ++i) { float v = (rand() % 17 - 8) / 8.f; A[i] = __float2half(v); Af[i] = v; }
    std::vector<__nv_bfloat16> Bb(N * 64);
    for (size_t i = 0; i < B.size(); ++i) { float v = (rand() % 13 - 6) / 4.f; B[i] = __float2half(v); Bb[i] = __float2bfloat16(v); Bf[i] = v; }
    __half *dA, *dB; float* dO;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, 128 * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, bfmt ? (void*)Bb.data() : (void*)B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dO, 0, 128 * N * 4);
    uint32_t idesc = umma_idesc_f16(128, N, afmt, bfmt, 0, 0);
    printf("  idesc 0x%08x\n", idesc);
    cudaFuncSetAttribute(k_mma_manual, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    k_mma_manual<<<1, 128, 60000>>>(dA, dB, dO, N, idesc, use_ld);
    report(id == 5 ? "mma N=128" : id == 6 ? "mma N=64" : id == 7 ? "mma fp16 x bf16" : "mma no tcgen05.ld");
    std::vector<float> O(128 * N); cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0; double maxerr = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
      float ref = 0; for (int k = 0; k < 64; ++k) ref += Af[m * 64 + k] * Bf[n * 64 + k];
      double e = fabs(ref - O[m * N + n]); if (e > maxerr) maxerr = e; if (e > 1e-3) ++bad;
    }
    printf("  mismatches %d maxerr %g\n", bad, maxerr);
    return 0;
  }
  if (id == 9 || id == 10) {
    const int sft = argc > 2 ? atoi(argv[2]) : 1, bo = argc > 3 ? atoi(argv[3]) : 0;
    srand(3);
    if (id == 9) {
      const int N = 64;
      std::vector<__half> A(176 * 64), B(N * 64); std::vector<float> Af(A.size()), Bf(B.size());
      const int sbo = argc > 4 ? atoi(argv[4]) : 1024;      // 1280: 8-row groups 10 rows apart (8-wide patch rows with a halo)
      for (size_t i = 0; i < A.size(); ++i) { float v = (rand() % 17 - 8) / 8.f; A[i] = __float2half(v); Af[i] = v; }
      for (size_t i = 0; i < B.size(); ++i) { float v = (rand() % 13 - 6) / 4.f; B[i] = __float2half(v); Bf[i] = v; }
      __half *dA, *dB; float* dO;
      cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, 128 * N * 4);
      cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
      cudaMemset(dO, 0, 128 * N * 4);
      cudaFuncSetAttribute(k_mma_kshift, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
      k_mma_kshift<<<1, 128, 60000>>>(dA, dB, dO, N, umma_idesc_f16(128, N, 0, 0, 0, 0), sft, bo, sbo);
      report("mma K-major row shift");
      std::vector<float> O(128 * N); cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        const int row = (m / 8) * (sbo / 128) + m % 8 + sft;
        float ref = 0; for (int k = 0; k < 64; ++k) ref += Af[row * 64 + k] * Bf[n * 64 + k];
        if (fabs(ref - O[m * N + n]) > 1e-3) ++bad;
      }
      printf("  K-major shift %d base_offset_mode %d sbo %d: mismatches %d of %d\n", sft, bo, sbo, bad, 128 * N);
    } else {
      std::vector<__half> A(80 * 128), B(64 * 64); std::vector<float> Af(A.size()), Bf(B.size());
      for (size_t i = 0; i < A.size(); ++i) { float v = (rand() % 17 - 8) / 8.f; A[i] = __float2half(v); Af[i] = v; }
      for (size_t i = 0; i < B.size(); ++i) { float v = (rand() % 13 - 6) / 4.f; B[i] = __float2half(v); Bf[i] = v; }
      __half *dA, *dB; float* dO;
      cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, 128 * 64 * 4);
      cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
      cudaMemset(dO, 0, 128 * 64 * 4);
      cudaFuncSetAttribute(k_mma_mnshift, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
      k_mma_mnshift<<<1, 128, 60000>>>(dA, dB, dO, umma_idesc_f16(128, 64, 0, 0, 1, 1), sft, bo);
      report("mma MN-major K-row shift");
      std::vector<float> O(128 * 64); cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
        float ref = 0; for (int k = 0; k < 64; ++k) ref += Af[(k + sft) * 128 + m] * Bf[k * 64 + n];
        if (fabs(ref - O[m * 64 + n]) > 1e-3) ++bad;
      }
      printf("  MN-major shift %d base_offset_mode %d: mismatches %d of %d\n", sft, bo, bad, 128 * 64);
    }
    return 0;
  }
  return 0;
}
