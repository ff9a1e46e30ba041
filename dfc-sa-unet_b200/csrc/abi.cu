// C-ABI plumbing: error strings, TMA descriptor encoding through the driver entry point, backend dispatch.
#include "common.cuh"
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <string.h>
#include <mutex>

namespace dfcsa {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return DFCSA_ERR_CUDA;
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      sms = n;
    else
      sms = 148;
  }
  return sms;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

int encode_tensor_map(CUtensorMap* map, int dtype, int rank, const void* base, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
  if (!g_encode) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return DFCSA_ERR_CUDA; }
  CUtensorMapDataType dt = dtype == DFCSA_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                         : dtype == DFCSA_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                              : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint64_t gdims[5]; cuuint64_t gstr[4]; cuuint32_t gbox[5]; cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) { gdims[i] = dims[i]; gbox[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = g_encode(map, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdims, gstr, gbox, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u] stride0 %llu",
              static_cast<int>(r), rank, (unsigned long long)gdims[0], (unsigned long long)(rank > 1 ? gdims[1] : 0),
              (unsigned long long)(rank > 2 ? gdims[2] : 0), (unsigned long long)(rank > 3 ? gdims[3] : 0),
              (unsigned long long)(rank > 4 ? gdims[4] : 0), gbox[0], rank > 1 ? gbox[1] : 0, rank > 2 ? gbox[2] : 0,
              rank > 3 ? gbox[3] : 0, rank > 4 ? gbox[4] : 0, (unsigned long long)(rank > 1 ? gstr[0] : 0));
    return DFCSA_ERR_CUDA;
  }
  return DFCSA_OK;
}

int conv_gemm_tc(const dfcsa_conv_params_t* p, cudaStream_t stream);
int conv_gemm_simt(const dfcsa_conv_params_t* p, cudaStream_t stream);
int conv_wgrad_tc(const dfcsa_wgrad_params_t* p, cudaStream_t stream);
int conv_wgrad_simt(const dfcsa_wgrad_params_t* p, cudaStream_t stream);
int conv_gemm_small(const dfcsa_conv_params_t* p, cudaStream_t stream, int* rc_out);
int conv_wgrad_small(const dfcsa_wgrad_params_t* p, cudaStream_t stream, int* rc_out);

}  // namespace dfcsa

using namespace dfcsa;

extern "C" int dfcsa_version(void) { return DFCSA_VERSION; }
extern "C" const char* dfcsa_last_error(void) { return g_err; }

extern "C" int dfcsa_wgrad_plan(int64_t items, int64_t pix_blocks, int32_t sms, int32_t* splits, int64_t* blocks_per_split) {
  DFCSA_CHECK_ARG(items > 0 && pix_blocks > 0 && sms > 0 && splits && blocks_per_split, "dfcsa_wgrad_plan: bad args");
  long long bps = 0;
  *splits = wgrad_pick_splits(items, pix_blocks, sms, 0, &bps);
  *blocks_per_split = bps;
  return DFCSA_OK;
}

extern "C" int dfcsa_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

extern "C" int dfcsa_conv_gemm(const dfcsa_conv_params_t* p, int backend, void* stream) {
  DFCSA_CHECK_ARG(p != nullptr, "dfcsa_conv_gemm: null params");
  DFCSA_CHECK_ARG(p->out != nullptr && p->w != nullptr, "dfcsa_conv_gemm: null out / w");
  for (int s = 0; s < p->n_seg && s < 3; ++s)
    DFCSA_CHECK_ARG(p->seg[s].ptr != nullptr, "dfcsa_conv_gemm: null segment %d", s);
  if (backend == DFCSA_BACKEND_TC) return conv_gemm_tc(p, static_cast<cudaStream_t>(stream));
  if (backend == DFCSA_BACKEND_SIMT) {
    DFCSA_CHECK_ARG(p->bn == nullptr, "dfcsa_conv_gemm: the BatchNorm fold needs DFCSA_BACKEND_TC (call dfcsa_bn_finalize)");
    DFCSA_CHECK_ARG(p->epi == nullptr || p->epi->mode == DFCSA_EPI_NONE,
                    "dfcsa_conv_gemm: the fused gate-mix / residual epilogue needs DFCSA_BACKEND_TC (call dfcsa_gate_mix_fwd / dfcsa_sum_out_fwd)");
    int rc = DFCSA_OK;   // the tiny-K / tiny-N layers (first conv, final conv) have bandwidth-shaped kernels of their own
    if (conv_gemm_small(p, static_cast<cudaStream_t>(stream), &rc)) return rc;
    return conv_gemm_simt(p, static_cast<cudaStream_t>(stream));
  }
  set_error("dfcsa_conv_gemm: unknown backend %d", backend);
  return DFCSA_ERR_BAD_ARG;
}

extern "C" int dfcsa_conv_wgrad(const dfcsa_wgrad_params_t* p, int backend, void* stream) {
  DFCSA_CHECK_ARG(p != nullptr, "dfcsa_conv_wgrad: null params");
  DFCSA_CHECK_ARG(p->x != nullptr && p->dy != nullptr && p->dw != nullptr, "dfcsa_conv_wgrad: null pointer");
  DFCSA_CHECK_ARG(!(p->x_tap_mode == DFCSA_TAP_3x3 && p->dy_tap_mode != DFCSA_TAP_1x1) && p->x_tap_mode != DFCSA_TAP_2x2S2 &&
                  p->dy_tap_mode != DFCSA_TAP_3x3, "dfcsa_conv_wgrad: unsupported tap mode combination");
  if (p->dy2 != nullptr) {
    DFCSA_CHECK_ARG(p->x_tap_mode == DFCSA_TAP_1x1 && p->dy_tap_mode == DFCSA_TAP_1x1 && p->dw2 != nullptr && p->N2 > 0 &&
                    p->c_begin2 >= 0 && p->c_begin2 < p->C, "dfcsa_conv_wgrad: bad second gradient (1x1 taps only)");
  }
  if (backend == DFCSA_BACKEND_TC) return conv_wgrad_tc(p, static_cast<cudaStream_t>(stream));
  if (backend == DFCSA_BACKEND_SIMT) {
    dfcsa_wgrad_params_t a = *p;
    a.dy2 = nullptr; a.dw2 = nullptr;
    int rc = DFCSA_OK;
    if (!conv_wgrad_small(&a, static_cast<cudaStream_t>(stream), &rc)) rc = conv_wgrad_simt(&a, static_cast<cudaStream_t>(stream));
    if (rc != DFCSA_OK || p->dy2 == nullptr) return rc;
    // the fp32 path runs the second gradient as a launch of its own over the channel slice [c_begin2, C) of x
    dfcsa_wgrad_params_t b = a;
    b.x = static_cast<const char*>(p->x) + static_cast<size_t>(p->c_begin2) * dtype_size(p->x_dtype);
    b.C = p->C - p->c_begin2;
    b.dy = p->dy2; b.ld_dy = p->ld_dy2; b.N = p->N2;
    b.dw = p->dw2; b.ld_dw = p->ld_dw2; b.alpha = p->alpha2;
    if (conv_wgrad_small(&b, static_cast<cudaStream_t>(stream), &rc)) return rc;
    return conv_wgrad_simt(&b, static_cast<cudaStream_t>(stream));
  }
  set_error("dfcsa_conv_wgrad: unknown backend %d", backend);
  return DFCSA_ERR_BAD_ARG;
}
