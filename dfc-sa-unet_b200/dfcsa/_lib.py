"""ctypes binding of libdfcsa.so (include/dfcsa.h).

The library is the product: there is no CPU or PyTorch fallback here.  If the shared object is missing the
import of any compute entry point raises, and every non-zero return code becomes a RuntimeError carrying
dfcsa_last_error().
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libdfcsa.so")

F32, F16, BF16 = 0, 1, 2
TAP_1x1, TAP_3x3, TAP_2x2S2 = 0, 1, 2
OUT_DIRECT, OUT_CONVT2x2 = 0, 1
BACKEND_TC, BACKEND_SIMT = 0, 1

_TORCH2DT = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}


def dt(t):
    return _TORCH2DT[t.dtype]


class Seg(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_int64), ("channels", C.c_int32), ("tap_mode", C.c_int32)]


class BnFold(C.Structure):        # dfcsa_bn_fold_t
    _fields_ = [("gamma", C.c_void_p), ("beta", C.c_void_p), ("conv_bias", C.c_void_p),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p),
                ("momentum", C.c_float), ("eps", C.c_float), ("count", C.c_int64),
                ("scale", C.c_void_p), ("shift", C.c_void_p), ("mean", C.c_void_p), ("invstd", C.c_void_p),
                ("ticket", C.c_void_p), ("channels", C.c_int32), ("pad_", C.c_int32)]


EPI_NONE, EPI_GATE_MIX, EPI_RESIDUAL = 0, 1, 2


class ConvEpi(C.Structure):       # dfcsa_conv_epi_t
    _fields_ = [("mode", C.c_int32), ("pad_", C.c_int32), ("p", C.c_void_p), ("q", C.c_void_p), ("ld", C.c_int64),
                ("scale", C.c_void_p)]


class ConvParams(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("n_seg", C.c_int32),
        ("seg", Seg * 3),
        ("src_dtype", C.c_int32), ("w_dtype", C.c_int32),
        ("w", C.c_void_p), ("N", C.c_int32), ("out_dtype", C.c_int32),
        ("out", C.c_void_p), ("ld_out", C.c_int64),
        ("out_mode", C.c_int32), ("accumulate", C.c_int32),
        ("bias", C.c_void_p), ("stats", C.c_void_p),
        ("shadow", C.c_void_p), ("ld_shadow", C.c_int64),
        ("act", C.c_int32), ("act_cols", C.c_int32), ("stats_cols", C.c_int32), ("pad_", C.c_int32),
        ("bn", C.c_void_p), ("epi", C.c_void_p),
    ]


class WgradParams(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("x", C.c_void_p), ("ld_x", C.c_int64), ("C", C.c_int32), ("x_dtype", C.c_int32), ("x_tap_mode", C.c_int32),
        ("dy", C.c_void_p), ("ld_dy", C.c_int64), ("N", C.c_int32), ("dy_dtype", C.c_int32), ("dy_tap_mode", C.c_int32),
        ("dw", C.c_void_p), ("ld_dw", C.c_int64),
        ("alpha", C.c_void_p),
        ("dy2", C.c_void_p), ("ld_dy2", C.c_int64), ("N2", C.c_int32), ("c_begin2", C.c_int32),
        ("dw2", C.c_void_p), ("ld_dw2", C.c_int64), ("alpha2", C.c_void_p),
    ]


class SgemmParams(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("A", C.c_void_p), ("a_b", C.c_int64), ("a_m", C.c_int64), ("a_k", C.c_int64),
        ("B", C.c_void_p), ("b_b", C.c_int64), ("b_k", C.c_int64), ("b_n", C.c_int64),
        ("C", C.c_void_p), ("c_b", C.c_int64), ("c_m", C.c_int64), ("c_n", C.c_int64),
        ("bias_n", C.c_void_p), ("bias_m", C.c_void_p),
        ("alpha", C.c_float), ("beta", C.c_float),
    ]


class BgemmParams(C.Structure):
    _fields_ = [("batch", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
                ("A", C.c_void_p), ("a_b", C.c_int64), ("ld_a", C.c_int64), ("a_mn_major", C.c_int32),
                ("B", C.c_void_p), ("b_b", C.c_int64), ("ld_b", C.c_int64), ("b_mn_major", C.c_int32),
                ("ab_dtype", C.c_int32),
                ("C", C.c_void_p), ("c_b", C.c_int64), ("ld_c", C.c_int64), ("c_dtype", C.c_int32),
                ("epi_mode", C.c_int32), ("rowstat", C.c_void_p), ("rowvec", C.c_void_p),
                ("aux", C.c_void_p), ("aux_dtype", C.c_int32)]


class PackJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("scale", C.c_void_p),
                ("D0", C.c_int64), ("D1", C.c_int64), ("D2", C.c_int64),
                ("s0", C.c_int64), ("s1", C.c_int64), ("s2", C.c_int64), ("ld_dst", C.c_int64),
                ("src_dtype", C.c_int32), ("dst_dtype", C.c_int32), ("flip1", C.c_int32), ("pad_", C.c_int32),
                ("row_scale", C.c_void_p)]


class ParamDesc(C.Structure):
    _fields_ = [("w", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("n", C.c_int64)]


# every symbol include/dfcsa.h declares (the CPU test checks the .so exports exactly these)
SYMBOLS = [
    "dfcsa_version", "dfcsa_last_error", "dfcsa_device_ok",
    "dfcsa_conv_gemm", "dfcsa_conv_wgrad", "dfcsa_wgrad_plan", "dfcsa_permute3", "dfcsa_pack_jobs", "dfcsa_sgemm", "dfcsa_bgemm", "dfcsa_bgemm_rowstat_parts", "dfcsa_lse_combine", "dfcsa_attn_pv_fused", "dfcsa_attn_bwd_fused", "dfcsa_attn_small_fwd", "dfcsa_attn_small_bwd",
    "dfcsa_softmax_rows", "dfcsa_softmax_rows_bwd", "dfcsa_softmax_rows_bwd_d", "dfcsa_rowdot",
    "dfcsa_bn_finalize", "dfcsa_bn_eval_affine",
    "dfcsa_bnrelu_pool_fwd", "dfcsa_branch_act_fwd", "dfcsa_gate_mix_fwd", "dfcsa_block_out_fwd", "dfcsa_sum_out_fwd",
    "dfcsa_block_out_bwd_reduce", "dfcsa_bn_bwd_apply", "dfcsa_gate_mix_bwd_reduce", "dfcsa_gate_mix_bwd_apply",
    "dfcsa_branch_bwd_reduce1", "dfcsa_branch_bwd_reduce2", "dfcsa_branch_bwd_apply", "dfcsa_bn_param_grads", "dfcsa_block_param_grads",
    "dfcsa_bn_bwd_reduce", "dfcsa_pool_window_terms",
    "dfcsa_nchw_to_nhwc", "dfcsa_nhwc_to_nchw", "dfcsa_colsum", "dfcsa_cast2d",
    "dfcsa_bce_dice_sums", "dfcsa_bce_dice_finalize", "dfcsa_bce_dice_bwd",
    "dfcsa_bce_dice_sums_batched", "dfcsa_bce_dice_finalize_batched",
    "dfcsa_grad_sumsq", "dfcsa_sgd_step", "dfcsa_accumulate",
    "dfcsa_preprocess", "dfcsa_preprocess_workspace_bytes",
    "dfcsa_resize_bilinear", "dfcsa_resize_bilinear_bwd", "dfcsa_adaptive_pool", "dfcsa_adaptive_pool_bwd",
]

_lib = None


def lib():
    """Load libdfcsa.so (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  dfcsa has no CPU or PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        l.dfcsa_last_error.restype = C.c_char_p
        l.dfcsa_version.restype = C.c_int
        l.dfcsa_preprocess_workspace_bytes.restype = C.c_int64
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().dfcsa_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libdfcsa {what} failed (code {rc}): {msg}")


class Sample(C.Structure):        # dfcsa_sample_t
    _fields_ = [("img", C.c_void_p), ("mask", C.c_void_p), ("h", C.c_int32), ("w", C.c_int32), ("rot_mode", C.c_int32),
                ("flip", C.c_int32), ("a", C.c_double * 6)]


class Profiler:
    """Optional per-entry-point timing with CUDA events on the launching stream (bench.py's roofline numbers)."""

    def __init__(self):
        self.records = []       # (tag, flops, start_event, end_event, desc)

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for tag, flops, e0, e1, _ in self.records:
            a = agg.setdefault(tag, {"ms": 0.0, "flops": 0.0, "launches": 0})
            a["ms"] += e0.elapsed_time(e1)
            a["flops"] += flops
            a["launches"] += 1
        return agg

    def detail(self, tags=("conv_tc", "wgrad_tc")):
        """per-shape totals for the GEMM kernels: {desc: {ms, flops, launches}}"""
        torch.cuda.synchronize()
        out = {}
        for tag, flops, e0, e1, desc in self.records:
            if tag in tags:
                a = out.setdefault(f"{tag} {desc}", {"ms": 0.0, "flops": 0.0, "launches": 0})
                a["ms"] += e0.elapsed_time(e1); a["flops"] += flops; a["launches"] += 1
        return out


PROF = None          # set to a Profiler() to time every ABI call
LAUNCHES = 0         # kernels enqueued through the ABI (bench.py's gpu_launches)
# entry points that enqueue more than one kernel
_KERNELS = {"dfcsa_bnrelu_pool_fwd": 2, "dfcsa_branch_bwd_reduce1": 3, "dfcsa_preprocess": 4, "dfcsa_attn_bwd_fused": 2, "dfcsa_branch_bwd_apply": 2}


def call(name, *args, tag=None, flops=0.0, desc=""):
    """Invoke one ABI entry point on the current stream; raise on a non-zero return code."""
    global LAUNCHES
    fn = getattr(lib(), name)
    if PROF is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        PROF.records.append((tag or name[6:], flops, e0, e1, desc))
    else:
        rc = fn(*args)
    LAUNCHES += _KERNELS.get(name, 1)
    check(rc, name)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("dfcsa kernels take CUDA tensors only (no CPU fallback)")
    return C.c_void_p(t.data_ptr())
