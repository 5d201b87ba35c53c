"""ctypes binding of ``include/diffassemble_b200.h`` (the C-ABI drop-in boundary).

There is no CPU path: ``load_library()`` raises if the shared library has not been
built, and every compute entry point returns ``DA_ERR_CUDA`` without an sm_100 GPU.
"""
import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libdiffassemble_b200.so"

DA_ABI_VERSION = 1
DA_OK, DA_ERR_INVALID, DA_ERR_CUDA, DA_ERR_UNSUPPORTED, DA_ERR_MISSING = 0, -1, -2, -3, -4
DA_HEAD_2D, DA_HEAD_SE3 = 0, 1
DA_ARCH_TRANSFORMER, DA_ARCH_EXOPHORMER = 0, 1
DA_GEMM_FP32_SIMT, DA_GEMM_BF16X3_UMMA = 0, 1
DA_ATTN_CSR, DA_ATTN_AUTO = 0, 1
DA_PRED_START_X, DA_PRED_EPSILON = 0, 1

GEMM_MODES = {"fp32": DA_GEMM_FP32_SIMT, "bf16x3": DA_GEMM_BF16X3_UMMA}
ATTN_MODES = {"csr": DA_ATTN_CSR, "auto": DA_ATTN_AUTO}

# every symbol include/diffassemble_b200.h declares (tests check the library exports each)
EXPORTED_SYMBOLS = [
    "da_abi_version", "da_create", "da_destroy", "da_last_error", "da_load_weights", "da_set_graph",
    "da_set_features", "da_forward", "da_ddpm_step", "da_ddim_step", "da_ddim_update", "da_workspace_bytes",
    "da_launch_count", "da_graph_stats", "da_set_profiling", "da_get_profile", "da_profile_tag_name",
    "da_op_linear", "da_op_graph_attention", "da_op_graph_attention_dense", "da_greedy_cost_assignment", "da_expander_edge_index", "da_graph_create", "da_graph_destroy",
    "da_op_graph_attention_fwd", "da_op_graph_attention_bwd", "da_op_linear_wgrad", "da_op_linear_ws", "da_op_segment_max", "da_adafactor_step", "da_graph_set_batch",
    "da_op_linear_workspace_bytes", "da_graph_plan_info", "da_ddpm_step_t", "da_ddim_step_t", "da_forward_attn", "da_op_conv2d_nhwc", "da_op_dwconv2d_nhwc", "da_op_spatial_mean", "da_op_channel_scale", "da_op_normalize_to_nhwc", "da_op_add_inplace",
]


class da_config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("feat_dim", C.c_int32), ("in_channels", C.c_int32),
        ("out_channels", C.c_int32), ("heads", C.c_int32), ("hidden", C.c_int32), ("n_layers", C.c_int32),
        ("steps", C.c_int32), ("mlp_hidden", C.c_int32), ("head_kind", C.c_int32), ("arch", C.c_int32),
        ("virt_nodes", C.c_int32), ("gemm_mode", C.c_int32), ("attn_mode", C.c_int32), ("reserved", C.c_int32 * 8),
    ]


class da_weight_desc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("rows", C.c_int64), ("cols", C.c_int64)]


class da_step_coef(C.Structure):
    _fields_ = [
        ("t", C.c_int32), ("t_index", C.c_int32), ("pred", C.c_int32), ("has_prev", C.c_int32),
        ("beta_t", C.c_float), ("sqrt_one_minus_acp", C.c_float), ("sqrt_recip_alpha", C.c_float),
        ("posterior_variance", C.c_float), ("acp", C.c_float), ("acp_prev", C.c_float),
        ("sqrt_recip_acp", C.c_float), ("sqrt_recipm1_acp", C.c_float), ("eta", C.c_float), ("cfg_w", C.c_float),
    ]


class da_schedule(C.Structure):
    _fields_ = [
        ("betas", C.c_void_p), ("alphas_cumprod", C.c_void_p), ("sqrt_one_minus_alphas_cumprod", C.c_void_p),
        ("sqrt_recip_alphas", C.c_void_p), ("posterior_variance", C.c_void_p), ("sqrt_recip_alphas_cumprod", C.c_void_p),
        ("sqrt_recipm1_alphas_cumprod", C.c_void_p), ("steps", C.c_int32), ("inference_ratio", C.c_int32),
    ]


class DiffAssembleError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"diffassemble_b200 C-ABI error {status}: {message}")
        self.status = status


_lib = None


def load_library():
    """Load the in-tree shared library; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m diffassemble_b200.build` "
            "(diffassemble_b200 has no CPU or PyTorch fallback)"
        )
    lib = C.CDLL(str(LIB_PATH))
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.da_abi_version.restype = C.c_int
    lib.da_create.argtypes = [C.POINTER(vp), C.POINTER(da_config)]
    lib.da_destroy.argtypes = [vp]
    lib.da_destroy.restype = None
    lib.da_last_error.argtypes = [vp]
    lib.da_last_error.restype = C.c_char_p
    lib.da_load_weights.argtypes = [vp, C.POINTER(da_weight_desc), i32]
    lib.da_set_graph.argtypes = [vp, vp, vp, i64, vp, i32, i32, vp, vp]
    lib.da_set_features.argtypes = [vp, vp, vp]
    lib.da_forward.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.da_forward_attn.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.da_ddpm_step.argtypes = [vp, vp, vp, C.POINTER(da_step_coef), vp, vp]
    lib.da_ddim_step.argtypes = [vp, vp, vp, C.POINTER(da_step_coef), vp, vp]
    lib.da_ddpm_step_t.argtypes = [vp, vp, vp, vp, i32, C.POINTER(da_schedule), vp, vp]
    lib.da_ddim_step_t.argtypes = [vp, vp, vp, vp, i32, C.c_float, C.POINTER(da_schedule), vp, vp]
    lib.da_ddim_update.argtypes = [vp, vp, vp, vp, C.POINTER(da_step_coef), vp, vp]
    lib.da_workspace_bytes.argtypes = [vp]
    lib.da_workspace_bytes.restype = C.c_size_t
    lib.da_launch_count.argtypes = [vp]
    lib.da_launch_count.restype = i64
    lib.da_graph_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    lib.da_graph_plan_info.argtypes = [vp, C.POINTER(i64), i32]
    lib.da_set_profiling.argtypes = [vp, i32]
    lib.da_get_profile.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64), i32, i32]
    lib.da_profile_tag_name.argtypes = [i32]
    lib.da_profile_tag_name.restype = C.c_char_p
    lib.da_op_linear.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.da_op_conv2d_nhwc.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.da_op_dwconv2d_nhwc.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.da_op_spatial_mean.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.da_op_channel_scale.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.da_op_normalize_to_nhwc.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
    lib.da_op_add_inplace.argtypes = [vp, vp, i64, vp]
    lib.da_op_segment_max.argtypes = [vp, i32, vp, i32, i32, vp, vp]
    lib.da_adafactor_step.argtypes = [vp, i32, C.c_float, C.c_float, C.c_float, C.c_float, vp]
    lib.da_op_graph_attention.argtypes = [vp, vp, vp, i64, i32, i32, i32, vp, vp, vp]
    lib.da_op_graph_attention_dense.argtypes = [vp, vp, vp, i64, vp, i32, i32, i32, vp, C.POINTER(i64), vp]
    lib.da_graph_create.argtypes = [C.POINTER(vp), vp, vp, i64, i32, vp]
    lib.da_graph_set_batch.argtypes = [vp, vp, vp, vp, vp]
    lib.da_graph_destroy.argtypes = [vp]
    lib.da_graph_destroy.restype = None
    lib.da_op_graph_attention_fwd.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    lib.da_op_graph_attention_bwd.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp]
    lib.da_op_linear_wgrad.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
    lib.da_op_linear_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.da_op_linear_workspace_bytes.restype = C.c_size_t
    lib.da_op_linear_ws.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, i32, vp, C.c_size_t, vp]
    lib.da_expander_edge_index.argtypes = [vp, i32, i32, i32, vp, vp, vp]
    lib.da_greedy_cost_assignment.argtypes = [vp, i32, vp, i32, vp, i32, i32, vp, vp]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("da_abi_version",):
            fn.restype = C.c_int
    if lib.da_abi_version() != DA_ABI_VERSION:
        raise RuntimeError("libdiffassemble_b200.so ABI version mismatch; rebuild the library")
    _lib = lib
    return lib
