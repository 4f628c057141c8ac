"""ctypes binding of libmgn_b200.so (include/mgn_b200.h) - the stand-in, in this Julia-less
image, for the ``ccall`` stubs of julia/GraphNetCoreB200.jl.  There is no fallback: a missing
library or a failing call raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmgn_b200.so")

MGN_OK = 0
COMPUTE_FP32 = 0
COMPUTE_BF16 = 1
STAGE_ENCODE, STAGE_DECODE = -1, -2
HALO_LATENT, HALO_GRAD = 0, 1
ROWS_PACK, ROWS_UNPACK, ROWS_ADD, ROWS_PACK_ZERO = 0, 1, 2, 3
NORM_FORWARD, NORM_INVERSE, NORM_FORWARD_VJP, NORM_INVERSE_VJP = 0, 1, 2, 3
DP_SUM, DP_MEAN = 0, 1
FEAT_AFFINE, FEAT_ONLINE = 0, 1
DP_UNIQUE_ID_BYTES = 128


class MgnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmgn_b200 error {code}: {msg}")
        self.code = code


class ModelConfig(C.Structure):
    _fields_ = [
        ("node_in", C.c_int32), ("edge_in", C.c_int32), ("out_dim", C.c_int32),
        ("latent", C.c_int32), ("mps", C.c_int32), ("hidden_layers", C.c_int32),
        ("ln_eps", C.c_float), ("compute_mode", C.c_int32),
        ("dense_layers", C.c_int32), ("ln_scale_first", C.c_int32), ("aggregate_post_residual", C.c_int32),
    ]


class AdamConfig(C.Structure):
    """mgn_adam_config of include/mgn_b200.h (mgn_backward_dp)."""
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("d_m", C.c_void_p), ("d_v", C.c_void_p), ("d_state16", C.c_void_p)]


class FeatureSeg(C.Structure):
    """mgn_feature_seg"""
    _fields_ = [("d_x", C.c_void_p), ("ld", C.c_int32), ("col", C.c_int32), ("width", C.c_int32), ("kind", C.c_int32),
                ("scale", C.c_float), ("shift", C.c_float), ("d_state", C.c_void_p), ("std_eps", C.c_float)]


class FusedIo(C.Structure):
    """mgn_fused_io"""
    _fields_ = [("n_node_segs", C.c_int32), ("node", FeatureSeg * 8), ("n_edge_segs", C.c_int32), ("edge", FeatureSeg * 8),
                ("n_out_segs", C.c_int32), ("out", FeatureSeg * 8), ("d_val_mask", C.c_void_p)]


class NormUpdate(C.Structure):
    """mgn_norm_update"""
    _fields_ = [("d_x", C.c_void_p), ("rows", C.c_int64), ("ld", C.c_int32), ("col", C.c_int32), ("features", C.c_int32),
                ("d_state", C.c_void_p), ("max_acc", C.c_float)]


class ParamEntry(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("offset", C.c_int64), ("rows", C.c_int32), ("cols", C.c_int32)]


_p = C.c_void_p
_i32, _i64, _f32, _sz = C.c_int32, C.c_int64, C.c_float, C.c_size_t

# name -> argtypes (every function returns int32_t)
SIGNATURES = {
    "mgn_abi_version": [],
    "mgn_last_error": [C.c_char_p, _sz],
    "mgn_device_count": [C.POINTER(_i32)],
    "mgn_one_hot": [_p, _i64, _i32, _i32, _p],
    "mgn_triangles_to_edges": [_p, _i64, _p, _p, C.POINTER(_i64)],
    "mgn_parse_edges": [_p, _i64, _p, _p],
    "mgn_shift_one_based": [_p, _p, _i64, C.POINTER(_i32)],
    "mgn_edge_features": [_p, _i64, _i32, _p, _p, _i64, _i32, _p],
    "mgn_one_hot_device": [_p, _i64, _i32, _i32, _p, _p],
    "mgn_triangles_to_edges_device": [_p, _i64, _p, _p, C.POINTER(_i64), _p],
    "mgn_parse_edges_device": [_p, _i64, _p, _p, _p],
    "mgn_shift_one_based_device": [_p, _p, _i64, C.POINTER(_i32), _p],
    "mgn_edge_features_device": [_p, _i64, _i32, _p, _p, _i64, _i32, _p, _p],
    "mgn_graph_create": [_i64, _i64, _p, _p, _i32, _p, C.POINTER(_p)],
    "mgn_graph_destroy": [_p],
    "mgn_graph_sizes": [_p, C.POINTER(_i64), C.POINTER(_i64)],
    "mgn_graph_get_index": [_p, _p, _p, _p, _p],
    "mgn_model_create": [C.POINTER(ModelConfig), C.POINTER(_p)],
    "mgn_model_destroy": [_p],
    "mgn_model_param_count": [_p, C.POINTER(_i64)],
    "mgn_model_param_layout": [_p, C.POINTER(ParamEntry), _i32, C.POINTER(_i32)],
    "mgn_workspace_bytes": [_p, _p, _i32, C.POINTER(_sz)],
    "mgn_forward": [_p, _p, _p, _p, _p, _p, _p, _sz, _i32, _p],
    "mgn_backward": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p],
    "mgn_forward_stage": [_p, _p, _p, _p, _p, _p, _p, _sz, _i32, _i32, _p],
    "mgn_backward_stage": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _i32, _p],
    "mgn_halo_row_bytes": [_p, _i32, C.POINTER(_sz)],
    "mgn_halo_rows": [_p, _p, _p, _sz, _i32, _i32, _i32, _p, _i64, _p, _i32, _p],
    "mgn_loss_mse_masked": [_p, _p, _i64, _i32, _p, _i64, _i32, _p, _p, _p],
    "mgn_adam_step": [_p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _i64, _p],
    "mgn_adam_step_device": [_p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _p, _p],
    "mgn_profile_begin": [_i32],
    "mgn_profile_end": [C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_f32), C.POINTER(_i64), _i32],
    "mgn_profile_tag_name": [_i32, C.c_char_p, _sz],
    "mgn_norm_online_update": [_p, _i64, _i32, _p, _f32, _p],
    "mgn_norm_online_apply": [_p, _i64, _i32, _p, _f32, _i32, _p, _i32, _i32, _p],
    "mgn_affine_apply": [_p, _i64, _i32, _f32, _f32, _p, _i32, _i32, _p],
    "mgn_ode_lincomb": [_p, _p, _p, _i32, _i64, _p, _p],
    "mgn_masked_overwrite": [_p, _p, _p, _i64, _p, _p],
    "mgn_vec_mul": [_p, _p, _i64, _p, _p],
    "mgn_norm_online_apply_ld": [_p, _i32, _i32, _i64, _i32, _p, _f32, _i32, _p, _i32, _i32, _p],
    "mgn_affine_apply_ld": [_p, _i32, _i32, _i64, _i32, _f32, _f32, _p, _i32, _i32, _p],
    "mgn_shooting_mse": [_p, _p, _p, _i64, _i64, _f32, _i32, _p, _p, _p],
    "mgn_shooting_continuity": [_p, _p, _i64, _f32, _p, _p, _p],
    "mgn_forward_fused": [_p, _p, _p, C.POINTER(FusedIo), _p, _p, _sz, _i32, _p],
    "mgn_backward_fused": [_p, _p, _p, C.POINTER(FusedIo), _p, _p, _p, _p, _sz, _p],
    "mgn_norm_online_update_multi": [C.POINTER(NormUpdate), _i32, _p],
    "mgn_dp_unique_id": [_p],
    "mgn_dp_init": [_p, _i32, _i32, C.POINTER(_p)],
    "mgn_dp_finalize": [_p],
    "mgn_dp_rank": [_p, C.POINTER(_i32), C.POINTER(_i32)],
    "mgn_dp_allreduce": [_p, _p, _i64, _i32, _p],
    "mgn_dp_allreduce_normaliser": [_p, _p, _p, _i32, _p],
    "mgn_halo_exchange": [_p, _p, C.POINTER(_i64), _p, C.POINTER(_i64), _i64, _p],
    "mgn_backward_dp": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p, C.POINTER(AdamConfig), _i32, _p],
    "mgn_library_release": [],
}

_lib = None


def load():
    """Loads the shared library (once) and declares every prototype of include/mgn_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MgnError(-1, f"{LIB_PATH} is missing - run `python meshgraphnets.jl_b200/build.py` "
                           "(there is no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = args
        fn.restype = _i32
    _lib = lib
    return lib


def last_error():
    buf = C.create_string_buffer(1024)
    load().mgn_last_error(buf, 1024)
    return buf.value.decode(errors="replace")


def check(status):
    if status != MGN_OK:
        raise MgnError(status, last_error())


def call(name, *args):
    check(getattr(load(), name)(*args))
