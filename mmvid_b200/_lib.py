"""ctypes binding of libmmvid_b200.so (the C ABI declared in include/mmvid_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmmvid_b200.so")

FP32, TF32, BF16, F16 = 0, 1, 2, 3
ACT_NONE, ACT_QUICKGELU, ACT_SWISH = 0, 1, 2
MASK_NONE, MASK_CAUSAL, MASK_PREV = 0, 1, 2
DT_F32, DT_BF16, DT_F16 = 0, 1, 2
PRECISIONS = {"fp32": FP32, "tf32": TF32, "bf16": BF16, "fp16": F16}
H16 = (BF16, F16)  # kind::f16 precisions: 16-bit operands and activations, fp32 accumulation / residual stream

_p, _ll, _i, _f = C.c_void_p, C.c_longlong, C.c_int, C.c_float

# Kernels that update parameters through raw pointers (optim.FusedAdam) do not advance torch's per-tensor version
# counters, so every cache of derived weights (axial tables, 16-bit copies, repacked conv weights, captured CUDA graphs)
# keys on (data_ptr, _version, weights_epoch()); such kernels call bump_weights_epoch() after their launch.
_weights_epoch = [0]


def weights_epoch():
    return _weights_epoch[0]


def bump_weights_epoch():
    _weights_epoch[0] += 1


class EmbedSegment(C.Structure):
    _fields_ = [("ids", _p), ("ids_bstride", _ll), ("n", _i), ("seq_off", _i), ("table", _p), ("table2", _p),
                ("pos", _p), ("pad_value", _ll), ("pad_base", _ll), ("use_pad", _i), ("table_rows", _ll)]


class DecodeLayer(C.Structure):
    _fields_ = [(n, _p) for n in ("ln1_w", "ln1_b", "in_w", "in_b", "out_w", "out_b", "ln2_w", "ln2_b", "fc_w", "fc_b",
                                  "proj_w", "proj_b", "kcache", "vcache")]


class DecodeLayer16(C.Structure):
    _fields_ = [(n, _p) for n in ("ln1_w", "ln1_b", "in_b", "out_b", "ln2_w", "ln2_b", "fc_b", "proj_b", "in_w", "out_w",
                                  "fc_w", "proj_w", "kcache", "vcache")]


class AdamTensor(C.Structure):
    _fields_ = [("p", _p), ("g", _p), ("m", _p), ("v", _p), ("n", _ll), ("skipped", _ll)]


class ConvParams(C.Structure):
    _fields_ = [("inp", _p), ("w", _p), ("bias", _p), ("residual", _p), ("out", _p)] + \
               [(n, _i) for n in ("N", "H", "W", "Cin", "Cout", "KH", "KW", "stride", "pad_t", "pad_l", "Ho", "Wo",
                                  "upsample", "in_nchw", "out_nchw", "pre_affine", "post_clamp", "precision")] + \
               [("gn_partial", _p), ("gn_groups", _i)]


# name -> (restype, argtypes); kept in sync with include/mmvid_b200.h (tests/test_abi.py checks every symbol)
SIGNATURES = {
    "mmvid_version": (_i, []),
    "mmvid_last_error": (C.c_char_p, []),
    "mmvid_launch_count": (_ll, []),
    "mmvid_reset_launch_count": (None, []),
    "mmvid_add_launch_count": (None, [_ll]),
    "mmvid_embed_gather": (_i, [_p, _i, _i, _i, C.POINTER(EmbedSegment), _i, _p]),
    "mmvid_axial_table": (_i, [_p, _i, _i, _p, _p, _p, C.POINTER(_i), _i, _p]),
    "mmvid_layernorm": (_i, [_p, _ll, _p, _p, _p, _i, _ll, _i, _f, _p]),
    "mmvid_linear": (_i, [_p, _i, _ll, _p, _i, _ll, _p, _p, _ll, _p, _i, _ll, _ll, _i, _i, _i, _i, _p]),
    "mmvid_gemm_batched_f32": (_i, [_p, _ll, _ll, _ll, _p, _ll, _ll, _ll, _ll, _p, _ll, _ll, _ll, _i, _i, _i, _i, _i,
                                    _f, _p]),
    "mmvid_softmax_rows": (_i, [_p, _ll, _i, _i, _ll, _i, _p, _i, _p]),
    "mmvid_attention": (_i, [_p, _p, _p, _p, _i, _ll, _i, _i, _i, _i, _i, C.POINTER(_i), _i, _i, _p]),
    "mmvid_linear_qkv": (_i, [_p, _i, _ll, _p, _i, _ll, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "mmvid_qkv_split": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "mmvid_decode_attention": (_i, [_p, _ll, _p, _p, _p, _ll, _i, _i, _i, _i, _p]),
    "mmvid_linear_small_m": (_i, [_p, _ll, _p, _ll, _p, _p, _ll, _p, _ll, _i, _i, _i, _i, _p]),
    "mmvid_kv_append": (_i, [_p, _ll, _p, _p, _i, _i, _i, _i, _p]),
    "mmvid_artv_decode_workspace_floats": (_ll, [_i, _i, _i]),
    "mmvid_artv_decode_step": (_i, [C.POINTER(DecodeLayer), _i, _p, _p, _i, _i, _i, _i, _i, _p]),
    "mmvid_artv_decode_persistent": (_i, [C.POINTER(DecodeLayer), _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "mmvid_artv_decode_fused": (_i, [C.POINTER(DecodeLayer), _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "mmvid_artv_decode_stream_workspace_floats": (_ll, [_i, _i, _i]),
    "mmvid_artv_decode_stream": (_i, [C.POINTER(DecodeLayer16), _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p]),
    "mmvid_vq_argmin": (_i, [_p, _p, _p, _p, _ll, _i, _i, _p]),
    "mmvid_codebook_gather": (_i, [_p, _p, _p, _ll, _i, _p]),
    "mmvid_conv2d": (_i, [C.POINTER(ConvParams), _p]),
    "mmvid_conv2d_gn_fusable": (_i, [C.POINTER(ConvParams)]),
    "mmvid_groupnorm_from_partials": (_i, [_p, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _f, _i, _p]),
    "mmvid_groupnorm": (_i, [_p, _p, _i, _p, _p, _p, _i, _i, _i, _i, _f, _i, _p]),
    "mmvid_groupnorm_scratch_floats": (_ll, [_i, _i]),
    "mmvid_groupnorm_stats": (_i, [_p, _p, _i, _i, _i, _i, _f, _p]),
    "mmvid_conv_out_fused": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _i, _p]),
    "mmvid_conv_out_fused_from_partials": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _i, _p]),
    "mmvid_upsample2x": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "mmvid_nchw_to_nhwc": (_i, [_p, _p, _i, _i, _i, _p]),
    "mmvid_nhwc_to_nchw": (_i, [_p, _p, _i, _i, _i, _p]),
    "mmvid_softmax_logits": (_i, [_p, _p, _f, _p, _ll, _i, _p]),
    "mmvid_act_forward": (_i, [_p, _p, _ll, _i, _p]),
    "mmvid_act_backward": (_i, [_p, _p, _p, _ll, _i, _p]),
    "mmvid_colsum": (_i, [_p, _p, _p, _ll, _i, _i, _p]),
    "mmvid_layernorm_backward": (_i, [_p, _p, _p, _p, _p, _ll, _i, _f, _p]),
    "mmvid_softmax_backward": (_i, [_p, _p, _ll, _i, _ll, _f, _p]),
    "mmvid_cross_entropy": (_i, [_p, _p, _p, _p, _p, _ll, _i, _p]),
    "mmvid_embed_backward": (_i, [_p, _i, _i, _i, C.POINTER(EmbedSegment), _p, _p, _p, _p]),
    "mmvid_transpose2d": (_i, [_p, _p, _i, _i, _p]),
    "mmvid_mp_sample": (_i, [_p, _ll, _i, _f, _p, _p, _p, C.c_ulonglong, C.c_ulonglong, _p, _p]),
    "mmvid_mp_keep": (_i, [_p, _p, _p, _i, _i, _i, _i, _ll, _p, _p, C.c_ulonglong, C.c_ulonglong, _p]),
    "mmvid_debug_attention_trace": (_i, [_p]),
    "mmvid_debug_decode_trace": (_i, [_p]),
    "mmvid_debug_mma_rate": (_i, [_i, _i, _p, _p]),
    "mmvid_debug_gemm_trace": (_i, [_p]),
    "mmvid_debug_pick_tile": (_i, [_ll, _i, _i, _i, _i]),
    "mmvid_grad_sqnorm": (_i, [_p, _p, _p, _i, _i, _p, _p, _p]),
    "mmvid_grad_clip": (_i, [_p, _p, _p, _i, _i, _p, _f, _p]),
    "mmvid_adam_step": (_i, [_p, _p, _p, _i, _i, _f, _f, _f, _f, _f, _i, _i, _p]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"mmvid_b200: CUDA library not found at {LIB_PATH}. Build it with `python -m mmvid_b200.build` "
            "(or __graft_entry__.build()). There is no CPU / PyTorch fallback for the hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drifted apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().mmvid_last_error().decode(errors="replace")
        raise RuntimeError(f"mmvid_b200 {what} failed (code {rc}): {msg}")


def launch_count():
    return int(load().mmvid_launch_count())


def reset_launch_count():
    load().mmvid_reset_launch_count()


def add_launch_count(n):
    """Kernels of this library replayed from a captured CUDA graph (they do not pass through the entry points)."""
    load().mmvid_add_launch_count(int(n))
