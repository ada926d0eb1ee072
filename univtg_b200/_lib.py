"""ctypes binding of include/univtg_b200.h (the C-ABI of the CUDA library).

There is no CPU fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# UNIVTG_LIB: A/B-test another build of the same ABI (profiling only)
LIB_PATH = os.environ.get("UNIVTG_LIB") or os.path.join(_HERE, "lib", "libunivtg_b200.so")
_lib = None

c_int = ctypes.c_int32
c_void_p = ctypes.c_void_p
c_size_t = ctypes.c_size_t
c_float = ctypes.c_float


class Config(ctypes.Structure):
    """univtg_config (include/univtg_b200.h)."""

    _fields_ = [
        ("hidden_dim", c_int),
        ("nheads", c_int),
        ("dim_feedforward", c_int),
        ("enc_layers", c_int),
        ("n_input_proj", c_int),
        ("v_feat_dim", c_int),
        ("t_feat_dim", c_int),
        ("operand_format", c_int),
    ]


class Shape(ctypes.Structure):
    """univtg_shape."""

    _fields_ = [("batch", c_int), ("l_vid", c_int), ("l_txt", c_int), ("training", c_int)]


class Rng(ctypes.Structure):
    """univtg_rng."""

    _fields_ = [("seed", ctypes.c_uint64), ("input_dropout", c_float), ("droppath", c_float)]


# symbol -> (restype, argtypes); every symbol declared in include/univtg_b200.h must be listed here
SIGNATURES = {
    "univtg_last_error": (ctypes.c_char_p, []),
    "univtg_abi_version": (c_int, []),
    "univtg_num_params": (c_int, [ctypes.POINTER(Config)]),
    "univtg_packed_bytes": (c_size_t, [ctypes.POINTER(Config)]),
    "univtg_pack_weights": (c_int, [ctypes.POINTER(Config), ctypes.POINTER(c_void_p), c_int, c_void_p, c_void_p]),
    "univtg_workspace_bytes": (c_size_t, [ctypes.POINTER(Config), ctypes.POINTER(Shape)]),
    "univtg_prepare_workspace": (c_int, [ctypes.POINTER(Config), ctypes.POINTER(Shape), c_void_p, c_int, c_void_p]),
    "univtg_plan_create": (c_int, [ctypes.POINTER(Config), ctypes.POINTER(Shape), c_void_p, c_void_p, c_void_p, c_void_p,
                                   ctypes.POINTER(c_void_p)]),
    "univtg_plan_destroy": (None, [c_void_p]),
    "univtg_forward": (c_int, [c_void_p] * 12),
    "univtg_forward_num_launches": (c_int, [c_void_p]),
    "univtg_launch_count": (ctypes.c_int64, []),
    "univtg_host_register": (c_int, [c_void_p, c_size_t, c_int]),
    "univtg_h2d_gather_batch": (c_int, [c_void_p] * 11 + [c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "univtg_host_assemble_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_int, c_int, c_int, c_int, c_int, c_int]),
    "univtg_train_workspace_bytes": (c_size_t, [ctypes.POINTER(Config), ctypes.POINTER(Shape)]),
    "univtg_forward_train": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     ctypes.POINTER(c_void_p), ctypes.POINTER(Rng), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    "univtg_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.POINTER(c_void_p), ctypes.POINTER(Rng),
                                c_void_p, c_void_p, c_void_p, c_void_p, c_float, ctypes.POINTER(c_void_p), c_int, c_void_p]),
    "univtg_dropout_mask": (c_int, [ctypes.POINTER(Rng), c_int, c_size_t, c_size_t, c_void_p, c_void_p]),
    "univtg_droppath_scales": (c_int, [ctypes.POINTER(Rng), c_int, c_int, c_void_p, c_void_p]),
    "univtg_loss_scratch_bytes": (c_size_t, [c_int, c_int]),
    "univtg_loss_forward": (c_int, [c_void_p] * 10 + [c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "univtg_loss_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    "univtg_backward_stages": (c_int, [ctypes.POINTER(Config), c_void_p, c_int]),
    "univtg_plan_set_grad_events": (c_int, [c_void_p, ctypes.POINTER(c_void_p), c_int]),
    "univtg_plan_set_backward_sm_budget": (c_int, [c_void_p, c_int]),
    "univtg_decode_mr": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    "univtg_temporal_nms": (c_int, [c_void_p, c_int, c_int, c_int, ctypes.c_double, c_int, c_void_p, c_void_p, c_void_p]),
    "univtg_adamw_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_size_t, c_float, c_float, c_float, c_float,
                                  c_float, c_int, c_float, c_int, c_void_p, ctypes.POINTER(Config), c_void_p, c_void_p]),
    "univtg_pack_vectors": (c_int, [ctypes.POINTER(Config), ctypes.POINTER(c_void_p), c_int, c_void_p, c_void_p]),
    "univtg_plan_set_profiling": (c_int, [c_void_p, c_int]),
    "univtg_plan_set_input_format": (c_int, [c_void_p, c_int]),
    "univtg_plan_read_profile": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "univtg_op_gemm": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                               c_float, c_void_p, c_void_p, c_void_p]),
    "univtg_op_gemm_cluster": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                       c_float, c_void_p, c_void_p, c_void_p]),
    "univtg_debug_gemm_timeline": (c_int, [c_void_p]),
    "univtg_debug_choose_tile": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "univtg_debug_tmem_ld_rate": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "univtg_debug_mma_rate": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "univtg_op_layernorm": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p, c_int,
                                    c_void_p]),
    "univtg_op_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                    c_void_p]),
    "univtg_op_attention_bwd": (c_int, [c_void_p] * 7 + [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
}


def load_library():
    """Load libunivtg_b200.so and attach signatures.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "univtg_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.univtg_abi_version() != 2:
        raise RuntimeError("univtg_b200: ABI version mismatch between header and library")
    _lib = lib
    return lib


def last_error():
    lib = load_library()
    msg = lib.univtg_last_error()
    return msg.decode() if msg else ""


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"univtg_b200: {what} failed (rc={rc}): {last_error()}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)
