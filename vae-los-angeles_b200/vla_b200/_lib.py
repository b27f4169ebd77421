"""ctypes binding of libvla_b200.so (include/vla_b200.h).  No CPU fallback: a missing library raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvla_b200.so")

KIND = {"multimodal": 0, "rna2dna": 1, "dna2rna": 2, "rna2dna_ae": 3, "dna2rna_ae": 4}
TENSOR_PARAM, TENSOR_BUFFER, TENSOR_COUNTER = 0, 1, 2

c_float_p = C.c_void_p  # device pointers travel as integers


class Config(C.Structure):
    _fields_ = [("kind", C.c_int), ("dim_a", C.c_int), ("dim_b", C.c_int), ("n_sites", C.c_int),
                ("latent", C.c_int), ("embed", C.c_int)]


class TensorInfo(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("kind", C.c_int), ("offset", C.c_longlong), ("ndim", C.c_int),
                ("shape", C.c_int * 2)]


class ForwardArgs(C.Structure):
    _fields_ = [("params", C.c_void_p), ("buffers", C.c_void_p), ("counters", C.c_void_p),
                ("x_a", C.c_void_p), ("x_b", C.c_void_p), ("site", C.c_void_p),
                ("batch", C.c_int), ("train", C.c_int), ("refresh_shadows", C.c_int),
                ("eps", C.c_void_p), ("keep_masks", C.POINTER(C.c_void_p)),
                ("seed", C.c_ulonglong), ("offset", C.c_ulonglong),
                ("recon_a", C.c_void_p), ("recon_b", C.c_void_p), ("recon_c", C.c_void_p),
                ("mu", C.c_void_p), ("logvar", C.c_void_p)]


class BackwardArgs(C.Structure):
    _fields_ = [("params", C.c_void_p), ("g_recon_a", C.c_void_p), ("g_recon_b", C.c_void_p),
                ("g_recon_c", C.c_void_p), ("g_mu", C.c_void_p), ("g_logvar", C.c_void_p),
                ("recon_b", C.c_void_p), ("grads", C.c_void_p)]


class LossArgs(C.Structure):
    _fields_ = [("recon_a", C.c_void_p), ("a", C.c_void_p), ("dim_a", C.c_int),
                ("recon_b", C.c_void_p), ("b", C.c_void_p), ("dim_b", C.c_int),
                ("recon_c", C.c_void_p), ("site", C.c_void_p), ("class_weights", C.c_void_p), ("n_sites", C.c_int),
                ("mu", C.c_void_p), ("logvar", C.c_void_p), ("latent", C.c_int),
                ("batch", C.c_int), ("beta", C.c_float), ("gamma", C.c_float),
                ("g_recon_a", C.c_void_p), ("g_recon_b", C.c_void_p), ("g_recon_c", C.c_void_p),
                ("g_mu", C.c_void_p), ("g_logvar", C.c_void_p),
                ("out", C.c_void_p), ("workspace", C.c_void_p)]


class AdamWArgs(C.Structure):
    _fields_ = [("params", C.c_void_p), ("grads", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("step", C.c_int)]


class TrainArgs(C.Structure):
    _fields_ = [("params", C.c_void_p), ("grads", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("buffers", C.c_void_p), ("counters", C.c_void_p),
                ("x_a", C.c_void_p), ("x_b", C.c_void_p), ("site", C.c_void_p), ("class_weights", C.c_void_p),
                ("batch", C.c_int), ("dataset_rows", C.c_longlong),
                ("eps", C.c_void_p), ("keep_masks", C.POINTER(C.c_void_p)), ("seed", C.c_ulonglong),
                ("beta1", C.c_float), ("beta2", C.c_float), ("adam_eps", C.c_float),
                ("recon_a", C.c_void_p), ("recon_b", C.c_void_p), ("recon_c", C.c_void_p),
                ("mu", C.c_void_p), ("logvar", C.c_void_p), ("loss_out", C.c_void_p), ("phases", C.c_int),
                ("dp", C.c_void_p), ("sync_bn", C.c_int)]


class MetricsArgs(C.Structure):
    _fields_ = [("y_true", C.c_void_p), ("y_pred", C.c_void_p), ("rows", C.c_longlong), ("dim", C.c_int),
                ("cosine", C.c_void_p), ("pearson", C.c_void_p), ("out", C.c_void_p), ("workspace", C.c_void_p)]


class ProfEntry(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("ms", C.c_float), ("flops", C.c_double), ("bytes", C.c_double)]


EXPORTS = {
    "vla_last_error": (C.c_char_p, []),
    "vla_abi_version": (C.c_int, []),
    "vla_model_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "vla_model_create_layout_only": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "vla_model_destroy": (None, [C.c_void_p]),
    "vla_model_reserve": (C.c_int, [C.c_void_p, C.c_int]),
    "vla_param_count": (C.c_longlong, [C.c_void_p]),
    "vla_buffer_count": (C.c_longlong, [C.c_void_p]),
    "vla_counter_count": (C.c_int, [C.c_void_p]),
    "vla_num_tensors": (C.c_int, [C.c_void_p]),
    "vla_tensor_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(TensorInfo)]),
    "vla_forward": (C.c_int, [C.c_void_p, C.POINTER(ForwardArgs), C.c_void_p]),
    "vla_refresh_shadows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vla_backward": (C.c_int, [C.c_void_p, C.POINTER(BackwardArgs), C.c_void_p]),
    "vla_loss_workspace_bytes": (C.c_longlong, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vla_loss": (C.c_int, [C.POINTER(LossArgs), C.c_void_p]),
    "vla_adamw": (C.c_int, [C.c_void_p, C.POINTER(AdamWArgs), C.c_void_p]),
    "vla_gather_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vla_scale_inplace": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.c_int, C.c_void_p, C.c_void_p]),
    "vla_metrics_workspace_bytes": (C.c_longlong, [C.c_longlong]),
    "vla_recon_metrics": (C.c_int, [C.POINTER(MetricsArgs), C.c_void_p]),
    "vla_train_step": (C.c_int, [C.c_void_p, C.POINTER(TrainArgs), C.c_void_p]),
    "vla_train_step_group": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.POINTER(TrainArgs)), C.c_int, C.c_void_p]),
    "vla_group_cached_plans": (C.c_int, [C.c_void_p]),
    "vla_set_hyper": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "vla_set_step": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "vla_dp_create": (C.c_int, [C.c_int, C.c_int, C.c_longlong, C.POINTER(C.c_void_p)]),
    "vla_dp_ipc_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vla_dp_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vla_dp_grads": (C.c_void_p, [C.c_void_p]),
    "vla_dp_losses": (C.c_void_p, [C.c_void_p]),
    "vla_dp_trace": (C.c_int, [C.c_void_p, C.POINTER(C.c_ulonglong)]),
    "vla_dp_disconnect": (C.c_int, [C.c_void_p]),
    "vla_dp_destroy": (None, [C.c_void_p]),
    "vla_profile_begin": (C.c_int, [C.c_void_p]),
    "vla_profile_collect": (C.c_int, [C.c_void_p, C.POINTER(ProfEntry), C.c_int]),
    "vla_profile_read": (C.c_int, [C.c_void_p, C.POINTER(ProfEntry), C.c_int]),
    "vla_profile_pause": (C.c_int, [C.c_void_p]),
    "vla_chain_timeline": (C.c_int, [C.c_void_p, C.c_int]),
    "vla_rowchain_timeline": (C.c_int, [C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vla_chain_count": (C.c_int, [C.c_void_p]),
    "vla_chain_cached_plans": (C.c_int, [C.c_void_p]),
    "vla_chain_info": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "vla_chain_phase_name": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "vla_chain_timeline_read": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_ulonglong)]),
    "vla_model_pin": (C.c_int, [C.c_void_p, C.c_int]),
    "vla_test_workspace": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]),
    "vla_test_set_timeline": (C.c_int, [C.c_void_p]),
    "vla_test_set_flags": (C.c_int, [C.c_int]),
    "vla_test_gemm": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    """The loaded library.  Raises RuntimeError when it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  vla_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().vla_last_error()
        raise RuntimeError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
