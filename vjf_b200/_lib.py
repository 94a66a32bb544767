"""ctypes binding of libvjf_b200.so (the C ABI declared in include/vjf_b200.h)."""
import ctypes as C
import os

from . import build as _build

MAX_LAYERS = 4
MAX_XDIM = 16
LIK = {"poisson": 0, "gaussian": 1}
FLAG_SGD, FLAG_UPDATE, FLAG_WARMUP, FLAG_DECODER_FROZEN, FLAG_PRIOR_Q0 = 1, 2, 4, 8, 16
ST_RECON_NONFINITE, ST_DYN_NONFINITE, ST_ENTROPY_NONFINITE, ST_MSE_NONFINITE, ST_CHOL_FAILED, ST_COMM_TIMEOUT = 1, 2, 4, 8, 16, 32
Y_F32, Y_U8 = 0, 1


class Config(C.Structure):
    _fields_ = [("ydim", C.c_int32), ("xdim", C.c_int32), ("udim", C.c_int32), ("n_rbf", C.c_int32),
                ("n_layers", C.c_int32), ("hidden", C.c_int32 * MAX_LAYERS), ("likelihood", C.c_int32),
                ("max_trials", C.c_int32)]


class Layout(C.Structure):
    _fields_ = [("lik_logvar", C.c_int64), ("dec_w", C.c_int64), ("dec_b", C.c_int64),
                ("mlp_w", C.c_int64 * MAX_LAYERS), ("mlp_b", C.c_int64 * MAX_LAYERS),
                ("head_m_w", C.c_int64), ("head_v_w", C.c_int64), ("head_v_b", C.c_int64), ("n_train", C.c_int64),
                ("prior_mean", C.c_int64), ("prior_logvar", C.c_int64), ("tr_logvar", C.c_int64),
                ("centroid", C.c_int64), ("logwidth", C.c_int64), ("w_mean", C.c_int64), ("w_chol", C.c_int64),
                ("w_precision", C.c_int64), ("w_pchol", C.c_int64), ("lik_n", C.c_int64), ("tr_n", C.c_int64),
                ("total", C.c_int64)]


_P, _I32, _I64, _U32, _U64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float

# name -> (restype, argtypes); mirrors include/vjf_b200.h one to one
SIGNATURES = {
    "vjf_last_error": (C.c_char_p, []),
    "vjf_version": (C.c_int, []),
    "vjf_get_layout": (C.c_int, [C.POINTER(Config), C.POINTER(Layout)]),
    "vjf_create": (C.c_int, [C.POINTER(Config), _P, C.POINTER(_P)]),
    "vjf_destroy": (C.c_int, [_P]),
    "vjf_init_state": (C.c_int, [_P, _P]),
    "vjf_step": (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _U64, _U64, _U32, _F, _P, _P, _P, _P]),
    "vjf_run": (C.c_int, [_P, _I32, _I32, _P, _I32, _P, _P, _P, _P, _U64, _U64, _U32, _F, _P, _P, _P, _P]),
    "vjf_run_host": (C.c_int, [_P, _I32, _I32, _P, _I32, _P, _P, _U64, _U64, _U32, _F, _P, _P, _P, _I32]),
    "vjf_reduce_size": (_I64, [_P]),
    "vjf_reduce_buffer": (_P, [_P]),
    "vjf_step_phase_a": (C.c_int, [_P, _I32, _I32, _P, _I32, _P, _P, _P, _P, _U64, _U64, _U64, _U32, _P, _P, _P]),
    "vjf_step_phase_b": (C.c_int, [_P, _I32, _U32, _F, _P, _P]),
    "vjf_comm_local_handle": (C.c_int, [_P, _P]),
    "vjf_comm_connect": (C.c_int, [_P, _I32, _I32, _P]),
    "vjf_run_sharded": (C.c_int, [_P, _I32, _I32, _I32, _U64, _P, _I32, _P, _P, _P, _P, _U64, _U64, _U32, _F, _P, _P, _P, _P]),
    "vjf_run_sharded_host": (C.c_int, [_P, _I32, _I32, _I32, _U64, _P, _I32, _P, _P, _U64, _U64, _U32, _F, _P, _P, _P, _I32]),
    "vjf_set_rls_precision": (C.c_int, [_P, _I32]),
    "vjf_bigr_buffer": (_P, [_P, _I32]),
    "vjf_wide_stamps": (_P, [_P]),
    "vjf_get_status": (C.c_int, [_P, _P, C.POINTER(_U32), _I32]),
    "vjf_philox_normal": (C.c_int, [_U64, _U64, _U64, _I32, _I32, _P, _P]),
    "vjf_launch_count": (_I64, []),
    "vjf_last_launch_kind": (_I32, []),
    "vjf_set_tile_mode": (C.c_int, [_I32]),
    "vjf_rls_initialize": (C.c_int, [_P, _I64, _P, _P, _P, _P]),
    "vjf_weight_kalman": (C.c_int, [_P, _I64, _P, _P, _P, _F, _F, _P]),
    "vjf_forecast": (C.c_int, [_P, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "vjf_kalman_predict_batched": (C.c_int, [_I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vjf_kalman_update_batched": (C.c_int, [_I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vjf_kalman_joseph_update_batched": (C.c_int, [_I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vjf_symmetrize_batched": (C.c_int, [_I32, _I32, _P, _P, _P]),
    "vjf_positivize_batched": (C.c_int, [_I32, _I32, _P, _F, _P, _P]),
}

_lib = None


def lib_path():
    return os.environ.get("VJF_B200_LIB") or _build.LIB  # VJF_B200_LIB: development builds (python -m vjf_b200.build --debug)


def load(build_if_missing=True):
    """Load the CUDA library.  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise RuntimeError(f"{path} is missing: run `python -m vjf_b200.build` (there is no CPU fallback)")
        _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError("vjf_b200: " + load().vjf_last_error().decode())


class DevBuf:
    """Exposes a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def make_config(ydim, xdim, udim, n_rbf, hidden_sizes, likelihood, max_trials):
    hs = list(hidden_sizes)
    if not 1 <= len(hs) <= MAX_LAYERS:
        raise ValueError(f"hidden_sizes must have 1..{MAX_LAYERS} entries")
    c = Config()
    c.ydim, c.xdim, c.udim, c.n_rbf, c.n_layers = ydim, xdim, udim, n_rbf, len(hs)
    for i, h in enumerate(hs):
        c.hidden[i] = int(h)
    c.likelihood = LIK[likelihood]
    c.max_trials = int(max_trials)
    return c


def get_layout(cfg):
    lay = Layout()
    check(load().vjf_get_layout(C.byref(cfg), C.byref(lay)))
    return lay
