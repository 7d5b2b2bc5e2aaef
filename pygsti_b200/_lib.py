"""ctypes binding of libb200fwdsim.so (the C ABI declared in include/b200_fwdsim.h).

The library is the product: if it is missing or cannot be loaded this module raises -- there is no
Python/numpy fallback for any compute entry point.
"""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200fwdsim.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "b200_fwdsim.h")

_lib = None

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
vp = C.c_void_p

# name -> (restype, argtypes); must list every function the header declares (tests check this)
SIGNATURES = {
    "b200_version": (C.c_int, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "b200_ctx_create": (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
    "b200_ctx_destroy": (C.c_int, [vp]),
    "b200_ctx_sync": (C.c_int, [vp]),
    "b200_ctx_launch_count": (C.c_int, [vp, c_i64p]),
    "b200_ctx_phase_timing": (C.c_int, [vp, C.c_int]),
    "b200_ctx_set_jtj_mode": (C.c_int, [vp, C.c_int]),
    "b200_ctx_phase_ms": (C.c_int, [vp, C.POINTER(C.c_double), c_i64p]),
    "b200_atom_upload": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int64, vp, vp, vp, vp, vp, C.c_int32, vp, vp, vp, C.c_int64,
                                   C.POINTER(vp)]),
    "b200_atom_free": (C.c_int, [vp, vp]),
    "b200_atom_info": (C.c_int, [vp, c_i64p]),
    "b200_atom_set_model": (C.c_int, [vp, vp, vp, vp, vp]),
    "b200_atom_set_model_factored": (C.c_int, [vp, vp, C.c_int32, vp, vp, vp, vp, vp, C.c_int64, vp, vp]),
    "b200_atom_set_derivs": (C.c_int, [vp, vp, C.c_int64, C.c_int32, C.c_int64, vp, vp, vp]),
    "b200_atom_set_derivs_factored": (C.c_int, [vp, vp, C.c_int64, C.c_int32, C.c_int64, vp, vp, vp]),
    "b200_atom_bind_params": (C.c_int, [vp, vp, C.c_int32, vp]),
    "b200_atom_set_params": (C.c_int, [vp, vp, C.c_int32, vp]),
    "b200_atom_set_params_dev": (C.c_int, [vp, vp, C.c_int32, vp]),
    "b200_atom_get_model": (C.c_int, [vp, vp, C.c_int64, vp]),
    "b200_lindblad_members": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp]),
    "b200_fill_probs": (C.c_int, [vp, vp, vp, C.c_int64]),
    "b200_fill_dprobs": (C.c_int, [vp, vp, vp, C.c_int64, vp, C.c_int64]),
    "b200_fill_dprobs_fd": (C.c_int, [vp, vp, C.c_double, vp, C.c_int64, vp, C.c_int64]),
    "b200_fill_hprobs_linear": (C.c_int, [vp, vp, C.c_int32, vp, C.c_int32, vp, vp]),
    "b200_fill_hprobs": (C.c_int, [vp, vp, C.c_int32, vp, C.c_int32, vp, C.c_int64, vp, vp, vp, vp, vp]),
    "b200_hessian_block": (C.c_int, [vp, vp, C.c_int32, vp, C.c_int32, vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp]),
    "b200_fill_dprobs_scaled": (C.c_int, [vp, vp, vp, vp, C.c_int64, vp, C.c_int64]),
    "b200_jtj": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "b200_jtj_dev": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "b200_fill_probs_dev": (C.c_int, [vp, vp, vp]),
    "b200_fill_dprobs_dev": (C.c_int, [vp, vp, vp, C.c_int64, vp]),
    "b200_peer_alloc": (C.c_int, [vp, C.c_int64, C.POINTER(vp), vp]),
    "b200_peer_open": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "b200_peer_close": (C.c_int, [vp, vp]),
    "b200_peer_free": (C.c_int, [vp, vp]),
    "b200_fill_dprobs_bcast_dev": (C.c_int, [vp, vp, vp, C.c_int64, vp, C.c_int, vp, vp]),
    "b200_host_alloc": (C.c_int, [C.POINTER(vp), C.c_int64]),
    "b200_host_free": (C.c_int, [vp]),
    "b200_host_register": (C.c_int, [vp, C.c_int64]),
    "b200_host_unregister": (C.c_int, [vp]),
}

E_INVALID, E_CUDA, E_NOMEM, E_STATE, E_UNSUPPORTED = -1, -2, -3, -4, -5


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200fwdsim error %d: %s" % (code, msg))
        self.code = code


def header_functions():
    """Names of all functions declared in include/b200_fwdsim.h."""
    with open(HEADER_PATH) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def load():
    """Load the shared library (raises if it is not built -- no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is not built; run `python -m pygsti_b200.build` (needs nvcc). "
                          "The B200 engine has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc == 0:
        return
    msg = load().b200_last_error().decode("utf-8", "replace")
    if rc == E_NOMEM:
        raise MemoryError("b200fwdsim: " + msg)
    raise B200Error(rc, msg)
