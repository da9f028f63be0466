"""ctypes binding of the C-ABI library (include/namaste_b200.h).

There is no CPU fallback: if the shared object is missing or a call fails, this raises."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnamaste_b200.so")
if os.environ.get("NMB_LIB"):  # diagnostics: A/B builds of the same library
    LIB_PATH = os.environ["NMB_LIB"]

NMB_OK, NMB_ERR_ARG, NMB_ERR_CUDA, NMB_ERR_PRECOND, NMB_ERR_NAN = 0, -1, -2, -3, -4
MODE_BRUTE, MODE_WINDOWED, MODE_FUSED = 0, 1, 2
F64, F32, F64_STRICT, F32_PURE = 64, 32, 66, 33
MODES = {"brute": MODE_BRUTE, "windowed": MODE_WINDOWED, "fused": MODE_FUSED,
         MODE_BRUTE: MODE_BRUTE, MODE_WINDOWED: MODE_WINDOWED, MODE_FUSED: MODE_FUSED}

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int64_p = ctypes.POINTER(ctypes.c_int64)
_vp = ctypes.c_void_p

# name -> (restype, argtypes); every symbol include/namaste_b200.h declares
SIGNATURES = {
    "nmb_last_error": (ctypes.c_char_p, []),
    "nmb_version": (ctypes.c_char_p, []),
    "nmb_lc_create": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(_vp)]),
    "nmb_lc_create_batch": (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                           ctypes.POINTER(_vp)]),
    "nmb_lc_destroy": (None, [_vp]),
    "nmb_lc_ncand": (ctypes.c_int64, [_vp]),
    "nmb_lc_npoints": (ctypes.c_int64, [_vp, ctypes.c_int64]),
    "nmb_lc_cadence": (ctypes.c_double, [_vp, ctypes.c_int64]),
    "nmb_lc_oversamp": (ctypes.c_int, [_vp]),
    "nmb_lnprob": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, _vp]),
    "nmb_lnprob_host": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int]),
    "nmb_gp_lnprob": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int, _vp, ctypes.c_int, _vp, _vp, _vp, _vp]),
    "nmb_gp_lnprob_host": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int, _vp, ctypes.c_int, _vp, _vp, _vp]),
    "nmb_model": (ctypes.c_int, [_vp, ctypes.c_int64, _vp, ctypes.c_int64, _vp, _vp]),
    "nmb_occultquad": (ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_double, _vp,
                                      _vp]),
    "nmb_lnprior": (ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_int, _vp, _vp, _vp]),
    "nmb_sampler_create": (ctypes.c_int, [_vp, ctypes.c_int64, _vp, ctypes.c_uint64, ctypes.c_double, ctypes.c_int,
                                          ctypes.POINTER(_vp)]),
    "nmb_sampler_create_gp": (ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_int, _vp, ctypes.c_int, _vp, ctypes.c_uint64,
                                             ctypes.c_double, ctypes.POINTER(_vp)]),
    "nmb_sampler_destroy": (None, [_vp]),
    "nmb_sampler_set_state": (ctypes.c_int, [_vp, _vp, _vp]),
    "nmb_sampler_run": (ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]),
    "nmb_sampler_halfstep": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]),
    "nmb_sampler_advance": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int]),
    "nmb_sampler_sync": (ctypes.c_int, [_vp]),
    "nmb_sampler_split_export": (ctypes.c_int, [_vp, _vp]),
    "nmb_sampler_split_attach": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp]),
    "nmb_sampler_run_split": (ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_int]),
    "nmb_sampler_pack_half": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, _vp]),
    "nmb_sampler_unpack_half": (ctypes.c_int, [_vp, ctypes.c_int, _vp]),
    "nmb_sampler_reset": (ctypes.c_int, [_vp]),
    "nmb_sampler_set_candidate_offset": (ctypes.c_int, [_vp, ctypes.c_int64]),
    "nmb_get_device": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "nmb_set_device": (ctypes.c_int, [ctypes.c_int]),
    "nmb_sampler_len": (ctypes.c_int64, [_vp]),
    "nmb_sampler_steps": (ctypes.c_int64, [_vp]),
    "nmb_sampler_get": (ctypes.c_int, [_vp, ctypes.c_int, _vp]),
    "nmb_sampler_summary": (ctypes.c_int, [_vp, ctypes.c_int64, _vp, ctypes.c_int, _vp, _vp, ctypes.POINTER(ctypes.c_int64),
                                           ctypes.POINTER(ctypes.c_int64)]),
    "nmb_sampler_devptr": (_vp, [_vp, ctypes.c_int]),
    "nmb_sampler_stream": (_vp, [_vp]),
    "nmb_flat_create": (ctypes.c_int, [ctypes.POINTER(_vp)]),
    "nmb_flat_destroy": (None, [_vp]),
    "nmb_anomcut_batch": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_double, _vp]),
    "nmb_flatten_batch": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
                                         ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_double, _vp]),
    "nmb_window_mask_batch": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, ctypes.c_int64, _vp, _vp, ctypes.c_int, _vp]),
    "nmb_measure_fp64_peak": (ctypes.c_int, [_c_double_p, _c_double_p]),
    "nmb_debug_trace": (ctypes.c_int, [_vp, _vp]),
    "nmb_last_kernels": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p)]),
    "nmb_stage_timing": (ctypes.c_int, [_vp, ctypes.c_int]),
    "nmb_stage_times": (ctypes.c_int, [_vp, _c_double_p, _c_double_p, ctypes.POINTER(ctypes.c_int64)]),
    "nmb_host_checksum": (ctypes.c_int, [_vp, ctypes.c_int64, _vp]),
    "nmb_launch_count": (ctypes.c_int64, []),
}

_lib = None


class NamasteB200Error(RuntimeError):
    pass


def load():
    """Load libnamaste_b200.so (built in-tree by namaste_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NamasteB200Error(
            "namaste_b200: %s is missing. Build it with `python -m namaste_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    """Translate a C-ABI status into the exception the reference's Python would raise."""
    if rc == NMB_OK:
        return
    msg = load().nmb_last_error().decode("utf-8", "replace")
    if rc in (NMB_ERR_ARG, NMB_ERR_PRECOND, NMB_ERR_NAN):
        raise ValueError("%s: %s" % (what or "namaste_b200", msg))
    raise NamasteB200Error("%s: %s (code %d)" % (what or "namaste_b200", msg, rc))


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a):
    """void* of a contiguous numpy array, a torch tensor, an int address or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    if isinstance(a, np.ndarray):
        return ctypes.c_void_p(a.ctypes.data)
    return ctypes.c_void_p(a.data_ptr())  # torch.Tensor
