"""ctypes binding of libbevr.so (the C ABI declared in include/bevr.h).

There is no CPU fallback: if the library is missing it is built in place with
nvcc (``make`` in csrc/), and if that fails the import raises.  The host side
only ever passes raw device pointers + sizes across this boundary.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# BEVR_LIB: tuning sweeps load a variant build of the same ABI (scripts/gpu_sweep_*.sh); unset in normal use
LIB_PATH = os.environ.get("BEVR_LIB") or os.path.join(_HERE, "libbevr.so")
CSRC = os.path.join(_HERE, "csrc")

_p, _i, _f, _z = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol of include/bevr.h
SIGNATURES = {
    "bevr_version": (_i, []),
    "bevr_last_error": (C.c_char_p, []),
    "bevr_msda_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p] + [_i] * 10 + [_p]),
    "bevr_msda_bwd_workspace_bytes": (_z, [_i] * 5),
    "bevr_msda_bwd": (_i, [_p] * 11 + [_i] * 10 + [_p]),
    "bevr_project_points": (_i, [_p] * 6 + [_i] * 3 + [_p]),
    "bevr_affine_warp": (_i, [_p] * 3 + [_i] * 4 + [_p]),
    "bevr_correlate_workspace_bytes": (_z, [_i] * 6),
    "bevr_correlate": (_i, [_p] * 5 + [_i] * 6 + [_p]),
    "bevr_rpe_attn_workspace_bytes": (_z, [_i] * 8),
    "bevr_rpe_attn_fwd": (_i, [_p] * 8 + [_i] * 9 + [_f, _i, _p]),
    "bevr_rpe_attn_bwd": (_i, [_p] * 15 + [_i] * 9 + [_f, _i, _p]),
    "bevr_rpe_attn_fwd_dropout": (_i, [_p] * 8 + [_i] * 9 + [_f, _i, _f, C.c_ulonglong, _p]),
    "bevr_rpe_attn_bwd_dropout": (_i, [_p] * 15 + [_i] * 9 + [_f, _i, _f, C.c_ulonglong, _p]),
    "bevr_recall_ranks": (_i, [_p, _p, _i, _p]),
    "bevr_offset_head_fwd": (_i, [_p] * 5 + [C.c_longlong, _p, _p] + [_i] * 7 + [_f, _f, _p]),
    "bevr_offset_head_bwd": (_i, [_p] * 5 + [C.c_longlong] + [_p] * 6 + [_i] * 7 + [_f, _f, _p]),
}


def build(force=False):
    """Compile csrc/*.cu for sm_100a into libbevr.so (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libbevr.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


_lib = None
EXPECTED_VERSION = 210          # BEVR_VERSION of include/bevr.h; bumped with every signature change


def _stale():
    """True if libbevr.so is missing or older than any of its sources (a stale .so with old signatures would be
    called with the new argtypes: silent memory corruption)."""
    if os.environ.get("BEVR_LIB"):
        return False
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")) or f == "Makefile"]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "bevr.h"))
    return any(os.path.getmtime(f) > t for f in srcs if os.path.exists(f))


def lib():
    global _lib
    if _lib is None:
        if _stale():
            build()
        handle = C.CDLL(LIB_PATH)
        handle.bevr_version.restype = C.c_int
        if handle.bevr_version() != EXPECTED_VERSION:
            raise RuntimeError(f"libbevr.so reports ABI version {handle.bevr_version()}, the host side expects "
                               f"{EXPECTED_VERSION}: rebuild with make -C bevrender_b200/csrc")
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


class BevrError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        raise BevrError(f"{what} failed (rc={rc}): {lib().bevr_last_error().decode()}")
