"""ctypes binding of include/pop_arrowhead.h (the C ABI).  There is no CPU fallback: if the
shared library is missing or no CUDA device is present the calls fail loudly."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("POP_ARROWHEAD_LIB") or os.path.join(_HERE, "libpop_arrowhead.so")   # the override serves A/B runs of two builds

POP_OK = 0
ERR_NAMES = {-1: "POP_ERR_ARG", -2: "POP_ERR_DIM", -3: "POP_ERR_STRUCT", -4: "POP_ERR_CUDA", -5: "POP_ERR_ALLOC", -6: "POP_ERR_NOGPU"}
BASIS_C1, BASIS_DIRICHLET = 0, 1
LEFT, RIGHT = 0, 1
GENERAL, SYM_UPPER = 0, 1
UPPER, LOWER = 0, 1
SOLVE_AUTO, SOLVE_STAGED, SOLVE_FUSED, SOLVE_FUSED_SMEM = 0, 1, 2, 3


class PopError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class DimensionMismatch(PopError, ValueError):
    pass


class StructureError(PopError, AssertionError):
    pass


class PosDefException(PopError, ArithmeticError):
    """Non-positive pivot at 1-based index `.code` (the reference throws PosDefException)."""


class pop_desc(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("m", "nh", "q", "nB", "nC", "lD", "uD", "uniform")]


def build(verbose=False):
    """Compile libpop_arrowhead.so for sm_100a in-tree (nvcc, no GPU needed)."""
    cmd = ["make", "-C", _HERE, "-j8"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libpop_arrowhead.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


_lib = None
dp = C.POINTER(C.c_double)
vp = C.c_void_p
i64 = C.c_int64

_PROTOS = {
    "pop_abi_version": (C.c_int, []),
    "pop_last_error": (C.c_char_p, []),
    "pop_create": (C.c_int, [C.POINTER(vp), C.c_int, vp]),
    "pop_destroy": (C.c_int, [vp]),
    "pop_set_stream": (C.c_int, [vp, vp]),
    "pop_sync": (C.c_int, [vp]),
    "pop_launch_count": (i64, [vp]),
    "pop_panel_alloc": (C.c_int, [vp, i64, i64, C.POINTER(vp), C.POINTER(i64)]),
    "pop_panel_free": (C.c_int, [vp, vp]),
    "pop_panel_upload": (C.c_int, [vp, vp, i64, vp, i64, i64, i64]),
    "pop_panel_download": (C.c_int, [vp, vp, i64, vp, i64, i64, i64]),
    "pop_arrowhead_upload": (C.c_int, [vp, C.POINTER(pop_desc), vp, vp, vp, vp, vp, vp, C.POINTER(vp)]),
    "pop_arrowhead_desc": (C.c_int, [vp, C.POINTER(pop_desc)]),
    "pop_arrowhead_download": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, vp, vp, vp]),
    "pop_arrowhead_free": (C.c_int, [vp, vp]),
    "pop_arrowhead_as_factor": (C.c_int, [vp, vp]),
    "pop_arrowhead_copy": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "pop_assemble": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(vp), C.POINTER(vp)]),
    "pop_assemble_mesh": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, C.POINTER(vp), C.POINTER(vp)]),
    "pop_axpby": (C.c_int, [vp, C.c_double, vp, C.c_double, vp, C.POINTER(vp)]),
    "pop_revchol": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "pop_revchol_inplace": (C.c_int, [vp, vp]),
    "pop_revchol_shifted_batch": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, C.POINTER(vp), vp]),
    "pop_mul": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_double, vp, i64, C.c_double, vp, i64, i64]),
    "pop_mul_host": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_double, vp, i64, C.c_double, vp, i64, i64]),
    "pop_trsm": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, i64, i64]),
    "pop_trsm_host": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, i64, i64]),
    "pop_solve": (C.c_int, [vp, vp, C.c_int, vp, i64, i64, C.c_int]),
    "pop_solve_host": (C.c_int, [vp, vp, C.c_int, vp, i64, i64, C.c_int]),
    "pop_solve_host_to": (C.c_int, [vp, vp, C.c_int, vp, i64, vp, i64, i64, C.c_int]),
    "pop_solve_shift_batch": (C.c_int, [vp, C.POINTER(vp), C.c_int, vp, i64, i64, i64, C.c_int]),
    "pop_solve_mul_axpy": (C.c_int, [vp, vp, vp, C.c_double, vp, i64, vp, i64, i64]),
    "pop_dist_handle_bytes": (C.c_int, []),
    "pop_dist_create": (C.c_int, [vp, C.c_int, C.c_int, i64, C.POINTER(vp)]),
    "pop_dist_get_handle": (C.c_int, [vp, vp]),
    "pop_dist_open": (C.c_int, [vp, vp]),
    "pop_dist_destroy": (C.c_int, [vp, vp]),
    "pop_dist_buffer": (vp, [vp, C.c_int]),
    "pop_dist_ld": (i64, [vp]),
    "pop_dist_lo": (i64, [vp, C.c_int]),
    "pop_dist_transpose": (C.c_int, [vp, vp, C.c_int, C.c_int]),
    "pop_adi_poisson_dist": (C.c_int, [vp, vp, vp, vp, vp, i64, vp, i64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]),
    "pop_heat_steps_dist": (C.c_int, [vp, vp, vp, vp, C.c_double, C.c_int, vp, i64]),
    "pop_dd_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "pop_dd_get_handle": (C.c_int, [vp, vp]),
    "pop_dd_open": (C.c_int, [vp, vp]),
    "pop_dd_destroy": (C.c_int, [vp, vp]),
    "pop_dd_nloc": (i64, [vp]),
    "pop_dd_ld": (i64, [vp]),
    "pop_dd_column": (i64, [vp, i64]),
    "pop_dd_row_half_step": (C.c_int, [vp, vp, vp, vp, C.c_double, vp, i64, vp, i64]),
    "pop_adi_poisson_dd": (C.c_int, [vp, vp, vp, vp, vp, i64, vp, i64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]),
    "pop_legendre_analysis_2d": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, i64, vp, i64]),
    "pop_weakform_apply": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_double, vp, C.c_int, vp, i64, vp, i64, i64]),
    "pop_legendre_to_c1": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, i64, vp, i64, i64, vp]),
    "pop_c1_to_legendre": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, i64, vp, i64, i64]),
    "pop_legendre_synthesis_2d": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, i64, vp, i64]),
    "pop_assemble_mesh_coef": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.POINTER(vp), C.POINTER(vp)]),
    "pop_transpose": (C.c_int, [vp, vp, i64, i64, i64, vp, i64]),
    "pop_adi_num_iterations": (C.c_int, [C.c_double] * 5),
    "pop_adi_shifts": (C.c_int, [C.c_int] + [C.c_double] * 4 + [vp, vp]),
    "pop_adi_poisson": (C.c_int, [vp, vp, vp, vp, i64, vp, i64] + [C.c_double] * 5 + [C.c_int]),
    "pop_adi_poisson_host": (C.c_int, [vp, vp, vp, vp, i64, vp, i64] + [C.c_double] * 5 + [C.c_int]),
    "pop_heat_steps": (C.c_int, [vp, vp, vp, C.c_double, C.c_int, C.c_int, vp, i64, i64]),
    "pop_eigvals_lanczos": (C.c_int, [vp, vp, vp, C.c_int, C.c_uint64, vp, vp, C.POINTER(C.c_int)]),
    "pop_heat_steps_theta": (C.c_int, [vp, vp, vp, C.c_double, C.c_double, C.c_int, C.c_int, vp, i64, i64]),
}
EXPORTED = tuple(_PROTOS)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (make -C {_HERE}); "
                               "this package has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            if "POP_ARROWHEAD_LIB" in os.environ and not hasattr(L, name):
                continue   # an older build under A/B comparison: calls to what it lacks fail with AttributeError
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what=""):
    if rc == POP_OK:
        return
    msg = lib().pop_last_error().decode()
    if rc > 0:
        raise PosDefException(rc, msg or f"{what}: matrix is not positive definite")
    if rc == -2:
        raise DimensionMismatch(rc, msg)
    if rc == -3:
        raise StructureError(rc, msg)
    raise PopError(rc, msg)


class Context:
    """One device + one stream (pop_ctx).  `stream` may be a raw cudaStream_t integer, e.g.
    torch.cuda.current_stream().cuda_stream."""

    def __init__(self, device=0, stream=None):
        self._h = vp()
        check(lib().pop_create(C.byref(self._h), int(device), vp(stream) if stream else None), "pop_create")
        self.device = int(device)

    @property
    def handle(self):
        return self._h

    def set_stream(self, stream):
        """Bind to a raw cudaStream_t handle; 0/None is the CUDA legacy default stream (torch's default)."""
        check(lib().pop_set_stream(self._h, vp(int(stream)) if stream else None))

    def sync(self):
        check(lib().pop_sync(self._h))

    @property
    def launches(self):
        return int(lib().pop_launch_count(self._h))

    def close(self):
        if self._h:
            lib().pop_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=None):
    """Context on the current torch device and stream if torch is initialised, else device 0."""
    if device is None:
        device = 0
        try:
            import torch
            if torch.cuda.is_available():
                device = torch.cuda.current_device()
        except Exception:
            pass
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
