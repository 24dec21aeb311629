"""ctypes binding of the C ABI in include/oi_b200.h (liboi_b200.so, built in-tree by
``optimalinterpolation_b200/csrc/Makefile``).  There is no CPU fallback: a missing library or a
missing sm_100 device raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "liboi_b200.so")

EXPORTS = ("oi_version", "oi_last_error", "oi_create", "oi_destroy", "oi_set_observations", "oi_set_cells",
           "oi_gather_neighbours", "oi_get_neighbours", "oi_nlml_grad", "oi_run", "oi_get_results",
           "oi_get_stats", "oi_gpr_day", "oi_set_stream", "oi_sizeof_params", "oi_sizeof_stats", "oi_set_time_window", "oi_debug_trace", "oi_get_debug_trace",
           "oi_set_shared_queue", "oi_unlink_shared_queue", "oi_get_owned")


class OiParams(C.Structure):
    _fields_ = [("radius_m", C.c_double), ("t_pred", C.c_double), ("prior_mean", C.c_double),
                ("n_hyp", C.c_int32), ("mode", C.c_int32), ("grad_convention", C.c_int32), ("maxiter", C.c_int32),
                ("x0", C.c_double * 6), ("gtol", C.c_double), ("scratch_gib", C.c_double),
                ("max_active", C.c_int32), ("n_groups", C.c_int32), ("engine", C.c_int32), ("group_size", C.c_int32),
                ("evals_per_launch", C.c_int32), ("optimiser", C.c_int32)]


class OiStats(C.Structure):
    _fields_ = [("ms_total", C.c_double), ("ms_gather", C.c_double), ("flops", C.c_double),
                ("n_evals", C.c_int64), ("n_launches", C.c_int64), ("n_iterations", C.c_int64),
                ("sum_n", C.c_int64), ("ms_factor", C.c_double), ("flops_factor", C.c_double),
                ("ms_build", C.c_double), ("ms_chol", C.c_double), ("ms_fwd", C.c_double), ("ms_trtri", C.c_double),
                ("ms_alpha", C.c_double), ("ms_lauum", C.c_double), ("ms_finalize", C.c_double),
                ("flops_chol", C.c_double), ("flops_trtri", C.c_double), ("flops_lauum", C.c_double),
                ("launches_chol", C.c_int64), ("launches_trtri", C.c_int64), ("launches_lauum", C.c_int64),
                ("n_groups", C.c_int64), ("group_size", C.c_int64), ("launches_persistent", C.c_int64), ("n_graph_captures", C.c_int64), ("n_graph_launches", C.c_int64), ("ms_graph", C.c_double),
                ("n_express_cells", C.c_int64),
                ("ms_persistent", C.c_double), ("cycles_phase", C.c_double * 8)]


_lib = None


def build(verbose: bool = False) -> str:
    """Compile liboi_b200.so for sm_100a with nvcc (works without a GPU)."""
    import subprocess
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout, out.stderr)
    if out.returncode:
        raise RuntimeError("building liboi_b200.so failed")
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(optimalinterpolation_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, dp, ip, lp = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    L.oi_version.restype = C.c_int
    L.oi_last_error.restype = C.c_char_p
    L.oi_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.oi_destroy.argtypes = [vp]
    L.oi_set_stream.argtypes = [vp, vp]
    L.oi_destroy.restype = None
    L.oi_set_observations.argtypes = [vp, dp, dp, dp, dp, C.c_int64]
    L.oi_set_cells.argtypes = [vp, dp, C.c_int64]
    L.oi_set_time_window.argtypes = [vp, C.c_double, C.c_double]
    L.oi_gather_neighbours.argtypes = [vp, C.c_double, ip]
    L.oi_get_neighbours.argtypes = [vp, lp, ip]
    L.oi_nlml_grad.argtypes = [vp, dp, C.c_int32, C.c_double, C.c_int32, dp, dp]
    L.oi_run.argtypes = [vp, C.POINTER(OiParams), dp]
    L.oi_get_results.argtypes = [vp, dp, ip, ip, ip]
    L.oi_get_stats.argtypes = [vp, C.POINTER(OiStats)]
    L.oi_gpr_day.argtypes = [vp, dp, dp, dp, dp, C.c_int64, dp, C.c_int64, C.POINTER(OiParams), dp, dp, ip, ip, ip]
    L.oi_debug_trace.argtypes = [vp, C.c_int64, C.c_int32]
    L.oi_set_shared_queue.argtypes = [vp, C.c_char_p]
    L.oi_unlink_shared_queue.argtypes = [C.c_char_p]
    L.oi_get_owned.argtypes = [vp, dp]
    L.oi_get_debug_trace.argtypes = [vp, dp, ip]
    L.oi_sizeof_params.restype = C.c_int
    L.oi_sizeof_stats.restype = C.c_int
    if L.oi_sizeof_params() != C.sizeof(OiParams) or L.oi_sizeof_stats() != C.sizeof(OiStats):
        raise RuntimeError("liboi_b200.so does not match this binding (struct sizes differ): rebuild the library")
    _lib = L
    return L
