"""ctypes driver for the host build of the optimiser state machine (test helper)."""
import ctypes, os, subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _lib():
    so = os.path.join(_HERE, "_build", "libcg_host.so")
    src = os.path.join(_HERE, "cg_host.cpp")
    hdrs = [os.path.join(_HERE, "..", "optimalinterpolation_b200", "csrc", h) for h in ("cg_scipy.h", "lbfgs_fast.h", "oi_shared_queue.h")]
    if (not os.path.exists(so)) or os.path.getmtime(so) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src, "-lrt", "-pthread"])
    L = ctypes.CDLL(so)
    L.cgh_new.restype = ctypes.c_void_p
    L.cgh_free.argtypes = [ctypes.c_void_p]
    L.cgh_init.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double]
    L.cgh_resume.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]
    L.cgh_resume.restype = ctypes.c_int
    for nm in ("cgh_req_x", "cgh_x"):
        getattr(L, nm).argtypes = [ctypes.c_void_p]
        getattr(L, nm).restype = ctypes.POINTER(ctypes.c_double)
    L.cgh_fval.argtypes = [ctypes.c_void_p]; L.cgh_fval.restype = ctypes.c_double
    for nm in ("cgh_status", "cgh_nit", "cgh_nfev"):
        getattr(L, nm).argtypes = [ctypes.c_void_p]; getattr(L, nm).restype = ctypes.c_int
    L.sq_new.restype = ctypes.c_void_p
    L.sq_free.argtypes = [ctypes.c_void_p]
    L.sq_attach.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    L.sq_unlink.argtypes = [ctypes.c_char_p]
    L.sq_begin.argtypes = [ctypes.c_void_p, ctypes.c_uint]
    L.sq_take.argtypes = [ctypes.c_void_p, ctypes.c_int]; L.sq_take.restype = ctypes.c_long
    L.lbh_new.restype = ctypes.c_void_p
    L.lbh_free.argtypes = [ctypes.c_void_p]
    L.lbh_init.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double]
    L.lbh_resume.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]
    L.lbh_resume.restype = ctypes.c_int
    for nm in ("lbh_req_x", "lbh_x"):
        getattr(L, nm).argtypes = [ctypes.c_void_p]
        getattr(L, nm).restype = ctypes.POINTER(ctypes.c_double)
    L.lbh_fval.argtypes = [ctypes.c_void_p]; L.lbh_fval.restype = ctypes.c_double
    for nm in ("lbh_status", "lbh_nit", "lbh_nfev"):
        getattr(L, nm).argtypes = [ctypes.c_void_p]; getattr(L, nm).restype = ctypes.c_int
    return L


def minimize_cg(fun, x0, maxiter=0, gtol=1e-5, max_evals=100000):
    """fun(x) -> (f, g).  Returns dict(x, fun, status, nit, nfev, trace)."""
    L = _lib()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    dim = len(x0)
    s = L.cgh_new()
    L.cgh_init(s, x0.ctypes.data, dim, maxiter, gtol)
    f, g = 0.0, np.zeros(dim)
    trace = []
    for _ in range(max_evals):
        rc = L.cgh_resume(s, float(f), g.ctypes.data)
        if rc != 0:
            break
        x = np.array([L.cgh_req_x(s)[i] for i in range(dim)])
        f, g = fun(x)
        f = float(np.asarray(f).reshape(-1)[0])
        g = np.ascontiguousarray(g, dtype=np.float64)
        trace.append(x)
    out = dict(x=np.array([L.cgh_x(s)[i] for i in range(dim)]), fun=L.cgh_fval(s),
               status=L.cgh_status(s), nit=L.cgh_nit(s), nfev=L.cgh_nfev(s), trace=trace)
    L.cgh_free(s)
    return out


def minimize_lbfgs(fun, x0, maxiter=0, pgtol=1e-5, max_evals=100000):
    """The fast-mode optimiser (csrc/lbfgs_fast.h).  fun(x) -> (f, g) with the TRUE gradient."""
    L = _lib()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    dim = len(x0)
    s = L.lbh_new()
    L.lbh_init(s, x0.ctypes.data, dim, maxiter, pgtol)
    f, g = 0.0, np.zeros(dim)
    trace = []
    for _ in range(max_evals):
        rc = L.lbh_resume(s, float(f), g.ctypes.data)
        if rc != 0:
            break
        x = np.array([L.lbh_req_x(s)[i] for i in range(dim)])
        f, g = fun(x)
        f = float(np.asarray(f).reshape(-1)[0])
        g = np.ascontiguousarray(g, dtype=np.float64)
        trace.append(x)
    out = dict(x=np.array([L.lbh_x(s)[i] for i in range(dim)]), fun=L.lbh_fval(s),
               status=L.lbh_status(s), nit=L.lbh_nit(s), nfev=L.lbh_nfev(s), trace=trace)
    L.lbh_free(s)
    return out
