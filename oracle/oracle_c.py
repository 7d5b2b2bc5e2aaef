"""ctypes wrappers of the CPU oracle libraries.  TEST / BASELINE INFRASTRUCTURE ONLY (see fwdsim_oracle.c).

  liboracle.so              plain-C restatement (oracle/fwdsim_oracle.c), built by oracle/Makefile
  _ref/libref_densitymx.so  the reference's own C++ reps (compiled from /root/reference) + ref_driver.cpp
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_ORACLE = os.path.join(HERE, "liboracle.so")
LIB_REF = os.path.join(HERE, "_ref", "libref_densitymx.so")


class _Atom(C.Structure):
    _fields_ = [("dim", C.c_int), ("n_ops", C.c_int), ("n_rho", C.c_int), ("n_eff", C.c_int),
                ("n_rows", C.c_int64), ("n_elements", C.c_int64), ("cache_size", C.c_int32),
                ("row_ptr", C.c_void_p), ("row_ops", C.c_void_p), ("row_istart", C.c_void_p),
                ("row_prep", C.c_void_p), ("row_icache", C.c_void_p),
                ("out_ptr", C.c_void_p), ("out_eff", C.c_void_p), ("out_el", C.c_void_p)]


def build(force=False):
    """(Re)build liboracle.so and, if /root/reference exists, oracle/_ref (make decides)."""
    if force or not os.path.exists(LIB_ORACLE) or (os.path.isdir("/root/reference") and not os.path.exists(LIB_REF)):
        subprocess.run(["make", "-C", HERE, "all"], check=True, capture_output=True)


def _p(a):
    return C.c_void_p(a.ctypes.data)


class _Holder:
    def __init__(self, t):
        self.arrs = [np.ascontiguousarray(x, dtype=np.int32) for x in
                     (t.row_ptr, t.row_ops, t.row_istart, t.row_prep, t.row_icache, t.out_ptr, t.out_eff, t.out_el)]
        self.s = _Atom(t.dim, t.n_ops, t.n_rho, t.n_eff, t.n_rows, t.n_elements, t.cache_size,
                       *[x.ctypes.data for x in self.arrs])


def csc_of(D):
    """CSC (cptr, crow, cval) of a packing.DerivMap, duplicates summed."""
    import scipy.sparse as sps
    m = sps.coo_matrix((D.vals, (D.rows, D.cols)), shape=(D.n_w, D.n_params)).tocsc()
    m.sum_duplicates()
    return (np.ascontiguousarray(m.indptr, dtype=np.int32), np.ascontiguousarray(m.indices, dtype=np.int32),
            np.ascontiguousarray(m.data, dtype=np.float64))


class Oracle:
    """which = 'port' (liboracle.so) or 'reference' (oracle/_ref/libref_densitymx.so)."""

    def __init__(self, which="port"):
        self.which = which
        path = LIB_ORACLE if which == "port" else LIB_REF
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.pre = "oracle_" if which == "port" else "ref_"

    def num_threads(self):
        return getattr(self.lib, self.pre + "num_threads")()

    def mapfill_probs(self, t, G, rho, E):
        h = _Holder(t)
        G, rho, E = (np.ascontiguousarray(x, dtype=np.float64) for x in (G, rho, E))
        out = np.full(t.n_elements, np.nan)
        rc = getattr(self.lib, self.pre + "mapfill_probs")(C.byref(h.s), _p(G), _p(rho), _p(E), _p(out))
        assert rc == 0
        return out

    def dprobs_fd(self, t, G, rho, E, D, p_lo=0, p_hi=None, eps=1e-7, n_threads=1, csc=None):
        h = _Holder(t)
        G, rho, E = (np.ascontiguousarray(x, dtype=np.float64) for x in (G, rho, E))
        cptr, crow, cval = csc if csc is not None else csc_of(D)
        p_hi = D.n_params if p_hi is None else p_hi
        out = np.full((t.n_elements, p_hi - p_lo), np.nan)
        probs = np.full(t.n_elements, np.nan)
        fn = getattr(self.lib, self.pre + "dprobs_fd")
        fn.argtypes = [C.c_void_p] * 7 + [C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        rc = fn(C.cast(C.byref(h.s), C.c_void_p), _p(G), _p(rho), _p(E), _p(cptr), _p(crow), _p(cval),
                int(p_lo), int(p_hi), float(eps), _p(out), int(p_hi - p_lo), _p(probs), int(n_threads))
        assert rc == 0
        return out, probs

    def dprobs_analytic(self, t, G, rho, E, D):
        assert self.which == "port", "the reference has no analytic Map path"
        h = _Holder(t)
        G, rho, E = (np.ascontiguousarray(x, dtype=np.float64) for x in (G, rho, E))
        cptr, crow, cval = csc_of(D)
        out = np.full((t.n_elements, D.n_params), np.nan)
        probs = np.full(t.n_elements, np.nan)
        fn = self.lib.oracle_dprobs_analytic
        fn.argtypes = [C.c_void_p] * 7 + [C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        rc = fn(C.cast(C.byref(h.s), C.c_void_p), _p(G), _p(rho), _p(E), _p(cptr), _p(crow), _p(cval),
                int(D.n_params), _p(out), int(D.n_params), _p(probs))
        assert rc == 0
        return out, probs
