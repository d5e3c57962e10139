"""Shared test helpers: build the oracle's Spec from the package's problem description and
call the C restatement (oracle/_ref/liboracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import hvi_b200  # noqa: E402
from oracle import hvi_oracle as O  # noqa: E402


def oracle_spec(prob) -> "O.Spec":
    d = prob.as_plain_dict()
    return O.Spec(**d)


_clib = None


def c_oracle_lib():
    global _clib
    if _clib is None:
        p = os.path.join(ROOT, "oracle", "_ref", "liboracle.so")
        if not os.path.exists(p):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        _clib = C.CDLL(p)
    return _clib


def c_oracle_value_and_grad(prob, phi, data, zP, zMs, nthreads=1, count_globals=True, want_grad=True, i_sites=None):
    lib = c_oracle_lib()
    ft = prob.np_dtype
    fn = getattr(lib, "hvi_oracle_neg_elbo_grad_" + prob.dtype)
    desc = prob.to_desc(0, count_globals)
    arr = lambda a: np.asfortranarray(np.asarray(a, dtype=ft))
    xM, xP, yo, yu, zP_, zMs_ = map(arr, (data["xM"], data["xP"], data["y_o"], data["y_unc"], zP, zMs))
    phi_ = np.ascontiguousarray(phi, dtype=ft)
    nb, M = xM.shape[1], zMs_.shape[0]
    comps = (C.c_double * 6)()
    nL = C.c_double()
    grad = np.zeros(phi_.shape, dtype=ft)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int64, C.c_int] + [C.c_void_p] * 7 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    isp = None if i_sites is None else np.ascontiguousarray(i_sites, dtype=np.int64)
    rc = fn(C.byref(desc), nb, M, p(phi_), p(xM), p(xP), p(yo), p(yu), p(zP_), p(zMs_),
            C.cast(comps, C.c_void_p), C.cast(C.byref(nL), C.c_void_p), p(grad) if want_grad else None, nthreads,
            None if isp is None else p(isp))
    assert rc == 0, rc
    return nL.value, np.array(list(comps)), grad
