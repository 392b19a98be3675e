"""ctypes binding of the C ABI in include/openifem_b200.h (the same stub a
reference-side maintainer would write). Loading never falls
back to anything else: a missing library raises ImportError-like RuntimeError and
every compute call fails with IFEM_ERR_NO_DEVICE when there is no GPU."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libopenifem_b200.so")


class InsControl(C.Structure):
    _fields_ = [("fgmres_rel", C.c_double), ("fgmres_floor", C.c_double), ("cg_mp_rel", C.c_double),
                ("cg_sm_rel", C.c_double), ("cg_floor", C.c_double), ("a_inv_rel", C.c_double),
                ("a_inv_max_it", C.c_int), ("basis_size", C.c_int), ("a_inv_fp32", C.c_int),
                ("cg_sm_fp32", C.c_int), ("supg_ilu", C.c_int)]


class NewtonRecord(C.Structure):
    _fields_ = [("timestep", C.c_uint), ("iteration", C.c_uint), ("abs_res", C.c_double), ("rel_res", C.c_double),
                ("gmres_its", C.c_int), ("gmres_res", C.c_double), ("cg_mp_its", C.c_int), ("cg_sm_its", C.c_int),
                ("a_inv_its", C.c_int), ("precond_applies", C.c_int), ("true_res", C.c_double)]


class SolidRecord(C.Structure):
    _fields_ = [("timestep", C.c_uint), ("iteration", C.c_uint), ("res_F", C.c_double), ("res_U", C.c_double), ("cg_its", C.c_int)]


_lib = None


class IfemError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m openifem_b200.build` (no fallback path exists)")
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.ifem_last_error.restype = C.c_char_p
    return _lib


def check(rc):
    if rc != 0:
        raise IfemError(lib().ifem_last_error().decode())


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def lptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))
