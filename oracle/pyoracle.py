"""ctypes wrapper around oracle/liboracle.so (the plain-C CPU restatement).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by scs_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class OrcMatrix(C.Structure):
    _fields_ = [("x", c_dp), ("i", c_ip), ("p", c_ip), ("m", C.c_int), ("n", C.c_int)]


class OrcCone(C.Structure):
    _fields_ = [("z", C.c_int), ("l", C.c_int), ("bu", c_dp), ("bl", c_dp), ("bsize", C.c_int),
                ("q", c_ip), ("qsize", C.c_int), ("s", c_ip), ("ssize", C.c_int)]


class OrcSettings(C.Structure):
    _fields_ = [("normalize", C.c_int), ("scale", C.c_double), ("adaptive_scale", C.c_int),
                ("rho_x", C.c_double), ("max_iters", C.c_int), ("eps_abs", C.c_double),
                ("eps_rel", C.c_double), ("eps_infeas", C.c_double), ("alpha", C.c_double),
                ("acceleration_lookback", C.c_int), ("acceleration_interval", C.c_int),
                ("acceleration_type_1", C.c_int), ("acceleration_regularization", C.c_double),
                ("acceleration_relaxation", C.c_double)]


class OrcInfo(C.Structure):
    _fields_ = [("iter", C.c_int), ("status_val", C.c_int), ("scale_updates", C.c_int),
                ("pobj", C.c_double), ("dobj", C.c_double), ("res_pri", C.c_double),
                ("res_dual", C.c_double), ("gap", C.c_double), ("scale", C.c_double),
                ("cg_iters", C.c_longlong), ("lin_sys_solves", C.c_longlong),
                ("accepted_accel_steps", C.c_int), ("rejected_accel_steps", C.c_int),
                ("solve_time_ms", C.c_double), ("setup_time_ms", C.c_double)]


def dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else c_dp()


def ip(a):
    return a.ctypes.data_as(c_ip) if a is not None else c_ip()


_lib = None


def load(build=True):
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) and build:
            subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
        lib = C.CDLL(LIB)
        lib.orc_linsys_init.restype = C.c_void_p
        lib.orc_linsys_init.argtypes = [C.POINTER(OrcMatrix), c_dp]
        lib.orc_linsys_free.argtypes = [C.c_void_p]
        lib.orc_linsys_solve.argtypes = [C.c_void_p, c_dp, c_dp, C.c_double]
        lib.orc_linsys_update_diag_r.argtypes = [C.c_void_p, c_dp]
        lib.orc_linsys_last_cg_its.argtypes = [C.c_void_p]
        lib.orc_cone_init.restype = C.c_void_p
        lib.orc_cone_init.argtypes = [C.POINTER(OrcCone), C.c_int]
        lib.orc_cone_free.argtypes = [C.c_void_p]
        lib.orc_proj_dual_cone.argtypes = [C.c_void_p, c_dp, c_dp]
        lib.orc_aa_init.restype = C.c_void_p
        lib.orc_aa_init.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                    C.c_double, C.c_double, C.c_int]
        lib.orc_aa_apply.restype = C.c_double
        lib.orc_aa_apply.argtypes = [C.c_void_p, c_dp, c_dp]
        lib.orc_aa_safeguard.argtypes = [C.c_void_p, c_dp, c_dp]
        lib.orc_aa_free.argtypes = [C.c_void_p]
        lib.orc_accum_by_a.argtypes = [C.POINTER(OrcMatrix), c_dp, c_dp]
        lib.orc_accum_by_atrans.argtypes = [C.POINTER(OrcMatrix), c_dp, c_dp]
        lib.orc_default_settings.argtypes = [C.POINTER(OrcSettings)]
        lib.orc_solve.argtypes = [C.POINTER(OrcMatrix), c_dp, c_dp, C.POINTER(OrcCone),
                                  C.POINTER(OrcSettings), c_dp, c_dp, c_dp, C.POINTER(OrcInfo)]
        _lib = lib
    return _lib


class Matrix:
    def __init__(self, A_csc):
        self.x = np.ascontiguousarray(A_csc[0], dtype=np.float64)
        self.i = np.ascontiguousarray(A_csc[1], dtype=np.int32)
        self.p = np.ascontiguousarray(A_csc[2], dtype=np.int32)
        self.m, self.n = int(A_csc[3][0]), int(A_csc[3][1])
        self.c = OrcMatrix(dp(self.x), ip(self.i), ip(self.p), self.m, self.n)


class Cone:
    def __init__(self, cone):
        self.keep = {}
        k = OrcCone()
        k.z, k.l = int(cone.get("z", 0)), int(cone.get("l", 0))
        if cone.get("bu") is not None and len(cone["bu"]):
            self.keep["bu"] = np.ascontiguousarray(cone["bu"], dtype=np.float64)
            self.keep["bl"] = np.ascontiguousarray(cone["bl"], dtype=np.float64)
            k.bu, k.bl, k.bsize = dp(self.keep["bu"]), dp(self.keep["bl"]), len(self.keep["bu"]) + 1
        else:
            k.bsize = int(cone.get("bsize", 0))
        if cone.get("q") is not None and len(cone["q"]):
            self.keep["q"] = np.ascontiguousarray(cone["q"], dtype=np.int32)
            k.q, k.qsize = ip(self.keep["q"]), len(self.keep["q"])
        if cone.get("s") is not None and len(cone["s"]):
            self.keep["s"] = np.ascontiguousarray(cone["s"], dtype=np.int32)
            k.s, k.ssize = ip(self.keep["s"]), len(self.keep["s"])
        self.c = k


def proj_dual_cone(cone, x, r_y=None, reps=1):
    lib = load()
    k = Cone(cone)
    w = lib.orc_cone_init(C.byref(k.c), len(x))
    out = None
    for _ in range(reps):
        out = np.array(x, dtype=np.float64)
        lib.orc_proj_dual_cone(w, dp(out), dp(r_y))
    lib.orc_cone_free(w)
    return out


def solve(prob, **over):
    lib = load()
    A = Matrix(prob["A"])
    k = Cone(prob["cone"])
    st = OrcSettings()
    lib.orc_default_settings(C.byref(st))
    for kk, vv in over.items():
        setattr(st, kk, vv)
    b = np.ascontiguousarray(prob["b"], dtype=np.float64)
    c = np.ascontiguousarray(prob["c"], dtype=np.float64)
    x, y, s = np.zeros(A.n), np.zeros(A.m), np.zeros(A.m)
    info = OrcInfo()
    status = lib.orc_solve(C.byref(A.c), dp(b), dp(c), C.byref(k.c), C.byref(st), dp(x), dp(y), dp(s),
                           C.byref(info))
    return status, info, x, y, s
