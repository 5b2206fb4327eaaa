"""Lean driver for ncu: build the C2 matrix, create the linsys workspace, run a few SpMV / CG launches."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scs_b200 import capi, problems

cfg = os.environ.get("CFG", "C2")
scale = float(os.environ.get("SCALE", "1.0"))
rng = np.random.default_rng(1234)
n = int(1_000_000 * scale)
m = 3 * n
A = problems.random_sparse_csc(m, n, 10, rng)
hp = capi.HostProblem(A, np.zeros(m), np.zeros(n), {"l": m})
lib = capi.load()
dr = np.empty(n + m + 1)
dr[:n] = 1e-6
dr[n:] = 10.0
w = lib.scs_init_lin_sys_work(C.byref(hp.A), None, capi.dptr(dr))
if os.environ.get("CHECK"):
    # full-size parity of both operators against numpy (CSC scatter / gather in fp64)
    x = rng.standard_normal(n)
    yv = rng.standard_normal(m)
    mine = np.zeros(m)
    assert lib.scs_b200_accum_by_a(w, capi.dptr(x), capi.dptr(mine), 0) == 0
    ref = problems.csc_matvec(A, x)
    e1 = np.abs(mine - ref).max() / np.abs(ref).max()
    mine = np.zeros(n)
    assert lib.scs_b200_accum_by_atrans(w, capi.dptr(yv), capi.dptr(mine), 0) == 0
    ref = problems.csc_rmatvec(A, yv)
    e2 = np.abs(mine - ref).max() / np.abs(ref).max()
    print(f"full-size parity vs numpy: A x {e1:.2e}  A'y {e2:.2e}")
    assert e1 < 1e-12 and e2 < 1e-12
ab = C.c_double()
reps = int(os.environ.get("REPS", "5"))
for op in (0, 1):
    ms = lib.scs_b200_time_spmv(w, op, reps, C.byref(ab))
    print(f"spmv op{op}: {ms*1e3:.1f} us  {ab.value/ms/1e6:.0f} GB/s")
ms = lib.scs_b200_time_cg_iter(w, reps, C.byref(ab))
print(f"cg iter: {ms*1e3:.1f} us  {ab.value/ms/1e6:.0f} GB/s")
lib.scs_free_lin_sys_work(w)
