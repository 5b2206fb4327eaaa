"""Generate the committed golden fixtures under tests/golden/ by running the
UNMODIFIED reference (oracle/_ref/libscsindir_ref.so, built from /root/reference
by `make -C oracle ref`) on seeded inputs.

Run here (where /root/reference exists):  python oracle/make_golden.py
The fixtures pin both the CPU restatement (tests/test_oracle_cpu.py) and the
CUDA path (tests/test_golden_gpu.py).  TEST INFRASTRUCTURE ONLY.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scs_b200 import capi, problems  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref", "libscsindir_ref.so")


def diag_r_for(n, m, z, scale=0.1, rho_x=1e-6):
    d = np.empty(n + m + 1)
    d[:n] = rho_x
    d[n:n + z] = 1.0 / (1000.0 * scale)
    d[n + z:n + m] = 1.0 / scale
    d[n + m] = 10.0
    return d


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = capi.load_reference(REF)
    ref._scs_init_cone.restype = C.c_void_p
    ref._scs_init_cone.argtypes = [C.POINTER(capi.ScsCone), C.c_int]
    ref._scs_proj_dual_cone.argtypes = [capi.c_double_p, C.c_void_p, C.c_void_p, capi.c_double_p]
    ref._scs_finish_cone.argtypes = [C.c_void_p]

    # ---------------- linsys: spmv + solve (cold, warm)
    rng = np.random.default_rng(2024)
    m, n, col = 600, 200, 8
    A = problems.random_sparse_csc(m, n, col, rng)
    hp = capi.HostProblem(A, np.zeros(m), np.zeros(n), {"l": m})
    dr = diag_r_for(n, m, 60)
    rhs = rng.standard_normal(n + m)
    warm = 0.01 * rng.standard_normal(n)
    xv = rng.standard_normal(n)
    yv = rng.standard_normal(m)
    Ax = np.zeros(m)
    ref._scs_accum_by_a(C.byref(hp.A), capi.dptr(xv), capi.dptr(Ax))
    Aty = np.zeros(n)
    ref._scs_accum_by_atrans(C.byref(hp.A), capi.dptr(yv), capi.dptr(Aty))
    w = ref.scs_init_lin_sys_work(C.byref(hp.A), None, capi.dptr(dr))
    cold = rhs.copy()
    ref.scs_solve_lin_sys(w, capi.dptr(cold), None, 1e-12)
    wsol = rhs.copy()
    ref.scs_solve_lin_sys(w, capi.dptr(wsol), capi.dptr(warm), 1e-9)
    ref.scs_free_lin_sys_work(w)
    np.savez_compressed(os.path.join(OUT, "linsys.npz"), Ax_data=A[0], Ai=A[1], Ap=A[2], m=m, n=n,
                        diag_r=dr, rhs=rhs, warm=warm, xv=xv, yv=yv, A_xv=Ax, At_yv=Aty,
                        sol_cold_tol1e12=cold, sol_warm_tol1e9=wsol)

    # ---------------- cone projections
    cones = {
        "zl": {"z": 5, "l": 9},
        "soc": {"z": 3, "l": 4, "q": [1, 2, 3, 5, 17, 64, 300]},
        "soc_big": {"q": [9000, 3]},
        "box": {"z": 2, "l": 3, "bl": [-1.0, -0.5, 0.0, -2.0] * 10, "bu": [1.0, 0.5, 3.0, 0.0] * 10},
        "psd": {"l": 1, "s": [1, 2, 3, 5, 12]},
        "mixed": {"z": 4, "l": 6, "bl": [-1.0] * 9, "bu": [2.0] * 9, "q": [4, 30], "s": [6, 3]},
    }
    blob = {}
    for name, cone in cones.items():
        mm = capi.cone_rows(cone)
        rng = np.random.default_rng(abs(hash(name)) % 1000 + 5)
        x = rng.standard_normal(mm) * 2.0
        r_y = np.full(mm, 10.0)
        r_y[: cone.get("z", 0)] = 1.0 / 100.0
        for tag, ry in (("id", None), ("ry", r_y)):
            k, keep = capi.make_cone(cone)
            cw = ref._scs_init_cone(C.byref(k), mm)
            out = x.copy()
            ryc = None if ry is None else ry.copy()
            ref._scs_proj_dual_cone(capi.dptr(out), cw, None, capi.dptr(ryc))
            ref._scs_finish_cone(cw)
            blob[f"{name}__{tag}__out"] = out
        blob[f"{name}__x"] = x
        blob[f"{name}__ry"] = r_y
    np.savez_compressed(os.path.join(OUT, "cones.npz"), **blob)
    import json
    with open(os.path.join(OUT, "cones.json"), "w") as f:
        json.dump(cones, f, indent=1)

    # ---------------- AA sequence
    ref.aa_init.restype = C.c_void_p
    ref.aa_init.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                            C.c_double, C.c_int, C.c_int]
    ref.aa_apply.restype = C.c_double
    ref.aa_apply.argtypes = [capi.c_double_p, capi.c_double_p, C.c_void_p]
    ref.aa_safeguard.argtypes = [capi.c_double_p, capi.c_double_p, C.c_void_p]
    ref.aa_finish.argtypes = [C.c_void_p]
    blob = {}
    for type1 in (1, 0):
        dim, mem = 400, 6
        rng = np.random.default_rng(77)
        d = rng.uniform(0.2, 0.97, size=dim)
        bb = rng.standard_normal(dim)
        a = ref.aa_init(dim, mem, mem, type1, 1e-8 if type1 else 1e-12, 1.0, 1.0, 1e10, 5, 0)
        x = np.zeros(dim)
        norms = []
        for i in range(25):
            if i > 0:
                norms.append(ref.aa_apply(capi.dptr(x), capi.dptr(xp), a))
            xp = x.copy()
            x = d * x + 0.02 * np.roll(x, 7) + bb
            ref.aa_safeguard(capi.dptr(x), capi.dptr(xp), a)
        ref.aa_finish(a)
        blob[f"t{type1}_x_final"] = x
        blob[f"t{type1}_norms"] = np.array(norms)
    blob["d"] = d
    blob["b"] = bb
    np.savez_compressed(os.path.join(OUT, "aa.npz"), **blob)

    # ---------------- whole solves (reference outcome on seeded problems)
    sols = {}
    specs = {
        "lp": (300, 100, 6, {"z": 30, "l": 270}),
        "socp": (400, 100, 8, {"z": 40, "l": 120, "q": [3, 7, 30, 200]}),
        "sdp": (60 + 21 + 36 + 10, 40, 8, {"l": 60, "s": [6, 8, 4]}),
    }
    for name, (mm, nn, cc, cone) in specs.items():
        prob = problems.make_problem(mm, nn, cc, cone, seed=11)
        hp = capi.HostProblem(prob["A"], prob["b"], prob["c"], prob["cone"])
        st = capi.default_settings(ref, verbose=0, eps_abs=1e-9, eps_rel=1e-9, max_iters=20000)
        x, y, s = np.zeros(nn), np.zeros(mm), np.zeros(mm)
        sol = capi.ScsSolution(capi.dptr(x), capi.dptr(y), capi.dptr(s))
        info = capi.ScsInfo()
        status = ref.scs(C.byref(hp.data), C.byref(hp.cone), C.byref(st), C.byref(sol), C.byref(info))
        sols[f"{name}__status"] = status
        sols[f"{name}__iter"] = info.iter
        sols[f"{name}__pobj"] = info.pobj
        sols[f"{name}__dobj"] = info.dobj
        sols[f"{name}__x"] = x
        sols[f"{name}__y"] = y
        sols[f"{name}__s"] = s
        print(name, status, info.iter, info.pobj, prob["opt"])
    # one ADMM iteration (max_iters=1): the first KKT solve runs at tol 1e-12, so the result is a
    # sharp end-to-end check of equilibration + KKT solve + cone projection + un-normalisation
    rngb = np.random.default_rng(5)
    specs1 = dict(specs)
    specs1["box"] = (300, 80, 6, {"z": 20, "l": 180, "bl": -rngb.uniform(0.5, 1.5, 99), "bu": rngb.uniform(0.5, 1.5, 99)})
    specs1["mixed"] = (None, 300, 12, {"z": 10, "l": 50, "bl": -rngb.uniform(0.5, 1.5, 19), "bu": rngb.uniform(0.5, 1.5, 19),
                                       "q": [5, 9000, 12], "s": [5, 1, 9]})
    for name, (mm, nn, cc, cone) in specs1.items():
        if mm is None:
            mm = capi.cone_rows(cone)
        prob = problems.make_problem(mm, nn, cc, cone, seed=11)
        for mi in (1, 3):
            hp = capi.HostProblem(prob["A"], prob["b"], prob["c"], prob["cone"])
            st = capi.default_settings(ref, verbose=0, max_iters=mi)
            x, y, s = np.zeros(nn), np.zeros(mm), np.zeros(mm)
            sol = capi.ScsSolution(capi.dptr(x), capi.dptr(y), capi.dptr(s))
            info = capi.ScsInfo()
            ref.scs(C.byref(hp.data), C.byref(hp.cone), C.byref(st), C.byref(sol), C.byref(info))
            sols[f"{name}__it{mi}_x"] = x
            sols[f"{name}__it{mi}_y"] = y
            sols[f"{name}__it{mi}_s"] = s
        if name in ("box", "mixed"):
            sols[f"{name}__bl"] = np.asarray(cone["bl"])
            sols[f"{name}__bu"] = np.asarray(cone["bu"])
    np.savez_compressed(os.path.join(OUT, "solves.npz"), **sols)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
