"""Golden record for the C3 (LP + box cone) converged solve: the UNMODIFIED reference (oracle/_ref/libscsindir_ref.so)
with its default settings on problems.config("C3", scale=0.002).  The reference needs ~3000 ADMM iterations = minutes
of host time for this, too long to repeat inside the GPU suite: its answer travels as tests/golden/c3_converged.json.

Run here:  python oracle/make_golden_c3.py          TEST INFRASTRUCTURE ONLY.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scs_b200 import capi, problems  # noqa: E402

SCALE = 0.002


def main():
    ref = capi.load_reference(os.path.join(ROOT, "oracle", "_ref", "libscsindir_ref.so"))
    prob = problems.config("C3", scale=SCALE)
    hp = capi.HostProblem(prob["A"], prob["b"], prob["c"], prob["cone"])
    st = capi.default_settings(ref, verbose=0)
    x, y, s = np.zeros(hp.n), np.zeros(hp.m), np.zeros(hp.m)
    sol = capi.ScsSolution(capi.dptr(x), capi.dptr(y), capi.dptr(s))
    info = capi.ScsInfo()
    status = ref.scs(C.byref(hp.data), C.byref(hp.cone), C.byref(st), C.byref(sol), C.byref(info))
    rec = {"config": "C3", "scale": SCALE, "settings": "defaults (eps 1e-4, acceleration_lookback 10)",
           "status": int(status), "status_str": info.status.decode(), "iter": int(info.iter),
           "pobj": float(info.pobj), "dobj": float(info.dobj), "res_pri": float(info.res_pri),
           "res_dual": float(info.res_dual), "gap": float(info.gap), "generator_opt": float(prob["opt"]),
           "x_inf_norm": float(np.abs(x).max()), "n": int(hp.n), "m": int(hp.m)}
    with open(os.path.join(ROOT, "tests", "golden", "c3_converged.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(rec)


if __name__ == "__main__":
    main()
