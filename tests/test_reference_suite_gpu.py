"""The reference's OWN minunit suite (test/run_tests.c, 57 tests: every cone type, QPs with P,
infeasible/unbounded certificates, warm starts, scs_update, AA options ...) with the reference's
unmodified driver, cones and AA, linked against OUR linear-system plugin
(scs_b200/libscs_b200_linsys.so) -- i.e. the link-time drop-in of INTEGRATION.md section 1.
The binary oracle/_ref/run_tests_b200 is built by `make -C oracle ref` where /root/reference exists."""
import os
import subprocess

import pytest

from conftest import REF_DIR

pytestmark = pytest.mark.gpu


def test_reference_suite_through_b200_linsys_plugin():
    exe = os.path.join(REF_DIR, "run_tests_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/run_tests_b200 not built")
    ob = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs"
    env = dict(os.environ, LD_LIBRARY_PATH=ob + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    p = subprocess.run([exe], cwd=REF_DIR, env=env, capture_output=True, text=True, timeout=1200)
    tail = "\n".join(p.stdout.splitlines()[-15:])
    print(tail)
    # the binary must really be linked against OUR plugin: (i) the dynamic loader resolves libscs_b200_linsys.so,
    # (ii) the solver header the reference prints (src/scs.c:127 "lin-sys:  %s") names our backend
    ldd = subprocess.run(["ldd", exe], env=env, capture_output=True, text=True).stdout
    assert "libscs_b200_linsys.so" in ldd, ldd
    assert "libscsindir" not in ldd, ldd
    assert "sparse-indirect-b200" in p.stdout, tail
    assert "sparse-indirect-scs" not in p.stdout
    assert "ALL TESTS PASSED" in p.stdout, tail + p.stderr[-2000:]
    assert "Tests run: 57" in p.stdout
