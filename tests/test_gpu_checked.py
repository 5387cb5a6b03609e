"""The bounds-checked build of the CUDA library (make -C gmx-acqueduct_b200 checked, -DWN_CHECKS: every queue,
bitmap, row-slot and computed global index is asserted on the device and traps) runs the sanitizer workload
(profiles/sanitize_run.py: every kernel family on small batches, results compared with the oracle) to completion.
compute-sanitizer itself is closed on the GPU pool (profiles/r02/sanitizer_unavailable.txt)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gmx-acqueduct_b200", "libwn_b200_checked.so")


@pytest.mark.gpu
def test_checked_build_runs_every_kernel_family_clean():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "gmx-acqueduct_b200"), "checked"])
    env = dict(os.environ, WN_B200_LIB=LIB)
    pr = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "sanitize_run.py")], env=env, capture_output=True,
                        text=True, timeout=900)
    assert pr.returncode == 0, (pr.stdout[-2000:], pr.stderr[-2000:])
    assert "WN_CHECK failed" not in pr.stdout + pr.stderr
    assert "sanitize_run: done" in pr.stdout


def test_checked_build_compiles_the_checks_in():
    """CPU: the checked library exists after build() and carries the check strings; the release library does not."""
    rel = os.path.join(ROOT, "gmx-acqueduct_b200", "libwn_b200.so")
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "gmx-acqueduct_b200"), "checked"])
    assert b"WN_CHECK failed" in open(LIB, "rb").read()
    assert b"WN_CHECK failed" not in open(rel, "rb").read()
