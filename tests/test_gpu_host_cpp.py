"""The C++ host mirror (CudaFindPairs, HydrogenBondEdges, FrameBatcher) on a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_host_mirror():
    exe = os.path.join(ROOT, "gmx-acqueduct_b200", "test_host")
    assert os.path.exists(exe), "build first: __graft_entry__.build()"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "test_host OK" in r.stdout, r.stdout + r.stderr


def test_cpp_host_mirror_fails_loudly_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    exe = os.path.join(ROOT, "gmx-acqueduct_b200", "test_host")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr
