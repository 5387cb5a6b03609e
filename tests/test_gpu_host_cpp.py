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


@pytest.mark.gpu
@pytest.mark.parametrize("pbc,gpu_text", [(False, 1), (True, 1), (False, 0)])
def test_standalone_driver_writes_the_reference_files(tmp_path, pbc, gpu_text):
    """wn_waternetwork on a synthetic .gro trajectory: edges file, nodes file and statistics rows equal
    what the oracle (open boundary) / the periodic oracle gives for the same parsed coordinates."""
    import driver_cases
    exe = os.path.join(ROOT, "gmx-acqueduct_b200", "wn_waternetwork")
    assert os.path.exists(exe), "build first: __graft_entry__.build()"
    driver_cases.check_reference_files(exe, tmp_path, pbc, gpu_text)


def test_standalone_driver_fails_loudly_without_gpu(tmp_path):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    exe = os.path.join(ROOT, "gmx-acqueduct_b200", "wn_waternetwork")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    gro = tmp_path / "t.gro"
    gro.write_text("w\n    3\n    1SOL     OW    1   0.100   0.100   0.100\n    1SOL    HW1    2   0.190   0.100   0.100\n"
                   "    1SOL    HW2    3   0.070   0.190   0.100\n   1.00000   1.00000   1.00000\n")
    r = subprocess.run([exe, "-f", str(gro), "-edges", str(tmp_path / "e"), "-nodes", str(tmp_path / "n"),
                        "-ohb", "", "-obw", ""], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr
    # the topology part runs before the device is touched: nodes file of the one water
    rows = (tmp_path / "n.dat").read_text().splitlines()
    assert rows[1].split() == ["2", "0", "OW", "0", "SOL", "2", "2"]


@pytest.mark.gpu
def test_standalone_driver_buried_lists(tmp_path):
    """-buried: a different water list per frame; ids in the edge file are those of the listed waters."""
    import driver_cases
    driver_cases.check_buried_lists(os.path.join(ROOT, "gmx-acqueduct_b200", "wn_waternetwork"), tmp_path)


def _device_count():
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "gmx-acqueduct_b200", "libwn_b200.so"))
    return int(lib.wn_device_count())


@pytest.mark.gpu
def test_standalone_driver_frame_sharding_over_gpus(tmp_path):
    """wn_waternetwork -gpus N (frame blocks round-robin over the GPUs, one context + worker thread each,
    NCCL sum of the statistics tables for the .xvg rows): every file byte-identical to the one-GPU run.
    Needs two GPUs in the box (gpurun --gpus 2); the one-GPU pipelined path is compared as well.  The same check
    runs on the CPU box against the stand-in C ABI with 8 "GPUs" (tests/test_host_mock_cpu.py)."""
    import driver_cases
    done = driver_cases.check_frame_sharding(os.path.join(ROOT, "gmx-acqueduct_b200", "wn_waternetwork"), tmp_path, _device_count())
    if not done:
        pytest.skip("one GPU in this box: sharded run not exercised (run under gpurun --gpus 2)")


@pytest.mark.gpu
def test_cpp_bench_binary_small():
    """wn_bench (the C++ end-to-end bench of FrameBatcher + CudaFindPairs) runs and delivers frames in order."""
    import json
    exe = os.path.join(ROOT, "gmx-acqueduct_b200", "wn_bench")
    assert os.path.exists(exe), "build first: __graft_entry__.build()"
    for layout, mode in (("16", "push"), ("12", "stage"), ("20", "push")):
        r = subprocess.run([exe, "-waters", "2000", "-batch", "8", "-batches", "3", "-warmup", "2", "-ring", "16",
                            "-layout", layout, "-mode", mode], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        d = json.loads(r.stdout.strip().splitlines()[-1])
        assert d["in_frame_order"] and d["timed_frames"] == 24 and d["e2e_cpp"] > 0 and d["find_pairs"] > 0
        assert d["stats_rows_reduced"] == 40 and d["layout_bytes_per_edge"] == int(layout)
