"""The C++ host mirror (CudaFindPairs, HydrogenBondEdges, FrameBatcher, the standalone driver) on a CPU box.

The product has no CPU path: libwn_b200.so fails loudly without a device.  These tests link the SAME host sources
against tests/mock/wn_mock_abi.cpp -- a stand-in for the C ABI with the test oracle behind it, test infrastructure
only -- so that the host logic around the C ABI is exercised here: batching, pipelining, frame order, frame blocks
round-robin over several "GPUs" (the N>1 path of SURVEY 8e in the C++ product), the summed statistics table, the
files the driver writes.  The GPU suite runs the same checks (tests/driver_cases.py, tests/cpp/test_host.cpp) against
the shipped binaries on a B200."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "mock")
BUILD = os.path.join(MOCK, "build")


@pytest.fixture(scope="module")
def mock_build():
    subprocess.check_call(["make", "-s", "-j8", "-C", MOCK, "all"])
    return BUILD


def _run(exe, timeout=300, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([exe], capture_output=True, text=True, timeout=timeout, env=e)


def test_cpp_host_mirror_on_the_stand_in(mock_build):
    """tests/cpp/test_host.cpp, unchanged: FindPairs drop-in, batched == unbatched, makeEdges on a pair list, periodic
    and triclinic boxes, GPU-text == host writer, components, site lists, compact layouts, several GPUs."""
    r = _run(os.path.join(mock_build, "test_host_mock"))
    assert r.returncode == 0 and "test_host OK" in r.stdout and "multi-GPU batcher OK on 4 GPUs" in r.stdout, r.stdout + r.stderr
    # one "GPU" in the box: the several-GPU branch is skipped, everything else still holds
    r = _run(os.path.join(mock_build, "test_host_mock"), env={"WN_MOCK_DEVICES": "1"})
    assert r.returncode == 0 and "test_host OK" in r.stdout and "multi-GPU" not in r.stdout, r.stdout + r.stderr


def test_frame_batcher_state_machine_several_gpus(mock_build):
    """FrameBatcher over 1-8 "GPUs" of uneven speed: 25 run shapes (batch size, layout, push/stage, pipelined or not,
    helper copy threads, a flush in the middle), frame order, delivery on the caller's thread, block-to-GPU assignment,
    the summed statistics table, overlap of the GPUs in time, boxes and site lists, batch hook, failing batch calls."""
    r = _run(os.path.join(mock_build, "test_batcher_mock"))
    assert r.returncode == 0 and "test_batcher_mock OK" in r.stdout, r.stdout + r.stderr


def _sanitizer_links(flag):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not cxx:
        return False
    p = subprocess.run([cxx, flag, "-x", "c++", "-", "-o", "/dev/null"], input="int main(){return 0;}", capture_output=True, text=True)
    return p.returncode == 0


def test_frame_batcher_under_thread_sanitizer(mock_build):
    """The worker threads, the helper copy threads and the caller share staging buffers, result sets and the lane
    state: the stress test under ThreadSanitizer reports no data race."""
    if not _sanitizer_links("-fsanitize=thread"):
        pytest.skip("no ThreadSanitizer runtime with this compiler")
    subprocess.check_call(["make", "-s", "-j8", "-C", MOCK, "tsan"])
    r = _run(os.path.join(mock_build, "test_batcher_tsan"), env={"TSAN_OPTIONS": "halt_on_error=1 exitcode=66"})
    assert r.returncode == 0 and "test_batcher_mock OK" in r.stdout and "ThreadSanitizer" not in r.stderr, r.stdout + r.stderr[-3000:]


def test_host_mirror_under_address_sanitizer(mock_build):
    """Result pointers are only valid until the next call (or the one after it with two result sets): the stand-in
    re-assigns its buffers exactly then, so a host that reads an expired result trips AddressSanitizer."""
    if not _sanitizer_links("-fsanitize=address"):
        pytest.skip("no AddressSanitizer runtime with this compiler")
    subprocess.check_call(["make", "-s", "-j8", "-C", MOCK, "asan"])
    for exe, ok in (("test_batcher_asan", "test_batcher_mock OK"), ("test_host_asan", "test_host OK")):
        r = _run(os.path.join(mock_build, exe), env={"ASAN_OPTIONS": "detect_leaks=1"})
        assert r.returncode == 0 and ok in r.stdout, exe + ": " + r.stdout + r.stderr[-3000:]


@pytest.mark.parametrize("pbc,gpu_text", [(False, 1), (True, 1), (False, 0)])
def test_standalone_driver_writes_the_reference_files(mock_build, tmp_path, pbc, gpu_text):
    import driver_cases
    driver_cases.check_reference_files(os.path.join(mock_build, "wn_waternetwork_mock"), tmp_path, pbc, gpu_text)


def test_standalone_driver_buried_lists(mock_build, tmp_path):
    import driver_cases
    driver_cases.check_buried_lists(os.path.join(mock_build, "wn_waternetwork_mock"), tmp_path)


def test_standalone_driver_frame_sharding_over_8_gpus(mock_build, tmp_path):
    """-gpus 2, 4 and 8: edge, node and statistics files byte-identical to the one-GPU run (open and periodic)."""
    import driver_cases
    assert driver_cases.check_frame_sharding(os.path.join(mock_build, "wn_waternetwork_mock"), tmp_path, 8) == [2, 4, 8]


def test_standalone_driver_buried_lists_sharded(mock_build, tmp_path):
    """Per-frame site lists through the sharded path: the ids of the listed waters, frames in order."""
    import driver_cases
    driver_cases.check_buried_lists(os.path.join(mock_build, "wn_waternetwork_mock"), tmp_path, extra=["-gpus", "3", "-batch", "1"])


def test_the_stand_in_is_not_part_of_the_product():
    """Nothing that ships includes, links or names the stand-in; it is built only by tests/mock/Makefile."""
    hits = []
    for base in ("gmx-acqueduct_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".so", ".o", ".a")) or "checked_obj" in dp:
                    continue
                p = os.path.join(dp, fn)
                try:
                    txt = open(p, errors="ignore").read()
                except OSError:
                    continue
                if "wn_mock" in txt or "tests/mock" in txt:
                    hits.append(p)
    for fn in ("bench.py", "__graft_entry__.py", "wn_pkg.py"):
        if "wn_mock" in open(os.path.join(ROOT, fn)).read():
            hits.append(fn)
    assert not hits, hits


def test_cpp_bench_binary_on_the_stand_in(mock_build):
    """wn_bench (host/wn-bench.cpp) linked against the stand-in: every option combination the GPU test and bench.py use
    runs to completion, delivers the frames in order and prints one JSON line (the numbers mean nothing here)."""
    import json
    exe = os.path.join(mock_build, "wn_bench_mock")
    for extra in (["-layout", "16", "-mode", "push"], ["-layout", "12", "-mode", "stage"], ["-layout", "20", "-mode", "push", "-copy-threads", "1"],
                  ["-layout", "16", "-mode", "push", "-gpus", "3"], ["-layout", "12", "-mode", "stage", "-gpus", "4", "-dynamic", "1"],
                  ["-layout", "16", "-mode", "push", "-gpus", "2", "-pipelined", "0"]):
        r = subprocess.run([exe, "-waters", "300", "-batch", "4", "-batches", "3", "-warmup", "2", "-ring", "8"] + extra,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, " ".join(extra) + ": " + r.stdout + r.stderr
        d = json.loads(r.stdout.strip().splitlines()[-1])
        g = d["gpus"]
        assert d["in_frame_order"] and d["timed_frames"] == 12 * g and d["e2e_cpp"] > 0 and d["find_pairs"] > 0
        assert d["stats_rows_reduced"] == 20 * g and sum(d["frames_per_gpu"]) == 12 * g
        assert d["layout_bytes_per_edge"] == int(extra[1])
        if "-dynamic" not in extra:
            assert d["frames_per_gpu"] == [12] * g            # round-robin: equal shares


def test_find_pairs_result_hand_over_large_lists(mock_build):
    """CudaFindPairs::find (by value), findInto (a vector the caller keeps) and findRaw on lists above the size where
    the hand-over goes parallel (huge-page hint + first touch from several threads, shared copy), growing, shrinking
    and empty -- identical to a brute loop; also under AddressSanitizer/UBSan (test_host_mirror_under_address_sanitizer
    builds the binary)."""
    r = _run(os.path.join(mock_build, "test_findpairs_mock"))
    assert r.returncode == 0 and "test_findpairs_mock OK" in r.stdout, r.stdout + r.stderr
    asan = os.path.join(mock_build, "test_findpairs_asan")
    if _sanitizer_links("-fsanitize=address"):
        subprocess.check_call(["make", "-s", "-C", MOCK, "build/test_findpairs_asan"])
        r = _run(asan)
        assert r.returncode == 0 and "test_findpairs_mock OK" in r.stdout, r.stdout + r.stderr[-3000:]


@pytest.mark.parametrize("bug_compat,extra", [(True, []), (False, []), (True, ["-gpus", "2"]), (True, ["-pbc", "-gpu-text", "0"])])
def test_standalone_driver_protein_and_water(mock_build, tmp_path, bug_compat, extra):
    """makeSites typing + protein sites end to end: nodes, edges and statistics files of a small
    protein in water against the oracle on the site table restated from the reference's text."""
    import driver_cases
    driver_cases.check_protein_and_water(os.path.join(mock_build, "wn_waternetwork_mock"), tmp_path, bug_compat, extra)


@pytest.mark.parametrize("pbc,batch", [(False, 2), (True, 2), (False, 1), (False, 64)])
def test_gromacs_module_runs_and_equals_the_driver(mock_build, tmp_path, pbc, batch):
    """the gated gmx::TrajectoryAnalysisModule source compiled, linked and RUN (GROMACS is absent: against
    a working miniature of the API calls it makes, tests/shim/gromacs_run, and the stand-in C ABI) on a protein-in-water
    trajectory: same nodes/edges files and data-set rows as the standalone driver, whose output is pinned to the oracle.
    Batch sizes below, equal to and above the frame count exercise the late data-set rows (ModuleData::finish)."""
    import driver_cases
    driver_cases.check_gromacs_module_equals_driver(os.path.join(mock_build, "wn_gmx_module_mock"),
                                                    os.path.join(mock_build, "wn_waternetwork_mock"), tmp_path, pbc, batch)


def test_driver_files_feed_the_reference_downstream_script(mock_build, tmp_path):
    """The files the drop-in writes are the input of the reference's python/network-analysis.py.  Its own reader
    (read_node_list, :62-76), its edge-row loop (main, :146-178: skip the header, '#' lines separate frames,
    whitespace split, weight = |energy|) and its two filters (filter_edge_by_weight :16-24, remove_node_degree :26-35)
    -- the function texts taken from the reference file itself, which needs matplotlib as a whole and cannot be
    imported here -- run on the nodes/edges files of a protein-in-water trajectory; the connected components networkx
    finds per frame equal the rows the driver's own -components writes (threshold 0.0 on the 4-decimal text == 5e-5 on
    the binary energy: no float lies between (float)5e-5 and 5e-5)."""
    import ast
    import driver_cases
    script = "/root/reference/python/network-analysis.py"
    if not os.path.exists(script):
        pytest.skip("reference tree absent")
    nx = pytest.importorskip("networkx")
    tree = ast.parse(open(script).read())
    wanted = {"filter_edge_by_weight", "remove_node_degree", "read_node_list"}
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    assert {f.name for f in fns} == wanted
    ns = {"nx": nx}
    exec(compile(ast.Module(body=fns, type_ignores=[]), script, "exec"), ns)

    exe = os.path.join(mock_build, "wn_waternetwork_mock")
    driver_cases.check_protein_and_water(exe, tmp_path, True, ["-components", "5e-5", "-ocomp", str(tmp_path / "comp.xvg")])
    with open(str(tmp_path / "nodes.dat")) as fh:
        nodes, residues = ns["read_node_list"](fh)
    assert len(nodes) > 125 and residues["OW 4"] in nodes and "NZ 1" in residues and nodes[residues["NZ 1"]] == "NZ 1 LYS"
    rows = []
    with open(str(tmp_path / "edges.dat")) as fh:
        next(fh)                                                   # :147
        G, started = nx.Graph(), False

        def close_frame(G):
            G = ns["filter_edge_by_weight"](G, threshold=0.0)
            G = ns["remove_node_degree"](G)
            comps = list(nx.connected_components(G))
            rows.append((len(comps), max((len(c) for c in comps), default=0), G.number_of_nodes(), G.number_of_edges()))

        for line in fh:
            if line[0] == "#":                                     # :151
                if started:
                    close_frame(G)
                G, started = nx.Graph(), True
                continue
            f = line.split()                                       # :176-178
            assert all(int(f[k]) in nodes for k in (0, 1))         # every printed id is a node of the nodes file
            G.add_edge(int(f[0]), int(f[1]), weight=abs(float(f[2])), length=float(f[3]), angle=float(f[4]))
        close_frame(G)
    got = [tuple(int(v) for v in l.split()[1:]) for l in open(str(tmp_path / "comp.xvg")).read().splitlines()]
    assert len(rows) == 3 and got == rows, (got, rows)


@pytest.mark.parametrize("extra", [[], ["-gpus", "2"], ["-gpu-text", "0"]])
def test_standalone_driver_triclinic_box_line(mock_build, tmp_path, extra):
    import driver_cases
    driver_cases.check_triclinic_box_line(os.path.join(mock_build, "wn_waternetwork_mock"), tmp_path, extra)


@pytest.mark.parametrize("extra", [[], ["-gpus", "3"], ["-gpu-text", "0", "-pbc"]])
def test_standalone_driver_frame_selection_and_resume(mock_build, tmp_path, extra):
    """SURVEY 5 (checkpoint / resume = the runner's -b/-e/-dt; outputs are per frame, so a resumed run appends)."""
    import driver_cases
    driver_cases.check_frame_selection_and_resume(os.path.join(mock_build, "wn_waternetwork_mock"), tmp_path, extra)


def test_standalone_driver_rejects_malformed_input_cleanly(mock_build, tmp_path):
    """Malformed trajectories and option values end the driver with exit code 1 and a message -- no crash, no
    out-of-bounds read (the AddressSanitizer/UBSan build of the driver is used where the runtime exists)."""
    exe = os.path.join(mock_build, "wn_waternetwork_mock")
    if _sanitizer_links("-fsanitize=address"):
        subprocess.check_call(["make", "-s", "-C", MOCK, "build/wn_waternetwork_asan"])
        exe = os.path.join(mock_build, "wn_waternetwork_asan")
    atom = "    1SOL     OW    1   0.100   0.100   0.100\n    1SOL    HW1    2   0.190   0.100   0.100\n    1SOL    HW2    3   0.070   0.190   0.100\n"
    good = "w t= 0.0\n    3\n" + atom + "   1.00000   1.00000   1.00000\n"
    cases = {
        "empty": "",
        "only_title": "title\n",
        "negative_count": "w\n   -3\n",
        "short_atom_line": "w\n    3\n    1SOL     OW    1   0.1\n",
        "ends_inside_frame": "w\n    3\n" + atom[:len(atom) // 2],
        "missing_box": "w\n    3\n" + atom,
        "count_changes": good + "w t= 1.0\n    6\n" + atom + atom.replace("    1SOL", "    2SOL") + "   1.00000   1.00000   1.00000\n",
        "not_whole_waters": "w\n    2\n" + "".join(atom.splitlines(True)[:2]) + "   1.00000   1.00000   1.00000\n",
        "huge_count": "w\n99999999\n" + atom + "   1.00000   1.00000   1.00000\n",
        "binary_garbage": "\x00\x01\x02\xff\n\x7f\x00\n" * 20,
    }
    for name, text in cases.items():
        p = tmp_path / (name + ".gro")
        p.write_bytes(text.encode("latin-1"))
        r = subprocess.run([exe, "-f", str(p), "-edges", str(tmp_path / "e"), "-nodes", str(tmp_path / "n"), "-ohb", "", "-obw", ""],
                           capture_output=True, text=True, timeout=120)
        assert r.returncode == 1 and "wn_waternetwork:" in r.stderr, (name, r.returncode, r.stderr[-400:])
        assert "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, (name, r.stderr[-2000:])
    ok = tmp_path / "ok.gro"
    ok.write_text(good)
    for bad_args in (["-nosuchoption", "1"], ["-batch"], ["-buried", str(tmp_path / "missing.txt")], ["-gpus", "2", "-components", "0.1"],
                     ["-gpus", "99"]):
        r = subprocess.run([exe, "-f", str(ok), "-edges", str(tmp_path / "e"), "-nodes", "", "-ohb", "", "-obw", ""] + bad_args,
                           capture_output=True, text=True, timeout=120)
        assert r.returncode == 1 and "wn_waternetwork:" in r.stderr, (bad_args, r.returncode, r.stderr[-400:])
        assert "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, (bad_args, r.stderr[-2000:])
    # a well-formed one-water trajectory runs (no pairs, no edges)
    r = subprocess.run([exe, "-f", str(ok), "-edges", str(tmp_path / "e"), "-nodes", "", "-ohb", "", "-obw", ""], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "1 frames, 0 hydrogen bonds" in r.stderr, r.stderr


def test_driver_module_and_bench_under_address_sanitizer(mock_build, tmp_path):
    """The standalone driver (sharded, site lists, frame selection + append), the GROMACS module on the API miniature
    and wn_bench (several GPUs, dynamic order, unpipelined) under AddressSanitizer + UBSan: same outputs, no report."""
    if not _sanitizer_links("-fsanitize=address"):
        pytest.skip("no AddressSanitizer runtime with this compiler")
    import driver_cases
    subprocess.check_call(["make", "-s", "-j8", "-C", MOCK, "asan"])
    drv, mod, bench = (os.path.join(mock_build, n) for n in ("wn_waternetwork_asan", "wn_gmx_module_asan", "wn_bench_asan"))
    for k, fn in enumerate((lambda d: driver_cases.check_gromacs_module_equals_driver(mod, drv, d, False, 2),
                            lambda d: driver_cases.check_frame_sharding(drv, d, 4),
                            lambda d: driver_cases.check_buried_lists(drv, d, ["-gpus", "3", "-batch", "1"]),
                            lambda d: driver_cases.check_frame_selection_and_resume(drv, d, []))):
        d = tmp_path / ("case%d" % k)
        d.mkdir()
        fn(d)
    for extra in (["-layout", "16", "-mode", "push"], ["-layout", "12", "-mode", "stage", "-gpus", "4", "-dynamic", "1"],
                  ["-layout", "20", "-mode", "push", "-gpus", "2", "-pipelined", "0"]):
        r = subprocess.run([bench, "-waters", "300", "-batch", "4", "-batches", "3", "-warmup", "2", "-ring", "8"] + extra,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "Sanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]


def test_make_sites_on_random_topologies_through_the_driver(mock_build, tmp_path):
    import driver_cases
    driver_cases.check_random_topologies(os.path.join(mock_build, "wn_waternetwork_mock"), tmp_path)
