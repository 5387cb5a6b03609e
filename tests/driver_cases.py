"""Checks of the standalone driver (wn_waternetwork) shared by the GPU tests (tests/test_gpu_host_cpp.py: the shipped
binary on a B200) and the CPU tests (tests/test_host_mock_cpu.py: the same driver source linked against the stand-in
C ABI of tests/mock).  Expectations come from the oracle on the coordinates the driver parses."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def write_gro(path, frames, box, with_time=True):
    """Multi-frame .gro of a pure water box (OW, HW1, HW2 per SOL residue), 8.3 coordinate fields."""
    nf, natoms, _ = frames.shape
    with open(path, "w") as fh:
        for f in range(nf):
            fh.write("synthetic water box%s\n" % (" t= %9.5f" % (2.0 * f) if with_time else ""))
            fh.write("%5d\n" % natoms)
            for a in range(natoms):
                w = a // 3
                fh.write("%5d%-5s%5s%5d%8.3f%8.3f%8.3f\n" % ((w + 1) % 100000, "SOL", ("OW", "HW1", "HW2")[a % 3],
                                                           (a + 1) % 100000, *frames[f, a]))
            fh.write("%10.5f%10.5f%10.5f\n" % (box, box, box))


def parse_back(frames):
    """What the driver reads: every coordinate through its 8.3 text field and (float)strtod."""
    import numpy as np
    txt = np.char.mod("%8.3f", frames.astype(np.float64))
    return txt.astype(np.float64).astype(np.float32)


EDGE_HEADER = "%6s%8s%6s%6s%8s%6s%8s%8s%8s" % ("#Id", "AtmId", "ResId", "Id", "AtmId", "ResId", "Energy", "Length", "Cos\n")


def check_reference_files(exe, tmp_path, pbc, gpu_text):
    """wn_waternetwork on a synthetic .gro trajectory: edges file, nodes file and statistics rows equal
    what the oracle (open boundary) / the periodic oracle gives for the same parsed coordinates."""
    import numpy as np
    import wn_oracle_py as o
    n, nf = 216, 3
    sb = o.synth_box(n, model=o.MODEL_SPC)
    frames = o.synth_frames(sb, 0, nf)
    box = float(sb.box)
    gro = str(tmp_path / "traj.gro")
    write_gro(gro, frames, box)
    args = [exe, "-f", gro, "-edges", str(tmp_path / "edges"), "-nodes", str(tmp_path / "nodes"),
            "-ohb", str(tmp_path / "ohb.xvg"), "-obw", str(tmp_path / "obw.xvg"), "-batch", "2", "-gpu-text", str(gpu_text),
            "-components", "0.5", "-ocomp", str(tmp_path / "comp.xvg")]
    if pbc:
        args.append("-pbc")
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Solvent has : 216 Sites of out 648 atoms." in r.stderr

    parsed = parse_back(frames)
    sa, ha, ty = o.water_site_table(n)
    boxf = np.float32(np.float64("%10.5f" % box))
    want = EDGE_HEADER
    ohb = ""
    comp = ""
    for f in range(nf):
        if pbc:
            e, st, _, _ = o.frame_hbonds_pbc(parsed[f], sa, ha, ty, np.array([boxf] * 3, np.float32))
        else:
            e, st, _, _ = o.frame_hbonds(parsed[f], sa, ha, ty)
        want += "#--- frame %d ---\n" % f
        # ids: solvent site w has id (w + 1) + nProtein + 1 = w + 2  (src/waternetwork.cpp:420-423)
        want += "".join("%6d%6d%8.4f%8.4f%8.4f\n" % (a + 2, b + 2, en, ln, cs)
                        for a, b, en, ln, cs in zip(e["i"], e["j"], e["energy"].astype(np.float64),
                                                    e["length"].astype(np.float64), e["cosine"].astype(np.float64)))
        ohb += " %11.3f %8.3f %8.3f %8.3f %8.3f %8.3f %8.3f\n" % (2.0 * f, n, len(e), len(e) / n, 0, 0, 0)
        _, cst = o.components(e, n, 0.5)
        comp += " %11.3f %8d %8d %8d %8d\n" % (2.0 * f, cst[0], cst[1], cst[2], cst[3])
    got = open(str(tmp_path / "edges.dat")).read()
    assert got == want
    assert open(str(tmp_path / "comp.xvg")).read() == comp
    assert open(str(tmp_path / "ohb.xvg")).read() == ohb
    assert open(str(tmp_path / "obw.xvg")).read() == "".join(" %11.3f %8.3f %8.3f\n" % (2.0 * f, n, 0) for f in range(nf))
    nodes = open(str(tmp_path / "nodes.dat")).read().splitlines()
    assert len(nodes) == 1 + n and nodes[0].split() == ["#Id", "AtmId", "Name", "ResId", "Res", "Type", "Hyd"]
    assert nodes[1].split()[:3] == ["2", "0", "OW"]


def check_buried_lists(exe, tmp_path, extra=()):
    """-buried: a different water list per frame; ids in the edge file are those of the listed waters."""
    import numpy as np
    import wn_oracle_py as o
    n, nf = 216, 3
    sb = o.synth_box(n, model=o.MODEL_SPC)
    frames = o.synth_frames(sb, 0, nf)
    gro = str(tmp_path / "traj.gro")
    write_gro(gro, frames, float(sb.box), with_time=False)
    rng = np.random.default_rng(2)
    lists = [np.sort(rng.choice(n, size=k, replace=False)) for k in (150, 0, 216)]
    (tmp_path / "buried.txt").write_text("".join(" ".join(str(int(w)) for w in l) + "\n" for l in lists))
    r = subprocess.run([exe, "-f", gro, "-edges", str(tmp_path / "edges"), "-nodes", "", "-ohb", str(tmp_path / "ohb.xvg"),
                        "-obw", str(tmp_path / "obw.xvg"), "-buried", str(tmp_path / "buried.txt"), "-batch", "4"] + list(extra),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    parsed = parse_back(frames)
    want = EDGE_HEADER
    obw = ""
    for f in range(nf):
        l = lists[f]
        sa = (3 * l).astype(np.int32)
        ha = np.stack([3 * l, 3 * l + 1], axis=1).astype(np.int32)
        ty = np.full(len(l), o.WATER, np.uint8)
        e, st, _, _ = o.frame_hbonds(parsed[f], sa, ha, ty)
        want += "#--- frame %d ---\n" % f
        want += "".join("%6d%6d%8.4f%8.4f%8.4f\n" % (l[a] + 2, l[b] + 2, en, ln, cs)
                        for a, b, en, ln, cs in zip(e["i"], e["j"], e["energy"].astype(np.float64),
                                                    e["length"].astype(np.float64), e["cosine"].astype(np.float64)))
        obw += " %11.3f %8.3f %8.3f\n" % (f, len(l), 0)
    assert open(str(tmp_path / "edges.dat")).read() == want
    assert open(str(tmp_path / "obw.xvg")).read() == obw


def check_frame_sharding(exe, tmp_path, ndev):
    """wn_waternetwork -gpus N (frame blocks round-robin over the GPUs, one context + worker thread each, the sum
    of the statistics tables for the .xvg rows): every file byte-identical to the one-GPU run; the one-GPU
    pipelined path is compared as well.  Returns the GPU counts that were exercised."""
    import wn_oracle_py as o
    n, nf = 216, 11
    sb = o.synth_box(n, model=o.MODEL_SPC)
    frames = o.synth_frames(sb, 0, nf)
    gro = str(tmp_path / "traj.gro")
    write_gro(gro, frames, float(sb.box))

    def run(tag, extra):
        d = tmp_path / tag
        d.mkdir()
        r = subprocess.run([exe, "-f", gro, "-edges", str(d / "edges"), "-nodes", str(d / "nodes"), "-ohb", str(d / "ohb.xvg"),
                            "-obw", str(d / "obw.xvg"), "-batch", "2"] + extra, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        return {k: open(str(d / k)).read() for k in ("edges.dat", "nodes.dat", "ohb.xvg", "obw.xvg")}

    base = run("one", ["-gpu-text", "0"])
    assert base["edges.dat"].count("#--- frame") == nf
    assert run("one_gputext", ["-gpu-text", "1"]) == base
    done = []
    if ndev < 2:
        return done
    for g in sorted({2, min(ndev, 4), min(ndev, 8)}):
        assert run("g%d" % g, ["-gpus", str(g)]) == base
        assert run("g%d_pbc" % g, ["-gpus", str(g), "-pbc"]) == run("one_pbc_%d" % g, ["-pbc"])
        done.append(g)
    return done


# ---- protein + water: the site rules of makeSites and the site order of analyzeFrame, restated from the reference text
_SITE_NAMES = {"OW", "O", "N", "SG", "NE2", "OE1", "OD1", "ND2", "OG", "OG1", "OH", "NE1", "NZ", "NH1", "NH2", "ND1"}   # :99-118
_DONOR, _ACCEPTOR, _WATER = 0, 1, 2
_SITE_TYPE = {"NE": _DONOR, "NH1": _DONOR, "NH2": _DONOR, "ND1": _DONOR, "ND2": _DONOR, "OD1": _DONOR, "OD2": _DONOR,
              "NE1": _DONOR, "NE2": _DONOR, "OE1": _ACCEPTOR, "OE2": _ACCEPTOR, "NZ": _DONOR, "OG": _DONOR, "OH": _DONOR,
              "O": _ACCEPTOR, "N": _DONOR, "SG": _DONOR, "OW": _WATER, "OCD": _ACCEPTOR}                               # :120-139


def _make_sites(selection, names, resnames_of_atom, resindex_of_atom):
    """makeSites (src/waternetwork.cpp:141-179) on plain lists: [(id, atmIndex, name, resIndex, resName, type, nbH)]."""
    out, count = [], 0
    for index in selection:
        name = names[index]
        if name not in _SITE_NAMES:
            continue
        count += 1
        nhb = 0
        while True:
            if index + nhb + 1 > len(selection):                # :157, topology index against selection size, as written
                break
            if index + nhb + 1 >= len(names):                   # (the reference would read past the topology)
                break
            probe = names[index + nhb + 1].lstrip("0123456789")[:1]      # element of a .gro atom: first letter of its name
            if "H" not in probe:
                break
            nhb += 1
        typ = _SITE_TYPE.get(name, _DONOR)                      # operator[] default-constructs DONOR (:168)
        if resnames_of_atom[index] == "WAT":                    # :170-173
            typ = _WATER
        out.append((count, index, name, resindex_of_atom[index], resnames_of_atom[index], typ, nhb))
    return out


def check_protein_and_water(exe, tmp_path, bug_compat, extra=()):
    """A small 'protein' (SER, LYS, THR, GLN with explicit hydrogens) in front of a water box: nodes file, edges file
    and statistics rows against the oracle run on the site table that the reference's makeSites (:95-180) and
    analyzeFrame (:547-576) would build -- protein sites first (hydrogens at proteinPoints[id + i + 1] with
    -bug-compat 1, :554; at atmIndex + i + 1 otherwise), then every water with hydrogen list {3w, 3w+1, ...}."""
    import numpy as np
    import wn_oracle_py as o
    n_w, nf = 125, 3
    sb = o.synth_box(n_w, model=o.MODEL_SPC)
    wframes = o.synth_frames(sb, 0, nf)
    residues = [("SER", ["N", "H", "CA", "HA", "CB", "HB1", "HB2", "OG", "HG", "C", "O"]),
                ("LYS", ["N", "H", "CA", "HA", "CB", "HB1", "CG", "CD", "CE", "NZ", "HZ1", "HZ2", "HZ3", "C", "O"]),
                ("THR", ["N", "H", "CA", "HA", "CB", "HB", "OG1", "HG1", "CG2", "1HG2", "C", "O"]),
                ("GLN", ["N", "H", "CA", "CB", "CG", "CD", "OE1", "NE2", "1HE2", "2HE2", "C", "O"])]
    names, resn, resi = [], [], []
    for r, (rn, atoms) in enumerate(residues):
        for a in atoms:
            names.append(a); resn.append(rn); resi.append(r)
    n_p = len(names)
    for w in range(n_w):
        for a in ("OW", "HW1", "HW2"):
            names.append(a); resn.append("SOL"); resi.append(len(residues) + w)
    natoms = len(names)
    rng = np.random.default_rng(17)
    frames = np.zeros((nf, natoms, 3), np.float32)
    box = float(sb.box)
    for f in range(nf):
        # heavy atoms on a jittered 0.3 nm grid in a slab just outside the +x face of the water box (a dense water box
        # has no cavity for them): 0.26 nm or more from every water oxygen and from each other, so the energies stay
        # inside the 8-character text field (its overflow glues two columns together in the reference's format);
        # hydrogens 0.1 nm from the atom before them
        pos = np.zeros((n_p, 3))
        slots = [(box + 0.28 + 0.3 * ix, 0.25 + 0.3 * iy, 0.25 + 0.3 * iz) for ix in range(3) for iy in range(4) for iz in range(4)]
        order = rng.permutation(len(slots))
        k = 0
        for a in range(n_p):
            if names[a].lstrip("0123456789").startswith("H"):
                d = rng.normal(size=3); d /= np.linalg.norm(d)
                pos[a] = pos[a - 1] + 0.1 * d
            else:
                pos[a] = np.array(slots[order[k]]) + rng.uniform(-0.02, 0.02, size=3)
                k += 1
        frames[f, :n_p] = pos
        frames[f, n_p:] = wframes[f]
    gro = str(tmp_path / "traj.gro")
    with open(gro, "w") as fh:
        for f in range(nf):
            fh.write("protein in water t= %9.5f\n" % (0.5 * f))
            fh.write("%5d\n" % natoms)
            for a in range(natoms):
                fh.write("%5d%-5s%5s%5d%8.3f%8.3f%8.3f\n" % (resi[a] + 1, resn[a], names[a], a + 1, *frames[f, a]))
            fh.write("%10.5f%10.5f%10.5f\n" % (box, box, box))
    r = subprocess.run([exe, "-f", gro, "-edges", str(tmp_path / "edges"), "-nodes", str(tmp_path / "nodes"),
                        "-ohb", str(tmp_path / "ohb.xvg"), "-obw", str(tmp_path / "obw.xvg"), "-batch", "2",
                        "-bug-compat", str(int(bug_compat))] + list(extra), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr

    protein_sel = list(range(n_p))
    solvent_sel = list(range(n_p, natoms))
    psites = _make_sites(protein_sel, names, resn, resi)
    ssites = _make_sites(solvent_sel, names, resn, resi)
    assert len(psites) >= 9 and len(ssites) == n_w
    assert "Protein has : %d Sites of out %d atoms." % (len(psites), n_p) in r.stderr
    n_ps = len(psites)
    ssites = [(sid + n_ps + 1,) + tuple(rest) for (sid, *rest) in ssites]              # :420-423
    # nodes file (:461-482, operator<< of SiteInfo: seven setw(7) fields)
    want_nodes = "%7s%7s%7s%7s%7s%7s%7s" % ("#Id", "AtmId", "Name", "ResId", "Res", "Type", "Hyd\n")
    for (sid, idx, nm, ri, rn, ty, nh) in psites + ssites:
        want_nodes += "%7d%7d%7s%7d%7s%7d%7d\n" % (sid, idx, nm, ri, rn, ty, nh)
    assert open(str(tmp_path / "nodes.dat")).read() == want_nodes
    # the site table analyzeFrame builds (:547-576): atom indices into the frame
    hmax = max(1, max(s[6] for s in psites + ssites))
    site_atom, h_atom, types, ids = [], [], [], []
    for (sid, idx, nm, ri, rn, ty, nh) in psites:
        base = sid if bug_compat else idx
        hs = [protein_sel[base + i + 1] for i in range(nh)]                      # proteinPoints.at(id + i + 1) (:554)
        site_atom.append(protein_sel[idx]); h_atom.append(hs + [-1] * (hmax - nh)); types.append(ty); ids.append(sid)
    for w, (sid, idx, nm, ri, rn, ty, nh) in enumerate(ssites):
        hs = [solvent_sel[3 * w + i] for i in range(nh)]                         # solventPoints.at(3*idx + i) (:569)
        site_atom.append(solvent_sel[3 * w]); h_atom.append(hs + [-1] * (hmax - nh)); types.append(ty); ids.append(sid)
    # the quirk of :157 shows: the last waters of a solvent selection that does not start at atom 0 lose hydrogens
    assert ssites[0][6] == 2 and ssites[-1][6] < 2
    sa = np.array(site_atom, np.int32); ha = np.array(h_atom, np.int32).reshape(len(sa), hmax); ty = np.array(types, np.uint8)
    parsed = parse_back(frames)
    want = EDGE_HEADER
    ohb = ""
    n_mixed = 0
    boxf = np.float32(np.float64("%10.5f" % box))
    for f in range(nf):
        if "-pbc" in extra:
            e, st, _, _ = o.frame_hbonds_pbc(parsed[f], sa, ha, ty, np.array([boxf] * 3, np.float32))
        else:
            e, st, _, _ = o.frame_hbonds(parsed[f], sa, ha, ty)
        n_mixed += int(np.sum((e["i"] < n_ps) != (e["j"] < n_ps)))
        want += "#--- frame %d ---\n" % f
        want += "".join("%6d%6d%8.4f%8.4f%8.4f\n" % (ids[a], ids[b], en, ln, cs)
                        for a, b, en, ln, cs in zip(e["i"], e["j"], e["energy"].astype(np.float64),
                                                    e["length"].astype(np.float64), e["cosine"].astype(np.float64)))
        ohb += " %11.3f %8.3f %8.3f %8.3f %8.3f %8.3f %8.3f\n" % (0.5 * f, len(sa), len(e), len(e) / len(sa), 0, 0, 0)
    assert n_mixed > 0                                                           # protein-water bonds really occur
    assert open(str(tmp_path / "edges.dat")).read() == want
    assert open(str(tmp_path / "ohb.xvg")).read() == ohb
    assert open(str(tmp_path / "obw.xvg")).read() == "".join(" %11.3f %8.3f %8.3f\n" % (0.5 * f, n_w, 0) for f in range(nf))


def check_gromacs_module_equals_driver(module_exe, driver_exe, tmp_path, pbc=False, batch=2):
    """The compile-gated GROMACS module (host/gromacs/waternetwork-b200.cpp, run on the working API miniature of
    tests/shim/gromacs_run) writes the same nodes, edges and statistics as the standalone driver -- which
    check_protein_and_water pins to the oracle -- for the same protein-in-water trajectory."""
    extra = ["-pbc", "-gpu-text", "0"] if pbc else []
    check_protein_and_water(driver_exe, tmp_path, True, extra)
    m = tmp_path / "module"
    m.mkdir()
    args = [module_exe, "-f", str(tmp_path / "traj.gro"), "-edges", str(m / "edges"), "-nodes", str(m / "nodes"),
            "-ohb", str(m / "ohb"), "-obw", str(m / "obw"), "-batch", str(batch)]
    if pbc:
        args.append("-usepbc")
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    # GROMACS completes plot-type file names with .xvg; the module appends .dat for nodes and edges (:461,:485)
    assert open(str(m / "edges.xvg.dat")).read() == open(str(tmp_path / "edges.dat")).read()
    assert open(str(m / "nodes.xvg.dat")).read() == open(str(tmp_path / "nodes.dat")).read()

    def rows(path):
        return [l for l in open(path).read().splitlines() if l and l[0] not in "#@"]

    assert rows(str(m / "ohb.xvg")) == rows(str(tmp_path / "ohb.xvg"))
    assert rows(str(m / "obw.xvg")) == rows(str(tmp_path / "obw.xvg"))
    assert len(rows(str(m / "ohb.xvg"))) == 3


def check_triclinic_box_line(exe, tmp_path, extra=()):
    """-pbc with a triclinic .gro box line (v1(x) v2(y) v3(z) v1(y) v1(z) v2(x) v2(z) v3(x) v3(y), rows of the GROMACS
    matrix = box vectors): the edge file equals the triclinic oracle on the matrix the nine numbers stand for."""
    import numpy as np
    import wn_oracle_py as o
    n, nf = 125, 2
    sb = o.synth_box(n, model=o.MODEL_SPC)
    frames = o.synth_frames(sb, 0, nf)
    L = float(sb.box)
    ax, by, cz, bx, cx, cy = L, L, L, 0.4, -0.3, 0.5
    gro = str(tmp_path / "traj.gro")
    with open(gro, "w") as fh:
        for f in range(nf):
            fh.write("skewed cell t= %9.5f\n" % float(f))
            fh.write("%5d\n" % (3 * n))
            for a in range(3 * n):
                fh.write("%5d%-5s%5s%5d%8.3f%8.3f%8.3f\n" % (a // 3 + 1, "SOL", ("OW", "HW1", "HW2")[a % 3], a + 1, *frames[f, a]))
            fh.write("%10.5f%10.5f%10.5f%10.5f%10.5f%10.5f%10.5f%10.5f%10.5f\n" % (ax, by, cz, 0.0, 0.0, bx, 0.0, cx, cy))
    r = subprocess.run([exe, "-f", gro, "-edges", str(tmp_path / "edges"), "-nodes", "", "-ohb", str(tmp_path / "ohb.xvg"), "-obw", "",
                        "-batch", "2", "-pbc"] + list(extra), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    f32 = lambda v: np.float32(np.float64("%10.5f" % v))
    box9 = np.array([f32(ax), 0, 0, f32(bx), f32(by), 0, f32(cx), f32(cy), f32(cz)], np.float32)
    parsed = parse_back(frames)
    sa, ha, ty = o.water_site_table(n)
    want = EDGE_HEADER
    n_rect = n_tric = 0
    for f in range(nf):
        e, st, _, _ = o.frame_hbonds_tric(parsed[f], sa, ha, ty, box9)
        n_tric += len(e)
        n_rect += len(o.frame_hbonds_pbc(parsed[f], sa, ha, ty, np.array([f32(L)] * 3, np.float32))[0])
        want += "#--- frame %d ---\n" % f
        want += "".join("%6d%6d%8.4f%8.4f%8.4f\n" % (a + 2, b + 2, en, ln, cs)
                        for a, b, en, ln, cs in zip(e["i"], e["j"], e["energy"].astype(np.float64),
                                                    e["length"].astype(np.float64), e["cosine"].astype(np.float64)))
    assert n_tric != n_rect                                   # the off-diagonal numbers matter
    assert open(str(tmp_path / "edges.dat")).read() == want


def check_frame_selection_and_resume(exe, tmp_path, extra=()):
    """-b / -e / -dt (the GROMACS runner's frame selection) and -frnr0 / -append (continue an earlier run): a run split in
    two writes the same bytes as one run over everything; -dt keeps every second frame, numbered among the analysed
    frames as analyzeFrame's frnr is."""
    import wn_oracle_py as o
    n, nf = 125, 9
    sb = o.synth_box(n, model=o.MODEL_SPC)
    frames = o.synth_frames(sb, 0, nf)
    gro = str(tmp_path / "traj.gro")
    write_gro(gro, frames, float(sb.box))                       # t = 0, 2, 4, ... 16

    def run(tag, more, fresh=True):
        d = tmp_path / tag
        if fresh:
            d.mkdir()
        r = subprocess.run([exe, "-f", gro, "-edges", str(d / "edges"), "-nodes", str(d / "nodes"), "-ohb", str(d / "ohb.xvg"),
                            "-obw", str(d / "obw.xvg"), "-batch", "2"] + more + list(extra), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        return {k: open(str(d / k)).read() for k in ("edges.dat", "nodes.dat", "ohb.xvg", "obw.xvg")}, r.stderr

    whole, _ = run("whole", [])
    assert whole["edges.dat"].count("#--- frame") == nf
    first, err = run("split", ["-e", "7.0"])                    # t = 0..6: four frames
    assert "4 frames," in err and first["edges.dat"].count("#--- frame") == 4
    both, err = run("split", ["-b", "7.5", "-frnr0", "4", "-append", "1"], fresh=False)
    assert "5 frames," in err
    assert both == whole                                        # every file, byte for byte
    # every second frame: the analysed frames are numbered 0, 1, 2, ... and carry the times 0, 4, 8, ...
    sel, err = run("dt", ["-dt", "4.0"])
    assert "5 frames," in err

    def split_frames(text):
        parts = text.split("#--- frame ")[1:]
        return [p.split(" ---\n", 1) for p in parts]

    w, s2 = split_frames(whole["edges.dat"]), split_frames(sel["edges.dat"])
    assert [int(h) for h, _ in s2] == [0, 1, 2, 3, 4]
    assert [rows for _, rows in s2] == [w[k][1] for k in (0, 2, 4, 6, 8)]
    assert [float(l.split()[0]) for l in sel["ohb.xvg"].splitlines()] == [0.0, 4.0, 8.0, 12.0, 16.0]
    # a window in the middle
    mid, err = run("mid", ["-b", "4.0", "-e", "8.0"])
    assert [rows for _, rows in split_frames(mid["edges.dat"])] == [w[k][1] for k in (2, 3, 4)]


def check_random_topologies(exe, tmp_path, n_cases=25, seed=7):
    """makeSites through the driver on random topologies: the nodes file (ids, atom indices, names, residues, types,
    hydrogen counts -- including names that are sites but have no entry in the type map (OG1 -> DONOR by
    default-construction), names in the map that are not sites (NE, OD2, OE2, OCD), a residue called WAT, hydrogens
    named with leading digits, the :157 bound on the solvent selection) equals the Python restatement of :95-180."""
    import numpy as np
    rng = np.random.default_rng(seed)
    pool = ["N", "H", "CA", "HA", "CB", "C", "O", "OG", "HG", "OG1", "HG1", "SG", "NE", "NE1", "NE2", "1HE2", "2HE2", "OE1", "OE2",
            "OD1", "OD2", "ND1", "ND2", "HD1", "NZ", "HZ1", "HZ2", "HZ3", "NH1", "NH2", "HH11", "OH", "HH", "OCD", "CG", "CD", "OW", "FE"]
    res_pool = ["ALA", "SER", "THR", "GLN", "LYS", "ARG", "HIS", "TYR", "CYS", "ASN", "TRP"]
    wat_names = ["SOL", "WAT", "HOH", "TIP3", "SPC"]
    for case in range(n_cases):
        n_res = int(rng.integers(0, 6))
        names, resn, resi = [], [], []
        for r in range(n_res):
            rn = res_pool[int(rng.integers(len(res_pool)))]
            for _ in range(int(rng.integers(1, 9))):
                names.append(pool[int(rng.integers(len(pool)))]); resn.append(rn); resi.append(r)
        n_p = len(names)
        n_w = int(rng.integers(1, 8))
        wn = wat_names[int(rng.integers(len(wat_names)))]
        for w in range(n_w):
            for a in ("OW", "HW1", "HW2"):
                names.append(a); resn.append(wn); resi.append(n_res + w)
        natoms = len(names)
        xyz = rng.uniform(0, 3, size=(natoms, 3))
        gro = tmp_path / ("topo%d.gro" % case)
        with open(str(gro), "w") as fh:
            fh.write("random topology\n%5d\n" % natoms)
            for a in range(natoms):
                fh.write("%5d%-5s%5s%5d%8.3f%8.3f%8.3f\n" % (resi[a] + 1, resn[a], names[a], a + 1, *xyz[a]))
            fh.write("%10.5f%10.5f%10.5f\n" % (3.0, 3.0, 3.0))
        nodes = tmp_path / ("nodes%d" % case)
        r = subprocess.run([exe, "-f", str(gro), "-edges", "", "-nodes", str(nodes), "-ohb", "", "-obw", "", "-bug-compat", "0"],
                           capture_output=True, text=True, timeout=120)
        psites = _make_sites(list(range(n_p)), names, resn, resi)
        ssites = _make_sites(list(range(n_p, natoms)), names, resn, resi)
        if len(ssites) * 3 != natoms - n_p:
            # a water whose OW is not followed by... cannot happen with OW/HW1/HW2 triplets; keep the guard honest
            assert r.returncode == 1, (case, r.stderr)
            continue
        assert r.returncode == 0, (case, r.stdout + r.stderr)
        ssites = [(sid + len(psites) + 1,) + tuple(rest) for (sid, *rest) in ssites]
        want = "%7s%7s%7s%7s%7s%7s%7s" % ("#Id", "AtmId", "Name", "ResId", "Res", "Type", "Hyd\n")
        for (sid, idx, nm, ri, rn, ty, nh) in psites + ssites:
            want += "%7d%7d%7s%7d%7s%7d%7d\n" % (sid, idx, nm, ri, rn, ty, nh)
        assert open(str(nodes) + ".dat").read() == want, (case, names[:n_p])
        assert "Protein has : %d Sites of out %d atoms." % (len(psites), n_p) in r.stderr
