"""CPU tests of the host-side logic: the C-ABI library loads and exports every declared symbol,
the sweep plan is valid on every mesh family, the builders agree with the oracle's geometry."""
import ctypes
import os
import re

import numpy as np
import pytest

import util
from oracle import pampa_oracle as orc
from pampa_b200 import _lib, problem as pb, synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pampa_sn.h")).read()
    declared = set(re.findall(r"\b(pampa_sn_[a-z_]+)\s*\(", hdr))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)


def test_no_cpu_fallback():
    """Without a CUDA device the compute path fails loudly (this container has no GPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    em, xs, quad, ls, z = util.load_golden("slabs_s2")
    with pytest.raises(pb.SNError, match="no CUDA device"):
        pb.SNDevice(em, xs, quad, ls)


@pytest.mark.parametrize("name", ["slabs_s2", "slabs_s4", "pwr_cartesian_s2", "pwr_unstructured_s2",
                                  "pwr_cartesian_s8_lsoff"])
def test_plan_on_reference_meshes(name):
    em, xs, quad, ls, z = util.load_golden(name)
    info = pb.plan_check(em, quad, int(z["G"]))
    assert info["num_cells"] == em.num_cells
    assert info["updates_per_sweep"] == em.num_cells * len(quad.weights) * int(z["G"])
    assert info["num_classes"] >= 2 and info["sweep_tasks"] > 0


@pytest.mark.parametrize("opts", [{}, {"patch_cells": 32}, {"patch_cells": 64, "z_chunk": 3}, {"dt_max": 3},
                                  {"tile_i": 8, "tile_j": 8}])
def test_plan_on_synthetic_meshes(opts):
    q8 = syn.level_symmetric(8)
    mesh, xs = syn.checkerboard_core(37, 29, 11, num_groups=3)
    info = pb.plan_check(mesh, q8, 3, **opts)
    assert info["num_classes"] == 8
    assert info["updates_per_sweep"] == 37 * 29 * 11 * 80 * 3
    hexm, hxs, _ = syn.hex_core(7, 6, num_groups=2)
    info = pb.plan_check(hexm, q8, 2, **opts)
    assert info["updates_per_sweep"] == hexm.num_cells * 80 * 2
    assert info["num_classes"] >= 8          # the hexagon normals split the octants further


def test_plan_lattice_tilings():
    """Unstructured meshes of congruent cells are recognised as lattices: a hexagonal core gets one rhombic tiling
    per pair of lattice directions and every ordering class finds one whose patch graph is acyclic for it, so the
    one-launch dataflow kernel sweeps them all; quadrilaterals written as polygons behave like a Cartesian mesh;
    hexagons cut into triangles sit on a finer lattice with a third of its points empty: still tiled, and swept by
    the general kernel where the dataflow kernel's limits (levels per patch, steps back) do not hold."""
    for rings, order in ((6, 8), (20, 12)):
        mesh, xs, _ = syn.hex_core(rings, 6, num_groups=2)
        info = pb.plan_check(mesh, syn.level_symmetric(order), 2)
        assert info["lattice"] == 1 and info["num_tilings"] == 3
        assert info["num_classes"] == 12 and info["flow_classes"] == 12 and info["tile_classes"] == 12
    em, xs, quad, ls, z = util.load_golden("pwr_unstructured_s2")
    info = pb.plan_check(em, quad, 2)
    assert info["lattice"] == 1 and info["num_tilings"] == 1 and info["flow_classes"] >= 3
    # small patches are an explicit request for the general path
    mesh, xs, _ = syn.hex_core(6, 4, num_groups=2)
    info = pb.plan_check(mesh, syn.level_symmetric(4), 2, patch_cells=32)
    assert info["lattice"] == 0
    # triangles: a honeycomb of centroids = a triangular lattice with holes
    pts, cells, _ = syn.hex_lattice(4, 1.0)
    pts = list(map(tuple, pts))
    tri = []
    for c in cells:
        cx = sum(pts[p][0] for p in c) / 6.0; cy = sum(pts[p][1] for p in c) / 6.0
        pts.append((cx, cy))
        for a in range(6):
            tri.append([c[a], c[(a + 1) % 6], len(pts) - 1])
    tmesh = syn.polygon_mesh(np.array(pts), tri, np.ones(3), np.zeros(3 * len(tri), dtype=np.int32))
    info = pb.plan_check(tmesh, syn.level_symmetric(4), 2)
    assert info["lattice"] == 1 and info["tile_classes"] == info["num_classes"] == 12


def test_plan_sharding_partitions_the_work():
    q8 = syn.level_symmetric(8)
    mesh, xs = syn.checkerboard_core(20, 20, 8, num_groups=4)
    for mode in (0, 1):
        tot = sum(pb.plan_check(mesh, q8, 4, rank=r, num_ranks=4, shard_mode=mode)["updates_per_sweep"]
                  for r in range(4))
        assert tot == 20 * 20 * 8 * 80 * 4


def test_plan_rejects_bad_input():
    em, xs, quad, ls, z = util.load_golden("slabs_s2")
    em.bc_types = [0, pb.BC_VACUUM, 3]
    with pytest.raises(pb.SNError, match="boundary condition not implemented"):
        pb.plan_check(em, quad, 2)


def test_level_symmetric_matches_oracle_tables():
    for order in (2, 4, 6, 8, 12):                    # S12: added table, pinned in test_oracle.py
        q = syn.level_symmetric(order)
        d, w, r = orc.quadrature(order)
        assert np.array_equal(q.directions, d) and np.array_equal(q.weights, w)
        assert np.array_equal(q.reflected, r)
    assert len(syn.level_symmetric(12).weights) == 168


def test_builders_agree_with_oracle_geometry():
    """synthetic.cartesian_mesh / polygon_mesh describe the same cells as the oracle's meshes."""
    rng = np.random.default_rng(3)
    dx, dy, dz = rng.uniform(1, 2, 5), rng.uniform(1, 2, 4), rng.uniform(1, 2, 3)
    mats = rng.integers(0, 2, size=(3, 4, 5)); mats[:, 3, 4] = -1
    em = syn.cartesian_mesh(dx, dy, dz, mats)
    names = ["-x", "+x", "-y", "+y", "-z", "+z"]
    om = orc.build_cartesian_mesh(dx, dy, dz, mats.reshape(-1), names, [0] + [orc.VACUUM] * 6)
    em2 = util.extruded_from_oracle(om)
    for f in ("xy_neighbor", "xy_face_fx", "xy_face_fy", "xy_area", "materials", "xy_ij"):
        assert np.allclose(getattr(em, f), getattr(em2, f)), f
    assert np.allclose(em.xy_face_cf, em2.xy_face_cf, atol=1e-14)
    mesh_d, xs, (points, cells) = syn.hex_core(3, 2, pitch=2.0, dz=3.0, num_groups=2)
    om = orc.build_unstructured_mesh(points, cells, np.full(2, 3.0), mesh_d.materials, ["-z", "+z", "exterior"],
                                     ["exterior"], [[]], 2, [0, 1, 1, 1], 3)
    em2 = util.extruded_from_oracle(om)
    assert np.array_equal(em2.xy_neighbor, mesh_d.xy_neighbor)
    for f in ("xy_face_fx", "xy_face_fy", "xy_area", "xy_face_cf"):
        assert np.allclose(getattr(mesh_d, f), getattr(em2, f), atol=1e-12), f


# ------------------------------------------------------------------------------ C++ host library
@pytest.fixture(scope="module")
def decks(tmp_path_factory):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_decks", os.path.join(ROOT, "tests", "decks", "make_decks.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.main(str(tmp_path_factory.mktemp("decks")))


def _host_lib():
    lib = ctypes.CDLL(os.path.join(ROOT, "pampa_b200", "lib", "libpampa.so"))
    lib.pampa_debug_describe.restype = ctypes.c_int
    lib.pampa_debug_describe.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_double)]
    return lib


def _oracle_digest(deck):
    m = deck.mesh
    out = np.zeros(16)
    out[0], out[1], out[2] = m.num_cells, m.num_dims, m.volumes.sum()
    out[3], out[4] = len(m.face_area), m.face_area.sum()
    out[5] = (m.face_neighbor < 0).sum()
    out[6] = m.face_neighbor[m.face_neighbor >= 0].sum()
    w = np.array([1.0, 2.0, 3.0])
    out[7], out[8], out[9] = (m.centroids @ w).sum(), (m.face_centroid @ w).sum(), (m.face_normal @ w).sum()
    out[10] = m.materials.sum()
    for x in deck.xs:
        G = x.G
        out[11] += x.sigma_total.sum(); out[13] += x.nu_sigma_fission.sum()
        out[14] += x.kappa_sigma_fission.sum(); out[15] += x.chi_effective.sum()
        out[12] += sum((1 + g + 2 * g2) * x.sigma_scattering[g, g2] for g in range(G) for g2 in range(G))
    return out


def _describe(path):
    lib = _host_lib()
    out = (ctypes.c_double * 16)()
    cwd = os.getcwd()
    os.chdir(os.path.dirname(path))                  # deck paths are relative to the cwd, as in the reference
    try:
        rc = lib.pampa_debug_describe(os.path.basename(path).encode(), out)
    finally:
        os.chdir(cwd)
    assert rc == 0
    return np.array(out[:])


@pytest.mark.parametrize("case", ["slab_s2", "slab_s4", "pwr_cartesian_s2", "pwr_unstructured_s2"])
def test_host_parser_and_meshes_match_oracle(decks, case):
    """The C++ Parser / CartesianMesh / UnstructuredExtrudedMesh / Material code builds the same
    cells, faces, neighbours and cross sections as the oracle's restatement of the reference."""
    path = os.path.join(decks, case, "input.pmp")
    got = _describe(path)
    want = _oracle_digest(orc.read_deck(path))
    assert np.allclose(got, want, rtol=1e-13, atol=1e-12), (got, want)


REF_DECKS = {"slabs/reflected-s2": "slab_s2", "slabs/reflected-s4": "slab_s4",
             "pwr-iaea-benchmark/cartesian-sn": "pwr_cartesian_s2",
             "pwr-iaea-benchmark/unstructured-sn": "pwr_unstructured_s2"}


@pytest.mark.skipif(not os.path.isdir("/root/reference/test"), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("ref", sorted(REF_DECKS))
def test_host_reads_the_reference_decks(decks, ref):
    """The reference's own input files parse to the same problem as the generated decks."""
    a = _describe(os.path.join("/root/reference/test", ref, "input.pmp"))
    b = _describe(os.path.join(decks, REF_DECKS[ref], "input.pmp"))
    assert np.array_equal(a, b)


@pytest.mark.parametrize("case", ["slab_s2", "pwr_cartesian_s2", "pwr_unstructured_s2"])
def test_host_mesh_vtk(decks, case, tmp_path):
    """Mesh::writeVTK in the reference's format (src/vtk.cxx:28-121, src/CartesianMesh.cxx:157-231,
    src/UnstructuredExtrudedMesh.cxx:160-207): point / cell counts, VTK cell types, 1-based materials, and the
    point lists really are the corners of the oracle's cells (their mean is the cell centroid on these meshes)."""
    lib = _host_lib()
    path = os.path.join(decks, case, "input.pmp")
    prefix = str(tmp_path / "mesh_data")
    cwd = os.getcwd()
    os.chdir(os.path.dirname(path))
    try:
        assert lib.pampa_debug_write_mesh_vtk(b"input.pmp", prefix.encode()) == 0
    finally:
        os.chdir(cwd)
    v = util.read_vtk(prefix + ".vtk")
    m = orc.read_deck(path).mesh
    assert len(v["cells"]) == m.num_cells == v["num_cell_data"]
    npts = {len(c) for c in v["cells"]}
    want_type = {2: 3, 3: 5, 4: 9, 6: 13, 8: 12, 12: 16}
    assert all(t == want_type[len(c)] for t, c in zip(v["types"], v["cells"]))
    assert npts <= ({2} if m.num_dims == 1 else {3, 4, 5, 6} if m.num_dims == 2 else {6, 8, 12})
    assert max(max(c) for c in v["cells"]) < len(v["points"])
    name, mats = v["scalars"][0]
    assert name == "materials" and len(v["scalars"]) == 1
    assert np.array_equal(mats.astype(int), m.materials + 1)
    cen = np.array([v["points"][c].mean(axis=0) for c in v["cells"]])
    assert np.allclose(cen[:, :m.num_dims], m.centroids[:, :m.num_dims], atol=1e-5)   # files carry 7 digits


def test_host_ptc_writer(tmp_path):
    """`petsc dump 1`: PETSc binary Vec files (big-endian class id 1211214, length, float64 values) as
    VecView writes them on a binary viewer (src/petsc.cxx:491-511)."""
    lib = _host_lib()
    v = np.random.default_rng(3).normal(size=70001)
    prefix = str(tmp_path / "angular_flux")
    lib.pampa_debug_write_ptc.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_long]
    assert lib.pampa_debug_write_ptc(prefix.encode(), 0, v.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), v.size) == 0
    raw = open(prefix + "_0.ptc", "rb").read()
    head = np.frombuffer(raw[:8], dtype=">i4")
    assert head[0] == 1211214 and head[1] == v.size and len(raw) == 8 + 8 * v.size
    assert np.array_equal(np.frombuffer(raw[8:], dtype=">f8"), v)


def test_host_c_api_exports():
    hdr = open(os.path.join(ROOT, "include", "pampa.h")).read()
    lib = _host_lib()
    for name in set(re.findall(r"\b(pampa_[a-z_]+)\s*\(", hdr)):
        assert hasattr(lib, name), name


def test_host_fails_loudly_without_gpu(decks):
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    exe = os.path.join(ROOT, "pampa_b200", "bin", "pampa")
    r = subprocess.run([exe, "input.pmp"], cwd=os.path.join(decks, "slab_s2"), capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.parametrize("case", ["slab_s2", "pwr_cartesian_s2", "pwr_unstructured_s2"])
def test_partitioned_mesh_round_trip(decks, case, tmp_path):
    """Mesh::writeData (src/Mesh.cxx:408-569) / `mesh partitioned` (src/PartitionedMesh.cxx:4-263): every array of a
    mesh written in the reference's plain-text format and read back gives the same discrete mesh -- the digest of the
    deck that points at the written file equals the digest of the original deck -- at full precision and, to the
    three decimals it keeps, in the reference's own fixed format."""
    import shutil
    lib = ctypes.CDLL(os.path.join(ROOT, "pampa_b200", "lib", "libpampa.so"))
    lib.pampa_debug_describe.restype = ctypes.c_int
    lib.pampa_debug_describe.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_double)]
    lib.pampa_debug_write_mesh_data.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    work = tmp_path / case
    shutil.copytree(os.path.join(decks, case), work)
    cwd = os.getcwd()
    os.chdir(work)
    try:
        ref = (ctypes.c_double * 16)()
        assert lib.pampa_debug_describe(b"input.pmp", ref) == 0
        for digits, tol in ((17, 1e-13), (-1, 2e-3)):
            assert lib.pampa_debug_write_mesh_data(b"input.pmp", b"mesh_data.pmp", digits) == 0
            written = open("mesh_data.pmp").read()
            assert written.startswith("points ") and "\ncells %d 0 %d\n" % (int(ref[0]), int(ref[0])) in written
            text = open("input.pmp").read()
            kind = "cartesian" if "mesh cartesian" in text else "unstructured"
            open("input_part.pmp", "w").write(text.replace("mesh %s mesh.pmp" % kind, "mesh partitioned mesh_data.pmp"))
            got = (ctypes.c_double * 16)()
            assert lib.pampa_debug_describe(b"input_part.pmp", got) == 0
            for a in range(16):
                assert abs(got[a] - ref[a]) <= tol * max(1.0, abs(ref[a])), (digits, a, got[a], ref[a])
        # one rank's part of a domain decomposition (ghost cells) is refused with an explanation
        text = open("mesh_data.pmp").read()
        n = int(ref[0])
        open("mesh_ghost.pmp", "w").write(text.replace("cells %d 0 %d" % (n, n), "cells %d 3 %d" % (n, 2 * n)))
        open("input_ghost.pmp", "w").write(open("input_part.pmp").read().replace("mesh_data.pmp", "mesh_ghost.pmp"))
        assert lib.pampa_debug_describe(b"input_ghost.pmp", got) != 0
    finally:
        os.chdir(cwd)
