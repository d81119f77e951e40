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
    for order in (2, 4, 6, 8):
        q = syn.level_symmetric(order)
        d, w, r = orc.quadrature(order)
        assert np.array_equal(q.directions, d) and np.array_equal(q.weights, w)
        assert np.array_equal(q.reflected, r)
    q12 = syn.level_symmetric(12)                     # added table: check the LQn moment conditions
    assert len(q12.weights) == 168
    assert abs(q12.weights.sum() - 1.0) < 1e-12
    for ax in range(3):
        assert abs((q12.weights * q12.directions[:, ax] ** 2).sum() - 1.0 / 3.0) < 1e-6
        assert abs((q12.weights * q12.directions[:, ax] ** 4).sum() - 1.0 / 5.0) < 1e-6


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
