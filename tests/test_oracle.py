"""CPU tests of the oracle: pinned to the reference's goldens, self-consistent, C port agrees."""
import math
import os

import numpy as np
import pytest

import util
from oracle import pampa_oracle as orc

REF = "/root/reference/test"
GOLDENS = {"slabs_s2": 0.970849, "slabs_s4": 0.982472, "pwr_cartesian_s2": 0.965761,
           "pwr_unstructured_s2": 0.965761}


@pytest.mark.parametrize("name", sorted(GOLDENS))
def test_fixture_matches_reference_golden(name):
    """The committed oracle solution prints the k-eff of test/check_ref.txt:32,53,234,415."""
    z = np.load(os.path.join(util.GOLDEN, name + ".npz"))
    assert float(z["golden_keff"]) == GOLDENS[name]
    assert "%.6f" % float(z["keff"]) == "%.6f" % GOLDENS[name]
    # normalisation of NeutronicSolver.cxx:46-78: total power 1
    assert abs(z["power"].sum() - 1.0) < 1e-12
    assert z["phi"].min() > 0.0


def _operator_from_fixture(name):
    """Rebuild the oracle operator from the committed arrays only (no reference tree needed)."""
    em, xs, quad, ls, z = util.load_golden(name)
    return em, xs, quad, ls, z


def test_oracle_resolves_slab_fixture():
    """Re-solve slabs_s2 from the fixture's arrays with the matrix-free oracle path: the stored
    solution is reproduced, i.e. the fixture is self-contained."""
    em, xs, quad, ls, z = util.load_golden("slabs_s2")
    nx = em.num_xy_cells
    dx = em.xy_area                                  # 1-D: base "area" = dx, unit face area
    mats = em.materials
    obcs = [0, orc.VACUUM, orc.VACUUM]
    mesh = orc.build_cartesian_mesh(dx, None, None, mats, ["-x", "+x"], obcs)
    op = orc.build_operator(mesh, util.xs_to_oracle(xs), 2, int(z["order"]), 1.0, "literal_zero_init", obcs)
    a = orc.solve_monolithic(op)
    b = orc.solve_matrix_free(op)
    assert abs(a.keff - float(z["keff"])) < 1e-11
    assert abs(b.keff - a.keff) < 1e-11
    assert util.rel_l2(a.phi, z["phi"]) < 1e-10
    assert util.rel_l2(b.phi, a.phi) < 1e-10
    assert util.rel_l2(a.psi, z["psi"]) < 1e-10


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("case,gold", [("slabs/reflected-s2", 0.970849), ("slabs/reflected-s4", 0.982472)])
def test_oracle_on_reference_decks(case, gold):
    sol = orc.solve_deck(os.path.join(REF, case, "input.pmp"))
    assert "%.6f" % sol.keff == "%.6f" % gold


def test_quadrature_tables():
    for order in (2, 4, 6, 8, 12):
        d, w, refl = orc.quadrature(order)
        assert len(w) == order * (order + 2)
        assert abs(w.sum() - 1.0) < 1e-6
        assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-6)
        for m in range(len(w)):
            for ax in range(3):
                r = d[m].copy(); r[ax] = -r[ax]
                assert np.allclose(d[refl[m, ax]], r)
    with pytest.raises(ValueError):
        orc.quadrature(16)
    # S12 is not in the reference (it stops at S8): its constants are pinned by the equations that define a
    # level-symmetric set.  (1) mu_i^2 is an arithmetic progression fixed by mu_1; (2) the quadrature integrates
    # every even moment through order 12 (7-digit constants: 6e-8), and not order 14 -- a wrong digit in any
    # of the five weights would show at the 1e-6 level.
    d, w, _ = orc.quadrature(12)
    mu = np.unique(np.round(np.abs(d[:, 0]), 7))
    assert len(mu) == 6
    step = 2.0 * (1.0 - 3.0 * mu[0] ** 2) / 10.0
    assert np.allclose(mu ** 2, mu[0] ** 2 + step * np.arange(6), atol=2e-7)
    for n in range(2, 13, 2):
        for ax in range(3):
            assert abs(np.sum(w * d[:, ax] ** n) - 1.0 / (n + 1)) < 1e-7, (n, ax)
    assert abs(np.sum(w * d[:, 0] ** 14) - 1.0 / 15) > 1e-6
    assert abs(np.sum(w * d[:, 0] ** 2 * d[:, 1] ** 2) - 1.0 / 15) < 1e-7
    assert abs(np.sum(w * d[:, 0] ** 2 * d[:, 1] ** 2 * d[:, 2] ** 2) - 1.0 / 105) < 1e-7


def test_c_port_matches_oracle():
    """oracle/sweep_cpu.c (the CPU baseline) solves the same discrete problem as the oracle."""
    from oracle import sweep_cpu
    from pampa_b200 import synthetic as syn
    rng = np.random.default_rng(1)
    nx, ny, nz, G = 8, 7, 6, 2
    dx, dy, dz = rng.uniform(1, 2, nx), rng.uniform(1, 2, ny), rng.uniform(1, 2, nz)
    mats = rng.integers(0, 2, size=(nz, ny, nx))
    xs = syn.synthetic_xs(G, seed=11)
    xs.nu_sigma_fission[0] *= 4
    quad = syn.level_symmetric(4)
    cpu = sweep_cpu.SweepCPU(dx, dy, dz, mats, xs.sigma_total, xs.sigma_scattering, xs.nu_sigma_fission,
                             xs.chi_effective, quad.directions, quad.weights)
    k, phi, it = cpu.solve(tol_k=1e-12, tol_phi=1e-10)
    names = ["-x", "+x", "-y", "+y", "-z", "+z"]
    obcs = [0] + [orc.VACUUM] * 6
    mesh = orc.build_cartesian_mesh(dx, dy, dz, mats.reshape(-1), names, obcs)
    op = orc.build_operator(mesh, util.xs_to_oracle(xs), G, 0, 1.0, "off", obcs, quad=util.quad_to_oracle(quad))
    sol = orc.solve_matrix_free(op)
    assert abs(k - sol.keff) < 1e-9
    ph = phi.reshape(G, -1).T
    ph = ph / np.sum(ph * op.kapsf * op.vol[:, None])
    assert util.rel_l2(ph, sol.phi) < 1e-7


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the
    contract's keys, produced by the oracle's C port on the host cores, no GPU involved."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--size", "48", "48", "48"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "updates/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "Cartesian core 48x48x48" in line["config"]["workload"]
