"""GPU parity tests: the CUDA path (through the C ABI) against the oracle.

Tolerances are the ones BASELINE.json's north_star states: k-eff within 1 pcm (1e-5), scalar flux
within 1e-5 relative L2 and 1e-4 max-relative per cell and group.  The solves are converged much
tighter than that (1e-10 / 1e-9), so the observed differences are ~1e-8."""
import math

import numpy as np
import pytest

import util
from oracle import pampa_oracle as orc
from pampa_b200 import problem as pb
from pampa_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

TOL_K, TOL_L2, TOL_MAX = 1.0e-5, 1.0e-5, 1.0e-4


def _solve(em, xs, quad, ls=None, **opts):
    dev = pb.SNDevice(em, xs, quad, ls, **opts)
    k, it = dev.solve_keff(tol_k=1e-11, tol_phi=1e-9, max_it=20000)
    return dev, k, it


def _check_solution(dev, k, sol_keff, sol_phi, sol_power):
    phi = dev.get("scalar-flux").reshape(sol_phi.shape)
    q = dev.get("power")
    assert abs(k - sol_keff) < TOL_K, (k, sol_keff)
    assert util.rel_l2(phi, sol_phi) < TOL_L2
    assert util.max_rel(phi, sol_phi) < TOL_MAX
    assert abs(q.sum() - sol_power.sum()) < 1e-9 * abs(sol_power.sum())
    assert util.rel_l2(q, sol_power) < TOL_L2


@pytest.mark.parametrize("name", ["slabs_s2", "slabs_s4", "pwr_cartesian_s2", "pwr_unstructured_s2",
                                  "pwr_cartesian_s2_lsoff", "pwr_cartesian_s8_lsoff", "pwr_cartesian_s8",
                                  "hex_core_s8_2g", "hex_core_s8_11g", "hex_core_tri_s4_2g"])
def test_reference_cases(name):
    """The SN cases the reference ships (test/check_ref.txt:32,53,234,415), BASELINE configs 2 (the PWR deck as
    shipped, LS on, at S8) and 3 (hex-core at S8 with the reference's 2- and 11-group data on its hex-cells mesh,
    and at S4 on its tri-cells mesh) and two variants."""
    em, xs, quad, ls, z = util.load_golden(name)
    dev, k, it = _solve(em, xs, quad, ls)
    gold = float(z["golden_keff"])
    if not math.isnan(gold):
        assert "%.6f" % k == "%.6f" % gold          # what check.sh diffs
    _check_solution(dev, k, float(z["keff"]), z["phi"], z["power"])
    P = dev.get("production-rate")
    assert util.rel_l2(P, z["production"]) < TOL_L2
    if "psi" in z:
        psi = dev.get("angular-flux").reshape(z["psi"].shape)
        assert util.rel_l2(psi, z["psi"]) < TOL_L2
    dev.close()


def _oracle_cart(dx, dy, dz, mats, bcs, xs, quad, G, delta=1.0):
    names = ["-x", "+x"] + (["-y", "+y"] if dy is not None else []) + (["-z", "+z"] if dz is not None else [])
    obcs = [0] + [{pb.BC_VACUUM: orc.VACUUM, pb.BC_REFLECTIVE: orc.REFLECTIVE}[(bcs or {}).get(n, pb.BC_VACUUM)]
                  for n in names]
    mesh = orc.build_cartesian_mesh(dx, dy, dz, np.asarray(mats).reshape(-1), names, obcs)
    op = orc.build_operator(mesh, util.xs_to_oracle(xs), G, 0, delta, "off", obcs, quad=util.quad_to_oracle(quad))
    return mesh, op


def test_single_sweep_cartesian_3d():
    """Kernel-1 unit parity: one sweep with a frozen source equals T^-1 q of the oracle."""
    rng = np.random.default_rng(7)
    nx, ny, nz, G = 21, 18, 11, 3
    dx, dy, dz = rng.uniform(0.5, 1.5, nx), rng.uniform(0.5, 1.5, ny), rng.uniform(0.5, 1.5, nz)
    mats = rng.integers(0, 2, size=(nz, ny, nx))
    xs = syn.synthetic_xs(G, seed=3)
    quad = syn.level_symmetric(4)
    em = syn.cartesian_mesh(dx, dy, dz, mats)
    mesh, op = _oracle_cart(dx, dy, dz, mats, None, xs, quad, G)
    import scipy.sparse.linalg as spla
    N, M = op.N, op.M
    phi0 = rng.uniform(0.5, 1.5, size=(N, G))
    keff = 0.9
    qd = np.einsum("nfg,nf->ng", op.sig_s, phi0) + op.chi * np.sum(op.nusf * phi0, axis=1)[:, None] / keff
    b = np.repeat((qd * op.vol[:, None]).reshape(N * G), M)
    psi = spla.splu(op.T.tocsc()).solve(b).reshape(N, G, M)
    # default = dataflow tile kernel; wave_launch = one launch per wavefront (old tile kernel); z_chunk and
    # generic_only = the general kernel; store_psi = 0 keeps only the patch-edge copies of psi
    for opts in ({}, {"wave_launch": 1}, {"z_chunk": 4}, {"tile_i": 8, "tile_j": 4}, {"generic_only": 1},
                 {"dt_max": 2, "z_chunk": 3}, {"dt_max": 3, "generic_only": 1}, {"dt_max": 4}, {"store_psi": 0},
                 {"store_psi": 0, "tile_i": 8, "tile_j": 8},
                 # groups a dataflow task sweeps back to back (default 4 -> one block of 3 here): no merging,
                 # a ragged last block, and merged rows holding only the edge copies
                 {"group_merge": 1}, {"group_merge": 2}, {"group_merge": 2, "store_psi": 0, "dt_max": 3},
                 # perimeter-first lane order: neighbouring patches read the psi rows themselves, no edge copies
                 {"inline_edges": 1}, {"inline_edges": 1, "store_psi": 0}, {"inline_edges": 1, "tile_i": 8, "tile_j": 4}):
        dev = pb.SNDevice(em, xs, quad, **opts)
        dev.set("flux-moments", phi0.reshape(-1))
        dev.source(keff)
        dev.sweep()
        dev.reduce()
        got_phi = dev.get("flux-moments").reshape(N, G)
        assert util.max_rel(got_phi, psi @ op.w) < 1e-11
        if opts.get("store_psi", 1):
            got_psi = dev.get("angular-flux").reshape(N, G, M)
            assert util.rel_l2(got_psi, psi) < 1e-12
        else:
            with pytest.raises(pb.SNError, match="store_psi"):
                dev.get("angular-flux")
        dev.close()


def test_single_sweep_cartesian_s12():
    """S12 (168 directions, 21 per octant: chunks of 5+4+4+4+4) on a small Cartesian mesh, 2 groups: one sweep against
    T^-1 q of the oracle, dataflow kernel and wavefront launches."""
    import scipy.sparse.linalg as spla
    rng = np.random.default_rng(12)
    nx, ny, nz, G = 12, 10, 7, 2
    dx, dy, dz = rng.uniform(0.5, 1.5, nx), rng.uniform(0.5, 1.5, ny), rng.uniform(0.5, 1.5, nz)
    mats = rng.integers(0, 2, size=(nz, ny, nx))
    xs = syn.synthetic_xs(G, seed=5)
    quad = syn.level_symmetric(12)
    em = syn.cartesian_mesh(dx, dy, dz, mats)
    mesh, op = _oracle_cart(dx, dy, dz, mats, None, xs, quad, G)
    N, M = op.N, op.M
    assert M == 168
    phi0 = rng.uniform(0.5, 1.5, size=(N, G))
    keff = 1.1
    qd = np.einsum("nfg,nf->ng", op.sig_s, phi0) + op.chi * np.sum(op.nusf * phi0, axis=1)[:, None] / keff
    b = np.repeat((qd * op.vol[:, None]).reshape(N * G), M)
    psi = spla.splu(op.T.tocsc()).solve(b).reshape(N, G, M)
    for opts in ({}, {"wave_launch": 1}):
        dev = pb.SNDevice(em, xs, quad, **opts)
        assert dev.info()["num_chunks"] == 8 * 5
        dev.set("flux-moments", phi0.reshape(-1))
        dev.source(keff)
        dev.sweep()
        dev.reduce()
        assert util.max_rel(dev.get("flux-moments").reshape(N, G), psi @ op.w) < 1e-11
        assert util.rel_l2(dev.get("angular-flux").reshape(N, G, M), psi) < 1e-12
        dev.close()


def test_keff_cartesian_3d_reflective():
    """3-D Cartesian core with void corner cells, reflective -x/-y/-z and vacuum +x/+y/+z, S4."""
    nx, ny, nz, G = 12, 12, 8, 2
    mats = np.zeros((nz, ny, nx), dtype=int)
    mats[:, :, 8:] = 1; mats[:, 8:, :] = 1; mats[6:] = 1
    mats[:, 10:, 10:] = -1
    xs = syn.synthetic_xs(G, seed=11)
    xs.nu_sigma_fission[0] *= 4.0; xs.kappa_sigma_fission[0] *= 4.0
    quad = syn.level_symmetric(4)
    bcs = {"-x": pb.BC_REFLECTIVE, "-y": pb.BC_REFLECTIVE, "-z": pb.BC_REFLECTIVE}
    h = np.full(nx, 2.0)
    em = syn.cartesian_mesh(h, h, np.full(nz, 2.5), mats, bcs)
    mesh, op = _oracle_cart(h, h, np.full(nz, 2.5), mats, bcs, xs, quad, G)
    sol = orc.solve_matrix_free(op)
    for opts in ({}, {"group_merge": 1}, {"generic_only": 1}, {"z_chunk": 3, "dt_max": 2}):
        dev, k, it = _solve(em, xs, quad, **opts)
        _check_solution(dev, k, sol.keff, sol.phi, sol.power)
        psi = dev.get("angular-flux").reshape(sol.psi.shape)
        assert util.rel_l2(psi, sol.psi) < TOL_L2
        dev.close()


def _hex_problem(nrings, nz, G, order, seed, delta=1.0):
    mesh_d, xs, (points, cells) = syn.hex_core(nrings, nz, pitch=2.0, dz=3.0, num_groups=G, seed=seed, delta=delta)
    xs.nu_sigma_fission[0] *= 4.0; xs.kappa_sigma_fission[0] *= 4.0
    quad = syn.level_symmetric(order)
    bnames = ["-z", "+z", "exterior"]
    obcs = [0, orc.VACUUM, orc.VACUUM, orc.VACUUM]
    omesh = orc.build_unstructured_mesh(points, cells, np.full(nz, 3.0), mesh_d.materials, bnames, ["exterior"],
                                        [[]], 2, obcs, 3)
    # the reference numbers the default boundary by its xy ordinal (quirk C.8): ordinal 2 -> index 3 here
    op = orc.build_operator(omesh, util.xs_to_oracle(xs), G, 0, delta, "off", obcs, quad=util.quad_to_oracle(quad))
    return mesh_d, xs, quad, op


def test_single_sweep_hex_3d():
    """Hexagonal prisms: several ordering classes per octant, level-chunk patches."""
    rng = np.random.default_rng(5)
    em, xs, quad, op = _hex_problem(6, 5, 2, 8, seed=2)
    import scipy.sparse.linalg as spla
    N, G, M = op.N, op.G, op.M
    phi0 = rng.uniform(0.5, 1.5, size=(N, G))
    qd = np.einsum("nfg,nf->ng", op.sig_s, phi0) + op.chi * np.sum(op.nusf * phi0, axis=1)[:, None]
    b = np.repeat((qd * op.vol[:, None]).reshape(N * G), M)
    psi = spla.splu(op.T.tocsc()).solve(b).reshape(N, G, M)
    # default: the mesh is recognised as a lattice, every class is swept by the dataflow kernel on a rhombic tiling
    # (three incoming faces, sources one and two steps back); wave_launch / z_chunk / generic_only: the general
    # kernel on the same tilings; small patches: k-d leaves and level chunks, >= 32 launches per sweep, replayed from
    # a CUDA graph unless no_graph is set
    for opts in ({}, {"dt_max": 3}, {"dt_max": 5}, {"dt_max": 6, "group_merge": 1}, {"store_psi": 0}, {"wave_launch": 1},
                 {"z_chunk": 2}, {"generic_only": 1}, {"tile_i": 8, "tile_j": 8}, {"tile_i": 4, "tile_j": 8, "store_psi": 0},
                 {"patch_cells": 32}, {"patch_cells": 64, "z_chunk": 2}, {"patch_cells": 32, "no_graph": 1}):
        dev = pb.SNDevice(em, xs, quad, **opts)
        for _ in range(3):                      # the graph is captured in the first sweep of each buffer parity
            dev.set("flux-moments", phi0.reshape(-1))
            dev.source(1.0)
            dev.sweep()
            dev.reduce()
            assert util.max_rel(dev.get("flux-moments").reshape(N, G), psi @ op.w) < 1e-11, opts
            if opts.get("store_psi", 1):
                got_psi = dev.get("angular-flux").reshape(N, G, M)
                assert util.rel_l2(got_psi, psi) < 1e-12, opts
        if not opts:
            info = dev.info()
            assert info["lattice"] == 1 and info["flow_classes"] == info["num_classes"] and info["sweep_launches"] <= 4
        dev.close()


def test_keff_hex_3d():
    em, xs, quad, op = _hex_problem(5, 6, 2, 4, seed=4)
    sol = orc.solve_matrix_free(op)
    for opts in ({}, {"store_psi": 0}, {"patch_cells": 64}, {"patch_cells": 64, "no_graph": 1}):
        dev, k, it = _solve(em, xs, quad, **opts)
        _check_solution(dev, k, sol.keff, sol.phi, sol.power)
        dev.close()


def test_hex_schedule_independence(monkeypatch):
    """Same as test_sweep_schedule_independence on a hexagonal core (three tilings, the three-face dataflow kernel):
    publishing the progress counter every row, one / three / eight groups per task and rows that hold only the edge
    copies must all give the same bits."""
    G = 8
    mesh, xs, _ = syn.hex_core(40, 48, pitch=1.0, dz=1.0, num_groups=G, seed=54321)
    quad = syn.level_symmetric(8)
    ref = None
    for opts, dbg in (({}, None), ({}, "16"), ({"group_merge": 1}, "16"), ({"group_merge": 3, "store_psi": 0}, "16"),
                      ({"group_merge": 8, "store_psi": 0}, None)):
        if dbg is None:
            monkeypatch.delenv("PAMPA_SN_DBG", raising=False)
        else:
            monkeypatch.setenv("PAMPA_SN_DBG", dbg)
        dev = pb.SNDevice(mesh, xs, quad, **opts)
        k = dev.iterate(3)
        phi = dev.get("flux-moments")
        dev.close()
        if ref is None:
            ref = (k, phi)
            assert phi.min() > 0.0
        else:
            assert k == ref[0], (opts, dbg)
            assert np.array_equal(phi, ref[1]), (opts, dbg)
    # and the general kernel on the same tilings agrees to rounding (different arithmetic order)
    monkeypatch.delenv("PAMPA_SN_DBG", raising=False)
    dev = pb.SNDevice(mesh, xs, quad, generic_only=1)
    k2 = dev.iterate(3)
    phi2 = dev.get("flux-moments")
    dev.close()
    assert abs(k2 - ref[0]) < 1e-12 * abs(k2)
    assert util.max_rel(phi2, ref[1]) < 1e-11


def test_delta_golden_pwr():
    """The reference's PWR deck with its default mixed-face-interpolation (0.1) and LS off: k and the scalar flux of
    the committed oracle eigenpair; the angular flux has negative values there (the reference fails that solve)."""
    em, xs, quad, ls, z = util.load_golden("pwr_cartesian_s2_delta01_lsoff")
    assert em.delta == 0.1 and float(z["psi_min"]) < 0.0
    dev, k, it = _solve(em, xs, quad, ls)
    assert abs(k - float(z["keff"])) < TOL_K
    phi = dev.get("scalar-flux").reshape(z["phi"].shape)
    assert util.rel_l2(phi, z["phi"]) < TOL_L2 and util.max_rel(phi, z["phi"]) < TOL_MAX
    assert abs(float(dev.get("angular-flux-min")[0]) - float(z["psi_min"])) < 1e-5 * float(z["psi_max"])
    dev.close()


def test_single_sweep_hex_s12_16g():
    """BASELINE config 5 in small: hexagonal prisms, S12 (168 directions), 16 groups.  One sweep with a frozen
    source against the sparse LU of the oracle's operator, and a k-eff solve against the oracle's eigenpair."""
    rng = np.random.default_rng(12)
    em, xs, quad, op = _hex_problem(3, 3, 16, 12, seed=54321)
    assert len(quad.weights) == 168
    import scipy.sparse.linalg as spla
    N, G, M = op.N, op.G, op.M
    phi0 = rng.uniform(0.5, 1.5, size=(N, G))
    qd = np.einsum("nfg,nf->ng", op.sig_s, phi0) + op.chi * np.sum(op.nusf * phi0, axis=1)[:, None]
    b = np.repeat((qd * op.vol[:, None]).reshape(N * G), M)
    psi = spla.splu(op.T.tocsc()).solve(b).reshape(N, G, M)
    for opts in ({}, {"patch_cells": 32}):
        dev = pb.SNDevice(em, xs, quad, **opts)
        dev.set("flux-moments", phi0.reshape(-1))
        dev.source(1.0)
        dev.sweep()
        dev.reduce()
        assert util.rel_l2(dev.get("angular-flux").reshape(N, G, M), psi) < 1e-12
        assert util.max_rel(dev.get("flux-moments").reshape(N, G), psi @ op.w) < 1e-11
        dev.close()
    sol = orc.solve_matrix_free(op)
    dev, k, it = _solve(em, xs, quad)
    _check_solution(dev, k, sol.keff, sol.phi, sol.power)
    dev.close()


def _check_delta(dev, op, sol, k):
    """delta < 1 eigenpairs may have negative angular fluxes (the reference fails those solves; the device layer
    reports the minimum): compare k, the scalar flux and the angular flux with the oracle's eigenvector."""
    phi = dev.get("scalar-flux").reshape(sol.phi.shape)
    assert abs(k - sol.keff) < TOL_K, (k, sol.keff)
    assert util.rel_l2(phi, sol.phi) < TOL_L2
    scale = np.abs(sol.phi).max()
    assert np.max(np.abs(phi - sol.phi)) < TOL_MAX * scale
    psi_min = float(dev.get("angular-flux-min")[0])
    assert abs(psi_min - min(0.0, sol.psi.min())) < 1e-6 * np.abs(sol.psi).max()
    try:
        psi = dev.get("angular-flux").reshape(sol.psi.shape)
        assert sol.psi.min() >= 0.0
    except pb.SNError as e:                     # the export reports negative values the way the reference does
        assert "negative values in the angular-flux solution" in str(e) and sol.psi.min() < 0.0
        return
    assert util.rel_l2(psi, sol.psi) < TOL_L2


@pytest.mark.parametrize("delta", [0.1, 0.5])
def test_delta_slabs(delta):
    """mixed-face-interpolation < 1 (the reference's default is 0.1, src/SNSolver.hxx:16) on the reference's
    slab problem: the deferred correction converges to the eigenpair of the reference's delta operator."""
    em, xs, quad, ls, z = util.load_golden("slabs_s2")
    obcs = [0, orc.VACUUM, orc.VACUUM]
    mesh = orc.build_cartesian_mesh(em.xy_area, None, None, em.materials, ["-x", "+x"], obcs)
    op = orc.build_operator(mesh, util.xs_to_oracle(xs), 2, 2, delta, "off", obcs)
    sol = orc.solve_monolithic(op, allow_negative=True)
    emd = syn.cartesian_mesh(em.xy_area, None, None, em.materials.reshape(1, 1, -1), delta=delta)
    for opts in ({}, {"anderson_depth": -1}):
        dev, k, it = _solve(emd, xs, quad, **opts)
        print("delta %.1f %s: k %.9f (oracle %.9f) in %d iterations" % (delta, opts, k, sol.keff, it))
        _check_delta(dev, op, sol, k)
        dev.close()


def test_delta_cartesian_3d():
    """delta = 0.1 on a 3-D Cartesian core with non-uniform spacings, reflective -x / -z, void cells."""
    rng = np.random.default_rng(3)
    nx, ny, nz, G = 10, 9, 7, 2
    dx, dy, dz = rng.uniform(0.8, 1.6, nx), rng.uniform(0.8, 1.6, ny), rng.uniform(0.8, 1.6, nz)
    mats = np.zeros((nz, ny, nx), dtype=int)
    mats[:, :, 6:] = 1; mats[:, 6:, :] = 1; mats[5:] = 1
    mats[:, 8:, 8:] = -1
    xs = syn.synthetic_xs(G, seed=11)
    xs.nu_sigma_fission[0] *= 4.0; xs.kappa_sigma_fission[0] *= 4.0
    quad = syn.level_symmetric(4)
    bcs = {"-x": pb.BC_REFLECTIVE, "-z": pb.BC_REFLECTIVE}
    for delta in (0.1, 0.6):
        em = syn.cartesian_mesh(dx, dy, dz, mats, bcs, delta=delta)
        mesh, op = _oracle_cart(dx, dy, dz, mats, bcs, xs, quad, G, delta=delta)
        sol = orc.solve_matrix_free(op, allow_negative=True)
        for opts in ({}, {"z_chunk": 3, "dt_max": 2}):
            dev, k, it = _solve(em, xs, quad, **opts)
            print("delta %.1f %s: k %.9f (oracle %.9f) in %d iterations" % (delta, opts, k, sol.keff, it))
            _check_delta(dev, op, sol, k)
            dev.close()


def test_delta_hex_3d():
    """delta = 0.1 on hexagonal prisms (level-chunk patches, three incoming faces)."""
    em, xs, quad, op = _hex_problem(4, 5, 2, 4, seed=4, delta=0.1)
    sol = orc.solve_matrix_free(op, allow_negative=True)
    dev, k, it = _solve(em, xs, quad, patch_cells=64)
    _check_delta(dev, op, sol, k)
    dev.close()


def test_reduced_c4_matches_cpu_port():
    """BASELINE config 4 itself has no independent answer at 216^3 (the oracle cannot run there): the same
    generator at 40^3 cells, S8, 8 groups, converged on the device (Anderson) and by the oracle's C port (plain
    power iteration on the host cores) must give the same k and flux."""
    from oracle import sweep_cpu
    n, G = 40, 8
    mesh, xs = syn.checkerboard_core(n, n, n, num_groups=G)
    quad = syn.level_symmetric(8)
    dev = pb.SNDevice(mesh, xs, quad)
    k, it = dev.solve_keff(tol_k=1e-11, tol_phi=1e-10, max_it=20000)
    phi = dev.get("flux-moments").reshape(n, n, n, G)
    dev.close()
    cpu = sweep_cpu.SweepCPU(np.ones(n), np.ones(n), np.ones(n), mesh.materials.reshape(n, n, n), xs.sigma_total,
                             xs.sigma_scattering, xs.nu_sigma_fission, xs.chi_effective, quad.directions,
                             quad.weights)
    kc, phic, itc = cpu.solve(tol_k=1e-11, tol_phi=1e-10, max_it=20000)
    phic = phic.transpose(1, 2, 3, 0)
    print("%d^3 core: device k %.10f in %d iterations, CPU port k %.10f in %d" % (n, k, it, kc, itc))
    assert abs(k - kc) < TOL_K * kc
    a, b = phi / np.linalg.norm(phi), phic / np.linalg.norm(phic)
    assert util.rel_l2(a, b) < TOL_L2
    assert util.max_rel(a, b) < TOL_MAX


def test_errors_are_loud():
    """Wrong inputs fail with the reference's messages instead of computing something else."""
    em, xs, quad, ls, z = util.load_golden("slabs_s2")
    em.bc_types = [0, pb.BC_VACUUM, 0]
    with pytest.raises(pb.SNError, match="boundary condition not implemented"):
        pb.SNDevice(em, xs, quad)
    em.bc_types = [0, pb.BC_VACUUM, pb.BC_VACUUM]
    xs.nu_sigma_fission[:] = 0.0; xs.kappa_sigma_fission[:] = 0.0
    dev = pb.SNDevice(em, xs, quad)
    with pytest.raises(pb.SNError):
        dev.solve_keff(max_it=5)
    dev.close()


def test_anderson_matches_plain_power_iteration():
    """The accelerated solve converges to the same eigenpair as the plain power iteration, faster."""
    em, xs, quad, ls, z = util.load_golden("pwr_cartesian_s2_lsoff")
    plain = pb.SNDevice(em, xs, quad, ls, anderson_depth=-1)
    k0, it0 = plain.solve_keff(tol_k=1e-11, tol_phi=1e-9, max_it=20000)
    acc = pb.SNDevice(em, xs, quad, ls)
    k1, it1 = acc.solve_keff(tol_k=1e-11, tol_phi=1e-9, max_it=20000)
    assert abs(k0 - k1) < 1e-8 and abs(k1 - float(z["keff"])) < TOL_K
    assert util.rel_l2(acc.get("scalar-flux"), plain.get("scalar-flux")) < 1e-6
    assert util.rel_l2(acc.get("angular-flux"), plain.get("angular-flux")) < 1e-6
    assert it1 < it0, (it1, it0)
    print("iterations: plain %d, Anderson %d" % (it0, it1))
    plain.close(); acc.close()


def test_sweep_schedule_independence(monkeypatch):
    """The sweep result must not depend on how it is scheduled: one launch per wavefront (stream order), the
    dataflow launch (patches coupled by progress counters) with 1 / 4 / 8 groups per task, publishing the counter
    every row (PAMPA_SN_DBG = 16: readers follow the writer as closely as possible).  Bit-exact: an unfenced
    counter store once let readers see rows of the previous sweep on a mesh of this shape."""
    G = 8
    mesh, xs = syn.checkerboard_core(64, 64, 96, num_groups=G)
    quad = syn.level_symmetric(4)
    ref = None
    # (the fused tail of the iteration sums the k integrals per un-shear CTA, the separate reduction pass -- which
    # graph-replayed plans such as wave_launch use -- per grid-stride block: same flux, last bits of k differ;
    # the schedules are compared under one reduction, test_fused_tail_matches_separate_passes compares the two)
    monkeypatch.setenv("PAMPA_SN_NO_FUSE", "1")
    for opts, dbg in (({"wave_launch": 1}, None), ({}, None), ({"group_merge": 8}, None), ({"group_merge": 8}, "16"),
                      ({"group_merge": 1}, "16"), ({"group_merge": 3, "store_psi": 0}, "16"),
                      ({"inline_edges": 1}, "16")):
        if dbg is None:
            monkeypatch.delenv("PAMPA_SN_DBG", raising=False)
        else:
            monkeypatch.setenv("PAMPA_SN_DBG", dbg)
        dev = pb.SNDevice(mesh, xs, quad, **opts)
        k = dev.iterate(3)
        phi = dev.get("flux-moments")
        dev.close()
        if ref is None:
            ref = (k, phi)
        else:
            assert k == ref[0], (opts, dbg)
            assert np.array_equal(phi, ref[1]), (opts, dbg)


@pytest.mark.parametrize("kind", ["cartesian", "hex"])
def test_fused_tail_matches_separate_passes(monkeypatch, kind):
    """Plain source iterations end with un-shear, reduction and rotation of the iterate; the dataflow path fuses the
    three into the last un-shear sweep of a column (and, sharded, the delivery to the peers); accelerated iterations
    fuse the reduction only.  One tiling (Cartesian): same flux moments bit for bit after one iteration, k to rounding
    (the integrals are summed in another order), and still after five.  Three tilings (hexagonal lattice): the fused
    tail runs the base tiling's pass last instead of first, so the moments agree to rounding too."""
    G = 8
    if kind == "cartesian":
        mesh, xs = syn.checkerboard_core(64, 48, 40, num_groups=G)
    else:
        mesh, xs, _ = syn.hex_core(24, 40, pitch=1.0, dz=1.0, num_groups=G, seed=54321)
    quad = syn.level_symmetric(4)
    out = {}
    # (fused?, slabs per column of the un-shear passes: a rank that owns few columns splits them in z, forced here)
    variants = ((True, None), (False, None), (True, "3"), (False, "2"))
    for fused, zsplit in variants:
        if fused:
            monkeypatch.delenv("PAMPA_SN_NO_FUSE", raising=False)
        else:
            monkeypatch.setenv("PAMPA_SN_NO_FUSE", "1")
        if zsplit is None:
            monkeypatch.delenv("PAMPA_SN_UNSHEAR_ZSPLIT", raising=False)
        else:
            monkeypatch.setenv("PAMPA_SN_UNSHEAR_ZSPLIT", zsplit)
        dev = pb.SNDevice(mesh, xs, quad)
        if kind == "hex":
            assert dev.info()["flow_classes"] > 0 and dev.info()["sweep_launches"] <= 4
        k1 = dev.iterate(1)
        p1 = dev.get("flux-moments")
        k5 = dev.iterate(4)
        p5 = dev.get("flux-moments")
        sol = dev.solve_keff(tol_k=1e-10, tol_phi=1e-9)
        dev.close()
        out[(fused, zsplit)] = (k1, p1, k5, p5, sol[0])
    b = out[(False, None)]
    for key in variants:
        a = out[key]
        # same order of the additions into a cell's moments: fused or not on one tiling, and whatever the slabs
        if kind == "cartesian" or key[0] is False:
            assert np.array_equal(a[1], b[1]), key
        else:
            assert util.max_rel(a[1], b[1]) < 1e-13, key
        assert abs(a[0] - b[0]) < 1e-13 * abs(b[0]), key
        assert abs(a[2] - b[2]) < 1e-13 * abs(b[2]) and util.max_rel(a[3], b[3]) < 1e-12, key
        assert abs(a[4] - b[4]) < 1e-9, key


def test_full_size_properties(monkeypatch):
    """BASELINE config 4 at its own size (216^3 cells, S8, 8 groups; the oracle cannot run there): properties that
    do not depend on the size.  The checkerboard core, its vacuum boundaries and the level-symmetric quadrature are
    invariant under x <-> y, x <-> z and the three reflections, so the flux moments of any number of source
    iterations from a flat start must be too (every octant's sweep maps onto another one's); they are positive;
    and the dataflow sweep must reproduce the stream-ordered wavefront sweep bit for bit."""
    n, G = 216, 8
    mesh, xs = syn.checkerboard_core(n, n, n, num_groups=G)
    quad = syn.level_symmetric(8)
    monkeypatch.setenv("PAMPA_SN_NO_FUSE", "1")          # one reduction order for the bit-for-bit comparison below
    dev = pb.SNDevice(mesh, xs, quad)
    assert dev.info()["updates_per_sweep"] == n ** 3 * 80 * G
    k = dev.iterate(2)
    phi = dev.get("flux-moments").reshape(n, n, n, G)       # [z][y][x][g]
    dev.close()
    assert phi.min() > 0.0
    scale = np.abs(phi).max()
    for name, other in (("x<->y", phi.transpose(0, 2, 1, 3)), ("x<->z", phi.transpose(2, 1, 0, 3)),
                        ("-x", phi[:, :, ::-1]), ("-y", phi[:, ::-1]), ("-z", phi[::-1])):
        assert np.abs(phi - other).max() < 1e-12 * scale, name
    dev = pb.SNDevice(mesh, xs, quad, wave_launch=1)
    k2 = dev.iterate(2)
    phi2 = dev.get("flux-moments").reshape(n, n, n, G)
    dev.close()
    assert k2 == k
    assert np.array_equal(phi2, phi)
