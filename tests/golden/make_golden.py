"""Generate the golden fixtures under tests/golden/ (run in the build container, where the
reference tree is mounted at /root/reference; the GPU box only sees the committed .npz files).

For each SN case the reference ships (check.sh / test/check_ref.txt) this parses the reference's
own input deck with the oracle, solves the reference's discrete eigenproblem, checks k-eff against
the 6-decimal golden printed in test/check_ref.txt, and stores the problem arrays (exactly what
the device layer is fed) together with the oracle solution.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pampa_oracle as orc   # noqa: E402
import util                               # noqa: E402

REF = "/root/reference/test"
# case -> (deck, check_ref.txt line, golden k-eff, LS mode that reproduces it, order override)
CASES = {
    "slabs_s2": ("slabs/reflected-s2/input.pmp", 32, 0.970849, "literal_zero_init", None),
    "slabs_s4": ("slabs/reflected-s4/input.pmp", 53, 0.982472, "literal_zero_init", None),
    "pwr_cartesian_s2": ("pwr-iaea-benchmark/cartesian-sn/input.pmp", 234, 0.965761, "reference_effective", None),
    "pwr_unstructured_s2": ("pwr-iaea-benchmark/unstructured-sn/input.pmp", 415, 0.965761, "reference_effective", None),
    # same decks, variants without a reference-run golden (oracle is the sole authority)
    "pwr_cartesian_s2_lsoff": ("pwr-iaea-benchmark/cartesian-sn/input.pmp", None, None, "off", None),
    "pwr_cartesian_s8_lsoff": ("pwr-iaea-benchmark/cartesian-sn/input.pmp", None, None, "off", 8),
    # BASELINE config 2 as named: the shipped deck (LS boundary interpolation on) at S8
    "pwr_cartesian_s8": ("pwr-iaea-benchmark/cartesian-sn/input.pmp", None, None, "reference_effective", 8),
}
# the reference's default mixed-face-interpolation (0.1, src/SNSolver.hxx:16): what a deck that omits the keyword
# runs with.  The eigenvector may have negative angular fluxes, which the reference reports as an error
# (src/SNSolver.cxx:329): the fixture records the minimum.
DELTA = {"pwr_cartesian_s2_delta01_lsoff": 0.1}
CASES["pwr_cartesian_s2_delta01_lsoff"] = ("pwr-iaea-benchmark/cartesian-sn/input.pmp", None, None, "off", None)


def hex_core_deck(groups, cells="hex-cells"):
    """BASELINE config 3: the reference ships no SN input for hex-core, so this authors one on its
    hex-cells mesh (test/hex-core/hex-cells-diffusion-3d/mesh.pmp: 163 hexagons x 16 layers) with the
    materials of that directory's input.pmp:5-10, S8, vacuum on every boundary, LS off."""
    base = os.path.join(REF, "hex-core")
    mesh = orc.read_unstructured_mesh(os.path.join(base, cells + "-diffusion-3d", "mesh.pmp"))
    names = ["fuel", "fuel", "fuel", "fuel", "reflector", "reflector"]
    xs = [orc.read_material(os.path.join(base, "materials", "%s-%d-groups.pmp" % (n, groups))) for n in names]
    bcs = [0] * (1 + len(mesh.boundaries))
    for b in ("exterior", "-z", "+z"):
        bcs[mesh.boundaries.index(b) + 1] = orc.VACUUM
    return orc.Deck(mesh=mesh, xs=xs, G=groups, order=8, delta=1.0, ls=False, power=1.0, bcs=bcs)


def main(only=None):
    ref_lines = open(os.path.join(REF, "check_ref.txt")).read().split("\n")
    cases = dict(CASES)
    cases["hex_core_s8_2g"] = ("hex-core:2", None, None, "off", None)
    # BASELINE config 3 with the reference's own 11-group graphite-moderated data (upscatter, c ~ 0.95)
    cases["hex_core_s8_11g"] = ("hex-core:11", None, None, "off", None)
    # the same core with every hexagon cut into six triangles (test/hex-core/tri-cells-diffusion-3d/mesh.pmp:
    # 978 triangles x 16 layers), S4 to keep the fixture small
    cases["hex_core_tri_s4_2g"] = ("hex-core:2:tri-cells", None, None, "off", 4)
    for name, (deck_path, line, gold, ls_mode, order) in cases.items():
        if only and name not in only:
            continue
        deck = (hex_core_deck(int(deck_path.split(":")[1]), *deck_path.split(":")[2:]) if deck_path.startswith("hex-core")
                else orc.read_deck(os.path.join(REF, deck_path)))
        if order is not None:
            deck.order = order
        if name in DELTA:
            deck.delta = DELTA[name]
        op = orc.build_operator(deck.mesh, deck.xs, deck.G, deck.order, deck.delta, ls_mode, deck.bcs)
        big = op.N * (op.G * op.M) ** 2 > 4e7
        sol = (orc.solve_matrix_free if big else orc.solve_monolithic)(op, deck.power, allow_negative=name in DELTA)
        if gold is not None:
            printed = ref_lines[line - 1].strip()
            assert printed == "Effective multiplication factor: %.6f." % gold, printed
            assert "%.6f" % sol.keff == "%.6f" % gold, (name, sol.keff, gold)
        em, xs, quad, ls = util.deck_problem(deck, ls_mode)
        out = dict(
            keff=sol.keff, phi=sol.phi, power=sol.power, production=sol.production,
            golden_keff=np.nan if gold is None else gold, ls_mode=ls_mode, order=deck.order, G=deck.G,
            delta=deck.delta, psi_min=sol.psi.min(), psi_max=sol.psi.max(),
            xy_num_faces=em.xy_num_faces, xy_neighbor=em.xy_neighbor, xy_face_fx=em.xy_face_fx,
            xy_face_fy=em.xy_face_fy, xy_face_cf=em.xy_face_cf, xy_area=em.xy_area, xy_cx=em.xy_cx,
            xy_cy=em.xy_cy, materials=em.materials, bc_types=np.array(em.bc_types),
            dz=np.zeros(0) if em.dz is None else em.dz, bc_z=np.array([em.bc_minus_z, em.bc_plus_z]),
            xy_ij=np.zeros((0, 2), dtype=np.int32) if em.xy_ij is None else em.xy_ij,
            sigma_total=xs.sigma_total, sigma_scattering=xs.sigma_scattering,
            nu_sigma_fission=xs.nu_sigma_fission, kappa_sigma_fission=xs.kappa_sigma_fission,
            chi_effective=xs.chi_effective)
        if em.xy_face_kout is not None:
            out.update(xy_face_kout=em.xy_face_kout, xy_face_kin=em.xy_face_kin)
        if ls is not None:
            out.update(ls_cell=ls.cell, ls_ptr=ls.ptr, ls_nbr=ls.nbr, ls_omega=ls.omega, ls_nvec=ls.nvec)
        if op.N * op.G * op.M < 100000:
            out["psi"] = sol.psi
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("%-28s keff %.9f  golden %s  N=%d G=%d M=%d" % (name, sol.keff, gold, op.N, op.G, op.M))


if __name__ == "__main__":
    main(sys.argv[1:] or None)
