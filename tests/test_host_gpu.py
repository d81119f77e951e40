"""GPU tests of the drop-in boundary: the C++ host library (Parser -> meshes -> SNSolver -> C-ABI
CUDA layer) run the way the reference's check.sh runs its cases, compared with the reference's
printed goldens (test/check_ref.txt) and with the oracle's fields."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the text check.sh diffs for an SN case (test/check_ref.txt:24-41), k-eff lines :32, :53, :234, :415
EXPECTED = """
Initialize...
Done.

--------------------------------

Solve steady state...

Effective multiplication factor: %s.
Power: 1.000e+00.

Done.

--------------------------------

Finalize...
Done.

"""
CASES = {"slab_s2": ("0.970849", "slabs_s2"), "slab_s4": ("0.982472", "slabs_s4"),
         "pwr_cartesian_s2": ("0.965761", "pwr_cartesian_s2"), "pwr_unstructured_s2": ("0.965761", "pwr_unstructured_s2")}


@pytest.fixture(scope="module")
def decks(tmp_path_factory):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_decks", os.path.join(ROOT, "tests", "decks", "make_decks.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.main(str(tmp_path_factory.mktemp("decks")))


@pytest.mark.parametrize("case", sorted(CASES))
def test_stdout_matches_check_ref(decks, case):
    exe = os.path.join(ROOT, "pampa_b200", "bin", "pampa")
    r = subprocess.run([exe, "input.pmp"], cwd=os.path.join(decks, case), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout == EXPECTED % CASES[case][0], r.stdout


def test_fields_through_the_c_api(decks):
    """pampa_initialize / solve / get_field / finalize from a host code (ctypes plays the C driver)."""
    lib = ctypes.CDLL(os.path.join(ROOT, "pampa_b200", "lib", "libpampa.so"))
    lib.pampa_get_field_size.restype = ctypes.c_long
    lib.pampa_get_keff.restype = ctypes.c_double
    err = ctypes.c_int(0)
    argv = (ctypes.c_char_p * 3)(b"pampa", b"input.pmp", b"-silent")
    cwd = os.getcwd()
    os.chdir(os.path.join(decks, "pwr_cartesian_s2"))
    try:
        lib.pampa_initialize_steady_state(3, argv, ctypes.byref(err)); assert err.value == 0
        lib.pampa_solve_steady_state(ctypes.byref(err)); assert err.value == 0
        z = np.load(os.path.join(util.GOLDEN, "pwr_cartesian_s2.npz"))
        k = lib.pampa_get_keff(ctypes.byref(err))
        assert abs(k - float(z["keff"])) < 1e-5
        for name, want in (("scalar-flux", z["phi"].reshape(-1)), ("power", z["power"]),
                           ("production-rate", z["production"])):
            n = lib.pampa_get_field_size(name.encode(), ctypes.byref(err))
            assert n == want.size
            buf = np.zeros(n)
            lib.pampa_get_field(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), name.encode(), ctypes.byref(err))
            assert err.value == 0
            assert util.rel_l2(buf, want) < 1e-5, name
        n = lib.pampa_get_field_size(b"angular-flux", ctypes.byref(err))
        assert n == z["phi"].size * 8
        psi = np.zeros(n)
        lib.pampa_get_field(psi.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), b"angular-flux", ctypes.byref(err))
        assert err.value == 0 and psi.min() >= 0.0
        lib.pampa_get_field(psi.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), b"no-such-field", ctypes.byref(err))
        assert err.value == 1
        lib.pampa_finalize_steady_state(ctypes.byref(err)); assert err.value == 0
    finally:
        os.chdir(cwd)


def test_vtk_output(decks, tmp_path):
    """`vtk 1` in the main input: output_0.vtk = the mesh followed by flux_<g>, flux_<g>_<m> and power blocks in the
    reference's order (src/PhysicsSolver.cxx:19-28, src/SNSolver.cxx:754-770, src/vtk.cxx:123-170), holding the same
    numbers pampa_get_field returns (to the 7 digits the file carries)."""
    import shutil
    case = tmp_path / "case"
    shutil.copytree(os.path.join(decks, "pwr_cartesian_s2"), case)
    with open(case / "input.pmp", "a") as f:
        f.write("\nvtk 1\n")
    exe = os.path.join(ROOT, "pampa_b200", "bin", "pampa")
    r = subprocess.run([exe, "input.pmp"], cwd=case, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    v = util.read_vtk(str(case / "output_0.vtk"))
    z = np.load(os.path.join(util.GOLDEN, "pwr_cartesian_s2.npz"))
    N, G = z["phi"].shape
    M = 8
    names = [n for n, _ in v["scalars"]]
    assert names == (["materials"] + ["flux_%d" % (g + 1) for g in range(G)] +
                     ["flux_%d_%d" % (g + 1, m + 1) for g in range(G) for m in range(M)] + ["power"])
    assert len(v["cells"]) == N
    vals = dict(v["scalars"])
    for g in range(G):
        assert util.rel_l2(vals["flux_%d" % (g + 1)], z["phi"][:, g]) < 1e-5
    assert util.rel_l2(vals["power"], z["power"]) < 1e-5
    # after the two normalisations (phi = 4 pi sum w psi scaled to the power, src/SNSolver.cxx:288 and
    # src/NeutronicSolver.cxx:63; psi scaled to the power without the 4 pi, src/SNSolver.cxx:324) the fields
    # satisfy sum_m w_m psi_m = phi; level-symmetric S2 has equal weights 1/8
    for g in range(G):
        s = sum(vals["flux_%d_%d" % (g + 1, m + 1)] for m in range(M)) / M
        assert util.rel_l2(s, vals["flux_%d" % (g + 1)]) < 1e-5
    # without the switch nothing is written
    assert not os.path.exists(os.path.join(decks, "pwr_cartesian_s2", "output_0.vtk"))


def test_default_face_interpolation(decks, tmp_path):
    """A deck that omits `mixed-face-interpolation` runs with the reference's default, 0.1 (src/SNSolver.hxx:16).
    (1) The slab problem between two reflective boundaries: the solve goes through and prints the k of the oracle's
    delta = 0.1 operator.  (2) With a vacuum boundary the linear face interpolation undershoots next to it: the
    eigenvector has negative angular fluxes, which the reference reports as an error in normalizeAngularFlux
    (src/SNSolver.cxx:329) -- same message, non-zero exit (PWR deck, committed oracle fixture: psi_min < 0)."""
    import shutil
    from oracle import pampa_oracle as orc
    exe = os.path.join(ROOT, "pampa_b200", "bin", "pampa")

    def variant(src, name, edit_mesh=None):
        case = tmp_path / name
        shutil.copytree(os.path.join(decks, src), case)
        text = open(case / "input.pmp").read()
        assert "   mixed-face-interpolation 1.0\n" in text
        text = text.replace("   mixed-face-interpolation 1.0\n", "")
        text = text.replace("least-squares-boundary-interpolation 1", "least-squares-boundary-interpolation 0")
        open(case / "input.pmp", "w").write(text)
        if edit_mesh:
            mesh_text = edit_mesh(open(case / "mesh.pmp").read())
            open(case / "mesh.pmp", "w").write(mesh_text)
        return case

    case = variant("slab_s2", "slab_reflected", lambda m: m.replace("bc -x vacuum", "bc -x reflective").replace("bc +x vacuum", "bc +x reflective"))
    deck = orc.read_deck(str(case / "input.pmp"))
    assert deck.delta == 0.1 and deck.bcs[1:] == [orc.REFLECTIVE, orc.REFLECTIVE]
    sol = orc.solve_monolithic(orc.build_operator(deck.mesh, deck.xs, deck.G, deck.order, deck.delta, "off", deck.bcs))
    assert sol.psi.min() > 0.0
    r = subprocess.run([exe, "input.pmp"], cwd=case, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout == EXPECTED % ("%.6f" % sol.keff), r.stdout

    z = np.load(os.path.join(util.GOLDEN, "pwr_cartesian_s2_delta01_lsoff.npz"))
    assert float(z["psi_min"]) < 0.0
    case = variant("pwr_cartesian_s2", "pwr_default_delta")
    r = subprocess.run([exe, "input.pmp"], cwd=case, capture_output=True, text=True)
    assert r.returncode != 0
    assert "negative values in the angular-flux solution" in r.stdout + r.stderr


def test_temperature_feedback(decks):
    """The CouplingSolver-facing half of the path: pampa_set_field("temperature") re-tabulates the cross sections per
    cell (linear interpolation between the tabulated temperatures, src/FeedbackNuclearData.hxx:62-140) for the next
    solve.  Slab deck whose fuel is a nuclear-data-set at 300 K / 900 K; three temperature fields in a row through
    the C API -- uniform cold, a profile with 7 distinct fuel temperatures (the number of (material, temperature)
    rows on the device grows from 3 to 9), uniform hot (shrinks back) -- each against the oracle's eigenpair of the
    same per-cell data."""
    from oracle import pampa_oracle as orc
    case = os.path.join(decks, "slab_s2_feedback")
    deck = orc.read_deck(os.path.join(case, "input.pmp"))
    names = ["reflector-left", "fuel", "reflector-right"]      # material order of the deck (ids 2 | 1 | 3 in the mesh)
    order_in_deck = [l.split()[1] for l in open(os.path.join(case, "input.pmp")) if l.startswith("material ")]
    tabs = {n: orc.read_material_tables(os.path.join(case, n + ".pmp")) for n in names}
    mesh = deck.mesh
    N = mesh.num_cells
    x = mesh.centroids[:, 0]
    fuel = np.array([order_in_deck[m] == "fuel" for m in mesh.materials])
    assert fuel.sum() == 1000

    def oracle_solve(T):
        xs, ids, cell = [], {}, np.zeros(N, dtype=np.int64)
        for i in range(N):
            key = (int(mesh.materials[i]), float(T[i]))
            if key not in ids:
                ids[key] = len(xs)
                t, tb = tabs[order_in_deck[key[0]]]
                xs.append(orc.xs_at_temperature(t, tb, key[1]))
            cell[i] = ids[key]
        m2 = orc.Mesh(**{**mesh.__dict__, "materials": cell})
        op = orc.build_operator(m2, xs, deck.G, deck.order, 1.0, "off", deck.bcs)
        return orc.solve_monolithic(op)

    lib = ctypes.CDLL(os.path.join(ROOT, "pampa_b200", "lib", "libpampa.so"))
    lib.pampa_get_keff.restype = ctypes.c_double
    err = ctypes.c_int(0)
    argv = (ctypes.c_char_p * 3)(b"pampa", b"input.pmp", b"-silent")
    pd = ctypes.POINTER(ctypes.c_double)
    cwd = os.getcwd()
    os.chdir(case)
    try:
        lib.pampa_initialize_steady_state(3, argv, ctypes.byref(err)); assert err.value == 0
        profile = np.where(fuel, 300.0 + 100.0 * np.floor((x - 20.0) / 100.0 * 7.0).clip(0, 6), 0.0)
        ks = []
        for T in (np.where(fuel, 300.0, 0.0), profile, np.full(N, 900.0)):
            T = np.ascontiguousarray(T, dtype=np.float64)
            lib.pampa_set_field(T.ctypes.data_as(pd), b"temperature", ctypes.byref(err)); assert err.value == 0
            lib.pampa_solve_steady_state(ctypes.byref(err)); assert err.value == 0
            k = lib.pampa_get_keff(ctypes.byref(err))
            phi = np.zeros(N * deck.G)
            lib.pampa_get_field(phi.ctypes.data_as(pd), b"scalar-flux", ctypes.byref(err)); assert err.value == 0
            sol = oracle_solve(T)
            assert abs(k - sol.keff) < 1e-5, (k, sol.keff)
            assert util.rel_l2(phi.reshape(N, deck.G), sol.phi) < 1e-5
            assert util.max_rel(phi.reshape(N, deck.G), sol.phi) < 1e-4
            ks.append(k)
        assert ks[0] > ks[1] > ks[2]                     # hotter fuel: more absorption, less fission
        assert abs(ks[0] - 0.9708) < 5e-3                # cold = the reference's slab problem without the LS term
        lib.pampa_finalize_steady_state(ctypes.byref(err)); assert err.value == 0
    finally:
        os.chdir(cwd)


@pytest.mark.parametrize("case", ["slab_s2", "pwr_cartesian_s2", "pwr_unstructured_s2"])
def test_partitioned_mesh_solve(decks, case, tmp_path):
    """`mesh partitioned <file>` on the whole-domain dump of a mesh (Mesh::writeData): the extruded structure is
    recovered from the face tables and the solve prints the reference's golden for the original deck."""
    import shutil
    lib = ctypes.CDLL(os.path.join(ROOT, "pampa_b200", "lib", "libpampa.so"))
    lib.pampa_debug_write_mesh_data.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    work = tmp_path / case
    shutil.copytree(os.path.join(decks, case), work)
    cwd = os.getcwd()
    os.chdir(work)
    try:
        assert lib.pampa_debug_write_mesh_data(b"input.pmp", b"mesh_data.pmp", 17) == 0
    finally:
        os.chdir(cwd)
    text = open(work / "input.pmp").read()
    kind = "cartesian" if "mesh cartesian" in text else "unstructured"
    open(work / "input.pmp", "w").write(text.replace("mesh %s mesh.pmp" % kind, "mesh partitioned mesh_data.pmp"))
    exe = os.path.join(ROOT, "pampa_b200", "bin", "pampa")
    r = subprocess.run([exe, "input.pmp"], cwd=work, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout == EXPECTED % CASES[case][0], r.stdout
