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
    """A deck that omits `mixed-face-interpolation` runs with the reference's default, 0.1 (src/SNSolver.hxx:16).  On
    the PWR deck that eigenvector has negative angular fluxes, and the reference fails such a solve in
    normalizeAngularFlux (src/SNSolver.cxx:329): same message, non-zero exit.  With 0.9 the solve goes through."""
    import shutil
    exe = os.path.join(ROOT, "pampa_b200", "bin", "pampa")
    for delta, ok in ((None, False), ("0.9", True)):
        case = tmp_path / ("case_%s" % delta)
        shutil.copytree(os.path.join(decks, "pwr_cartesian_s2"), case)
        text = open(case / "input.pmp").read()
        assert "mixed-face-interpolation 1.0" in text
        text = text.replace("   mixed-face-interpolation 1.0\n", "" if delta is None else "   mixed-face-interpolation %s\n" % delta)
        text = text.replace("least-squares-boundary-interpolation 1", "least-squares-boundary-interpolation 0")
        open(case / "input.pmp", "w").write(text)
        r = subprocess.run([exe, "input.pmp"], cwd=case, capture_output=True, text=True)
        if ok:
            assert r.returncode == 0, r.stdout + r.stderr
            assert "Effective multiplication factor: 0.96" in r.stdout
        else:
            assert r.returncode != 0
            assert "negative values in the angular-flux solution" in r.stdout + r.stderr
