"""Write the input decks used by the host-library tests (same grammar as the reference's inputs).

The geometry and data are those of the benchmark problems the reference's own SN tests use
(three-region slab: reflector 20 cm | fuel 100 cm | reflector 15 cm at dx = 0.1 cm; 2-D IAEA PWR
quarter core, 17 x 17 assemblies of 10 cm refined 4 x 4), written by this script from their
definitions -- tests/test_host_cpu.py checks, where the reference tree is mounted, that they give
the same discrete problem as the reference's decks.

    python tests/decks/make_decks.py
"""
import os

HERE = os.path.dirname(os.path.abspath(__file__))

SLAB_MATERIALS = {
    "fuel": dict(st=[0.030, 0.080], nsf=[0.000, 0.135], ss=[[0.0, 0.020], [0.0, 0.0]], chi=[1.0, 0.0], fuel=1),
    "reflector-left": dict(st=[0.040, 0.010], ss=[[0.0, 0.040], [0.0, 0.0]]),
    "reflector-right": dict(st=[0.050, 0.005], ss=[[0.0, 0.050], [0.0, 0.0]]),
}

PWR_PRECURSORS = dict(lam=[0.0124, 0.0305, 0.1110, 0.3010, 1.1400, 3.0100],
                      beta=[0.000215, 0.001424, 0.001274, 0.002568, 0.000748, 0.000273])
PWR_MATERIALS = {
    "fuel1": dict(st=[0.03012, 0.080032], nsf=[0.0, 0.135], ss=[[0.0, 0.020], [0.0, 0.0]], chi=[1.0, 0.0],
                  prec=PWR_PRECURSORS, fuel=1),
    "fuel2": dict(st=[0.03012, 0.085032], nsf=[0.0, 0.135], ss=[[0.0, 0.020], [0.0, 0.0]], chi=[1.0, 0.0],
                  prec=PWR_PRECURSORS, fuel=1),
    "fuel2+rod": dict(st=[0.03012, 0.130032], nsf=[0.0, 0.135], ss=[[0.0, 0.020], [0.0, 0.0]], chi=[1.0, 0.0],
                      prec=PWR_PRECURSORS, fuel=1),
    "reflector": dict(st=[0.04016, 0.010024], ss=[[0.0, 0.040], [0.0, 0.0]]),
}

# IAEA 2-D PWR quarter core, one digit per 10 cm assembly half... (17 x 17 map of 10 cm cells):
# 1 fuel1, 2 fuel2, 3 fuel2+rod, 4 reflector, 0 outside the core
PWR_MAP = [
    "32222223322221144", "22222222222221144", "22222222222221144", "22222222222111144", "22222222222111144",
    "22222222222114444", "22222222222114444", "32222223311114400", "32222223311114400", "22222221111444400",
    "22222221111444400", "22211111144440000", "22211111144440000", "11111444444000000", "11111444444000000",
    "44444440000000000", "44444440000000000",
]


# temperature feedback: the slab fuel tabulated at two temperatures (nuclear-data-set, the format of
# test/hex-core/materials/fuel-2-groups.pmp); absorption grows and fission drops with temperature
SLAB_FUEL_HOT = dict(st=[0.031, 0.086], nsf=[0.000, 0.128], ss=[[0.0, 0.019], [0.0, 0.0]], chi=[1.0, 0.0], fuel=1)
SLAB_FEEDBACK_MATERIALS = dict(SLAB_MATERIALS)
SLAB_FEEDBACK_MATERIALS["fuel"] = dict(SLAB_MATERIALS["fuel"], set=[(300.0, SLAB_MATERIALS["fuel"]), (900.0, SLAB_FUEL_HOT)])


def write_material(path, m):
    G = len(m["st"])
    row = lambda v: " ".join("%.6g" % x for x in v)

    def block(f, t, ind):
        f.write("%snuclear-data {\n%s   energy-groups %d\n%s   sigma-total\n%s   %s\n" % (ind, ind, G, ind, ind, row(t["st"])))
        if "nsf" in t:
            f.write("%s   nu-sigma-fission\n%s   %s\n" % (ind, ind, row(t["nsf"])))
        f.write("%s   sigma-scattering\n" % ind)
        for r in t["ss"]:
            f.write("%s   %s\n" % (ind, row(r)))
        if "chi" in t:
            f.write("%s   fission-spectrum\n%s   %s\n" % (ind, ind, row(t["chi"])))
        f.write("%s}\n" % ind)

    with open(path, "w") as f:
        if "set" in m:
            f.write("nuclear-data-set {\n   temperature %d\n   %s\n" % (len(m["set"]), row([T for T, _ in m["set"]])))
            for _, t in m["set"]:
                block(f, t, "   ")
            f.write("}\n")
        else:
            block(f, m, "")
        if "prec" in m:
            f.write("precursor-data {\n   precursor-groups %d\n   lambda\n   %s\n   beta\n   %s\n}\n" % (
                len(m["prec"]["lam"]), row(m["prec"]["lam"]), row(m["prec"]["beta"])))
        if m.get("fuel"):
            f.write("fuel 1\n")


OUT = HERE


def write_deck(case, mesh_kind, materials, order, ls, extra=""):
    d = os.path.join(OUT, case)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "input.pmp"), "w") as f:
        f.write("# %s\nmesh %s mesh.pmp\n\n" % (case, mesh_kind))
        for name in materials:
            f.write("material %s %s.pmp\n" % (name, name))
        f.write("\nsolver sn {\n   energy-groups 2\n   order %d\n   mixed-face-interpolation 1.0\n"
                "   least-squares-boundary-interpolation %d\n%s}\n" % (order, ls, extra))
    for name, m in materials.items():
        write_material(os.path.join(d, name + ".pmp"), m)
    return d


def slab_mesh(d):
    layout = [2] * 200 + [1] * 1000 + [3] * 150
    with open(os.path.join(d, "mesh.pmp"), "w") as f:
        f.write("dx -%d\n0.100\n\nbc -x vacuum\nbc +x vacuum\n\nmaterials %d\n%s\n" % (
            len(layout), len(layout), " ".join(str(m) for m in layout)))


def pwr_layout(n=4):
    rows = []
    for line in PWR_MAP:
        r = [int(c) for c in line for _ in range(n)]
        rows += [r] * n
    return rows


def pwr_cartesian_mesh(d, n=4):
    rows = pwr_layout(n)
    nx = len(rows)
    h = " ".join(["%.3f" % (10.0 / n)] * nx)
    with open(os.path.join(d, "mesh.pmp"), "w") as f:
        f.write("dx %d\n%s\n\ndy %d\n%s\n\n" % (nx, h, nx, h))
        f.write("bc -x reflective\nbc +x vacuum\nbc -y reflective\nbc +y vacuum\n\n")
        f.write("materials %d\n\n" % (nx * nx))
        for r in rows:
            f.write(" ".join(str(m) for m in r) + "\n")


def pwr_unstructured_mesh(d, n=4):
    rows = pwr_layout(n)
    nx = len(rows)
    h = 10.0 / n
    exterior, interior = [], []
    with open(os.path.join(d, "mesh.pmp"), "w") as f:
        f.write("points %d\n" % ((nx + 1) * (nx + 1)))
        for j in range(nx + 1):
            for i in range(nx + 1):
                f.write("%.3f %.3f\n" % (i * h, j * h))
                p = j * (nx + 1) + i
                if i == 0 or j == 0:
                    interior.append(p)
                around = [rows[jj][ii] for jj in (j - 1, j) for ii in (i - 1, i) if 0 <= jj < nx and 0 <= ii < nx]
                if any(around) and (not all(around) or i == nx or j == nx):
                    exterior.append(p)
        cells = [(i, j) for j in range(nx) for i in range(nx) if rows[j][i]]
        f.write("\ncells %d %d\n" % (len(cells), 4 * len(cells)))
        for i, j in cells:
            f.write("%d %d %d %d\n" % (i + j * (nx + 1), i + 1 + j * (nx + 1), i + 1 + (j + 1) * (nx + 1),
                                       i + (j + 1) * (nx + 1)))
        for name, pts in (("exterior", exterior), ("interior", interior)):
            f.write("\nboundary %s %d\n" % (name, len(pts)))
            f.write("\n".join(str(p) for p in pts) + "\n")
        f.write("\nbc exterior vacuum\nbc interior reflective\n\nmaterials %d\n\n" % len(cells))
        for j in range(nx):
            f.write(" ".join(str(rows[j][i]) for i in range(nx) if rows[j][i]) + "\n")


def main(out=None):
    """Write all decks under `out` (default: next to this script); returns the directory."""
    global OUT
    OUT = out or HERE
    slab_mesh(write_deck("slab_s2", "cartesian", SLAB_MATERIALS, 2, 1))
    slab_mesh(write_deck("slab_s4", "cartesian", SLAB_MATERIALS, 4, 1))
    pwr_cartesian_mesh(write_deck("pwr_cartesian_s2", "cartesian", PWR_MATERIALS, 2, 1))
    pwr_unstructured_mesh(write_deck("pwr_unstructured_s2", "unstructured", PWR_MATERIALS, 2, 1))
    pwr_cartesian_mesh(write_deck("pwr_cartesian_s8_lsoff", "cartesian", PWR_MATERIALS, 8, 0))
    slab_mesh(write_deck("slab_s2_feedback", "cartesian", SLAB_FEEDBACK_MATERIALS, 2, 0))
    return OUT


if __name__ == "__main__":
    import sys
    print(main(sys.argv[1] if len(sys.argv) > 1 else None))
