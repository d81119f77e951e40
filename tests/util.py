"""Test helpers: convert between the oracle's objects and the device layer's arrays.

The oracle (oracle/pampa_oracle.py) is the checker; these helpers only feed the SAME discrete
problem to both sides."""
from __future__ import annotations

import math
import os

import numpy as np

from oracle import pampa_oracle as orc
from pampa_b200 import problem as pb

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def extruded_from_oracle(mesh: orc.Mesh, delta: float = 1.0) -> pb.ExtrudedMesh:
    """Oracle generic face tables -> extruded SoA (lateral faces of layer 0 + dz); delta < 1 adds the
    deferred-correction face weights of include/pampa_sn.h."""
    dz = mesh.ext.get("dz")
    nz = 1 if dz is None else len(dz)
    nxy = mesh.num_cells // nz
    nf = np.zeros(nxy, dtype=np.int32)
    rows = []
    for i in range(nxy):
        f0, f1 = mesh.face_ptr[i], mesh.face_ptr[i + 1]
        lat = [f for f in range(f0, f1) if abs(mesh.face_normal[f][2]) < 0.5]
        rows.append(lat); nf[i] = len(lat)
    F = int(nf.max())
    nb = np.zeros((nxy, F), dtype=np.int32)
    fx = np.zeros((nxy, F)); fy = np.zeros((nxy, F)); cf = np.ones((nxy, F))
    kout = np.zeros((nxy, F)); kin = np.zeros((nxy, F))
    h0 = 1.0 if dz is None else dz[0]
    for i, lat in enumerate(rows):
        for a, f in enumerate(lat):
            i2 = mesh.face_neighbor[f]
            nb[i, a] = i2
            fx[i, a] = mesh.face_normal[f][0] * mesh.face_area[f] / h0
            fy[i, a] = mesh.face_normal[f][1] * mesh.face_area[f] / h0
            if i2 >= 0:
                r1 = np.linalg.norm(mesh.face_centroid[f] - mesh.centroids[i])
                r2 = np.linalg.norm(mesh.face_centroid[f] - mesh.centroids[i2])
                r12 = np.linalg.norm(mesh.centroids[i] - mesh.centroids[i2])
                cf[i, a] = (r1 + r2) / r12
                kout[i, a] = (1.0 - delta) * r1 / r12
                kin[i, a] = (1.0 - delta) * r2 / r12
    bz = (mesh.boundaries.index("-z") + 1, mesh.boundaries.index("+z") + 1) if dz is not None else (0, 0)
    ij = None
    if mesh.ext.get("kind") == "cartesian":
        jj, ii = np.nonzero(mesh.ext["phys_xy"])
        ij = np.stack([ii, jj], axis=1).astype(np.int32)
    return pb.ExtrudedMesh(xy_num_faces=nf, xy_neighbor=nb, xy_face_fx=fx, xy_face_fy=fy, xy_face_cf=cf,
                           xy_area=mesh.volumes[:nxy] / h0, xy_cx=mesh.centroids[:nxy, 0].copy(),
                           xy_cy=mesh.centroids[:nxy, 1].copy(), materials=mesh.materials.astype(np.int32),
                           bc_types=[0], dz=None if dz is None else np.asarray(dz, dtype=float),
                           bc_minus_z=bz[0], bc_plus_z=bz[1], xy_ij=ij, delta=delta,
                           xy_face_kout=kout if delta < 1.0 else None, xy_face_kin=kin if delta < 1.0 else None)


def with_bcs(em: pb.ExtrudedMesh, mesh: orc.Mesh, bcs) -> pb.ExtrudedMesh:
    t = [0] * (1 + len(mesh.boundaries))
    for b in range(1, len(t)):
        v = bcs[b] if b < len(bcs) else 0
        t[b] = {orc.VACUUM: pb.BC_VACUUM, orc.REFLECTIVE: pb.BC_REFLECTIVE}.get(v, pb.BC_NONE)
    em.bc_types = t
    return em


def ls_from_oracle(mesh: orc.Mesh, mode: str, bcs) -> pb.LSCorrection | None:
    """The lagged LS boundary correction entries (SNSolver.cxx:485-518) for 1-D / 2-D meshes."""
    if mode == "off":
        return None
    cbc = orc.ls_boundary_coefs(mesh, mode)
    cell, ptr, nbr, om, nv = [], [0], [], [], []
    for i in sorted(cbc):
        f0, f1 = mesh.face_ptr[i], mesh.face_ptr[i + 1]
        n0 = len(nbr)
        for f in range(f0, f1):
            i2 = mesh.face_neighbor[f]
            if i2 >= 0 or bcs[-i2] != orc.VACUUM:
                continue
            dp = mesh.face_centroid[f] - mesh.centroids[i]
            for f2 in range(f0, f1):
                i3 = mesh.face_neighbor[f2]
                if i3 < 0:
                    continue
                nbr.append(i3); om.append(float(dp @ cbc[i][f2 - f0]))
                nv.append(mesh.face_normal[f] * mesh.face_area[f] / mesh.volumes[i])
        if len(nbr) > n0:
            cell.append(i); ptr.append(len(nbr))
    if not cell:
        return None
    return pb.LSCorrection(np.array(cell), np.array(ptr), np.array(nbr), np.array(om), np.array(nv))


def xs_from_oracle(xs_list) -> pb.CrossSections:
    return pb.CrossSections(
        np.array([x.sigma_total for x in xs_list]), np.array([x.sigma_scattering for x in xs_list]),
        np.array([x.nu_sigma_fission for x in xs_list]), np.array([x.kappa_sigma_fission for x in xs_list]),
        np.array([x.chi_effective for x in xs_list]), np.zeros(len(xs_list)))


def xs_to_oracle(xs: pb.CrossSections):
    out = []
    for m in range(xs.num_materials):
        x = orc.XS(G=xs.num_groups, sigma_total=xs.sigma_total[m], nu_sigma_fission=xs.nu_sigma_fission[m],
                   kappa_sigma_fission=xs.kappa_sigma_fission[m], sigma_scattering=xs.sigma_scattering[m],
                   chi_prompt=xs.chi_effective[m], chi_delayed=xs.chi_effective[m],
                   chi_effective=xs.chi_effective[m])
        out.append(x)
    return out


def quad_from_oracle(order) -> pb.Quadrature:
    d, w, r = orc.quadrature(order)
    return pb.Quadrature(d, w, r.astype(np.int32))


def quad_to_oracle(q: pb.Quadrature):
    return q.directions, q.weights, q.reflected.astype(np.int64)


def deck_problem(deck: orc.Deck, ls_mode: str):
    """(ExtrudedMesh, CrossSections, Quadrature, LSCorrection) of a parsed reference deck."""
    em = with_bcs(extruded_from_oracle(deck.mesh, deck.delta), deck.mesh, deck.bcs)
    return em, xs_from_oracle(deck.xs), quad_from_oracle(deck.order), ls_from_oracle(deck.mesh, ls_mode, deck.bcs)


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def max_rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))


def load_golden(name):
    """A committed fixture -> (ExtrudedMesh, CrossSections, Quadrature, LSCorrection, npz)."""
    from pampa_b200 import synthetic as syn
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    em = pb.ExtrudedMesh(
        xy_num_faces=z["xy_num_faces"], xy_neighbor=z["xy_neighbor"], xy_face_fx=z["xy_face_fx"],
        xy_face_fy=z["xy_face_fy"], xy_face_cf=z["xy_face_cf"], xy_area=z["xy_area"], xy_cx=z["xy_cx"],
        xy_cy=z["xy_cy"], materials=z["materials"], bc_types=[int(v) for v in z["bc_types"]],
        dz=z["dz"] if len(z["dz"]) else None, bc_minus_z=int(z["bc_z"][0]), bc_plus_z=int(z["bc_z"][1]),
        xy_ij=z["xy_ij"] if len(z["xy_ij"]) else None, delta=float(z["delta"]) if "delta" in z else 1.0,
        xy_face_kout=z["xy_face_kout"] if "xy_face_kout" in z else None,
        xy_face_kin=z["xy_face_kin"] if "xy_face_kin" in z else None)
    xs = pb.CrossSections(z["sigma_total"], z["sigma_scattering"], z["nu_sigma_fission"],
                          z["kappa_sigma_fission"], z["chi_effective"], np.zeros(len(z["sigma_total"])))
    quad = syn.level_symmetric(int(z["order"]))
    ls = None
    if "ls_cell" in z:
        ls = pb.LSCorrection(z["ls_cell"], z["ls_ptr"], z["ls_nbr"], z["ls_omega"], z["ls_nvec"])
    return em, xs, quad, ls, z


def read_vtk(path):
    """Parse a legacy ASCII .vtk file as the reference writes it (src/vtk.cxx): points, ragged cells, cell types
    and the ordered list of (name, values) SCALARS blocks."""
    tok = open(path).read().split("\n")
    assert tok[0] == "# vtk DataFile Version 3.0" and tok[2] == "ASCII" and tok[3] == "DATASET UNSTRUCTURED_GRID"
    i = 4
    out = {"scalars": []}
    while i < len(tok):
        line = tok[i].split()
        i += 1
        if not line:
            continue
        if line[0] == "POINTS":
            n = int(line[1])
            out["points"] = np.array([[float(x) for x in tok[i + a].split()] for a in range(n)])
            i += n
        elif line[0] == "CELLS":
            n = int(line[1])
            rows = [[int(x) for x in tok[i + a].split()] for a in range(n)]
            assert sum(len(r) for r in rows) == int(line[2])
            assert all(r[0] == len(r) - 1 for r in rows)
            out["cells"] = [r[1:] for r in rows]
            i += n
        elif line[0] == "CELL_TYPES":
            n = int(line[1])
            out["types"] = np.array([int(tok[i + a]) for a in range(n)])
            i += n
        elif line[0] == "CELL_DATA":
            out["num_cell_data"] = int(line[1])
        elif line[0] == "SCALARS":
            assert line[2:] == ["double", "1"] and tok[i] == "LOOKUP_TABLE default"
            n = out["num_cell_data"]
            out["scalars"].append((line[1], np.array([float(tok[i + 1 + a]) for a in range(n)])))
            i += 1 + n
        else:
            raise AssertionError("unexpected line in %s: %r" % (path, tok[i - 1]))
    return out
