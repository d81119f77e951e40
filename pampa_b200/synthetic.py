"""Mesh, quadrature and cross-section builders (numpy) for the benchmark and the parity tests.

These produce the arrays of include/pampa_sn.h directly, following the geometry conventions of
the reference meshes (face order and boundary numbering of src/CartesianMesh.cxx:146-414 and
src/UnstructuredExtrudedMesh.cxx:153-364) so that the same description fed to the oracle gives
the same discrete problem.  Independent of the oracle: nothing here imports it.
"""
from __future__ import annotations

import math

import numpy as np

from .problem import BC_REFLECTIVE, BC_VACUUM, CrossSections, ExtrudedMesh, Quadrature

# ------------------------------------------------------------------------------ quadrature
# first-octant level-symmetric tables: (mu values, index triplets, point weights summing to 1).
# S2..S8 carry the 7-digit constants of src/AngularQuadratureSet.cxx:15-140; S12 is the standard LQ12
# table (the reference stops at S8, src/AngularQuadratureSet.cxx:153), 7 digits and used as they are,
# like the others; its constants are pinned by the level-symmetric defining equations (even moments
# 2..12 exact to 6e-8, tests/test_oracle.py::test_quadrature_tables).
_LQ = {
    2: ([1.0 / math.sqrt(3.0)], [(0, 0, 0)], [1.0]),
    4: ([0.3500212, 0.8688903], [(0, 0, 1), (0, 1, 0), (1, 0, 0)], [1.0 / 3.0] * 3),
    6: ([0.2666355, 0.6815076, 0.9261808],
        [(0, 0, 2), (0, 2, 0), (2, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0)],
        [0.1761263] * 3 + [0.1572071] * 3),
    8: ([0.2182179, 0.5773503, 0.7867958, 0.9511897],
        [(0, 0, 3), (0, 3, 0), (3, 0, 0), (0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1),
         (2, 1, 0), (1, 1, 1)],
        [0.1209877] * 3 + [0.0907407] * 6 + [0.0925926]),
}


def _lqn_general(order, mu, wclass):
    """All (i,j,k) with i+j+k = order/2 - 1; weight class by the sorted index triplet."""
    n = order // 2
    idx, w = [], []
    for i in range(n):
        for j in range(n - i):
            k = n - 1 - i - j
            idx.append((i, j, k))
            w.append(wclass[tuple(sorted((i, j, k)))])
    return mu, idx, w


_LQ[12] = _lqn_general(
    12, [0.1672126, 0.4595476, 0.6280191, 0.7600210, 0.8722706, 0.9716377],
    {(0, 0, 5): 0.0707626, (0, 1, 4): 0.0558811, (0, 2, 3): 0.0373377, (1, 1, 3): 0.0502819,
     (1, 2, 2): 0.0258513})


def level_symmetric(order: int) -> Quadrature:
    """Directions over the 8 octants (octant o flips x if o&1, y if o&2, z if o&4), weights
    normalised to sum 1, and the reflection map (src/AngularQuadratureSet.cxx:161-205)."""
    if order not in _LQ:
        raise ValueError("SN order not implemented")
    mu, idx, w = _LQ[order]
    per = len(idx)
    M = 8 * per
    d = np.zeros((M, 3)); wt = np.zeros(M)
    for o in range(8):
        for m in range(per):
            v = [mu[idx[m][0]], mu[idx[m][1]], mu[idx[m][2]]]
            if o & 1: v[0] = -v[0]
            if o & 2: v[1] = -v[1]
            if o & 4: v[2] = -v[2]
            d[o * per + m] = v
            wt[o * per + m] = w[m] / 8.0
    refl = np.zeros((M, 3), dtype=np.int32)
    for o in range(8):
        for m in range(per):
            for ax in range(3):
                refl[o * per + m, ax] = (o ^ (1 << ax)) * per + m
    return Quadrature(d, wt, refl)


# ------------------------------------------------------------------------------ Cartesian
CART_BOUNDARIES = ["-x", "+x", "-y", "+y", "-z", "+z"]


def cartesian_mesh(dx, dy=None, dz=None, materials=None, bcs=None, delta=1.0) -> ExtrudedMesh:
    """Rectilinear mesh; `materials` is [nz,ny,nx] 0-based with -1 = void (the same xy pattern
    in every layer); `bcs` maps "-x".."+z" to BC_VACUUM / BC_REFLECTIVE (default vacuum); `delta` < 1:
    mixed-face-interpolation weights (src/SNSolver.cxx:193-198) for the deferred correction."""
    dx = np.asarray(dx, dtype=float)
    nx = len(dx)
    ny = 0 if dy is None else len(dy)
    nz = 0 if dz is None else len(dz)
    nyy, nzz = max(ny, 1), max(nz, 1)
    dyv = np.asarray(dy, dtype=float) if ny else np.ones(1)
    mats = np.asarray(materials).reshape(nzz, nyy, nx)
    phys = mats[0] != -1
    if not np.all((mats != -1) == phys[None]):
        raise ValueError("wrong material definition")
    names = ["-x", "+x"] + (["-y", "+y"] if ny else []) + (["-z", "+z"] if nz else [])
    bidx = {n: i + 1 for i, n in enumerate(names)}            # 1-based boundary index
    bc_types = [0] + [int((bcs or {}).get(n, BC_VACUUM)) for n in names]
    nxy = int(phys.sum())
    cid = np.full((nyy, nx), -1, dtype=np.int64)
    cid[phys] = np.arange(nxy)
    jj, ii = np.nonzero(phys)

    def nbr(dj, di, name):
        j2, i2 = jj + dj, ii + di
        ok = (j2 >= 0) & (j2 < nyy) & (i2 >= 0) & (i2 < nx)
        out = np.full(nxy, -bidx[name], dtype=np.int64)
        sel = ok.copy()
        sel[ok] = phys[j2[ok], i2[ok]]
        out[sel] = cid[j2[sel], i2[sel]]
        return out

    if ny:       # face order -y, +x, +y, -x
        F = 4
        nb = np.stack([nbr(-1, 0, "-y"), nbr(0, 1, "+x"), nbr(1, 0, "+y"), nbr(0, -1, "-x")], axis=1)
        z = np.zeros(nxy)
        fx = np.stack([z, dyv[jj], z, -dyv[jj]], axis=1)
        fy = np.stack([-dx[ii], z, dx[ii], z], axis=1)
        area = dx[ii] * dyv[jj]
    else:        # 1-D: +x, -x with unit face area
        F = 2
        nb = np.stack([nbr(0, 1, "+x"), nbr(0, -1, "-x")], axis=1)
        fx = np.stack([np.ones(nxy), -np.ones(nxy)], axis=1)
        fy = np.zeros((nxy, 2))
        area = dx[ii].copy()
    x0 = np.concatenate([[0.0], np.cumsum(dx)])
    y0 = np.concatenate([[0.0], np.cumsum(dyv)]) if ny else np.zeros(2)
    cx = x0[ii] + 0.5 * dx[ii]
    cy = (y0[jj] + 0.5 * dyv[jj]) if ny else np.zeros(nxy)
    kout = kin = None
    if delta < 1.0:
        # r_if = half the cell width across the face, r_i2f = half the neighbour's, r_ii2 = their sum
        hx, hy = dx[ii], dyv[jj]
        mine = np.stack([hy, hx, hy, hx], axis=1) if ny else np.stack([hx, hx], axis=1)
        theirs = np.zeros_like(mine)
        ok = nb >= 0
        nbc = np.where(ok, nb, 0)
        theirs = np.stack([hy[nbc[:, 0]], hx[nbc[:, 1]], hy[nbc[:, 2]], hx[nbc[:, 3]]], axis=1) if ny \
            else np.stack([hx[nbc[:, 0]], hx[nbc[:, 1]]], axis=1)
        kout = np.where(ok, (1.0 - delta) * mine / (mine + theirs), 0.0)
        kin = np.where(ok, (1.0 - delta) * theirs / (mine + theirs), 0.0)
    return ExtrudedMesh(
        xy_num_faces=np.full(nxy, F, dtype=np.int32), xy_neighbor=nb.astype(np.int32),
        xy_face_fx=fx, xy_face_fy=fy, xy_face_cf=np.ones((nxy, F)), xy_area=area, xy_cx=cx, xy_cy=cy,
        materials=mats[:, phys].reshape(-1).astype(np.int32), bc_types=bc_types,
        dz=np.asarray(dz, dtype=float) if nz else None,
        bc_minus_z=bidx.get("-z", 0), bc_plus_z=bidx.get("+z", 0),
        xy_ij=np.stack([ii, jj], axis=1).astype(np.int32), delta=delta, xy_face_kout=kout, xy_face_kin=kin)


# ------------------------------------------------------------------------------ polygons
def polygon_mesh(points, cells, dz=None, materials=None, boundary_points=None, bc_of_boundary=None,
                 bc_z=(BC_VACUUM, BC_VACUUM), delta=1.0) -> ExtrudedMesh:
    """Extruded 2-D polygon mesh (CCW point lists).  `boundary_points` maps a boundary name to
    the set of points on it; an edge whose two points lie on a boundary gets that boundary,
    other unmatched edges get the default boundary ("exterior", listed first)."""
    pts = np.asarray(points, dtype=float)
    nxy = len(cells)
    nz = 0 if dz is None else len(dz)
    names = (["-z", "+z"] if nz else []) + list((boundary_points or {"exterior": None}).keys())
    bc_of_boundary = bc_of_boundary or {}
    bc_types = [0]
    for n in names:
        if n == "-z": bc_types.append(int(bc_z[0]))
        elif n == "+z": bc_types.append(int(bc_z[1]))
        else: bc_types.append(int(bc_of_boundary.get(n, BC_VACUUM)))
    F = max(len(c) for c in cells)
    edge_owner = {}
    for i, c in enumerate(cells):
        n = len(c)
        for f in range(n):
            edge_owner.setdefault((min(c[f], c[(f + 1) % n]), max(c[f], c[(f + 1) % n])), []).append(i)
    bsets = {n: set(p) for n, p in (boundary_points or {}).items() if p is not None}
    default_b = next((n for n, p in (boundary_points or {"exterior": None}).items() if p is None), None)
    nb = np.full((nxy, F), 0, dtype=np.int32)
    fx = np.zeros((nxy, F)); fy = np.zeros((nxy, F)); cf = np.ones((nxy, F))
    kout = np.zeros((nxy, F)); kin = np.zeros((nxy, F))
    fcx = np.zeros((nxy, F)); fcy = np.zeros((nxy, F))
    area = np.zeros(nxy); cx = np.zeros(nxy); cy = np.zeros(nxy)
    nf = np.zeros(nxy, dtype=np.int32)
    for i, c in enumerate(cells):
        n = len(c); nf[i] = n
        p = pts[list(c)]
        q = np.roll(p, -1, axis=0)
        da = p[:, 0] * q[:, 1] - q[:, 0] * p[:, 1]
        a = 0.5 * da.sum()
        area[i] = a
        cx[i] = ((p[:, 0] + q[:, 0]) * da).sum() / (6.0 * a)
        cy[i] = ((p[:, 1] + q[:, 1]) * da).sum() / (6.0 * a)
        fx[i, :n] = q[:, 1] - p[:, 1]           # (dy, -dx): outward normal x edge length
        fy[i, :n] = p[:, 0] - q[:, 0]
        fcx[i, :n] = 0.5 * (p[:, 0] + q[:, 0]); fcy[i, :n] = 0.5 * (p[:, 1] + q[:, 1])
        for f in range(n):
            e = (min(c[f], c[(f + 1) % n]), max(c[f], c[(f + 1) % n]))
            other = [o for o in edge_owner[e] if o != i]
            if other:
                nb[i, f] = other[0]
            else:
                name = next((bn for bn, s in bsets.items() if e[0] in s and e[1] in s), default_b)
                if name is None:
                    raise ValueError("wrong mesh connectivity")
                nb[i, f] = -(names.index(name) + 1)
    for i in range(nxy):                        # upwind face weights (src/SNSolver.cxx:193-198)
        for f in range(nf[i]):
            j = nb[i, f]
            if j >= 0:
                r1 = math.hypot(fcx[i, f] - cx[i], fcy[i, f] - cy[i])
                r2 = math.hypot(fcx[i, f] - cx[j], fcy[i, f] - cy[j])
                r12 = math.hypot(cx[i] - cx[j], cy[i] - cy[j])
                cf[i, f] = (r1 + r2) / r12
                kout[i, f] = (1.0 - delta) * r1 / r12
                kin[i, f] = (1.0 - delta) * r2 / r12
    mats = np.zeros(nxy * max(nz, 1), dtype=np.int32) if materials is None else np.asarray(materials, dtype=np.int32)
    return ExtrudedMesh(xy_num_faces=nf, xy_neighbor=nb, xy_face_fx=fx, xy_face_fy=fy, xy_face_cf=cf,
                        xy_area=area, xy_cx=cx, xy_cy=cy, materials=mats, bc_types=bc_types,
                        dz=np.asarray(dz, dtype=float) if nz else None,
                        bc_minus_z=names.index("-z") + 1 if nz else 0,
                        bc_plus_z=names.index("+z") + 1 if nz else 0, delta=delta,
                        xy_face_kout=kout if delta < 1.0 else None, xy_face_kin=kin if delta < 1.0 else None)


def hex_lattice(nrings: int, pitch: float = 1.0):
    """Points and CCW cells of a hexagonal core: regular hexagons (flat-to-flat `pitch`) filling
    a hexagon of `nrings` rings around the central cell.  Returns (points, cells, ring_of_cell)."""
    r = pitch / math.sqrt(3.0)                   # circumradius
    pid, points, cells, ring = {}, [], [], []
    for qa in range(-nrings, nrings + 1):
        for ra in range(max(-nrings, -qa - nrings), min(nrings, -qa + nrings) + 1):
            cx0 = pitch * (qa + 0.5 * ra)
            cy0 = pitch * (math.sqrt(3.0) / 2.0) * ra
            c = []
            for v in range(6):
                ang = math.pi / 6.0 + v * math.pi / 3.0          # pointy-top, CCW
                key = (round((cx0 + r * math.cos(ang)) / pitch * 1e6), round((cy0 + r * math.sin(ang)) / pitch * 1e6))
                if key not in pid:
                    pid[key] = len(points)
                    points.append((cx0 + r * math.cos(ang), cy0 + r * math.sin(ang)))
                c.append(pid[key])
            cells.append(c)
            ring.append(max(abs(qa), abs(ra), abs(-qa - ra)))
    return np.array(points), cells, np.array(ring)


# ------------------------------------------------------------------------------ cross sections
def synthetic_xs(num_groups: int, seed: int = 12345) -> CrossSections:
    """Two materials (0 fuel, 1 moderator) per SURVEY.md section 8(d): sigma_t ~ U(0.2,1),
    scattering ratio c ~ U(0.5,0.9) mostly down-scatter, weak up-scatter in the last 3 groups,
    nu-sigma-f ~ U(0,0.05) in the fuel, decaying fission spectrum in groups 0-3."""
    rng = np.random.default_rng(seed)
    G = num_groups
    st = rng.uniform(0.2, 1.0, size=(2, G))
    ss = np.zeros((2, G, G))
    for m in range(2):
        for g in range(G):
            c = rng.uniform(0.5, 0.9)
            prof = np.zeros(G)
            for g2 in range(g, G):
                prof[g2] = 0.5 ** (g2 - g)
            for g2 in range(max(0, g - 2), g):
                if g >= G - 3:
                    prof[g2] = 1.0e-2 * rng.uniform(0.1, 1.0)
            ss[m, g] = c * st[m, g] * prof / prof.sum()
    nusf = np.zeros((2, G))
    nusf[0] = rng.uniform(0.0, 0.05, size=G)
    chi = np.zeros((2, G))
    ng = min(4, G)
    spec = 0.5 ** np.arange(ng)
    chi[0, :ng] = spec / spec.sum()
    return CrossSections(st, ss, nusf, nusf * (3.2e-11 / 2.4355), chi, np.zeros(2))


def checkerboard_core(nx, ny, nz, h=1.0, assembly=8, num_groups=8, seed=12345, bcs=None):
    """Synthetic 3-D Cartesian core: `assembly`^3-cell blocks alternating fuel / moderator."""
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    mats = ((i // assembly + j // assembly + k // assembly) % 2).astype(np.int32)
    mesh = cartesian_mesh(np.full(nx, h), np.full(ny, h), np.full(nz, h), mats, bcs)
    return mesh, synthetic_xs(num_groups, seed)


def hex_core(nrings, nz, pitch=1.0, dz=1.0, num_groups=16, seed=54321, delta=1.0):
    """Synthetic hexagonal-prism core: fuel / moderator alternating by ring and axial block."""
    points, cells, ring = hex_lattice(nrings, pitch)
    nxy = len(cells)
    mats = np.zeros((nz, nxy), dtype=np.int32)
    for kk in range(nz):
        mats[kk] = (ring // 4 + kk // 8) % 2
    mesh = polygon_mesh(points, cells, np.full(nz, dz), mats.reshape(-1), delta=delta)
    return mesh, synthetic_xs(num_groups, seed), (points, cells)
