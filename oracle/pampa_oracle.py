"""CPU oracle for pampa's SN k-eigenvalue path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch numpy/scipy restatement of the *discrete problem* the
reference assembles and solves; nothing in the product path imports it.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it.

Parity status: PINNED for the four SN cases the reference ships (k-eff to the six
printed decimals of test/check_ref.txt:32,53,234,415 -- see tests/test_oracle.py).
Scalar-flux vectors are not pinned by any reference artefact (the reference checks
none in); for flux, S8, hex-core and the synthetic cases the oracle is the sole
authority ("parity pinned by oracle restatement only").

The reference's arithmetic for this path lives in third-party PETSc 3.12 /
SLEPc 3.12 (Krylov-Schur + shift-and-invert + LU, src/petsc.cxx:175-204, :428-433),
which is absent from /root/reference.  The result is mathematically defined (the
eigenpair of R x = (1/k) F x with largest k), so scipy SuperLU + ARPACK are an
exact stand-in.

What follows which reference lines:
  tokenizer           src/input.cxx:4-60
  materials           src/Material.cxx:18-134, src/ConstantNuclearData.cxx:4-177,
                      src/FeedbackNuclearData.hxx:62-90 (T=0 -> first table),
                      src/PrecursorData.cxx:4-60 (beta_total)
  Cartesian mesh      src/CartesianMesh.cxx:19-414
  unstructured mesh   src/UnstructuredExtrudedMesh.cxx:19-364, src/math.cxx:4-79
  quadrature          src/AngularQuadratureSet.cxx:4-209
  face weights        src/SNSolver.cxx:159-208
  LS boundary scheme  src/SNSolver.cxx:212-269 (+ SURVEY.md App. C.2/C.3 quirks)
  operator R, F       src/SNSolver.cxx:344-631 (steady branch)
  post-processing     src/SNSolver.cxx:272-341, src/NeutronicSolver.cxx:46-116
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

VACUUM, REFLECTIVE = 1, 2          # BC::VACUUM, BC::REFLECTIVE (src/utils.hxx:113)
BC_NAMES = {"vacuum": 1, "reflective": 2, "robin": 3, "dirichlet": 4,
            "adiabatic": 5, "convection": 6}
DBL_TOL = 1.0e-6                   # src/utils.hxx:35
KAPPA_OVER_NU = 3.2e-11 / 2.4355   # src/ConstantNuclearData.hxx:27


# --------------------------------------------------------------------------- tokenizer
class Lines:
    """Line tokenizer: collapse tabs/double spaces, trim, skip blank and '#' lines."""

    def __init__(self, path):
        with open(path, "r") as f:
            self.raw = f.read().split("\n")
        self.pos = 0

    def next(self):
        while self.pos < len(self.raw):
            s = self.raw[self.pos].replace("\t", " ")
            self.pos += 1
            while "  " in s:
                s = s.replace("  ", " ")
            s = s.strip(" ")
            if not s or s[0] == "#":
                continue
            return s.split(" ")
        return []

    def numbers(self, n, conv=float):
        out = []
        while len(out) < n:
            line = self.next()
            if not line:
                raise ValueError("missing data")
            out.extend(conv(t) for t in line)
        if len(out) > n:
            raise ValueError("out-of-bounds data")
        return out


# --------------------------------------------------------------------------- materials
@dataclass
class XS:
    G: int = -1
    sigma_total: np.ndarray | None = None
    nu_sigma_fission: np.ndarray | None = None
    kappa_sigma_fission: np.ndarray | None = None
    sigma_scattering: np.ndarray | None = None   # [from, to]
    chi_prompt: np.ndarray | None = None
    chi_delayed: np.ndarray | None = None
    chi_effective: np.ndarray | None = None


def _read_nuclear_data(L: Lines) -> XS:
    xs = XS()
    while True:
        line = L.next()
        if not line or line[0] == "}":
            break
        k = line[0]
        if k == "energy-groups":
            xs.G = int(line[1])
        elif k == "sigma-total":
            xs.sigma_total = np.array(L.numbers(xs.G))
        elif k == "nu-sigma-fission":
            xs.nu_sigma_fission = np.array(L.numbers(xs.G))
        elif k == "kappa-sigma-fission":
            xs.kappa_sigma_fission = np.array(L.numbers(xs.G))
        elif k == "sigma-scattering":
            rows = []
            for _ in range(xs.G):
                row = L.next()
                if len(row) != xs.G:
                    raise ValueError("wrong scattering row")
                rows.append([float(t) for t in row])
            xs.sigma_scattering = np.array(rows)
        elif k in ("fission-spectrum", "fission-spectrum-prompt"):
            xs.chi_prompt = np.array(L.numbers(xs.G))
        elif k == "fission-spectrum-delayed":
            xs.chi_delayed = np.array(L.numbers(xs.G))
        elif k in ("sigma-transport", "diffusion-coefficient", "neutron-velocity"):
            L.numbers(xs.G)
        else:
            raise ValueError("unrecognized keyword '%s'" % k)
    return xs


def _finish_nuclear_data(xs: XS, beta_total: float) -> None:
    """Derived data, src/ConstantNuclearData.cxx:129-177."""
    G = xs.G
    if xs.nu_sigma_fission is None:
        xs.nu_sigma_fission = np.zeros(G)
    if xs.kappa_sigma_fission is None:
        xs.kappa_sigma_fission = xs.nu_sigma_fission * KAPPA_OVER_NU
    if xs.chi_prompt is None:
        xs.chi_prompt = np.zeros(G)
    if xs.chi_delayed is None:
        xs.chi_delayed = xs.chi_prompt.copy()
    if beta_total > 0.0:
        xs.chi_effective = (1.0 - beta_total) * xs.chi_prompt + beta_total * xs.chi_delayed
    else:
        xs.chi_effective = xs.chi_prompt.copy()


def read_material(path) -> XS:
    """Material file -> the cross-section table a standalone SN run sees (T = 0)."""
    return read_material_tables(path)[1][0]


def xs_at_temperature(temperatures, tables, T) -> XS:
    """Linear interpolation between the tabulated temperatures, clamped outside the table
    (src/FeedbackNuclearData.hxx:62-140).  T equal to the first tabulated temperature is read as that table (the
    reference's lowerBound would index entry -1 there)."""
    n = len(tables)
    if n < 2 or T <= temperatures[0]:
        return tables[0]
    if T > temperatures[-1]:
        return tables[-1]
    i2 = int(np.searchsorted(temperatures, T, side="left"))
    i1 = i2 - 1
    f = (T - temperatures[i1]) / (temperatures[i2] - temperatures[i1])
    a, b = tables[i1], tables[i2]
    mix = lambda u, v: (1.0 - f) * u + f * v
    return XS(G=a.G, sigma_total=mix(a.sigma_total, b.sigma_total),
              nu_sigma_fission=mix(a.nu_sigma_fission, b.nu_sigma_fission),
              kappa_sigma_fission=mix(a.kappa_sigma_fission, b.kappa_sigma_fission),
              sigma_scattering=mix(a.sigma_scattering, b.sigma_scattering),
              chi_prompt=mix(a.chi_prompt, b.chi_prompt), chi_delayed=mix(a.chi_delayed, b.chi_delayed),
              chi_effective=mix(a.chi_effective, b.chi_effective))


def read_material_tables(path):
    """Material file -> (temperatures, [XS per tabulated temperature]); constant data: ([0.0], [XS])."""
    L = Lines(path)
    tables, beta_total = [], 0.0
    temperatures = [0.0]
    while True:
        line = L.next()
        if not line or line[0] == "}":
            break
        k = line[0]
        if k == "nuclear-data":
            tables = [_read_nuclear_data(L)]
        elif k == "nuclear-data-set":
            tables = []
            while True:
                sub = L.next()
                if not sub or sub[0] == "}":
                    break
                if sub[0] == "temperature":
                    temperatures = L.numbers(int(sub[1]))
                elif sub[0] == "nuclear-data":
                    tables.append(_read_nuclear_data(L))
                else:
                    raise ValueError("unrecognized keyword '%s'" % sub[0])
        elif k == "precursor-data":
            while True:
                sub = L.next()
                if not sub or sub[0] == "}":
                    break
                if sub[0] == "precursor-groups":
                    npg = int(sub[1])
                elif sub[0] == "lambda":
                    L.numbers(npg)
                elif sub[0] == "beta":
                    for b in L.numbers(npg):
                        beta_total += b
        elif k in ("thermal-properties", "fuel", "bc", "split"):
            pass
        else:
            raise ValueError("unrecognized keyword '%s'" % k)
    for xs in tables:                  # (T = 0 clamps to the first table)
        _finish_nuclear_data(xs, beta_total)
    return np.array(temperatures, dtype=float), tables


# --------------------------------------------------------------------------- mesh
@dataclass
class Mesh:
    """Generic finite-volume mesh: ragged per-cell face tables (reference face order)."""
    num_dims: int
    volumes: np.ndarray
    centroids: np.ndarray            # [N,3]
    materials: np.ndarray            # [N] 0-based
    face_ptr: np.ndarray             # [N+1]
    face_area: np.ndarray
    face_centroid: np.ndarray        # [nf,3]
    face_normal: np.ndarray          # [nf,3]
    face_neighbor: np.ndarray        # >=0 cell, <0 = -(1-based boundary index)
    boundaries: list
    bcs: list                        # 1-based: bcs[0] unused; entries BC type ints (0 = none)
    # extruded description (what the device layer consumes)
    ext: dict = field(default_factory=dict)

    @property
    def num_cells(self):
        return len(self.volumes)


def _read_axis(L, tok):
    n = int(tok)
    if n > 0:
        return np.array(L.numbers(n))
    return np.full(-n, L.numbers(1)[0])


def read_cartesian_mesh(path) -> Mesh:
    L = Lines(path)
    dx = np.array([1.0]); dy = None; dz = None
    boundaries, bc_lines, mats, num_dims = [], [], None, 0
    while True:
        line = L.next()
        if not line:
            break
        k = line[0]
        if k == "dx":
            dx = _read_axis(L, line[1]); num_dims += 1
            boundaries += ["-x", "+x"]
        elif k == "dy":
            dy = _read_axis(L, line[1])
            if len(dy) > 1: num_dims += 1
            boundaries += ["-y", "+y"]
        elif k == "dz":
            dz = _read_axis(L, line[1])
            if len(dz) > 1: num_dims += 1
            boundaries += ["-z", "+z"]
        elif k == "bc":
            bc_lines.append(line)
        elif k == "materials":
            n = int(line[1])
            mats = np.array(L.numbers(n, int)) - 1
        elif k == "nodal-indices":
            L.numbers(int(line[1]), int)
        else:
            raise ValueError("unrecognized keyword '%s'" % k)
    bcs = []
    if bc_lines:
        bcs = [0] * (1 + len(boundaries))
        for line in bc_lines:
            bcs[boundaries.index(line[1]) + 1] = BC_NAMES[line[2]]
    return build_cartesian_mesh(dx, dy, dz, mats, boundaries, bcs, num_dims)


def build_cartesian_mesh(dx, dy, dz, mats, boundaries, bcs, num_dims=None) -> Mesh:
    """Faces in the reference order (-y,+x,+y,-x,-z,+z); void cells (material -1) dropped."""
    nx = len(dx); ny = 0 if dy is None else len(dy); nz = 0 if dz is None else len(dz)
    if num_dims is None:
        num_dims = 1 + (ny > 1) + (nz > 1)
    nyy, nzz = max(ny, 1), max(nz, 1)
    x = np.concatenate([[0.0], np.cumsum(dx)])
    y = np.concatenate([[0.0], np.cumsum(dy)]) if ny else np.zeros(2)
    z = np.concatenate([[0.0], np.cumsum(dz)]) if nz else np.zeros(2)
    dyv = dy if ny else np.zeros(1)
    dzv = dz if nz else np.zeros(1)
    mats3 = np.asarray(mats).reshape(nzz, nyy, nx)
    phys = mats3 != -1
    if not np.all(phys == phys[0:1]):
        raise ValueError("wrong material definition")
    cid = np.full(mats3.shape, -1, dtype=np.int64)
    cid[phys] = np.arange(phys.sum())
    bidx = {name: boundaries.index(name) for name in boundaries}

    def nb(k, j, i, name, dk, dj, di):
        kk, jj, ii = k + dk, j + dj, i + di
        if 0 <= kk < nzz and 0 <= jj < nyy and 0 <= ii < nx and phys[kk, jj, ii]:
            return cid[kk, jj, ii]
        return -bidx[name] - 1

    vol, cen, mat = [], [], []
    fptr, farea, fcen, fnor, fnei = [0], [], [], [], []
    for k in range(nzz):
        for j in range(nyy):
            for i in range(nx):
                if not phys[k, j, i]:
                    continue
                dxi, dyj, dzk = dx[i], dyv[j], dzv[k]
                vol.append(dxi * dyj * dzk if nz else (dxi * dyj if ny else dxi))
                cx, cy, cz = x[i] + 0.5 * dxi, y[j] + 0.5 * dyj, z[k] + 0.5 * dzk
                cen.append((cx, cy, cz)); mat.append(mats3[k, j, i])
                a_x = dyj * dzk if nz else (dyj if ny else 1.0)
                a_y = dxi * dzk if nz else dxi
                if ny:
                    farea.append(a_y); fcen.append((cx, y[j], cz)); fnor.append((0., -1., 0.))
                    fnei.append(nb(k, j, i, "-y", 0, -1, 0))
                farea.append(a_x); fcen.append((x[i] + dxi, cy, cz)); fnor.append((1., 0., 0.))
                fnei.append(nb(k, j, i, "+x", 0, 0, 1))
                if ny:
                    farea.append(a_y); fcen.append((cx, y[j] + dyj, cz)); fnor.append((0., 1., 0.))
                    fnei.append(nb(k, j, i, "+y", 0, 1, 0))
                farea.append(a_x); fcen.append((x[i], cy, cz)); fnor.append((-1., 0., 0.))
                fnei.append(nb(k, j, i, "-x", 0, 0, -1))
                if nz:
                    farea.append(dxi * dyj); fcen.append((cx, cy, z[k])); fnor.append((0., 0., -1.))
                    fnei.append(nb(k, j, i, "-z", -1, 0, 0))
                    farea.append(dxi * dyj); fcen.append((cx, cy, z[k] + dzk)); fnor.append((0., 0., 1.))
                    fnei.append(nb(k, j, i, "+z", 1, 0, 0))
                fptr.append(len(farea))
    m = Mesh(num_dims, np.array(vol), np.array(cen), np.array(mat, dtype=np.int64),
             np.array(fptr), np.array(farea), np.array(fcen), np.array(fnor),
             np.array(fnei, dtype=np.int64), list(boundaries), list(bcs))
    m.ext = dict(kind="cartesian", dx=np.asarray(dx), dy=None if dy is None else np.asarray(dy),
                 dz=None if dz is None else np.asarray(dz), phys_xy=phys[0])
    return m


def _poly_area(pts, ids):
    a = 0.0
    n = len(ids)
    for i in range(n):
        p1, p2 = pts[ids[i]], pts[ids[(i + 1) % n]]
        a += p1[0] * p2[1] - p2[0] * p1[1]
    return 0.5 * a


def _poly_centroid(pts, ids, a):
    cx = cy = 0.0
    n = len(ids)
    for i in range(n):
        p1, p2 = pts[ids[i]], pts[ids[(i + 1) % n]]
        da = p1[0] * p2[1] - p2[0] * p1[1]
        cx += (p1[0] + p2[0]) * da
        cy += (p1[1] + p2[1]) * da
    return cx * (1.0 / (6.0 * a)), cy * (1.0 / (6.0 * a))


def read_unstructured_mesh(path) -> Mesh:
    L = Lines(path)
    pts = None; cells = None; dz = None; nz = 0
    boundaries, xy_b_names, xy_b_pts, xy_default, bc_lines, mats = [], [], [], -1, [], None
    num_dims = 0
    while True:
        line = L.next()
        if not line:
            break
        k = line[0]
        if k == "points":
            n = int(line[1])
            pts = np.array([[float(t) for t in L.next()] for _ in range(n)])
        elif k == "cells":
            n = int(line[1])
            cells = [[int(t) for t in L.next()] for _ in range(n)]
            num_dims += 2
        elif k == "dz":
            dz = _read_axis(L, line[1]); nz = len(dz)
            if nz > 1: num_dims += 1
            boundaries += ["-z", "+z"]
        elif k == "boundary":
            name, npts = line[1], int(line[2])
            boundaries.append(name); xy_b_names.append(name)
            if npts > 0:
                xy_b_pts.append(L.numbers(npts, int))
            else:
                xy_default = len(xy_b_pts)
                xy_b_pts.append([])
        elif k == "bc":
            bc_lines.append((line, len(boundaries)))
        elif k == "materials":
            mats = np.array(L.numbers(int(line[1]), int)) - 1
        elif k == "nodal-indices":
            L.numbers(int(line[1]), int)
        else:
            raise ValueError("unrecognized keyword '%s'" % k)
    bcs = []
    if bc_lines:
        # the array is sized when the first bc line is met (UnstructuredExtrudedMesh.cxx:116)
        bcs = [0] * (1 + bc_lines[0][1])
        for line, _ in bc_lines:
            bcs[boundaries.index(line[1]) + 1] = BC_NAMES[line[2]]
    return build_unstructured_mesh(pts, cells, dz, mats, boundaries, xy_b_names, xy_b_pts,
                                   xy_default, bcs, num_dims)


def build_unstructured_mesh(pts, cells, dz, mats, boundaries, xy_b_names, xy_b_pts,
                            xy_default, bcs, num_dims) -> Mesh:
    nxy = len(cells)
    nz = 0 if dz is None else len(dz)
    nzz = max(nz, 1)
    dzv = dz if nz else np.zeros(1)
    z = np.concatenate([[0.0], np.cumsum(dzv)])
    # point -> list of cells / boundary tags (UnstructuredExtrudedMesh.cxx:230-246)
    pc = [[] for _ in range(len(pts))]
    for i, c in enumerate(cells):
        for p in c:
            pc[p].append(i)
    for b, name in enumerate(xy_b_names):
        tag = -boundaries.index(name) - 1
        for p in xy_b_pts[b]:
            pc[p].append(tag)
    xy_nei = []
    for i, c in enumerate(cells):
        row = []
        n = len(c)
        for f in range(n):
            p1, p2 = c[f], c[(f + 1) % n]
            found, val = False, 0
            for i1 in pc[p1]:
                for i2 in pc[p2]:
                    if i1 == i2 and i1 != i:
                        val = i1; found = True       # later matches overwrite (quirk C.8)
            if not found and xy_default >= 0:
                val = -xy_default - 1; found = True  # xy ordinal, not global index (quirk C.8)
            if not found:
                raise ValueError("wrong mesh connectivity")
            row.append(val)
        xy_nei.append(row)
    areas = [_poly_area(pts, c) for c in cells]
    cents = [_poly_centroid(pts, c, a) for c, a in zip(cells, areas)]
    iz_minus = boundaries.index("-z") if nz else -1
    iz_plus = boundaries.index("+z") if nz else -1
    vol, cen = [], []
    fptr, farea, fcen, fnor, fnei = [0], [], [], [], []
    for k in range(nzz):
        for i, c in enumerate(cells):
            n = len(c)
            a = areas[i]
            vol.append(a * dzv[k] if nz else a)
            zc = z[k] + 0.5 * dzv[k]
            cen.append((cents[i][0], cents[i][1], zc))
            for f in range(n):
                p1, p2 = pts[c[f]], pts[c[(f + 1) % n]]
                length = math.sqrt((p2[0] - p1[0]) ** 2 + (p2[1] - p1[1]) ** 2)
                farea.append(length * dzv[k] if nz else length)
                fcen.append((0.5 * (p1[0] + p2[0]), 0.5 * (p1[1] + p2[1]), zc))
                n0, n1 = p2[1] - p1[1], p1[0] - p2[0]
                nrm = math.sqrt(n0 * n0 + n1 * n1)
                fnor.append((n0 / nrm, n1 / nrm, 0.0))
                nbv = xy_nei[i][f]
                fnei.append(nbv + k * nxy if nbv >= 0 else nbv)
            if nz:
                ic = k * nxy + i
                farea.append(a); fcen.append((cents[i][0], cents[i][1], z[k])); fnor.append((0., 0., -1.))
                fnei.append(-iz_minus - 1 if k == 0 else ic - nxy)
                farea.append(a); fcen.append((cents[i][0], cents[i][1], z[k] + dzv[k])); fnor.append((0., 0., 1.))
                fnei.append(-iz_plus - 1 if k == nz - 1 else ic + nxy)
            fptr.append(len(farea))
    m = Mesh(num_dims, np.array(vol), np.array(cen), np.asarray(mats, dtype=np.int64),
             np.array(fptr), np.array(farea), np.array(fcen), np.array(fnor),
             np.array(fnei, dtype=np.int64), list(boundaries), list(bcs))
    m.ext = dict(kind="unstructured", points=np.asarray(pts), cells=cells,
                 dz=None if dz is None else np.asarray(dz))
    return m


# --------------------------------------------------------------------------- quadrature
_LQ = {
    2: ([1.0 / math.sqrt(3.0)], [(0, 0, 0)], [1.0]),
    4: ([0.3500212, 0.8688903], [(0, 0, 1), (0, 1, 0), (1, 0, 0)], [1.0 / 3.0] * 3),
    6: ([0.2666355, 0.6815076, 0.9261808],
        [(0, 0, 2), (0, 2, 0), (2, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0)],
        [0.1761263] * 3 + [0.1572071] * 3),
    8: ([0.2182179, 0.5773503, 0.7867958, 0.9511897],
        [(0, 0, 3), (0, 3, 0), (3, 0, 0), (0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0),
         (2, 0, 1), (2, 1, 0), (1, 1, 1)],
        [0.1209877] * 3 + [0.0907407] * 6 + [0.0925926]),
}


def _lq12():
    """LQ12 (BASELINE config 5; the reference itself stops at S8, AngularQuadratureSet.cxx:153): points (i,j,k)
    with i+j+k = 5 in the order i, then j; weight by the sorted index triplet.  Constants checked against the
    level-symmetric defining equations in tests/test_oracle.py (mu_i^2 arithmetic progression, even moments)."""
    mu = [0.1672126, 0.4595476, 0.6280191, 0.7600210, 0.8722706, 0.9716377]
    wc = {(0, 0, 5): 0.0707626, (0, 1, 4): 0.0558811, (0, 2, 3): 0.0373377, (1, 1, 3): 0.0502819,
          (1, 2, 2): 0.0258513}
    idx = [(i, j, 5 - i - j) for i in range(6) for j in range(6 - i)]
    return mu, idx, [wc[tuple(sorted(t))] for t in idx]


_LQ[12] = _lq12()


def quadrature(order):
    """Level-symmetric set: directions [M,3], weights [M] (sum 1), reflection map [M,3]."""
    if order not in _LQ:
        raise ValueError("SN order not implemented")
    mu, idx, w = _LQ[order]
    per = len(idx)
    M = order * (order + 2)
    assert M == 8 * per
    d = np.zeros((M, 3)); wt = np.zeros(M)
    for o in range(8):
        for m in range(per):
            v = [mu[idx[m][0]], mu[idx[m][1]], mu[idx[m][2]]]
            if o & 1: v[0] = -v[0]
            if o & 2: v[1] = -v[1]
            if o & 4: v[2] = -v[2]
            d[o * per + m] = v
            wt[o * per + m] = w[m] / 8.0
    refl = np.full((M, 3), -1, dtype=np.int64)
    for ax in range(3):
        for m in range(M):
            r = d[m].copy()
            r[ax] -= 2.0 * d[m, ax]
            hits = np.nonzero(d @ r > 1.0 - DBL_TOL)[0]
            if len(hits) != 1:
                raise ValueError("reflected direction not found")
            refl[m, ax] = hits[0]
    return d, wt, refl


# --------------------------------------------------------------------------- input deck
@dataclass
class Deck:
    mesh: Mesh
    xs: list
    G: int
    order: int
    delta: float = 0.1
    ls: bool = False
    power: float = 1.0
    bcs: list = field(default_factory=list)


def read_deck(path) -> Deck:
    """Parse a main input file; paths inside are relative to the deck's directory
    (the reference resolves them against the cwd, and its tests cd into the case dir)."""
    base = os.path.dirname(os.path.abspath(path))
    L = Lines(path)
    mesh, xs, deck = None, [], None
    while True:
        line = L.next()
        if not line:
            break
        k = line[0]
        if k == "mesh":
            fn = os.path.join(base, line[2])
            if line[1] == "cartesian":
                mesh = read_cartesian_mesh(fn)
            elif line[1] == "unstructured":
                mesh = read_unstructured_mesh(fn)
            else:
                raise ValueError("wrong mesh type")
        elif k == "material":
            xs.append(read_material(os.path.join(base, line[2])))
        elif k == "solver":
            if line[1] != "sn":
                raise ValueError("oracle only restates 'solver sn'")
            deck = Deck(mesh=mesh, xs=xs, G=-1, order=-1)
            while True:
                sub = L.next()
                if not sub or sub[0] == "}":
                    break
                s = sub[0]
                if s == "energy-groups": deck.G = int(sub[1])
                elif s == "order": deck.order = int(sub[1])
                elif s == "mixed-face-interpolation": deck.delta = float(sub[1])
                elif s == "least-squares-boundary-interpolation": deck.ls = bool(int(sub[1]))
                elif s == "power": deck.power = float(sub[1])
                elif s == "bc":
                    if not deck.bcs:
                        deck.bcs = [0] * (1 + len(mesh.boundaries))
                    deck.bcs[mesh.boundaries.index(sub[1]) + 1] = BC_NAMES[sub[2]]
                elif s == "convergence": pass
                else: raise ValueError("unrecognized keyword '%s'" % s)
        elif k in ("vtk", "petsc", "dt"):
            if k == "dt":
                raise ValueError("transient decks are out of scope")
        else:
            raise ValueError("unrecognized keyword '%s'" % k)
    if not deck.bcs:
        deck.bcs = list(mesh.bcs)
    return deck


# --------------------------------------------------------------------------- LS scheme
def ls_boundary_coefs(mesh: Mesh, mode: str):
    """c_bc[cell] -> array [num_faces,3] for boundary cells (SNSolver.cxx:212-269).

    mode 'literal_zero_init': the code as written with G zero-initialised (mis-indexed d,
      App. C.2).  mode 'reference_effective': the closed form that reproduces what the
      uninitialised G evidently did on the reference machine in 2-D (App. C.3)."""
    nd = mesh.num_dims
    out = {}
    for i in range(mesh.num_cells):
        f0, f1 = mesh.face_ptr[i], mesh.face_ptr[i + 1]
        nei = mesh.face_neighbor[f0:f1]
        if not np.any(nei < 0):
            continue
        nf = f1 - f0
        d = np.zeros((nf, nd))
        for f in range(nf):
            c2 = mesh.face_centroid[f0 + f] if nei[f] < 0 else mesh.centroids[nei[f]]
            d[f] = c2[:nd] - mesh.centroids[i, :nd]
        v = d.reshape(-1)                      # row-major flat storage of Array2D(nf, nd)
        coefs = np.zeros((nf, 3))
        if mode == "reference_effective" and nd > 1:
            g00 = sum(v[f] * v[f * nd] for f in range(nf))
            for f in range(nf):
                coefs[f, 0] = v[f] / g00
        else:
            Gm = np.zeros((nd, nd))
            for jg in range(nd):
                for ig in range(nd):
                    Gm[jg, ig] = sum(v[jg * nd + f] * v[f * nd + ig] for f in range(nf))
            Gi = np.linalg.inv(Gm)
            for f in range(nf):
                for idd in range(nd):
                    coefs[f, idd] = sum(Gi[idd, jd] * v[jd * nd + f] for jd in range(nd))
        out[i] = coefs
    return out


# --------------------------------------------------------------------------- operator
@dataclass
class Operator:
    """Pieces of R x = (1/k) F x in the reference's unknown order (i*G + g)*M + m."""
    N: int; G: int; M: int
    T: sp.csr_matrix               # streaming + collision + boundary terms (block-diag in g)
    sig_s: np.ndarray              # [N, G(from), G(to)]
    chi: np.ndarray                # [N, G]
    nusf: np.ndarray               # [N, G]
    kapsf: np.ndarray              # [N, G]
    vol: np.ndarray
    w: np.ndarray

    def scatter_matrix(self):
        """V * sigma_s(g2->g) * w_m2 coupling (SNSolver.cxx:417-431), positive sign."""
        return self._cell_block(self.sig_s.transpose(0, 2, 1))          # [i, g, g2]

    def fission_matrix(self):
        """F (SNSolver.cxx:434-439)."""
        return self._cell_block(self.chi[:, :, None] * self.nusf[:, None, :])

    def _cell_block(self, blk):
        N, G, M = self.N, self.G, self.M
        # entry (i,g,m ; i,g2,m2) = V_i * blk[i,g,g2] * w[m2]
        val = (self.vol[:, None, None, None, None] * blk[:, :, None, :, None]
               * self.w[None, None, None, None, :])
        val = np.broadcast_to(val, (N, G, M, G, M))
        base = (np.arange(N) * G * M)[:, None, None, None, None]
        rows = base + (np.arange(G) * M)[None, :, None, None, None] + np.arange(M)[None, None, :, None, None]
        cols = base + (np.arange(G) * M)[None, None, None, :, None] + np.arange(M)[None, None, None, None, :]
        rows = np.broadcast_to(rows, val.shape).ravel()
        cols = np.broadcast_to(cols, val.shape).ravel()
        n = N * G * M
        return sp.csr_matrix((val.ravel(), (rows, cols)), shape=(n, n))


def build_operator(mesh: Mesh, xs: list, G: int, order: int, delta: float, ls_mode: str,
                   bcs: list, quad=None) -> Operator:
    d, w, refl = quad if quad is not None else quadrature(order)
    M = len(w)
    N = mesh.num_cells
    sig_t = np.array([x.sigma_total for x in xs])[mesh.materials]          # [N,G]
    sig_s = np.array([x.sigma_scattering for x in xs])[mesh.materials]      # [N,G,G]
    chi = np.array([x.chi_effective for x in xs])[mesh.materials]
    nusf = np.array([x.nu_sigma_fission for x in xs])[mesh.materials]
    kapsf = np.array([x.kappa_sigma_fission for x in xs])[mesh.materials]
    cbc = ls_boundary_coefs(mesh, ls_mode) if ls_mode != "off" else None

    gm = (np.arange(G) * M)[:, None] + np.arange(M)[None, :]               # [G,M] offsets
    rows, cols, vals = [], [], []

    def add(i, i2, coef_m, m2=None):
        """coef_m [M] added at (i,g,m ; i2,g,m or m2[m]) for all g."""
        r = i * G * M + gm
        c = i2 * G * M + (gm if m2 is None else (np.arange(G) * M)[:, None] + m2[None, :])
        rows.append(r.ravel()); cols.append(c.ravel())
        vals.append(np.broadcast_to(coef_m[None, :], (G, M)).ravel())

    diag = sig_t[:, :, None] * mesh.volumes[:, None, None] * np.ones((1, 1, M))   # [N,G,M]
    for i in range(N):
        f0, f1 = mesh.face_ptr[i], mesh.face_ptr[i + 1]
        for f in range(f0, f1):
            i2 = mesh.face_neighbor[f]
            A = mesh.face_area[f]
            wm = d @ mesh.face_normal[f]                                     # [M]
            out = wm > 0.0
            if i2 >= 0:
                r_if = np.linalg.norm(mesh.face_centroid[f] - mesh.centroids[i])
                r_i2f = np.linalg.norm(mesh.face_centroid[f] - mesh.centroids[i2])
                r_ii2 = np.linalg.norm(mesh.centroids[i] - mesh.centroids[i2])
                c0 = (r_i2f + delta * r_if) / r_ii2
                c1 = (1.0 - delta) * r_if / r_ii2
                c2 = (1.0 - delta) * r_i2f / r_ii2
                c3 = (r_if + delta * r_i2f) / r_ii2
                diag[i] += (np.where(out, c0, c2) * wm * A)[None, :]
                add(i, i2, np.where(out, c1, c3) * wm * A)
            else:
                bc = bcs[-i2] if -i2 < len(bcs) else 0
                if bc == VACUUM:
                    diag[i] += (np.where(out, wm, 0.0) * A)[None, :]
                    if cbc is not None:
                        dp = mesh.face_centroid[f] - mesh.centroids[i]
                        for f2 in range(f0, f1):
                            i3 = mesh.face_neighbor[f2]
                            if i3 < 0:
                                continue
                            w_i3 = float(dp @ cbc[i][f2 - f0])
                            term = np.where(out, w_i3 * wm * A, 0.0)
                            diag[i] -= term[None, :]
                            add(i, i3, term)
                elif bc == REFLECTIVE:
                    diag[i] += (np.where(out, wm, 0.0) * A)[None, :]
                    ax = [a for a in range(3) if abs(mesh.face_normal[f][a]) > 1.0 - DBL_TOL]
                    if not ax:
                        raise ValueError("reflected direction not found")
                    add(i, i, np.where(out, 0.0, wm * A), m2=refl[:, ax[-1]])
                else:
                    raise ValueError("boundary condition not implemented")
    n = N * G * M
    rows.append(np.arange(n)); cols.append(np.arange(n)); vals.append(diag.ravel())
    T = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(n, n))
    T.sum_duplicates()
    return Operator(N, G, M, T, sig_s, chi, nusf, kapsf, mesh.volumes, w)


# --------------------------------------------------------------------------- solves
@dataclass
class Solution:
    keff: float
    phi: np.ndarray        # scalar flux [N,G], reference normalisation
    psi: np.ndarray        # angular flux [N,G,M], reference normalisation
    power: np.ndarray      # q_i [N]
    production: np.ndarray  # P_i [N]


def postprocess(op: Operator, keff: float, psi: np.ndarray, power: float, allow_negative=False) -> Solution:
    """SNSolver.cxx:272-341 and NeutronicSolver.cxx:46-116.  The reference fails the solve on a negative
    scalar or angular flux (possible with mixed-face-interpolation < 1 or the LS boundary term);
    allow_negative=True returns the eigenvector anyway, for tests of the operator itself."""
    psi = psi.reshape(op.N, op.G, op.M)
    phi = 4.0 * math.pi * (psi @ op.w)
    p0 = float(np.sum(phi * op.kapsf * op.vol[:, None]))
    phi = phi * (power / p0)
    p0a = float(np.sum((psi @ op.w) * op.kapsf * op.vol[:, None]))
    psi = psi * (power / p0a)
    if phi.min() < 0.0 and not allow_negative:
        raise ValueError("negative values in the scalar-flux solution")
    if psi.min() < 0.0 and not allow_negative:
        raise ValueError("negative values in the angular-flux solution")
    q = np.sum(phi * op.kapsf, axis=1) * op.vol
    P = np.sum(phi * op.nusf, axis=1) * op.vol / keff
    return Solution(keff, phi, psi, q, P)


def solve_monolithic(op: Operator, power=1.0, tol=1e-12, allow_negative=False) -> Solution:
    """The reference algorithm: LU of the monolithic R, Arnoldi on R^-1 F (petsc.cxx:193-197)."""
    R = (op.T - op.scatter_matrix()).tocsc()
    F = op.fission_matrix().tocsr()
    lu = spla.splu(R)
    n = R.shape[0]
    A = spla.LinearOperator((n, n), matvec=lambda x: lu.solve(F @ x), dtype=float)
    v0 = np.ones(n)
    vals, vecs = spla.eigs(A, k=1, which="LM", tol=tol, v0=v0, ncv=24)
    keff = float(vals[0].real)
    return postprocess(op, keff, np.real(vecs[:, 0]), power, allow_negative)


def solve_matrix_free(op: Operator, power=1.0, tol=1e-12, inner_tol=1e-13, allow_negative=False) -> Solution:
    """Same eigenpair without forming the dense (G*M)^2 cell blocks: Arnoldi on the
    fission-source operator s -> P R^-1 E chi s, with R^-1 applied by GMRES
    preconditioned with the LU of T.  Used where the monolithic R does not fit."""
    N, G, M = op.N, op.G, op.M
    n = N * G * M
    luT = spla.splu(op.T.tocsc())
    w = op.w

    def moments(x):                       # [N,G]
        return x.reshape(N, G, M) @ w

    def expand(q):                        # isotropic source density*V -> rows
        return np.repeat((q * op.vol[:, None]).reshape(N * G), M)

    def scatter_src(x):
        return expand(np.einsum("nfg,nf->ng", op.sig_s, moments(x)))

    def apply_R(x):
        return op.T @ x - scatter_src(x)

    Rop = spla.LinearOperator((n, n), matvec=apply_R, dtype=float)
    Pre = spla.LinearOperator((n, n), matvec=luT.solve, dtype=float)

    def fixed_source(b):
        x, info = spla.gmres(Rop, b, M=Pre, rtol=inner_tol, atol=0.0, restart=60, maxiter=50)
        if info != 0:
            raise RuntimeError("oracle inner GMRES did not converge")
        return x

    state = {}

    def fis_op(s):
        b = expand(op.chi * s[:, None])
        x = fixed_source(b)
        state["psi"] = x
        return np.sum(op.nusf * moments(x), axis=1)

    A = spla.LinearOperator((N, N), matvec=fis_op, dtype=float)
    vals, vecs = spla.eigs(A, k=1, which="LM", tol=tol, v0=np.ones(N), ncv=min(N - 1, 24))
    keff = float(vals[0].real)
    s = np.real(vecs[:, 0])
    if s.sum() < 0:
        s = -s
    psi = fixed_source(expand(op.chi * s[:, None])) / keff
    return postprocess(op, keff, psi, power, allow_negative)


def solve_deck(path, ls_mode=None, method="auto", order=None) -> Solution:
    deck = read_deck(path)
    if order is not None:
        deck.order = order
    if ls_mode is None:
        ls_mode = "off" if not deck.ls else (
            "reference_effective" if deck.mesh.num_dims == 2 else "literal_zero_init")
    op = build_operator(deck.mesh, deck.xs, deck.G, deck.order, deck.delta, ls_mode, deck.bcs)
    if method == "auto":
        method = "monolithic" if op.N * (op.G * op.M) ** 2 < 4e7 else "matrix_free"
    solve = solve_monolithic if method == "monolithic" else solve_matrix_free
    return solve(op, deck.power)
