"""Host-side problem description for the C-ABI CUDA layer and the `SNDevice` handle wrapper.

The arrays mirror include/pampa_sn.h one to one (the structs pampa_sn_mesh / _xs / _quadrature /
_ls).  The reference-facing host code (Parser, meshes, materials, SNSolver) is the C++ library in
pampa_b200/host; this module is what bench.py and the parity tests use to drive the device
layer directly with numpy arrays.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib

BC_NONE, BC_VACUUM, BC_REFLECTIVE = 0, 1, 2


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


@dataclass
class ExtrudedMesh:
    """2-D polygon mesh x layers (pampa_sn_mesh)."""
    xy_num_faces: np.ndarray        # [nxy]
    xy_neighbor: np.ndarray         # [nxy,F]  >=0 xy cell, <0 -(1-based bc)
    xy_face_fx: np.ndarray          # [nxy,F]  outward normal * lateral length
    xy_face_fy: np.ndarray
    xy_face_cf: np.ndarray          # [nxy,F]
    xy_area: np.ndarray             # [nxy]
    xy_cx: np.ndarray
    xy_cy: np.ndarray
    materials: np.ndarray           # [nz*nxy] 0-based
    bc_types: list                  # 1-based (entry 0 unused)
    dz: np.ndarray | None = None    # None: no z faces (1-D / 2-D mesh)
    bc_minus_z: int = 0
    bc_plus_z: int = 0
    xy_ij: np.ndarray | None = None  # [nxy,2]
    # mixed-face-interpolation delta < 1: deferred-correction weights of the lateral faces (include/pampa_sn.h)
    delta: float = 1.0
    xy_face_kout: np.ndarray | None = None   # [nxy,F]
    xy_face_kin: np.ndarray | None = None

    @property
    def num_xy_cells(self):
        return len(self.xy_area)

    @property
    def num_layers(self):
        return 1 if self.dz is None else len(self.dz)

    @property
    def num_cells(self):
        return self.num_xy_cells * self.num_layers


@dataclass
class CrossSections:
    """Per-material multigroup data (pampa_sn_xs); scattering is [mat][from][to]."""
    sigma_total: np.ndarray
    sigma_scattering: np.ndarray
    nu_sigma_fission: np.ndarray
    kappa_sigma_fission: np.ndarray
    chi_effective: np.ndarray
    beta_total: np.ndarray | None = None

    @property
    def num_materials(self):
        return self.sigma_total.shape[0]

    @property
    def num_groups(self):
        return self.sigma_total.shape[1]


@dataclass
class Quadrature:
    directions: np.ndarray          # [M,3]
    weights: np.ndarray             # [M]
    reflected: np.ndarray           # [M,3]


@dataclass
class LSCorrection:
    cell: np.ndarray
    ptr: np.ndarray
    nbr: np.ndarray
    omega: np.ndarray
    nvec: np.ndarray                # [nnz,3]


class SNError(RuntimeError):
    pass


class _Packed:
    """Keeps the numpy buffers alive next to the ctypes structs that point into them."""

    def __init__(self):
        self.keep = []

    def f64(self, a):
        a = _f64(a); self.keep.append(a)
        return a.ctypes.data_as(_lib.p_f64)

    def i32(self, a):
        a = _i32(a); self.keep.append(a)
        return a.ctypes.data_as(_lib.p_i32)


def pack_mesh(m: ExtrudedMesh, pk: _Packed) -> _lib.Mesh:
    nxy, F = m.xy_neighbor.shape
    cm = _lib.Mesh()
    cm.num_xy_cells, cm.num_layers = nxy, m.num_layers
    cm.has_z_faces = 0 if m.dz is None else 1
    cm.max_xy_faces = F
    cm.xy_num_faces = pk.i32(m.xy_num_faces)
    cm.xy_neighbor = pk.i32(m.xy_neighbor)
    cm.xy_face_fx, cm.xy_face_fy = pk.f64(m.xy_face_fx), pk.f64(m.xy_face_fy)
    cm.xy_face_cf = pk.f64(m.xy_face_cf)
    cm.xy_area, cm.xy_cx, cm.xy_cy = pk.f64(m.xy_area), pk.f64(m.xy_cx), pk.f64(m.xy_cy)
    cm.xy_ij = pk.i32(m.xy_ij) if m.xy_ij is not None else None
    cm.dz = pk.f64(m.dz) if m.dz is not None else None
    cm.materials = pk.i32(m.materials)
    cm.bc_minus_z, cm.bc_plus_z = int(m.bc_minus_z), int(m.bc_plus_z)
    cm.num_bcs = len(m.bc_types) - 1
    cm.bc_types = pk.i32(m.bc_types)
    cm.face_interpolation_delta = float(m.delta)
    cm.xy_face_kout = pk.f64(m.xy_face_kout) if m.xy_face_kout is not None else None
    cm.xy_face_kin = pk.f64(m.xy_face_kin) if m.xy_face_kin is not None else None
    return cm


def pack_xs(x: CrossSections, pk: _Packed) -> _lib.XS:
    cx = _lib.XS()
    cx.num_materials, cx.num_groups = x.num_materials, x.num_groups
    cx.sigma_total = pk.f64(x.sigma_total)
    cx.sigma_scattering = pk.f64(x.sigma_scattering)
    cx.nu_sigma_fission = pk.f64(x.nu_sigma_fission)
    cx.kappa_sigma_fission = pk.f64(x.kappa_sigma_fission)
    cx.chi_effective = pk.f64(x.chi_effective)
    cx.beta_total = pk.f64(x.beta_total if x.beta_total is not None else np.zeros(x.num_materials))
    return cx


def pack_quadrature(q: Quadrature, pk: _Packed) -> _lib.Quadrature:
    cq = _lib.Quadrature()
    cq.num_directions = len(q.weights)
    cq.directions, cq.weights, cq.reflected = pk.f64(q.directions), pk.f64(q.weights), pk.i32(q.reflected)
    return cq


def pack_ls(ls: LSCorrection | None, pk: _Packed):
    if ls is None or len(ls.cell) == 0:
        return None
    cl = _lib.LS()
    cl.num_cells = len(ls.cell)
    cl.cell, cl.ptr, cl.nbr = pk.i32(ls.cell), pk.i32(ls.ptr), pk.i32(ls.nbr)
    cl.omega, cl.nvec = pk.f64(ls.omega), pk.f64(ls.nvec)
    return cl


def make_options(**kw) -> _lib.Options:
    o = _lib.Options()
    _lib.load().pampa_sn_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError("unknown option '%s'" % k)
        setattr(o, k, int(v))
    return o


def plan_check(mesh: ExtrudedMesh, quad: Quadrature, num_groups: int, **opts) -> dict:
    """Host-only validation of the sweep plan (no GPU needed)."""
    lib = _lib.load()
    pk = _Packed()
    cm, cq, o, info = pack_mesh(mesh, pk), pack_quadrature(quad, pk), make_options(**opts), _lib.Info()
    if lib.pampa_sn_plan_check(C.byref(cm), C.byref(cq), num_groups, C.byref(o), C.byref(info)):
        raise SNError(lib.pampa_sn_last_error(None).decode())
    return {k: getattr(info, k) for k, _ in _lib.Info._fields_}


class SNDevice:
    """A device-resident SN problem (pampa_sn_handle)."""

    def __init__(self, mesh: ExtrudedMesh, xs: CrossSections, quad: Quadrature,
                 ls: LSCorrection | None = None, **opts):
        self.lib = _lib.load()
        pk = _Packed()
        cm, cx, cq = pack_mesh(mesh, pk), pack_xs(xs, pk), pack_quadrature(quad, pk)
        cl = pack_ls(ls, pk)
        o = make_options(**opts)
        h = C.c_void_p()
        rc = self.lib.pampa_sn_create(C.byref(h), C.byref(cm), C.byref(cx), C.byref(cq),
                                      C.byref(cl) if cl is not None else None, C.byref(o))
        if rc:
            raise SNError(self.lib.pampa_sn_last_error(None).decode())
        self.h = h
        self.num_cells, self.G, self.M = mesh.num_cells, xs.num_groups, len(quad.weights)

    def _check(self, rc):
        if rc:
            raise SNError(self.lib.pampa_sn_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.pampa_sn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update_xs(self, xs: CrossSections):
        pk = _Packed()
        cx = pack_xs(xs, pk)
        self._check(self.lib.pampa_sn_update_xs(self.h, C.byref(cx)))

    def update_materials(self, xs: CrossSections, materials):
        """New cross-section rows and a new cell -> row map (temperature feedback)."""
        pk = _Packed()
        cx = pack_xs(xs, pk)
        self._check(self.lib.pampa_sn_update_materials(self.h, C.byref(cx), pk.i32(materials)))

    def source(self, keff: float):
        self._check(self.lib.pampa_sn_source(self.h, keff))

    def sweep(self):
        self._check(self.lib.pampa_sn_sweep(self.h))

    def reduce(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._check(self.lib.pampa_sn_reduce(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def iterate(self, n: int) -> float:
        k = C.c_double()
        self._check(self.lib.pampa_sn_iterate(self.h, n, C.byref(k)))
        return k.value

    def iterate_timed(self, n: int):
        """n source iterations; returns (keff, total device ms, sweep-kernel device ms)."""
        k, a, b = C.c_double(), C.c_double(), C.c_double()
        self._check(self.lib.pampa_sn_iterate_timed(self.h, n, C.byref(k), C.byref(a), C.byref(b)))
        return k.value, a.value, b.value

    def solve_keff(self, tol_k=1e-9, tol_phi=1e-8, max_it=20000, power=1.0):
        k, it = C.c_double(), C.c_int32()
        self._check(self.lib.pampa_sn_solve_keff(self.h, tol_k, tol_phi, max_it, power, C.byref(k), C.byref(it)))
        return k.value, it.value

    def get(self, name: str, out: np.ndarray | None = None) -> np.ndarray:
        """Field in the reference layout; `out` may be a caller-owned (e.g. pinned) float64 buffer."""
        n = self.lib.pampa_sn_field_size(self.h, name.encode())
        if n < 0:
            raise SNError("unable to find field '%s'" % name)
        if out is None:
            out = np.empty(n, dtype=np.float64)
        elif out.dtype != np.float64 or out.size < n or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous float64 buffer of at least %d elements" % n)
        self._check(self.lib.pampa_sn_get(self.h, name.encode(), out.ctypes.data_as(_lib.p_f64)))
        return out[:n]

    def field_size(self, name: str) -> int:
        """Length of a field as get / set see it (this rank's part when the fields are partitioned)."""
        n = self.lib.pampa_sn_field_size(self.h, name.encode())
        if n < 0:
            raise SNError("unable to find field '%s'" % name)
        return int(n)

    def set(self, name: str, values):
        v = _f64(values)
        self._check(self.lib.pampa_sn_set(self.h, name.encode(), v.ctypes.data_as(_lib.p_f64)))

    def info(self) -> dict:
        info = _lib.Info()
        self._check(self.lib.pampa_sn_get_info(self.h, C.byref(info)))
        return {k: getattr(info, k) for k, _ in _lib.Info._fields_}

    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self.lib.pampa_sn_comm_init(self.h, buf, 128))

    def device_ptr(self, name: str):
        n = C.c_int64()
        p = self.lib.pampa_sn_device_ptr(self.h, name.encode(), C.byref(n))
        return p, n.value


def nccl_unique_id() -> bytes:
    lib = _lib.load()
    buf = C.create_string_buffer(128)
    if lib.pampa_sn_comm_unique_id(buf, 128):
        raise SNError(lib.pampa_sn_last_error(None).decode())
    return buf.raw
