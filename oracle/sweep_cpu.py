"""ctypes wrapper of oracle/sweep_cpu.c (test / CPU-baseline infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsweep_cpu.so")


class _Problem(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("G", C.c_int), ("M", C.c_int),
                ("nmat", C.c_int), ("dx", C.c_void_p), ("dy", C.c_void_p), ("dz", C.c_void_p),
                ("mats", C.c_void_p), ("sigma_t", C.c_void_p), ("sigma_s", C.c_void_p), ("nusf", C.c_void_p),
                ("chi", C.c_void_p), ("dirs", C.c_void_p), ("w", C.c_void_p), ("psi", C.c_void_p)]


def build(force=False):
    src = os.path.join(HERE, "sweep_cpu.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B"], check=True, capture_output=True)
    return LIB


class SweepCPU:
    """Cartesian-core source iteration on the host cores (fp64, OpenMP)."""

    def __init__(self, dx, dy, dz, mats, sigma_t, sigma_s, nusf, chi, dirs, w, threads=None, store_psi=False):
        """threads: OpenMP threads to use (None: every core this process may run on, whatever OMP_NUM_THREADS
        says -- torch.distributed.run exports OMP_NUM_THREADS=1); store_psi: keep the angular flux
        [G][M][nz][ny][nx] in host memory, as the GPU arm does by default."""
        self.lib = C.CDLL(build())
        self.lib.sweep_cpu_set_threads.argtypes = [C.c_int]
        if threads is None:
            try:
                threads = len(os.sched_getaffinity(0))
            except AttributeError:
                threads = os.cpu_count() or 1
        self.lib.sweep_cpu_set_threads(int(threads))
        self.lib.sweep_cpu_iterate.restype = C.c_double
        self.lib.sweep_cpu_iterate.argtypes = [C.POINTER(_Problem), C.c_void_p, C.c_double, C.c_int]
        self.lib.sweep_cpu_solve.restype = C.c_double
        self.lib.sweep_cpu_solve.argtypes = [C.POINTER(_Problem), C.c_void_p, C.c_double, C.c_double, C.c_int,
                                             C.POINTER(C.c_int)]
        self.lib.sweep_cpu_threads.restype = C.c_int
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.keep = [f(dx), f(dy), f(dz), np.ascontiguousarray(mats, dtype=np.int32), f(sigma_t), f(sigma_s),
                     f(nusf), f(chi), f(dirs), f(w)]
        p = _Problem()
        p.nx, p.ny, p.nz = len(dx), len(dy), len(dz)
        p.G, p.M, p.nmat = sigma_t.shape[1], len(w), sigma_t.shape[0]
        for name, a in zip(["dx", "dy", "dz", "mats", "sigma_t", "sigma_s", "nusf", "chi", "dirs", "w"], self.keep):
            setattr(p, name, a.ctypes.data)
        self.n = p.nx * p.ny * p.nz
        self.psi = None
        p.psi = None
        if store_psi:
            self.psi = np.zeros(p.G * p.M * self.n)
            p.psi = self.psi.ctypes.data
        self.p = p
        self.threads = self.lib.sweep_cpu_threads()

    def iterate(self, phi, keff, iters):
        return self.lib.sweep_cpu_iterate(C.byref(self.p), phi.ctypes.data, keff, iters)

    def solve(self, tol_k=1e-10, tol_phi=1e-9, max_it=20000):
        phi = np.ones(self.p.G * self.n)
        it = C.c_int()
        k = self.lib.sweep_cpu_solve(C.byref(self.p), phi.ctypes.data, tol_k, tol_phi, max_it, C.byref(it))
        return k, phi.reshape(self.p.G, self.p.nz, self.p.ny, self.p.nx), it.value
