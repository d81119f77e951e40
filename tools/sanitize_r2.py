"""Round-2 code paths on small problems, meant to run under compute-sanitizer (tools/gpu_r2_call21.sh): the fused
tail of the iteration with the un-shear passes split in z, the accelerated iteration (in-place fused reduction, streamed
mix pass), the three-face dataflow kernel on a hexagonal lattice (three tilings), the delta < 1 deferred correction."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pampa_b200 import problem as pb, synthetic as syn


def run(tag, mesh, xs, quad, iterations=3, solve_its=12, **opts):
    dev = pb.SNDevice(mesh, xs, quad, **opts)
    k = dev.iterate(iterations)
    phi = dev.get("flux-moments")
    info = dev.info()
    try:
        dev.solve_keff(tol_k=1e-14, tol_phi=1e-14, max_it=solve_its)      # a few accelerated iterations
    except pb.SNError as e:
        assert "did not converge" in str(e), e
    k2 = float(dev.get("keff")[0])
    dev.close()
    print("%-28s k(%d it) %.12f  k(+%d accelerated) %.12f  min phi %.3e  launches/sweep %d tilings %d flow classes %d" % (
        tag, iterations, k, solve_its, k2, phi.min(), info["sweep_launches"], info["num_tilings"], info["flow_classes"]))
    return k, phi


G = 4
quad = syn.level_symmetric(4)
if len(sys.argv) > 1 and sys.argv[1] == "hex-small":
    # (racecheck: one small hexagonal lattice, the three-face dataflow kernel and the per-tiling layout passes)
    hmesh, hxs, _ = syn.hex_core(8, 16, pitch=1.0, dz=1.0, num_groups=2, seed=54321)
    run("hex lattice (small)", hmesh, hxs, quad, iterations=2, solve_its=3)
    print("SANITIZE_R2_DONE")
    sys.exit(0)
mesh, xs = syn.checkerboard_core(40, 36, 32, assembly=4, num_groups=G)
os.environ["PAMPA_SN_UNSHEAR_ZSPLIT"] = "1"
k1, p1 = run("cartesian zsplit 1", mesh, xs, quad)
os.environ["PAMPA_SN_UNSHEAR_ZSPLIT"] = "2"
k2, p2 = run("cartesian zsplit 2", mesh, xs, quad)
# (k is summed per un-shear CTA: its last bits, and through 1/k those of the later iterates, depend on the slabs)
assert np.abs(p1 - p2).max() < 1e-13 * np.abs(p1).max() and abs(k1 - k2) < 1e-13 * abs(k1)
os.environ.pop("PAMPA_SN_UNSHEAR_ZSPLIT")
hmesh, hxs, _ = syn.hex_core(12, 32, pitch=1.0, dz=1.0, num_groups=G, seed=54321)
kh, ph = run("hex lattice", hmesh, hxs, quad)
os.environ["PAMPA_SN_UNSHEAR_ZSPLIT"] = "2"
kh2, ph2 = run("hex lattice zsplit 2", hmesh, hxs, quad)
assert np.abs(ph - ph2).max() < 1e-13 * np.abs(ph).max() and abs(kh - kh2) < 1e-13 * abs(kh)
os.environ.pop("PAMPA_SN_UNSHEAR_ZSPLIT")
nx, ny, nz = 24, 20, 16
kk, jj, ii = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
dmesh = syn.cartesian_mesh(np.full(nx, 1.2), np.full(ny, 1.1), np.full(nz, 1.3), ((ii // 4 + jj // 4 + kk // 4) % 2).astype(int),
                           {"-x": pb.BC_REFLECTIVE}, delta=0.1)
run("cartesian delta 0.1", dmesh, syn.synthetic_xs(G, seed=11), quad, solve_its=6)
print("SANITIZE_R2_DONE")
