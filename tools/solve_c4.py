"""Wall time of a full k-eff solve on the benchmark core (not the bench; informational)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pampa_b200 import problem as pb, synthetic as syn
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs=3, default=[216, 216, 216])
ap.add_argument("--groups", type=int, default=8)
ap.add_argument("--order", type=int, default=8)
ap.add_argument("--tol", type=float, default=1e-7)
ap.add_argument("--max-it", type=int, default=5000)
ap.add_argument("--opts", default="{}")
ap.add_argument("--pre", type=int, default=0, help="plain iterations before resetting the flux (perturbs the initial k)")
a = ap.parse_args()
mesh, xs = syn.checkerboard_core(*a.n, num_groups=a.groups)
dev = pb.SNDevice(mesh, xs, syn.level_symmetric(a.order), **json.loads(a.opts))
if a.pre > 0:
    dev.iterate(a.pre)
    import numpy as np
    dev.set("flux-moments", np.ones(mesh.num_cells * a.groups))
t0 = time.time()
try:
    k, it = dev.solve_keff(tol_k=a.tol, tol_phi=a.tol, max_it=a.max_it)
    print("keff %.8f iterations %d wall %.2f s (%.1f ms/iteration)" % (k, it, time.time() - t0, (time.time() - t0) / it * 1e3))
except pb.SNError as e:
    print("solve failed:", e, "wall %.2f s" % (time.time() - t0))
