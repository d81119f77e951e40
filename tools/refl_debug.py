"""The sharded-parity problem of bench.py (48^3 checkerboard core, reflective -x / -y) on one GPU with several
acceleration depths: k, iterations, smallest flux."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pampa_b200 import problem as pb, synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
bcs = {"-x": pb.BC_REFLECTIVE, "-y": pb.BC_REFLECTIVE}
mesh, xs = syn.checkerboard_core(n, n, n, num_groups=8, bcs=bcs)
quad = syn.level_symmetric(8)
for depth in (0, 3, -1):
    dev = pb.SNDevice(mesh, xs, quad, anderson_depth=depth, verbose=2 if depth == 0 else 1)
    t0 = time.time()
    try:
        k, it = dev.solve_keff(tol_k=1e-10, tol_phi=1e-9)
        phi = dev.get("scalar-flux")
        print("depth", depth, "k %.12f" % k, "iterations", it, "min/max phi %.3e" % (phi.min() / phi.max()), "%.2fs" % (time.time() - t0), flush=True)
    except pb.SNError as e:
        print("depth", depth, "FAILED:", e, "%.2fs" % (time.time() - t0), flush=True)
    dev.close()
