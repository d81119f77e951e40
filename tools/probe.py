"""Quick device probe: time the three kernels on a synthetic Cartesian core (not the bench)."""
import argparse
import json
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pampa_b200 import problem as pb, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs=3, default=[128, 128, 128])
ap.add_argument("--groups", type=int, default=8)
ap.add_argument("--order", type=int, default=8)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--opts", type=str, default="{}")
a = ap.parse_args()
nx, ny, nz = a.n
t0 = time.time()
mesh, xs = syn.checkerboard_core(nx, ny, nz, num_groups=a.groups)
quad = syn.level_symmetric(a.order)
t1 = time.time()
dev = pb.SNDevice(mesh, xs, quad, verbose=1, **json.loads(a.opts))
t2 = time.time()
info = dev.info()
print("build %.2fs create %.2fs  bytes %.2f GB launches/sweep %d tasks %d" % (
    t1 - t0, t2 - t1, info["device_bytes"] / 1e9, info["sweep_launches"], info["sweep_tasks"]))
U = info["updates_per_sweep"]
for it in range(a.iters):
    dev.source(1.0); dev.sweep(); dev.reduce()
    i = dev.info()
    print("iter %d sweep %.3f ms (%.3e upd/s, %.1f%% of 6551.7 GB/s at 16.2 B) source %.3f ms reduce %.3f ms" % (
        it, i["last_sweep_ms"], U / i["last_sweep_ms"] * 1e3, U / i["last_sweep_ms"] * 1e3 * (16 + 16 / len(quad.weights)) / 6551.7e9 * 100,
        i["last_source_ms"], i["last_reduce_ms"]))
t3 = time.time()
k = dev.iterate(5)
print("5 iterations: %.3f s, keff %.6f" % (time.time() - t3, k))
