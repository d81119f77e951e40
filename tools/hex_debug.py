"""Dataflow-kernel variants on a small hexagonal core against the general kernel (each case in its own process
with a timeout, so that a stuck launch is reported instead of hanging the run)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r'''
import sys, json, numpy as np
sys.path.insert(0, %r)
from pampa_b200 import problem as pb, synthetic as syn
rings, nz, G, order, opts, its = json.loads(sys.argv[1])
if isinstance(rings, str):            # a committed fixture
    sys.path.insert(0, %r + "/tests")
    import util
    mesh, xs, quad, ls, z = util.load_golden(rings)
else:
    mesh, xs, _ = syn.hex_core(rings, nz, num_groups=G)
    quad = syn.level_symmetric(order)
ref = pb.SNDevice(mesh, xs, quad, generic_only=1)
kr = ref.iterate(its); pr = ref.get("flux-moments"); ref.close()
dev = pb.SNDevice(mesh, xs, quad, **opts)
info = dev.info()
k = dev.iterate(its); p = dev.get("flux-moments"); dev.close()
err = float(np.max(np.abs(p - pr) / np.abs(pr)))
print(json.dumps({"k": k, "kref": kr, "max_rel": err, "flow_classes": info["flow_classes"], "launches": info["sweep_launches"], "chunks": info["num_chunks"]}))
''' % (ROOT, ROOT)

cases = [("hex_core_s8_2g", 0, 0, 0, {}, 3), ("hex_core_s8_11g", 0, 0, 0, {}, 3), ("pwr_unstructured_s2", 0, 0, 0, {}, 3)]
for dt in (1, 2, 3, 4, 5, 6):
    cases.append((24, 12, 4, 8, {"dt_max": dt}, 2))
cases += [(24, 12, 2, 8, {}, 2), (24, 12, 16, 12, {}, 2), (24, 12, 16, 12, {"group_merge": 4}, 2),
          (24, 12, 11, 8, {}, 2), (7, 16, 2, 8, {}, 3), (7, 16, 11, 8, {}, 3), (24, 40, 8, 8, {"store_psi": 0}, 2),
          (60, 30, 8, 8, {}, 2), (60, 30, 8, 8, {"dt_max": 3}, 2), (40, 20, 16, 12, {}, 2)]
for c in cases:
    try:
        r = subprocess.run([sys.executable, "-c", CHILD, json.dumps(c)], capture_output=True, text=True, timeout=90)
        out = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ("rc %d " % r.returncode) + r.stderr[-400:]
    except subprocess.TimeoutExpired:
        out = "TIMEOUT (stuck launch?)"
    print(c, "->", out, flush=True)
