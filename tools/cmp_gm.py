"""Development check: the flux moments after a few source iterations must not depend on how the sweep is
scheduled (wave launches, dataflow launch, groups per task)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pampa_b200 import problem as pb, synthetic as syn
n = [int(x) for x in sys.argv[1:4]]
cases = json.loads(sys.argv[4])
G = 8
mesh, xs = syn.checkerboard_core(*n, num_groups=G)
quad = syn.level_symmetric(8)
ref = None
for opts in cases:
    env = opts.pop("dbg", None)
    if env is not None:
        os.environ["PAMPA_SN_DBG"] = str(env)
    else:
        os.environ.pop("PAMPA_SN_DBG", None)
    dev = pb.SNDevice(mesh, xs, quad, **opts)
    k = dev.iterate(3)
    phi = dev.get("flux-moments").reshape(n[2], n[1], n[0], G)
    dev.close()
    if ref is None:
        ref = phi
        print(opts, "k", repr(k))
        continue
    rel = np.abs(phi - ref) / np.abs(ref)
    bad = np.argwhere(rel > 1e-13)
    print(opts, "dbg", env, "k", repr(k), "max rel diff %.3g" % rel.max(), "cells off", len(bad))
    if len(bad):
        print("  by group", np.bincount(bad[:, 3], minlength=G).tolist(),
              "z", bad[:, 0].min(), bad[:, 0].max(), "y", bad[:, 1].min(), bad[:, 1].max(), "x", bad[:, 2].min(), bad[:, 2].max())
