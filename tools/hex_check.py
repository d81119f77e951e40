"""Small hexagonal core through the dataflow kernel (for compute-sanitizer runs): a few source iterations with the
default options, edge-only rows and the general kernel; prints k of each."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pampa_b200 import problem as pb, synthetic as syn

mesh, xs, _ = syn.hex_core(24, 12, num_groups=4)
quad = syn.level_symmetric(4)
for opts in ({}, {"store_psi": 0}, {"generic_only": 1}):
    dev = pb.SNDevice(mesh, xs, quad, **opts)
    print(opts, dev.iterate(2), dev.info()["flow_classes"])
    dev.close()
