"""Development probe: host<->device field transfer times against torch's own pinned copies."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pampa_b200 import problem as pb, synthetic as syn
n = (216, 216, 216); G = 8
mesh, xs = syn.checkerboard_core(*n, num_groups=G)
dev = pb.SNDevice(mesh, xs, syn.level_symmetric(8), store_psi=0)
N = mesh.num_cells * G
pin = torch.empty(N, dtype=torch.float64, pin_memory=True); pin.fill_(1.0)
pag = np.ones(N)
d = torch.empty(N, dtype=torch.float64, device="cuda")
for name, src in (("pinned", pin),):
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(src, non_blocking=True); torch.cuda.synchronize()
        print("torch H2D", name, "%.2f ms" % ((time.perf_counter() - t0) * 1e3))
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); src.copy_(d, non_blocking=True); torch.cuda.synchronize()
        print("torch D2H", name, "%.2f ms" % ((time.perf_counter() - t0) * 1e3))
for _ in range(3):
    t0 = time.perf_counter(); dev.set("flux-moments", pin.numpy()); print("set pinned %.2f ms" % ((time.perf_counter() - t0) * 1e3))
t0 = time.perf_counter(); dev.set("flux-moments", pag); print("set pageable %.2f ms" % ((time.perf_counter() - t0) * 1e3))
out = pin.numpy()
for _ in range(3):
    t0 = time.perf_counter(); dev.get("flux-moments", out=out); print("get pinned %.2f ms" % ((time.perf_counter() - t0) * 1e3))
powb = torch.empty(mesh.num_cells, dtype=torch.float64, pin_memory=True).numpy()
for _ in range(2):
    t0 = time.perf_counter(); dev.get("power", out=powb); print("get power pinned %.2f ms" % ((time.perf_counter() - t0) * 1e3))
