"""Worker for the sharded-run tests: launched with torch.distributed.run, one process per rank.

  --backend nccl : each rank owns a GPU; the sharded k-eff solve (NCCL allreduce of the flux moments
                   inside the C-ABI layer) must reproduce the one-GPU solve.
  --backend gloo : CPU-only check of the host-side sharding logic (plan ownership, id broadcast)."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pampa_b200 import problem as pb, synthetic as syn  # noqa: E402


def problem(kind="cartesian"):
    if kind == "hex":
        # hexagonal lattice: three rhombic tilings, the three-face dataflow kernel, the fused tail of the iteration
        # with the base tiling's pass last (vacuum boundaries: the largest eigenvalue is well separated)
        G = 4
        mesh, xs, _ = syn.hex_core(14, 12, pitch=1.2, dz=1.5, num_groups=G, seed=7)
        xs.nu_sigma_fission[0] *= 4.0; xs.kappa_sigma_fission[0] *= 4.0
        return mesh, xs, syn.level_symmetric(4), G
    nx, ny, nz, G = 24, 20, 12, 4
    mats = np.zeros((nz, ny, nx), dtype=int)
    mats[:, :, 16:] = 1; mats[:, 14:, :] = 1; mats[9:] = 1
    xs = syn.synthetic_xs(G, seed=21)
    xs.nu_sigma_fission[0] *= 4.0; xs.kappa_sigma_fission[0] *= 4.0
    bcs = {"-x": pb.BC_REFLECTIVE, "-y": pb.BC_REFLECTIVE}
    mesh = syn.cartesian_mesh(np.full(nx, 1.5), np.full(ny, 1.5), np.full(nz, 2.0), mats, bcs)
    return mesh, xs, syn.level_symmetric(4), G


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--shard-mode", type=int, default=0)
    ap.add_argument("--mesh", default="cartesian", choices=["cartesian", "hex"])
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    mesh, xs, quad, G = problem(a.mesh)
    total = mesh.num_cells * len(quad.weights) * G
    if a.backend == "gloo":
        dist.init_process_group("gloo")
        info = pb.plan_check(mesh, quad, G, rank=rank, num_ranks=world, shard_mode=a.shard_mode)
        t = torch.tensor([info["updates_per_sweep"]], dtype=torch.int64)
        dist.all_reduce(t)
        assert int(t[0]) == total, (int(t[0]), total)
        assert 0 < info["updates_per_sweep"] < total
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid[:] = torch.arange(128, dtype=torch.uint8)
        dist.broadcast(uid, 0)
        assert bytes(uid.numpy().tobytes()) == bytes(range(128))
        dist.barrier()
        if rank == 0:
            print("GLOO_OK")
        dist.destroy_process_group()
        return
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ref = None
    if rank == 0:
        one = pb.SNDevice(mesh, xs, quad, device=local)
        # a few plain source iterations from a given iterate on the fresh handle: the path the benchmark times (in a
        # group-sharded run the reduction pass delivers the flux moments to the peers itself, peer-to-peer)
        one.set("flux-moments", np.linspace(0.5, 1.5, mesh.num_cells * G))
        plain = [one.iterate(5), one.get("flux-moments")]
        k1, it1 = one.solve_keff(tol_k=1e-11, tol_phi=1e-9)
        ref = [k1, one.get("scalar-flux"), one.get("power")] + plain
        one.close()
    dev = pb.SNDevice(mesh, xs, quad, device=local, rank=rank, num_ranks=world, shard_mode=a.shard_mode)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(pb.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    dev.comm_init(bytes(uid.cpu().numpy().tobytes()))
    dev.set("flux-moments", np.linspace(0.5, 1.5, mesh.num_cells * G))
    kp = dev.iterate(5)
    phip = dev.get("flux-moments")
    k, it = dev.solve_keff(tol_k=1e-11, tol_phi=1e-9)
    phi, q = dev.get("scalar-flux"), dev.get("power")
    ks = torch.tensor([k], dtype=torch.float64, device="cuda")
    dist.all_reduce(ks, op=dist.ReduceOp.MAX)
    assert abs(float(ks[0]) - k) < 1e-14          # every rank holds the same k
    if rank == 0:
        assert abs(k - ref[0]) < 1e-9, (k, ref[0])
        assert np.linalg.norm(phi - ref[1]) / np.linalg.norm(ref[1]) < 1e-7
        assert np.linalg.norm(q - ref[2]) / np.linalg.norm(ref[2]) < 1e-7
        assert abs(kp - ref[3]) < 1e-11 * abs(ref[3]), (kp, ref[3])
        perr = np.linalg.norm(phip - ref[4]) / np.linalg.norm(ref[4])
        assert perr < 1e-9, perr
        print("NCCL_OK keff %.9f iterations %d (1 GPU: %.9f); 5 plain iterations: k diff %.1e, flux diff %.1e" % (
            k, it, ref[0], abs(kp - ref[3]), perr))
    dev.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
