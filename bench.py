#!/usr/bin/env python
"""Benchmark of the SN k-eigenvalue hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, host cores)

Workload (BASELINE.json configs[3]): synthetic 3-D Cartesian core, 216 x 216 x 216 = 10 077 696
cells, S8 (80 directions), 8 energy groups, vacuum boundaries, checkerboard of fuel / moderator
assemblies, cross sections from seed 12345.  One "step" is one source
iteration: scattering + fission source, transport sweep of every direction and group, flux-moment /
k-eff reduction (and the NCCL allreduce of the flux moments when sharded over N GPUs).

metric = cell*angle*group updates per second = cells * directions * groups * steps / time.
Sharded over N ranks by energy group when G divides by N (peer-to-peer exchange of the group slabs of the
flux moments, sharded source / reduction), else by angle set (allreduce of the flux moments); the total work
is fixed => scaling "strong".  Before the timed region a sharded run solves a small problem on every rank
and compares k-eff and the scalar flux with a one-GPU solve of rank 0 ("sharded_parity" in the line).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cell-angle-group updates/s per source iteration"
UNIT = "updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, nargs=3, default=[216, 216, 216])
    ap.add_argument("--mesh", default="cartesian", choices=["cartesian", "hex"],
                    help="hex: synthetic unstructured extruded hexagonal core (BASELINE config 5 family); "
                         "informational, the bench line of record is the Cartesian default")
    ap.add_argument("--rings", type=int, default=120, help="hex mesh: rings of hexagons around the centre")
    ap.add_argument("--groups", type=int, default=8)
    ap.add_argument("--order", type=int, default=8)
    ap.add_argument("--cpu-scale", type=int, default=2, help="CPU sample: mesh edge divided by this")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-solve", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the sharded-vs-one-GPU check of multi-GPU runs")
    ap.add_argument("--opts", default="{}", help="JSON of pampa_sn_options overrides")
    ap.add_argument("--config", default=None, choices=["c4", "c5"],
                    help="c4: the default (BASELINE configs[3], the line of record); c5: BASELINE configs[4], synthetic "
                         "unstructured extruded hexagonal core, 200 467 hexagons x 200 layers = 40.1M cells, S12, 16 groups, "
                         "meant for --gpus 8 (energy-group sharding, 2 groups per GPU); the angular flux is not kept "
                         "(store_psi = 0: 860 GB of psi do not fit 8 x 180 GB next to the step-major arrays) and the "
                         "k-eff solve is skipped (its history vectors would not fit either)")
    a = ap.parse_args()
    if a.config == "c5":
        a.mesh, a.rings, a.order, a.groups = "hex", 258, 12, 16
        a.size = [1, 1, 200]
        o = json.loads(a.opts)
        o.setdefault("store_psi", 0)
        a.opts = json.dumps(o)
        a.no_solve = True
    return a


def workload_name(a):
    if a.mesh == "hex":
        nxy = 3 * a.rings * (a.rings + 1) + 1
        return "synthetic 3D unstructured extruded hex core %d hexagons x %d layers = %d cells, S%d, %d groups" % (
            nxy, a.size[2], nxy * a.size[2], a.order, a.groups)
    return "synthetic 3D extruded Cartesian core %dx%dx%d=%d cells, S%d, %d groups" % (
        a.size[0], a.size[1], a.size[2], a.size[0] * a.size[1] * a.size[2], a.order, a.groups)


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.stop_flag, self.index = [], False, index
        self.thread = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                self.samples.append([t.strip() for t in out.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        self.thread.join(timeout=6)
        sm = [int(s[0]) for s in self.samples if len(s) >= 6 and s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if len(s) >= 6 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples if len(s) >= 6 for n, v in zip(names, s[2:6]) if v == "Active"})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU arm
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_problem(a, scale):
    from oracle import sweep_cpu
    from pampa_b200 import synthetic as syn
    n = [max(8, s // scale) for s in a.size]
    mesh, xs = syn.checkerboard_core(n[0], n[1], n[2], num_groups=a.groups)
    quad = syn.level_symmetric(a.order)
    mats = mesh.materials.reshape(n[2], n[1], n[0])
    # every host core this process may use (torch.distributed.run exports OMP_NUM_THREADS=1: not inherited), and
    # the angular flux is stored, as in the GPU arm (store_psi = 1), when it fits a bounded share of host memory
    ncell = n[0] * n[1] * n[2]
    psi_bytes = ncell * len(quad.weights) * a.groups * 8
    store_psi = psi_bytes <= 16e9
    cpu = sweep_cpu.SweepCPU(np.ones(n[0]), np.ones(n[1]), np.ones(n[2]), mats, xs.sigma_total,
                             xs.sigma_scattering, xs.nu_sigma_fission, xs.chi_effective, quad.directions,
                             quad.weights, threads=host_threads(), store_psi=store_psi)
    updates = ncell * len(quad.weights) * a.groups
    return cpu, updates, n


def cpu_time_steps(a, scale, steps, warmup):
    cpu, updates, n = cpu_problem(a, scale)
    phi = np.ones(a.groups * n[0] * n[1] * n[2])
    k = 1.0
    for _ in range(warmup):
        k = cpu.iterate(phi, k, 1)
    t0 = time.perf_counter()
    for _ in range(steps):
        k = cpu.iterate(phi, k, 1)
    dt = time.perf_counter() - t0
    sample = "%dx%dx%d sub-core of the same workload (1/%d of the cells), %d source iterations, fp64, psi %s" % (
        n[0], n[1], n[2], scale ** 3, steps, "stored" if cpu.psi is not None else "not stored")
    return updates * steps / dt, dt / steps * 1e3, cpu.threads, sample


def reference_algorithm_timing():
    """SURVEY 8(d) baseline 1: the reference's own algorithm -- monolithic R, sparse LU, Arnoldi on R^-1 F
    (src/petsc.cxx:193-197) -- restated with scipy (oracle.solve_monolithic) on the reference's slab cases,
    rebuilt from the committed fixtures.  Real PETSc / SLEPc cannot be installed here; one process, wall time."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    from oracle import pampa_oracle as orc
    out = []
    for name in ("slabs_s2", "slabs_s4"):
        em, xs, quad, ls, z = util.load_golden(name)
        obcs = [0, orc.VACUUM, orc.VACUUM]
        mesh = orc.build_cartesian_mesh(em.xy_area, None, None, em.materials, ["-x", "+x"], obcs)
        t0 = time.perf_counter()
        op = orc.build_operator(mesh, util.xs_to_oracle(xs), xs.num_groups, int(z["order"]), 1.0,
                                "literal_zero_init", obcs)
        sol = orc.solve_monolithic(op)
        dt = time.perf_counter() - t0
        n = op.N * op.G * op.M
        out.append({"case": name, "unknowns": n, "wall_s": dt, "keff": sol.keff,
                    "golden_keff": float(z["golden_keff"]), "unknowns_per_s": n / dt})
    return {"kind": "restatement of reference algorithm (scipy SuperLU + ARPACK, 1 process)", "cores": 1,
            "note": "assembly + LU + eigen-solve of the monolithic system; C4 has 6.4e9 unknowns", "cases": out}


def run_reference(a, rank):
    """CPU arm: the oracle's C/OpenMP port of the same discrete operator on all host cores.  The
    reference's own PETSc/SLEPc path cannot be built here and cannot form its monolithic matrix at
    this size, so kind = "port"."""
    if rank != 0:
        return
    # a bounded sample sized so that warmup + steps end within a few minutes
    scale = max(a.cpu_scale, 2)
    per_step_budget_s = 240.0 / max(1, a.steps + a.warmup)
    while True:
        n = [max(8, s // scale) for s in a.size]
        est = n[0] * n[1] * n[2] * a.order * (a.order + 2) * a.groups / 1.5e8      # ~1.5e8 updates/s on 8 cores
        if est <= per_step_budget_s or scale >= 16:
            break
        scale += 1
    value, ms, threads, sample = cpu_time_steps(a, scale, a.steps, min(a.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "cpu_sample": sample, "same_config": scale == 1,
                       "sample_fraction_of_cells": 1.0 / scale ** 3},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
def comm_setup(dev, dist, rank):
    import torch
    from pampa_b200 import problem as pb
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(pb.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    dev.comm_init(bytes(uid.cpu().numpy().tobytes()))


def sharded_parity(a, dist, rank, world, local_rank, shard_opts):
    """Sharded-versus-single-GPU check carried by every multi-GPU bench line: a reduced copy of the workload
    (48^3 cells, same generator, same order and groups, reflective -x / -y so that the boundary-flux exchange is
    exercised too) is solved to 1e-10 on rank 0 alone and then by all ranks with the run's sharding."""
    import torch
    from pampa_b200 import problem as pb, synthetic as syn
    n = 48
    bcs = {"-x": pb.BC_REFLECTIVE, "-y": pb.BC_REFLECTIVE}
    mesh, xs = syn.checkerboard_core(n, n, n, num_groups=a.groups, bcs=bcs)
    quad = syn.level_symmetric(a.order)
    ref = None
    if rank == 0:
        one = pb.SNDevice(mesh, xs, quad, device=local_rank)
        k1, it1 = one.solve_keff(tol_k=1e-10, tol_phi=1e-9)
        ref = (k1, one.get("scalar-flux"), it1)
        one.close()
    dev = pb.SNDevice(mesh, xs, quad, device=local_rank, rank=rank, num_ranks=world, **shard_opts)
    comm_setup(dev, dist, rank)
    # the iterate goes in through the (partitioned) field interface too: every rank sets its own range of cells
    dev.set("flux-moments", np.ones(dev.field_size("flux-moments")))
    k, it = dev.solve_keff(tol_k=1e-10, tol_phi=1e-9)
    phi = dev.get("scalar-flux")                        # this rank's range of cells when the fields are partitioned
    dev.close()
    # every rank checks its own part against the one-GPU flux of rank 0
    nphi = mesh.num_cells * a.groups
    full = torch.zeros(nphi, dtype=torch.float64, device="cuda")
    if rank == 0:
        full.copy_(torch.from_numpy(ref[1]))
    dist.broadcast(full, 0)
    full = full.cpu().numpy()
    per = -(-mesh.num_cells // world) * a.groups if shard_opts.get("partition_fields") else nphi
    i0 = min(nphi, per * rank) if shard_opts.get("partition_fields") else 0
    mine = full[i0:i0 + phi.size]
    err = torch.tensor([float(np.sum((phi - mine) ** 2)), float(np.sum(mine ** 2))], dtype=torch.float64, device="cuda")
    mx = torch.tensor([float(np.max(np.abs(phi - mine) / np.abs(mine))), k, -k], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(phi.size)], dtype=torch.float64, device="cuda")
    dist.all_reduce(err)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(cnt)
    spread = float(mx[1] + mx[2])                       # max k - min k over the ranks
    out = None
    if rank == 0:
        l2 = float(torch.sqrt(err[0] / err[1]))
        covered = float(cnt[0]) / nphi                  # 1 with partitioned fields (each cell checked once), else N
        out = {"problem": "%d^3 cells, S%d, %d groups, reflective -x/-y" % (n, a.order, a.groups),
               "keff_1gpu": ref[0], "keff_sharded": k, "keff_diff": k - ref[0], "keff_spread_over_ranks": spread,
               "phi_rel_l2": l2, "phi_max_rel": float(mx[0]), "iterations": [ref[2], it],
               "field_coverage": covered,
               "ok": bool(abs(k - ref[0]) < 1e-7 and l2 < 1e-6 and float(mx[0]) < 1e-5 and spread < 1e-12)}
    return out


def run_b200(a, rank, world, local_rank):
    import torch
    from pampa_b200 import problem as pb, synthetic as syn

    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    nx, ny, nz = a.size
    if a.mesh == "hex":
        mesh, xs, _ = syn.hex_core(a.rings, nz, pitch=1.0, dz=1.0, num_groups=a.groups, seed=54321)
        a.no_cpu_baseline = True          # the oracle's C port is Cartesian only
    else:
        mesh, xs = syn.checkerboard_core(nx, ny, nz, num_groups=a.groups)
    quad = syn.level_symmetric(a.order)
    M = len(quad.weights)
    # sharding: by energy group when the groups divide evenly over the ranks (allgather of the group
    # slabs, sharded source / reduction), else by angle set (allreduce of the flux moments)
    # partition_fields: every rank moves its own contiguous range of cells between host and device (the
    # local-length vectors the reference hands its MPI ranks), 1/N of the host bytes each
    opts = dict(device=local_rank, rank=rank, num_ranks=world,
                shard_mode=1 if (world > 1 and a.groups % world == 0) else 0,
                partition_fields=1 if world > 1 else 0)
    opts.update(json.loads(a.opts))
    parity = None
    if world > 1 and not a.no_parity:
        parity = sharded_parity(a, dist, rank, world, local_rank,
                                {k: v for k, v in opts.items() if k not in ("device", "rank", "num_ranks")})
    dev = pb.SNDevice(mesh, xs, quad, **opts)
    if world > 1:
        comm_setup(dev, dist, rank)

    info0 = dev.info()
    U_total = mesh.num_cells * M * a.groups          # whole-job updates per step
    U_own = info0["updates_per_sweep"]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    dev.iterate(a.warmup)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = dev.info()["kernel_launches"]
    k, total_ms, sweep_ms = dev.iterate_timed(a.steps)
    barrier()
    launches = dev.info()["kernel_launches"] - l0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([total_ms, sweep_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, sweep_ms = float(t[0]), float(t[1])

    value = U_total * a.steps / (total_ms * 1e-3)

    # roofline of the dominant kernel (the sweep), algorithmic bytes per update 16 + 16/M
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    b_alg = 16.0 + 16.0 / M
    # the dominant kernel timed alone: CUDA events on the launching stream around the sweep-kernel
    # launches of every timed step (after the shear pass, before the un-shear pass)
    inf = dev.info()
    kernel_ms = inf["timed_kernel_ms"]
    phase = [inf["timed_source_ms"], sweep_ms, kernel_ms, inf["timed_reduce_ms"], inf["timed_exchange_ms"]]
    if dist is not None:
        t = torch.tensor([kernel_ms] + phase, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kernel_ms = float(t[0])
        phase = [float(x) for x in t[1:]]
    # device time per step of each phase (CUDA events on the launching stream, max over the ranks)
    step_phases_ms = {"source": phase[0] / a.steps, "shear + sweep + un-shear": phase[1] / a.steps,
                      "sweep kernel alone": phase[2] / a.steps, "reduce + scalars + exchange": phase[3] / a.steps,
                      "flux-moment exchange alone": phase[4] / a.steps}
    achieved = U_own * a.steps * b_alg / (kernel_ms * 1e-3) / 1e9
    achieved_sweep = U_own * a.steps * b_alg / (sweep_ms * 1e-3) / 1e9
    # DRAM bytes per launch from the committed ncu --set full capture of this very configuration (one GPU, default
    # options, Cartesian workload); it cannot be measured inside an un-profiled run, and it is NOT extrapolated to
    # sharded runs or other options, whose padding and task shapes differ: null there
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if os.path.exists(tpath) and world == 1 and a.mesh == "cartesian" and not json.loads(a.opts) \
            and list(a.size) == [216, 216, 216] and a.order == 8 and a.groups == 8:
        tj = json.load(open(tpath))
        per_update = tj.get("dram_bytes_per_update")
        if per_update:
            traffic = per_update * U_own / max(1, info0["sweep_launches"])
            traffic_src = "ncu capture %s (dram__bytes_read.sum + dram__bytes_write.sum per launch)" % tj.get("source", "profiles/")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "sn_sweep_flow_kernel" if info0["flow_classes"] > 0 else "sn_sweep_kernel (generic)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_update": b_alg, "launches_per_step": info0["sweep_launches"],
                "algorithmic_bytes_per_launch": b_alg * U_own / max(1, info0["sweep_launches"]),
                "kernel_ms_per_launch": kernel_ms / a.steps / max(1, info0["sweep_launches"]),
                # the same algorithmic bytes over shear + sweep kernel + un-shear (the layout passes
                # the step-major arrays cost), and that fraction of the peak
                "sweep_ms_per_step": sweep_ms / a.steps, "frac_with_layout_passes": achieved_sweep / peak}
    if not int(json.loads(a.opts).get("store_psi", 1)):
        # the angular flux is not kept: the 8 B psi store of the formula is not made (only the edge copies other
        # patches read), so `achieved` is a throughput stated on the psi-stored byte count, not a bandwidth claim
        roofline["psi_stored"] = False
        roofline["note"] = ("store_psi = 0: achieved / frac are updates/s x the psi-stored 16 + 16/M bytes, for comparison "
                            "with psi-stored runs; the kernel does not write the psi rows, so it is not a DRAM-bandwidth figure")

    # end to end through the C ABI with host buffers: one solve-like call = upload of the cross
    # sections and of the flux iterate, K source iterations, download of scalar flux and power
    e2e = None
    if not a.no_e2e:
        n_phi = dev.field_size("flux-moments")          # this rank's part of the field when sharded
        host_in = torch.empty(n_phi, dtype=torch.float64, pin_memory=True).numpy()
        host_phi = torch.empty(n_phi, dtype=torch.float64, pin_memory=True).numpy()
        host_pow = torch.empty(dev.field_size("power"), dtype=torch.float64, pin_memory=True).numpy()
        host_in[:] = 1.0
        # untimed warm-up of the same call sequence (first use allocates the device staging buffer)
        dev.set("flux-moments", host_in)
        dev.iterate(1)
        dev.get("scalar-flux", out=host_phi)
        dev.get("power", out=host_pow)
        barrier()
        t0 = time.perf_counter()
        dev.update_xs(xs)
        t1 = time.perf_counter()
        dev.set("flux-moments", host_in)
        t2 = time.perf_counter()
        dev.iterate(a.steps)
        t3 = time.perf_counter()
        phi_out = dev.get("scalar-flux", out=host_phi)
        pow_out = dev.get("power", out=host_pow)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        phases = {"update_xs": (t1 - t0) * 1e3, "set": (t2 - t1) * 1e3, "iterate": (t3 - t2) * 1e3,
                  "get": (t0 + dt - t3) * 1e3}
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        xs_bytes = sum(v.nbytes for v in (xs.sigma_total, xs.sigma_scattering, xs.nu_sigma_fission,
                                          xs.kappa_sigma_fission, xs.chi_effective))
        h2d, d2h = host_in.nbytes + xs_bytes, phi_out.nbytes + pow_out.nbytes
        if dist is not None:                              # whole-job bytes: summed over the ranks
            t = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            h2d, d2h = float(t[0]), float(t[1])
        e2e = {"value": U_total * a.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": h2d / a.steps,
               "d2h_bytes_per_step": d2h / a.steps,
               "call": "update_xs + set(flux-moments) + %d source iterations + get(scalar-flux, power)" % a.steps,
               "phases_ms": phases}

    # the other half of the headline metric: wall time of a full k-eff solve (Anderson-accelerated
    # source iteration to |dk| < 1e-7 and a relative flux change < 1e-7)
    keff_solve = None
    if not a.no_solve:
        # cold start: flat flux, k = 1 (the timed steps above leave a settled k estimate behind, which saves
        # ~40 of ~200 accelerated iterations)
        dev.set("flux-moments", np.ones(dev.field_size("flux-moments")))
        dev.set("keff", np.ones(1))
        barrier()
        solve_sampler = ClockSampler(local_rank)
        if rank == 0:
            solve_sampler.start()
        t0 = time.perf_counter()
        try:
            ks, its = dev.solve_keff(tol_k=1e-7, tol_phi=1e-7, max_it=20000)
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            keff_solve = {"wall_s": wall, "iterations": its, "keff": ks, "tol_k": 1e-7,
                          "tol_phi": 1e-7, "start": "flat flux, k = 1", "ms_per_iteration": wall * 1e3 / max(1, its),
                          "device_s": dev.info()["last_solve_ms"] * 1e-3,
                          "clocks": solve_sampler.stop() if rank == 0 else None}
        except pb.SNError as e:          # informational half of the metric: never lose the bench line over it
            keff_solve = {"error": str(e)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, ms, threads, sample = cpu_time_steps(a, max(a.cpu_scale, 2), 2, 1)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                        "same_config": False}
        try:
            cpu_baseline["reference_algorithm"] = reference_algorithm_timing()
        except Exception as e:           # informational: never lose the bench line over it
            cpu_baseline["reference_algorithm"] = {"error": str(e)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(a), "parallelism": "%s sharding x%d" % ("energy-group" if opts["shard_mode"] == 1 else "angle-set", world),
                           "l2_policy": "working set (%.1f GB) far exceeds the 126 MB L2" % (info0["device_bytes"] / 1e9),
                           "keff_after_steps": k, "options": json.loads(a.opts)},
                "roofline": roofline, "step_phases_ms": step_phases_ms, "cpu_baseline": cpu_baseline, "e2e": e2e,
                "keff_solve": keff_solve,
                "sharded_parity": parity,
                "gpu_launches": launches,
                "clocks": clocks}
        print(json.dumps(line), flush=True)
    dev.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank)
        return
    if world == 1 and a.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_b200(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
