#!/usr/bin/env python
"""Benchmark of the SCATTER hot path on B200: FP64 assembly + explicit time integration of a structured hexa8 soil box.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference ...                          # the CPU path (oracle port) on the host cores

Metric (BASELINE.json): DOF*timesteps/s of the time loop (+ assembly GB/s), next to the HBM roofline and the CPU path.

* workload: hexa8 box of `--size`^3 elements per GPU (default 255^3 -> 50.3 M DOF per GPU), bottom fixed, roller
  sides, lognormal per-element Young's modulus, Rayleigh damping, heaviside point load on the top surface,
  explicit central difference with lumped mass at dt = 0.5 h / vp.  N GPUs: the box grows along z (weak scaling),
  one z-slab per rank, one halo exchange (NCCL send/recv of the two interface planes) per time step.
* one bench "step" = one solver stage of `--stage` time steps (`sc_run_central_difference`; default 100 = the
  `output_interval` of the reference's run scripts, run_scatter_rose_2D.py:19).
  `value`   : stages with everything resident in HBM, no host traffic; timed with CUDA events inside the library on
              the launching stream, max over ranks.
  `e2e`     : the same stage through the reference-facing solver object (`CentralDifferenceSolver.update/calculate`):
              initial u, v copied from pinned host arrays, one output row (u, v, a) copied back per stage.
* inputs are far larger than L2 (K alone is ~49 GB vs 126 MB), so no explicit L2 flush is needed between steps.
"""
from __future__ import annotations

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

H = 0.5
RHO, NU, E_MEAN, E_STD = 1500.0, 0.2, 30e6, 1e6
DAMPING = [1, 0.01, 30, 0.01]


def stable_dt():
    ec = (E_MEAN + 6 * E_STD) * (1 - NU) / ((1 + NU) * (1 - 2 * NU))
    return 0.3 * H / np.sqrt(ec / RHO)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                                          str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
def cpu_port_cd(size: int, steps: int, seed: int = 0):
    """The CPU path (oracle port: numpy assembly + scipy CSR SpMV central difference) on a bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fem_np as oracle                               # checker / baseline only
    from scatter_b200 import boxmesh
    model = boxmesh.box_model(size, size, size, H, "hexa8")
    model.connectivities()
    om = oracle.Model(nodes=model.nodes, elem=model.elem, materials_index=model.materials_index, materials=model.materials,
                      element_type="hexa8", dimension=3, BC=model.BC, BC_dir=model.BC_dir, eq_nb_dof=model.eq_nb_dof,
                      type_BC=model.type_BC, number_eq=model.number_eq, eq_nb_elem=model.eq_nb_elem, nb_nodes_elem=8)
    om.extra["node_rows"] = model.node_rows()
    ne = len(model.elem)
    E = boxmesh.lognormal_young(ne, E_MEAN, E_STD)
    t0 = time.perf_counter()
    K, M = oracle.assemble_global(om, E, np.full(ne, NU), np.full(ne, RHO), 2)
    t_asm = time.perf_counter() - t0
    c0, c1 = oracle.rayleigh_coefficients(DAMPING)
    m = oracle.lump_rows(M)
    c = c0 * m + c1 * np.asarray(K.sum(axis=1)).ravel()
    dt = stable_dt()
    a0, a1 = 1 / dt ** 2, 1 / (2 * dt)
    inv_d = 1 / (a0 * m + a1 * c)
    alpha = 2 * a0 * m * inv_d
    n = model.number_eq
    u = np.zeros(n); up = np.zeros(n)
    f = np.zeros(n)
    f[int(model.eq_nb_dof[boxmesh.top_centre_node(size, size, size) - 1, 1])] = -1000.0
    t0 = time.perf_counter()
    for _ in range(steps):
        un = inv_d * (f - K @ u) + alpha * u - (alpha - 1) * up
        up, u = u, un
    t_loop = time.perf_counter() - t0
    return {"n_eq": n, "n_elem": ne, "nnz": int(K.nnz), "assembly_s": t_asm, "loop_s": t_loop, "steps": steps,
            "dof_steps_per_s": n * steps / t_loop, "elem_per_s": ne / t_asm, "checksum": float(np.abs(u).sum())}


def _cpu_worker(args):
    return cpu_port_cd(*args)


def run_reference(args):
    """`--impl reference`: the CPU path on the host cores; P independent replicas of the bounded sample."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    size, steps = args.cpu_size, args.cpu_steps
    with mp.get_context("spawn").Pool(procs) as pool:
        for _ in range(max(args.warmup, 0) and 1):
            pool.map(_cpu_worker, [(min(size, 12), 2)] * procs)
        times, res = [], None
        for _ in range(args.steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [(size, steps)] * procs)
            times.append(time.perf_counter() - t0)
    # throughput of the time loop only (assembly reported separately), all replicas running concurrently
    loop = max(r["loop_s"] for r in res)
    value = sum(r["n_eq"] * r["steps"] for r in res) / loop
    line = {"impl": "reference", "metric": "dof_timesteps_per_s", "value": value, "unit": "DOF*steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": "DOF*steps/s", "cores": procs, "kind": "port",
                             "sample": f"{procs} concurrent replicas of a {size}^3-element hexa8 box ({res[0]['n_eq']} DOF each), "
                                       f"{steps} central-difference steps with scipy CSR SpMV; numpy/scipy restatement of the "
                                       "reference path (oracle/fem_np.py) -- the reference's own per-element Python loop assembles "
                                       "~355 elem/s (SURVEY.md 6)",
                             "assembly_elem_per_s": sum(r["elem_per_s"] for r in res)},
            "e2e": {"value": value, "unit": "DOF*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, n_gpus):
    s = args.size
    return {"workload": f"hexa8 soil box {s}x{s}x{s * n_gpus} elements ({s}^3 per GPU), explicit central difference "
                        f"(lumped mass, Rayleigh damping), {args.stage} time steps per bench step",
            "element_type": "hexa8", "elements_per_gpu": s ** 3, "integrator": "central_difference", "dt": stable_dt(),
            "stage_steps": args.stage, "partition": f"z-slabs x{n_gpus}", "l2": "inputs (~49 GB of CSR per GPU at 255^3) exceed L2; no flush needed"}


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from scatter_b200 import _lib, boxmesh, partition, solvers, system_matrix

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    s = args.size
    t_host0 = time.perf_counter()
    dom = partition.slab_partition(s, s, s, rank, world, H, "hexa8")
    model = dom.model
    ne = len(model.elem)
    E = boxmesh.lognormal_young(ne, E_MEAN, E_STD, seed=26021981 + rank)
    t_mesh = time.perf_counter() - t_host0

    mx = system_matrix.GenerateMatrix(model.number_eq, 2, device=local_rank)
    ctx = mx.ctx
    if world > 1:
        uid = [_lib.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.dist_init(rank, world, uid[0])
    # --- assembly: pattern + K + lumped M (timed separately) --------------------------------------------------------
    rows = model.node_rows()
    eq = model.equation_table_int()
    ctx.set_mesh("hexa8", model.nodes[:, 1:], rows, eq, model.number_eq, dom.active if world > 1 else None)
    ctx.set_materials(E, np.full(ne, NU), np.full(ne, RHO))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nnz = ctx.build_pattern()
    t_pattern = time.perf_counter() - t0
    asm_times = []
    for _ in range(3):
        asm_times.append(ctx.assemble(2, _lib.ASM_K | _lib.ASM_M_LUMPED))
    t_asm = min(asm_times)
    mx.damping_Rayleigh(DAMPING)
    if world > 1:
        ctx.set_halo(dom.neighbor_rank, dom.send_ptr, dom.send_idx, dom.recv_ptr, dom.recv_idx)
    n_owned = int(len(dom.owned_eq))
    # algorithmic bytes of the assembly: xyz 24 B/node + conn 32 B/elem + material 24 B/elem +
    # K values 8 B/nnz + lumped mass 8 B/dof
    asm_bytes = 24 * len(model.nodes) + (32 + 24) * ne + 8 * nnz + 8 * n_owned
    # --- loads: heaviside on the top-centre node of the global box (owner rank only) ---------------------------------
    dt = stable_dt()
    total_steps = (args.steps + args.warmup + 4) * args.stage * 2 + 16
    nzg = s * world
    kc = nzg // 2                                                # global plane of the loaded node
    p0 = rank * s
    p1 = (rank + 1) * s + (1 if rank == world - 1 else 0)
    ptr = np.zeros(total_steps + 1, dtype=np.int64)
    dofs = np.zeros(0, dtype=np.int64); vals = np.zeros(0)
    if p0 <= kc < p1:
        z0 = max(p0 - 1, 0)
        node_row = (s // 2) + (s + 1) * (s + (s + 1) * (kc - z0))
        d = int(eq[node_row, 1])
        ramp = np.ones(total_steps); ramp[:5] = np.linspace(0, 1, 5)
        ptr = np.arange(total_steps + 1, dtype=np.int64)
        dofs = np.full(total_steps, d, dtype=np.int64); vals = -1000.0 * ramp
    ctx.set_load_schedule(ptr, dofs, vals)
    ctx.set_state(None, None)

    peak, peak_src = measured_peak()
    step_bytes = 12 * nnz + (8 + 40) * n_owned                # CSR values+cols, rowptr, 5 vector passes

    # --- device-resident stages ---------------------------------------------------------------------------------------
    t_cur = 0
    def stage_device():
        nonlocal t_cur
        _, _, _, st = ctx.run_central_difference(dt, t_cur, args.stage, args.stage, store=False)
        t_cur += args.stage
        return st

    for _ in range(max(args.warmup, 3)):
        stage_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = ctx.kernel_launches()
    dev_s = 0.0
    w0 = time.perf_counter()
    for _ in range(args.steps):
        st = stage_device()
        dev_s += st["seconds_device"]
    if st.get("step_bytes", 0) > 0:
        step_bytes = int(st["step_bytes"])
    barrier()
    wall = time.perf_counter() - w0
    launches = ctx.kernel_launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    dev_s = max_over_ranks(dev_s)
    wall = max_over_ranks(wall)
    total_dof = sum_over_ranks(float(n_owned))
    n_ts = args.steps * args.stage
    value = total_dof * n_ts / wall
    kernel_s = dev_s / n_ts                                    # one fused SpMV+update launch per time step
    achieved = step_bytes / kernel_s / 1e9

    # --- end-to-end stages through the solver object (host buffers) -----------------------------------------------------
    num = solvers.CentralDifferenceSolver()
    num.output_interval = args.stage
    n_stage_total = args.steps + max(args.warmup, 3)
    time_arr = np.arange(0, (n_stage_total + 1) * args.stage + 1) * dt
    num.number_equations = model.number_eq
    num.time = time_arr
    num.output_time = time_arr[::args.stage]
    n_out = len(num.output_time)
    num.u = _lib.pinned_zeros((n_out, model.number_eq)); num.v = _lib.pinned_zeros((n_out, model.number_eq))
    num.a = _lib.pinned_zeros((n_out, model.number_eq))
    num.u0 = num.u[0]; num.v0 = num.v[0]
    num.bind(mx)
    num.load_schedule = (ptr, dofs, vals)
    h2d = 2 * 8 * model.number_eq + int(ptr.nbytes + dofs.nbytes + vals.nbytes)
    d2h = 2 * 3 * 8 * model.number_eq                          # rows at both ends of the stage (t0 and t0+stage)

    def stage_e2e(k):
        num.update(k * args.stage)                             # u0, v0 <- stored host row (restart hook, scatter.py:158)
        num.calculate(None, None, None, None, k * args.stage, (k + 1) * args.stage)
        return float(num.u[k + 1, n_owned // 2])               # device->host result read

    for k in range(max(args.warmup, 3)):
        stage_e2e(k)
    barrier()
    w0 = time.perf_counter()
    for k in range(max(args.warmup, 3), max(args.warmup, 3) + args.steps):
        stage_e2e(k)
    barrier()
    e2e_wall = max_over_ranks(time.perf_counter() - w0)
    e2e_value = total_dof * n_ts / e2e_wall

    # --- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            r = cpu_port_cd(args.cpu_size, args.cpu_steps)
            cpu = {"value": r["dof_steps_per_s"], "unit": "DOF*steps/s", "cores": 1, "kind": "port",
                   "sample": f"{args.cpu_size}^3-element hexa8 box ({r['n_eq']} DOF), {r['steps']} central-difference steps with scipy CSR "
                             f"SpMV (single thread) after numpy assembly at {r['elem_per_s']:.0f} elem/s; oracle/fem_np.py",
                   "assembly_elem_per_s": r["elem_per_s"], "host_cores": os.cpu_count()}
        except Exception as exc:                      # the headline line must survive a failure of the CPU leg
            cpu = {"error": repr(exc)}

    # --- random material field at scale: Gaussian SRF, 1000 modes, at every element centroid of this
    #     rank's box; outside the timed region, reported next to the assembly (the headline E stays the seeded lognormal)
    rf = None
    if rank == 0 and args.random_field:
        try:
            from scatter_b200 import random_fields
            sf = random_fields.SpectralField("Gaussian", 3, var=np.log((E_STD / E_MEAN) ** 2 + 1),
                                             mean=np.log(E_MEAN ** 2 / np.sqrt(E_MEAN ** 2 + E_STD ** 2)),
                                             len_scale=[10 * 2.0, 2.0, 10 * 2.0], angles=0.0, seed=26021981)
            t0 = time.perf_counter()
            cen = random_fields.RF.centroids(model.nodes, model.elem)
            t_cen = time.perf_counter() - t0
            t0 = time.perf_counter()
            field = sf(cen, lognormal=True, ctx=ctx)
            t_all = time.perf_counter() - t0
            rf = {"points": int(len(cen)), "modes": sf.mode_no, "kernel_seconds": sf.seconds_device, "call_seconds": t_all,
                  "centroid_host_seconds": t_cen, "sincos_per_s": len(cen) * sf.mode_no / max(sf.seconds_device, 1e-12),
                  "mean": float(field.mean()), "std": float(field.std())}
            del cen, field
        except Exception as exc:
            rf = {"error": repr(exc)}

    secondary = None
    if rank == 0 and world == 1 and args.secondary:
        info0 = ctx.device_info()
        try:
            del num
            ctx.close()
            secondary = run_secondary_newmark(args, local_rank)
        except Exception as exc:                      # the headline line must survive a failure of the secondary workload
            secondary = {"error": repr(exc)}
    if rank == 0:
        info = info0 if secondary is not None else ctx.device_info()
        sm_max_hz = 1e6 * float(clocks.get("sm_max_mhz") or 1965.0)
        line = {"metric": "dof_timesteps_per_s", "value": value, "unit": "DOF*steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
                "dof_total": total_dof, "dof_per_gpu": n_owned, "nnz_per_gpu": nnz, "time_steps_timed": n_ts,
                "roofline": {"bound": "hbm", "kernel": st.get("step_kernel", "?") + ": fused SpMV + central-difference update",
                             "bytes_per_launch_plain_csr": 12 * nnz + 48 * n_owned,
                             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                             "bytes_per_launch": step_bytes, "kernel_ms": 1e3 * kernel_s, "traffic": TRAFFIC.get(s),
                             # the same launch time against the plain-CSR byte count BASELINE.md 4 derives its roofline from
                             # (1 020 B/DOF/step): how close the time loop is to what an ideal plain-CSR SpMV could reach
                             "frac_plain_csr_equivalent": (12 * nnz + 48 * n_owned) / kernel_s / 1e9 / peak},
                "assembly": {"seconds": t_asm, "pattern_seconds": t_pattern, "gbs": asm_bytes / t_asm / 1e9,
                             "frac_hbm": asm_bytes / t_asm / 1e9 / peak, "elements_per_s": ne / t_asm, "algorithmic_bytes": asm_bytes,
                             "host_mesh_seconds": t_mesh,
                             # the kernel is FP64 bound, not HBM bound: modelled FMA count of
                             # k_assemble_blk per hexa8 element against 64 FMA/clk/SM at the maximum SM clock
                             "fma_per_element": ASM_FMA_PER_HEXA8,
                             "frac_fp64_peak": ne * ASM_FMA_PER_HEXA8 / t_asm / (64.0 * info["sm_count"] * sm_max_hz)},
                "random_field": rf,
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "DOF*steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1e3 * e2e_wall / args.steps},
                "gpu_launches": int(launches), "clocks": clocks, "device": info["name"], "secondary": secondary}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_secondary_newmark(args, local_rank):
    """BASELINE.json config 4: structured hexa20 box, ~10 M DOF, Newmark (beta=1/4, gamma=1/2) with Jacobi-PCG on one B200."""
    from scatter_b200 import _lib, boxmesh, system_matrix
    s = args.size20
    t0 = time.perf_counter()
    model = boxmesh.box_model(s, s, s, H, "hexa20")
    ne = len(model.elem)
    E = boxmesh.lognormal_young(ne, E_MEAN, E_STD)
    t_mesh = time.perf_counter() - t0
    mx = system_matrix.GenerateMatrix(model.number_eq, 2, device=local_rank)
    ctx = mx.ctx
    ctx.set_mesh("hexa20", model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), model.number_eq, None)
    ctx.set_materials(E, np.full(ne, NU), np.full(ne, RHO))
    t0 = time.perf_counter()
    nnz = ctx.build_pattern()
    t_pat = time.perf_counter() - t0
    t_asm = min(ctx.assemble(2, _lib.ASM_K | _lib.ASM_M_FULL) for _ in range(2))
    mx.damping_Rayleigh(DAMPING)
    n = model.number_eq
    dt = 5e-4
    nsteps = args.steps20
    total = nsteps * 3 + 8
    d = int(model.eq_nb_dof[boxmesh.top_centre_node(s, s, s) - 1, 1])
    ramp = np.ones(total); ramp[:5] = np.linspace(0, 1, 5)
    ctx.set_load_schedule(np.arange(total + 1, dtype=np.int64), np.full(total, d, dtype=np.int64), -1000.0 * ramp)
    ctx.set_state(None, None)
    rtol = 1e-10
    ctx.run_newmark(dt, 0, 2, 1, rtol=rtol, store=False)                       # warm-up (also builds Khat)
    _, _, _, st = ctx.run_newmark(dt, 2, nsteps, 1, rtol=rtol, store=False)
    its = st["pcg_iterations"] / max(nsteps, 1)
    # bytes per step: rhs two-matrix SpMV (values of M and K, CSR columns once) + per PCG iteration one SpMV (node-blocked
    # index: one column list per node) + 14 vector passes (SpMV x/y, fused update, direction update)
    ps = ctx.pattern_stats()
    idx_bytes = ps["node_col_entries"] * 4 + ps["n_nodes"] * 24 if ps["node_blocked"] else nnz * 4 + n * 8
    rhs_bytes = nnz * 20 + n * 8 * 12
    it_bytes = nnz * 8 + idx_bytes + n * 8 * 14
    step_bytes = rhs_bytes + its * it_bytes
    sec = st["seconds_device"] / nsteps
    peak, _ = measured_peak()
    out = {"workload": f"hexa20 soil box {s}^3 elements, Newmark + Jacobi-PCG (rtol {rtol:g}), dt {dt}", "dof": n, "nnz": nnz,
           "dof_timesteps_per_s": n / sec, "ms_per_time_step": 1e3 * sec, "pcg_iterations_per_step": its,
           "roofline": {"bound": "hbm", "achieved": step_bytes / sec / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": step_bytes / sec / 1e9 / peak, "bytes_per_step": step_bytes},
           "assembly": {"seconds": t_asm, "elements_per_s": ne / t_asm, "pattern_seconds": t_pat, "host_mesh_seconds": t_mesh},
           "last_residual": st["last_residual"]}
    ctx.close()
    return out


# FP64 FMAs k_assemble_blk spends per hexa8 element (order 2): per Gauss point 121 for J, J^-1, detJ evaluated ~4 times per
# element (once per block that sees it) and 84 in each of the 16 pair lanes
ASM_FMA_PER_HEXA8 = 8 * (4 * 121 + 16 * 84)

# dram__bytes_read.sum + dram__bytes_write.sum of one fused central-difference launch (k_spmv_node<2,2,2> with the column
# dictionary) from the committed ncu capture, by box size
TRAFFIC = {255: 35245933000 + 408624640}      # profiles/r1_v5_k_spmv_node_dict_255cube.txt (1 GPU)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=255, help="elements per box edge per GPU")
    ap.add_argument("--stage", type=int, default=100, help="time steps per bench step = output interval of the e2e arm "
                                                             "(the reference's run scripts store every 100th step, run_scatter_rose_2D.py:19)")
    ap.add_argument("--cpu-size", type=int, default=40)
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--secondary", type=int, default=1, help="also run the hexa20 Newmark/PCG workload (N = 1 only)")
    ap.add_argument("--random-field", type=int, default=1, help="also time the random-field sampler on this rank's elements")
    ap.add_argument("--size20", type=int, default=94, help="hexa20 box edge (elements) of the secondary workload")
    ap.add_argument("--steps20", type=int, default=5)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
