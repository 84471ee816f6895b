#!/usr/bin/env python
"""Benchmark of the SCATTER hot path on B200: FP64 assembly + explicit time integration of a structured hexa8 soil box.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference ...                          # the reference's own CPU path on the host

Metric (BASELINE.json): DOF*timesteps/s of the time loop (+ assembly GB/s), next to the HBM roofline and the CPU path.

* workload: hexa8 box of `--size`^3 elements per GPU (default 255^3 -> 49.9 M DOF per GPU), bottom fixed, roller
  sides, lognormal per-element Young's modulus, Rayleigh damping [1, 0.01, 30, 0.01] (mass-proportional part diagonal,
  stiffness-proportional part lagged through the SpMV), heaviside point load on the top surface, explicit central
  difference with lumped mass at dt = 0.3 h / vp.  N GPUs: the box grows along z (weak scaling), one z-slab per rank,
  one halo exchange (NCCL send/recv of the two interface planes) per time step.
* one bench "step" = one solver stage of `--stage` time steps (`sc_run_central_difference`; default 100 = the
  `output_interval` of the reference's run scripts, run_scatter_rose_2D.py:19).
  `value`   : stages with everything resident in HBM, no host traffic; timed with CUDA events inside the library on
              the launching stream, max over ranks.
  `e2e`     : the same stage through the reference-facing solver object (`CentralDifferenceSolver.update/calculate`):
              initial u, v copied from pinned host arrays every stage, the output rows (u, v, a) copied back.
              `e2e_by_output_interval` repeats that at output intervals 100, 10 and 1 (the reference's default, scatter.py:133-137)
              and with an output selection (`sc_set_output_dofs`: only the equations of a few monitored nodes leave the GPU).
* `parity_check`: OUTSIDE the timed region, at every rank count: a 24^3 box decomposed over the N ranks with the same slab
  partition / halo plan (and, for N > 1, a recursive-coordinate-bisection partition), K rows, 40 central-difference and 10
  Newmark steps against the single-domain CPU oracle (checker only).  A failed check exits non-zero.
* inputs are far larger than L2 (K alone is ~32 GB vs 126 MB), so no explicit L2 flush is needed between steps.
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


def _oracle():
    """The CPU oracle -- checker (parity_check) and CPU-baseline legs only; never on the measured GPU path."""
    odir = os.path.join(ROOT, "oracle")
    if odir not in sys.path:
        sys.path.insert(0, odir)
    import fem_np
    return fem_np


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(size: int, steps: int):
    """the reference's OWN assembly code, imported unchanged (baseline/_ref = pip install --target of the
    reference tree; /root/reference when present), on a `size`^3 hexa8 box written as a gmsh file and read by the reference's
    own ReadMesh: read_gmsh / read_bc / mapping / connectivities -> GenerateMatrix.generate_stiffness_and_mass ->
    absorbing_boundaries -> damping_Rayleigh (scatter.py:68-114).  The time loop is the SURVEY.md 3.3 restatement of the
    reference's default solver (incremental Newmark, beta 1/4, gamma 1/2, scipy splu factorised once): the real solver
    package, PuggleSolvers 1.0.1, is not part of the reference tree and cannot be installed here."""
    import tempfile
    import warnings
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_import
    from scatter_b200 import boxmesh
    ref = ref_import.load_reference()
    oracle = _oracle()
    warnings.filterwarnings("ignore")
    path = os.path.join(tempfile.mkdtemp(prefix="scatter_ref_"), "box.msh")
    boxmesh.write_box_msh(path, size, size, size, H, "hexa8")
    bc = boxmesh.box_boundaries(size, size, size, H)
    mat = {"solid": {"density": RHO, "Young": E_MEAN, "poisson": NU}}
    t0 = time.perf_counter()
    m = ref.mesher.ReadMesh(path)
    m.read_gmsh(); m.read_bc(bc); m.mapping(); m.connectivities()
    t_mesh = time.perf_counter() - t0
    t0 = time.perf_counter()
    mx = ref.system_matrix.GenerateMatrix(m.number_eq, 2)
    mx.generate_stiffness_and_mass(m, mat)
    t_asm = time.perf_counter() - t0
    t0 = time.perf_counter()
    mx.absorbing_boundaries(m, mat, [1, 1], 1e3)
    mx.damping_Rayleigh(DAMPING)
    t_damp = time.perf_counter() - t0
    n = int(m.number_eq)
    d = int(m.eq_nb_dof[boxmesh.top_centre_node(size, size, size) - 1, 1])

    def force(t):
        f = np.zeros(n)
        f[d] = -1000.0 * min(1.0, t / 4.0)
        return f
    t0 = time.perf_counter()
    U = oracle.newmark(mx.M, mx.C, mx.K, force, np.arange(steps + 1) * 5e-4, steps)[0]
    t_loop = time.perf_counter() - t0
    ne = len(m.elem)
    nnz = int(mx.K.nnz)
    asm_bytes = 24 * len(m.nodes) + (32 + 24) * ne + 8 * nnz + 8 * nnz / 9.0
    return {"size": size, "n_eq": n, "n_elem": ne, "nnz": nnz, "mesher_s": t_mesh, "assembly_s": t_asm, "damping_s": t_damp,
            "loop_s": t_loop, "steps": steps, "dof_steps_per_s": n * steps / t_loop, "elem_per_s": ne / t_asm,
            "assembly_gbs": asm_bytes / t_asm / 1e9, "checksum": float(np.abs(U[-1]).sum())}


def cpu_port_cd(size: int, steps: int):
    """Second CPU number: the numpy/scipy restatement (oracle/fem_np.py) running the benchmarked explicit scheme -- vectorised
    assembly + scipy CSR SpMV central difference with lagged stiffness-proportional damping -- single thread."""
    oracle = _oracle()
    from scatter_b200 import boxmesh
    model = boxmesh.box_model(size, size, size, H, "hexa8")
    model.connectivities()
    om = oracle.model_from_readmesh(model)
    ne = len(model.elem)
    E = boxmesh.lognormal_young(ne, E_MEAN, E_STD)
    t0 = time.perf_counter()
    K, M = oracle.assemble_global(om, E, np.full(ne, NU), np.full(ne, RHO), 2)
    t_asm = time.perf_counter() - t0
    c0, c1 = oracle.rayleigh_coefficients(DAMPING)
    m = oracle.lump_rows(M)
    dt = stable_dt()
    a0, a1, g = 1 / dt ** 2, 1 / (2 * dt), c1 / dt
    inv_d = 1 / (a0 * m + a1 * c0 * m)
    alpha = 2 * a0 * m * inv_d
    n = model.number_eq
    u = np.zeros(n); up = np.zeros(n)
    f = np.zeros(n)
    f[int(model.eq_nb_dof[boxmesh.top_centre_node(size, size, size) - 1, 1])] = -1000.0
    t0 = time.perf_counter()
    for _ in range(steps):
        un = inv_d * (f - K @ ((1 + g) * u - g * up)) + alpha * u - (alpha - 1) * up
        up, u = u, un
    t_loop = time.perf_counter() - t0
    return {"n_eq": n, "n_elem": ne, "nnz": int(K.nnz), "assembly_s": t_asm, "loop_s": t_loop, "steps": steps,
            "dof_steps_per_s": n * steps / t_loop, "elem_per_s": ne / t_asm, "checksum": float(np.abs(u).sum())}


def _ref_worker(args):
    return cpu_reference_sample(*args)


def reference_available():
    return os.path.isdir("/root/reference/scatter") or os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "scatter"))


def run_reference(args):
    """`--impl reference`: the reference's CPU path on the host; one bench step = one bounded sample (assembly of a
    `--cpu-size`^3 box with the reference's own code + `--cpu-steps` Newmark/splu steps).  The reference is single-process,
    single-threaded Python: `value` is that one process; `replicas_all_cores` adds the aggregate of one
    independent replica per host core as an upper bound of what the host could do with it."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    size, steps = args.cpu_size, args.cpu_steps
    if not reference_available():
        # the reference tree did not travel: time the numpy port instead and say so
        res = [cpu_port_cd(max(size, 24), 20) for _ in range(max(args.steps, 1))]
        r = res[-1]
        value = float(np.mean([x["dof_steps_per_s"] for x in res]))
        workload = f"hexa8 box {max(size, 24)}^3 ({r['n_eq']} DOF), numpy port of the reference path, 20 explicit steps, 1 thread"
        line = {"impl": "reference", "metric": "dof_timesteps_per_s", "value": value, "unit": "DOF*steps/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["loop_s"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": workload},
                "cpu_baseline": {"value": value, "unit": "DOF*steps/s", "cores": 1, "kind": "port", "sample": workload},
                "e2e": {"value": value, "unit": "DOF*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return
    for _ in range(max(args.warmup, 0)):
        cpu_reference_sample(min(size, 6), 5)
    res, times = [], []
    for _ in range(max(args.steps, 1)):
        t0 = time.perf_counter()
        res.append(cpu_reference_sample(size, steps))
        times.append(time.perf_counter() - t0)
    r = res[-1]
    value = float(np.mean([x["dof_steps_per_s"] for x in res]))
    cores = os.cpu_count() or 1
    rep = None
    try:
        procs = max(1, min(cores, 32))
        with mp.get_context("spawn").Pool(procs) as pool:
            rr = pool.map(_ref_worker, [(size, steps)] * procs)
        rep = {"processes": procs, "dof_steps_per_s": sum(x["n_eq"] * x["steps"] for x in rr) / max(x["loop_s"] for x in rr),
               "elem_per_s": sum(x["n_elem"] for x in rr) / max(x["assembly_s"] for x in rr)}
    except Exception as exc:
        rep = {"error": repr(exc)}
    workload = (f"hexa8 soil box {size}x{size}x{size} elements ({r['n_eq']} DOF, {r['nnz']} nnz): the reference's own ReadMesh + "
                f"GenerateMatrix assembly (unmodified code), then {steps} Newmark steps (beta 1/4, gamma 1/2, scipy splu once; "
                "SURVEY 3.3 restatement of the un-vendored PuggleSolvers default solver), 1 thread")
    port = cpu_port_cd(24, 20)
    line = {"impl": "reference", "metric": "dof_timesteps_per_s", "value": value, "unit": "DOF*steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "element_type": "hexa8", "elements": size ** 3, "integrator": "newmark_splu",
                       "dt": 5e-4, "time_steps": steps,
                       "note": "the GPU arm times 255^3 elements per GPU with the explicit scheme; the reference cannot: its assembler "
                               "runs at a few hundred elements/s and its solver factorises K-hat (BASELINE.md 3)"},
            "cpu_baseline": {"value": value, "unit": "DOF*steps/s", "cores": 1, "kind": "reference", "sample": workload,
                             "assembly_elem_per_s": float(np.mean([x["elem_per_s"] for x in res])),
                             "assembly_gbs": float(np.mean([x["assembly_gbs"] for x in res])),
                             "assembly_seconds": r["assembly_s"], "mesher_seconds": r["mesher_s"], "loop_seconds": r["loop_s"],
                             "host_cores": cores, "replicas_all_cores": rep,
                             "port": {"kind": "port", "dof_steps_per_s": port["dof_steps_per_s"], "elem_per_s": port["elem_per_s"],
                                      "sample": f"oracle/fem_np.py, 24^3 box ({port['n_eq']} DOF), 20 explicit steps, 1 thread"}},
            "e2e": {"value": value, "unit": "DOF*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, n_gpus):
    s = args.size
    return {"workload": f"hexa8 soil box {s}x{s}x{s * n_gpus} elements ({s}^3 per GPU), explicit central difference "
                        f"(lumped mass, Rayleigh damping {DAMPING}: c0 M diagonal, c1 K lagged through the SpMV), {args.stage} time steps per bench step",
            "element_type": "hexa8", "elements_per_gpu": s ** 3, "integrator": "central_difference", "dt": stable_dt(),
            "stage_steps": args.stage, "partition": f"z-slabs x{n_gpus}", "l2": "inputs (~32 GB of K values per GPU at 255^3) exceed L2; no flush needed"}


# ---------------------------------------------------------------------------------------------------------------------
# parity check of the benchmarked configuration (outside the timed region; the oracle is the checker)
# ---------------------------------------------------------------------------------------------------------------------
PARITY_BOX = 24


def _owned_rows_vs_oracle(ctx, _lib, l2g, owned_eq, geq_owned, Ko):
    """max |K_local - K_oracle| over the rows this rank owns (columns mapped to global equations) / max |K_oracle|; -1 if the
    pattern of a row differs."""
    import scipy.sparse as sp
    rowptr, col = ctx.get_pattern()
    vals = ctx.get_values(_lib.MAT_K)
    loc = sp.csr_matrix((vals, l2g[col], rowptr), shape=(len(rowptr) - 1, Ko.shape[1]))[owned_eq]
    ref = Ko[geq_owned]
    loc.sort_indices(); ref.sort_indices()
    if not (np.array_equal(loc.indptr, ref.indptr) and np.array_equal(loc.indices, ref.indices)):
        return -1.0
    return float(np.abs(loc.data - ref.data).max() / np.abs(Ko.data).max()) if len(ref.data) else 0.0


def parity_check(rank, world, local_rank, dist):
    """Collective.  Returns the dict on rank 0 (None elsewhere) and whether it passed (same on every rank)."""
    import scipy.sparse as sp
    from scatter_b200 import _lib, boxmesh, partition, system_matrix
    S = PARITY_BOX
    gm = boxmesh.box_model(S, S, S, H, "hexa8")                  # the global model is tiny: every rank builds it
    gm.connectivities()
    n_g = gm.number_eq
    geq_tab = gm.equation_table_int()
    E_g = boxmesh.lognormal_young(S ** 3, E_MEAN, E_STD, seed=7)
    top = boxmesh.top_centre_node(S, S, S) - 1
    load_geq = int(geq_tab[top, 1])
    dt_cd, n_cd, dt_nm, n_nm = stable_dt(), 40, 5e-4, 10

    def gather(dom_owned_geq, arrays):
        payload = (np.asarray(dom_owned_geq, dtype=np.int64), [np.ascontiguousarray(a) for a in arrays])
        if world == 1:
            bucket = [payload]
        else:
            bucket = [None] * world
            dist.all_gather_object(bucket, payload)
        out = []
        for k in range(len(arrays)):
            g = np.full((arrays[k].shape[0], n_g), np.nan)
            for geq, parts in bucket:
                g[:, geq] = parts[k]
            out.append(g)
        return out

    def run_domain(dom, E_loc, l2g, geq_owned, newmark):
        model = dom.model
        mx = system_matrix.GenerateMatrix(model.number_eq, 2, device=local_rank)
        ctx = mx.ctx
        ctx.set_mesh("hexa8", model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), model.number_eq,
                     dom.active if world > 1 else None)
        ctx.set_materials(E_loc, np.full(len(E_loc), NU), np.full(len(E_loc), RHO))
        if world > 1:
            uid = [_lib.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            ctx.dist_init(rank, world, uid[0])
        ctx.build_pattern()
        ctx.assemble(2, _lib.ASM_K | _lib.ASM_M_LUMPED | (_lib.ASM_M_FULL if newmark else 0))
        mx.damping_Rayleigh(DAMPING)
        if world > 1:
            ctx.set_halo(dom.neighbor_rank, dom.send_ptr, dom.send_idx, dom.recv_ptr, dom.recv_idx)
        k_rel = _owned_rows_vs_oracle(ctx, _lib, l2g, dom.owned_eq, geq_owned, Ko_box[0])
        nt = max(n_cd, n_nm) + 2
        ptr = np.zeros(nt + 1, dtype=np.int64); dofs = np.zeros(0, dtype=np.int64); vals = np.zeros(0)
        hit = np.where(geq_owned == load_geq)[0]
        if len(hit):
            ptr = np.arange(nt + 1, dtype=np.int64)
            dofs = np.full(nt, dom.owned_eq[hit[0]], dtype=np.int64)
            vals = -1000.0 * np.minimum(1.0, np.arange(nt) / 4.0)
        ctx.set_load_schedule(ptr, dofs, vals)
        ctx.set_state(None, None)
        u, v, _, _ = ctx.run_central_difference(dt_cd, 0, n_cd, 10)
        res = [u[:, dom.owned_eq], v[:, dom.owned_eq]]
        if world > 1:
            # the optional overlap of the halo exchange with the interior tiles must give the same history bit for bit
            ctx.set_option("halo_overlap", 1)
            ctx.set_state(None, None)
            u2, v2, _, st2 = ctx.run_central_difference(dt_cd, 0, n_cd, 10)
            if not (np.array_equal(u2, u) and np.array_equal(v2, v)):
                raise RuntimeError("halo overlap changed the central-difference history")
            ctx.set_option("halo_overlap", 0)
        if newmark:
            ctx.set_state(None, None)
            u, v, _, _ = ctx.run_newmark(dt_nm, 0, n_nm, 5, rtol=1e-12)
            res += [u[:, dom.owned_eq], v[:, dom.owned_eq]]
        ctx.close()
        return k_rel, res

    # oracle (rank 0 computes the histories; every rank needs K for its own row check -- 24^3 assembles in seconds)
    oracle = _oracle()
    om = oracle.model_from_readmesh(gm)
    Ko, Mo = oracle.assemble_global(om, E_g, np.full(S ** 3, NU), np.full(S ** 3, RHO), 2)
    Ko = sp.csr_matrix(Ko); Mo = sp.csr_matrix(Mo)
    Ko_box = [Ko]
    # --- slab partition (the benchmark's) -------------------------------------------------------------------------------------
    per = S // world
    dom = partition.slab_partition(S, S, per, rank, world, H, "hexa8")
    p0 = rank * per
    p1 = (rank + 1) * per + (1 if rank == world - 1 else 0)
    z0, z1 = max(p0 - 1, 0), min(p1, S)
    npl = (S + 1) * (S + 1)
    g_rows = np.arange(len(dom.model.nodes)) + z0 * npl
    leq = dom.model.equation_table_int()
    l2g = np.zeros(dom.model.number_eq, dtype=np.int64)
    l2g[leq[leq >= 0]] = geq_tab[g_rows][leq >= 0]
    owned_mask = (leq >= 0) & (dom.active[:, None] == 1)
    geq_owned = geq_tab[g_rows][owned_mask]
    k_slab, res_slab = run_domain(dom, E_g[z0 * S * S:z1 * S * S], l2g, geq_owned, True)
    g_slab = gather(geq_owned, res_slab)
    # --- recursive coordinate bisection (general partitioner), N > 1 ------------------------------------------------------
    k_rcb, g_rcb = None, None
    if world > 1:
        owner = partition.owner_by_rcb(gm, world)
        dom2 = partition.partition_model(gm, owner, rank)
        leq2 = dom2.model.equation_table_int()
        l2g2 = np.zeros(dom2.model.number_eq, dtype=np.int64)
        l2g2[leq2[leq2 >= 0]] = geq_tab[dom2.global_nodes][leq2 >= 0]
        rows_g = gm.node_rows()
        sel = np.where((owner == rank)[rows_g].any(axis=1))[0]     # the elements partition_model kept, in the same order
        k_rcb, res_rcb = run_domain(dom2, E_g[sel], l2g2, dom2.global_eq_of_owned, False)
        g_rcb = gather(dom2.global_eq_of_owned, res_rcb)
    ks = [k_slab] + ([k_rcb] if k_rcb is not None else [])
    if world > 1:
        allk = [None] * world
        dist.all_gather_object(allk, ks)
        ks = [max(x[i] if x[i] >= 0 else np.inf for x in allk) for i in range(len(ks))]
    out, ok = None, True
    if rank == 0:
        c0, c1 = oracle.rayleigh_coefficients(DAMPING)
        ml = oracle.lump_rows(Mo)

        def force(t):
            f = np.zeros(n_g)
            f[load_geq] = -1000.0 * min(1.0, t / 4.0)
            return f
        Ucd, Vcd, _, _ = oracle.central_difference(sp.diags(ml), sp.diags(ml) * c0 + Ko * c1, Ko, force, np.arange(n_cd + 1) * dt_cd, 10, c1=c1)
        Unm, Vnm, _, _ = oracle.newmark(Mo, Mo * c0 + Ko * c1, Ko, force, np.arange(n_nm + 1) * dt_nm, 5)

        def rel(a, b):
            return float(np.linalg.norm(a - b) / np.linalg.norm(b)) if not np.isnan(a).any() else float("inf")
        out = {"n_ranks": world, "box": f"{S}^3 hexa8 ({n_g} DOF), z-slabs x{world}", "k_rel": ks[0],
               "cd_rel_l2": max(rel(g_slab[0], Ucd), rel(g_slab[1], Vcd)), "cd_steps": n_cd,
               "nm_rel_l2": max(rel(g_slab[2], Unm), rel(g_slab[3], Vnm)), "nm_steps": n_nm, "nm_pcg_rtol": 1e-12,
               "tolerances": {"k_rel": 1e-12, "history_rel_l2": 1e-8}}
        ok = 0 <= out["k_rel"] <= 1e-12 and out["cd_rel_l2"] <= 1e-8 and out["nm_rel_l2"] <= 1e-8
        if g_rcb is not None:
            out["rcb"] = {"partition": f"recursive coordinate bisection x{world}", "k_rel": ks[1],
                          "cd_rel_l2": max(rel(g_rcb[0], Ucd), rel(g_rcb[1], Vcd))}
            ok = ok and 0 <= ks[1] <= 1e-12 and out["rcb"]["cd_rel_l2"] <= 1e-8
        out["passed"] = bool(ok)
    if world > 1:
        flag = [ok]
        dist.broadcast_object_list(flag, src=0)
        ok = flag[0]
    return out, ok


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
    t0 = time.perf_counter()
    ctx.set_mesh("hexa8", model.nodes[:, 1:], rows, eq, model.number_eq, dom.active if world > 1 else None)
    ctx.set_materials(E, np.full(ne, NU), np.full(ne, RHO))
    torch.cuda.synchronize()
    t_h2d = time.perf_counter() - t0
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
    total_steps = (args.steps + args.warmup + 8) * args.stage * 2 + 16
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
    halo_s = max_over_ranks(st.get("seconds_halo", 0.0))
    total_dof = sum_over_ranks(float(n_owned))
    n_ts = args.steps * args.stage
    value = total_dof * n_ts / wall
    kernel_s = dev_s / n_ts                                    # one fused SpMV+update launch per time step
    achieved = step_bytes / kernel_s / 1e9

    # --- end-to-end stages through the solver object (host buffers) -----------------------------------------------------
    # one pinned pool of output rows, reused by every e2e variant (cudaMallocHost of ~10 GB takes seconds)
    n_eq = model.number_eq
    pool_rows = max(args.steps + max(args.warmup, 3) + 1, 10)
    pool = [_lib.pinned_zeros((pool_rows, n_eq)) for _ in range(3)]

    def make_solver(interval, n_steps_total, output_dofs=None):
        num = solvers.CentralDifferenceSolver()
        num.output_interval = interval
        num.resume_on_device = False                           # every stage starts from host rows, like the reference protocol
        time_arr = np.arange(0, n_steps_total + 1) * dt
        num.output_dofs = output_dofs
        num.number_equations = n_eq
        num.time = time_arr
        idx = np.arange(0, len(time_arr), interval)
        if idx[-1] != len(time_arr) - 1:
            idx = np.append(idx, len(time_arr) - 1)
        num.output_time_indices, num.output_time = idx, time_arr[idx]
        ncol = n_eq if output_dofs is None else len(output_dofs)
        assert len(idx) * ncol <= pool[0].size
        num.u, num.v, num.a = (p.reshape(-1)[:len(idx) * ncol].reshape(len(idx), ncol) for p in pool)
        num.u0 = pool[0][0] if output_dofs is None else np.zeros(n_eq)
        num.v0 = pool[1][0] if output_dofs is None else np.zeros(n_eq)
        num.bind(mx)
        num.load_schedule = (ptr, dofs, vals)
        return num

    def time_stages(num, stage_steps, n_warm, n_timed):
        def stage(k):
            if num.output_dofs is None:
                num.update(k * stage_steps)                    # u0, v0 <- stored host row (restart hook, scatter.py:158)
            else:
                num.resume_on_device = True                    # only selected columns come back: stages continue on the device
                if k == 0:
                    num.update(0)
            num.calculate(None, None, None, None, k * stage_steps, (k + 1) * stage_steps)
            return float(num.u[-1, 0])                         # device->host result read
        for k in range(n_warm):
            stage(k)
        barrier()
        w = time.perf_counter()
        for k in range(n_warm, n_warm + n_timed):
            stage(k)
        barrier()
        return max_over_ranks(time.perf_counter() - w)

    n_warm = max(args.warmup, 3)
    num = make_solver(args.stage, (n_warm + args.steps) * args.stage)
    e2e_wall = time_stages(num, args.stage, n_warm, args.steps)
    e2e_value = total_dof * n_ts / e2e_wall
    h2d = 2 * 8 * n_eq                                         # u0, v0 (the load schedule is uploaded once, outside)
    d2h = 3 * 8 * n_eq * (1 + 1)                               # output rows at both ends of the stage (t0 and t0 + stage)
    e2e_by = {str(args.stage): {"dof_steps_per_s": e2e_value, "ms_per_time_step": 1e3 * e2e_wall / n_ts,
                                "d2h_bytes_per_time_step": d2h / args.stage, "h2d_bytes_per_stage": h2d}}
    if world == 1 and args.e2e_sweep:
        for interval, stage_steps, n_st in ((10, 40, 1), (1, 4, 1)):
            num = make_solver(interval, (1 + n_st) * stage_steps)
            wsec = time_stages(num, stage_steps, 1, n_st)
            e2e_by[str(interval)] = {"dof_steps_per_s": total_dof * n_st * stage_steps / wsec, "ms_per_time_step": 1e3 * wsec / (n_st * stage_steps),
                                     "d2h_bytes_per_time_step": 3 * 8 * n_eq * (stage_steps // interval + 1) / stage_steps,
                                     "h2d_bytes_per_stage": h2d}
        # output selection: the y-displacements of the 1000 free nodes nearest to the load leave the GPU at EVERY time step
        sel = np.unique(eq[:, 1][eq[:, 1] >= 0])[-1000:].astype(np.int64)
        num = make_solver(1, 3 * 100, output_dofs=sel)
        if True:
            wsec = time_stages(num, 100, 1, 2)
            e2e_by["1_selected_dofs"] = {"dof_steps_per_s": total_dof * 200 / wsec, "ms_per_time_step": 1e3 * wsec / 200,
                                         "selected_dofs": int(len(sel)), "d2h_bytes_per_time_step": 3 * 8 * len(sel)}
        ctx.set_output_dofs(None)
    del num

    # --- parity of this rank count against the oracle (outside every timed region) -----------------------------------------
    parity, parity_ok = (None, True)
    if args.parity:
        parity, parity_ok = parity_check(rank, world, local_rank, dist)

    # --- CPU baseline (rank 0, N = 1 only): the reference's own code on a bounded sample ---------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            if reference_available():
                r = cpu_reference_sample(args.cpu_size, args.cpu_steps)
                port = cpu_port_cd(24, 20)
                cpu = {"value": r["dof_steps_per_s"], "unit": "DOF*steps/s", "cores": 1, "kind": "reference",
                       "sample": f"{args.cpu_size}^3-element hexa8 box ({r['n_eq']} DOF): the reference's own ReadMesh + GenerateMatrix "
                                 f"(unmodified code, {r['elem_per_s']:.0f} elem/s, {r['assembly_s']:.1f} s) and {r['steps']} Newmark/splu steps "
                                 f"({r['loop_s']:.1f} s; SURVEY 3.3 restatement of the un-vendored solver), single thread",
                       "assembly_elem_per_s": r["elem_per_s"], "assembly_gbs": r["assembly_gbs"], "host_cores": os.cpu_count(),
                       "port": {"kind": "port", "dof_steps_per_s": port["dof_steps_per_s"], "elem_per_s": port["elem_per_s"],
                                "sample": f"oracle/fem_np.py (numpy restatement), 24^3 box ({port['n_eq']} DOF), 20 explicit steps, 1 thread"}}
            else:
                r = cpu_port_cd(40, 20)
                cpu = {"value": r["dof_steps_per_s"], "unit": "DOF*steps/s", "cores": 1, "kind": "port",
                       "sample": f"40^3-element hexa8 box ({r['n_eq']} DOF), {r['steps']} explicit steps with scipy CSR SpMV after numpy "
                                 f"assembly at {r['elem_per_s']:.0f} elem/s; oracle/fem_np.py (baseline/_ref not present)",
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

    info0 = ctx.device_info()
    secondary = scatter_e2e = config5 = newmark_multi = None
    extra = (rank == 0 and world == 1 and (args.secondary or args.scatter_e2e)) or (world > 1 and (args.config5 or args.nm_size))
    if extra:
        del pool
        ctx.close()
    if rank == 0 and world == 1 and args.scatter_e2e:
        try:
            scatter_e2e = run_scatter_e2e(args, local_rank)
        except Exception as exc:
            scatter_e2e = {"error": repr(exc)}
    if rank == 0 and world == 1 and args.secondary:
        try:
            secondary = run_secondary_newmark(args, local_rank)
        except Exception as exc:                      # the headline line must survive a failure of the secondary workload
            secondary = {"error": repr(exc)}
    if world > 1 and args.config5:
        try:
            config5 = run_config5(args, rank, world, local_rank, dist, barrier, max_over_ranks, sum_over_ranks)
        except Exception as exc:
            config5 = {"error": repr(exc)}
    if world > 1 and args.nm_size:
        try:
            newmark_multi = run_newmark_multi(args, rank, world, local_rank, dist, barrier, max_over_ranks, sum_over_ranks)
        except Exception as exc:
            newmark_multi = {"error": repr(exc)}
    if rank == 0:
        info = info0
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
                "halo_seconds_per_step": halo_s / max(args.stage, 1),
                "halo_overlapped_with_interior_tiles": bool(st.get("halo_overlapped", False)) if world > 1 else None,
                "assembly": {"seconds": t_asm, "pattern_seconds": t_pattern, "gbs": asm_bytes / t_asm / 1e9,
                             "frac_hbm": asm_bytes / t_asm / 1e9 / peak, "elements_per_s": ne / t_asm, "algorithmic_bytes": asm_bytes,
                             "host_mesh_seconds": t_mesh, "h2d_seconds": t_h2d,
                             # the kernels are FP64 / issue bound, not HBM bound: modelled FP64 operation count
                             # of k_elem_records + k_assemble_tma per hexa8 element against 64 FMA/clk/SM at the maximum SM clock
                             "kernels": "k_elem_records + k_assemble_tma (persistent, TMA-fed)",
                             "fma_per_element": ASM_FMA_PER_HEXA8,
                             "frac_fp64_peak": ne * ASM_FMA_PER_HEXA8 / t_asm / (64.0 * info["sm_count"] * sm_max_hz)},
                "random_field": rf,
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "DOF*steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1e3 * e2e_wall / args.steps, "output_interval": args.stage},
                "e2e_by_output_interval": e2e_by,
                "parity_check": parity,
                "options": dict(a.split("=") for a in args.option) or None,
                "gpu_launches": int(launches), "clocks": clocks, "device": info["name"], "secondary": secondary,
                "scatter_e2e": scatter_e2e, "config5": config5, "newmark_weak_scaling": newmark_multi}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if not parity_ok:
        raise SystemExit(3)


def run_scatter_e2e(args, local_rank):
    """The whole `scatter(...)` pipeline (scatter.py:36-171) on the benchmark box, from arrays to the pickle file, stage by
    stage with host wall-clock times: what a user of the entry point waits for."""
    import tempfile
    from scatter_b200 import boxmesh
    from scatter_b200.scatter import Pipeline, Solver
    s = args.size
    n_steps = 2 * args.stage
    dt = stable_dt()
    out = tempfile.mkdtemp(prefix="scatter_b200_e2e_")
    t = {}
    t0 = time.perf_counter()
    model = boxmesh.box_model(s, s, s, H, "hexa8")
    t["mesh_arrays_bc_numbering"] = time.perf_counter() - t0
    top = boxmesh.top_centre_node(s, s, s)
    mats = {"solid": {"density": RHO, "Young": E_MEAN, "poisson": NU}}
    sett = {"int_order": 2, "damping": DAMPING, "absorbing_BC": [1, 1], "absorbing_BC_stiff": 1e3, "pickle": True,
            "pickle_nodes": [top, top - 1, top + 1], "VTK": False, "VTK_binary": True, "output_interval": args.stage}
    load = {"force": [0, -1000, 0], "node": [top], "time": n_steps * dt, "type": "heaviside", "ini_steps": 5}
    run = Pipeline(mats, boxmesh.box_boundaries(s, s, s, H), sett, load, dt, Solver.CENTRAL_DIFFERENCE, False, local_rank)
    for name, fn in (("mesh_stage", lambda: run.mesh(model)), ("matrices", run.matrices), ("solver_init", run.solver),
                     ("loads", run.loads), ("integrate", run.integrate), ("export", lambda: run.export(out))):
        t0 = time.perf_counter()
        res = fn()
        t[name] = time.perf_counter() - t0
    stats = run.numerical.stats[-1]
    n = model.number_eq
    total = sum(t.values())
    run.matrix.ctx.close()
    return {"workload": f"scatter(...) from arrays: hexa8 {s}^3, central difference, {n_steps} steps, output every {args.stage}, pickle of 3 nodes",
            "dof": n, "seconds": t, "seconds_total": total, "time_loop_device_seconds": stats["seconds_device"],
            "dof_steps_per_s_whole_call": n * n_steps / total, "dof_steps_per_s_integrate_stage": n * n_steps / t["integrate"],
            "matrices_detail": getattr(run.matrix, "timings", None),
            "top_displacement": float(res.dis[-1, int(model.eq_nb_dof[top - 1, 1])])}


def run_config5(args, rank, world, local_rank, dist, barrier, max_over_ranks, sum_over_ranks):
    """BASELINE.json config 5 as named: ONE hexa8 box of `--total`^3 elements (405^3 = 200.8 M DOF) cut into N z-slabs --
    fixed total size (strong scaling over N = 2, 4, 8; it does not fit one GPU), explicit central difference."""
    from scatter_b200 import _lib, boxmesh, partition, system_matrix
    T = args.total
    t0 = time.perf_counter()
    dom = partition.slab_partition(T, T, None, rank, world, H, "hexa8", nz_total=T)
    model = dom.model
    ne = len(model.elem)
    E = boxmesh.lognormal_young(ne, E_MEAN, E_STD, seed=405 + rank)
    t_mesh = time.perf_counter() - t0
    mx = system_matrix.GenerateMatrix(model.number_eq, 2, device=local_rank)
    ctx = mx.ctx
    uid = [_lib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.dist_init(rank, world, uid[0])
    ctx.set_mesh("hexa8", model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), model.number_eq, dom.active)
    ctx.set_materials(E, np.full(ne, NU), np.full(ne, RHO))
    nnz = ctx.build_pattern()
    t_asm = ctx.assemble(2, _lib.ASM_K | _lib.ASM_M_LUMPED)
    mx.damping_Rayleigh(DAMPING)
    ctx.set_halo(dom.neighbor_rank, dom.send_ptr, dom.send_idx, dom.recv_ptr, dom.recv_idx)
    n_owned = int(len(dom.owned_eq))
    dt = stable_dt()
    stage, n_st = args.stage, 3
    total_steps = (n_st + 3) * stage + 8
    # load on the first owned free y-dof of rank world // 2 (a point load somewhere inside the box)
    ptr = np.zeros(total_steps + 1, dtype=np.int64); dofs = np.zeros(0, dtype=np.int64); vals = np.zeros(0)
    if rank == world // 2:
        d = int(dom.owned_eq[len(dom.owned_eq) // 2])
        ptr = np.arange(total_steps + 1, dtype=np.int64); dofs = np.full(total_steps, d, dtype=np.int64)
        vals = -1000.0 * np.minimum(1.0, np.arange(total_steps) / 4.0)
    ctx.set_load_schedule(ptr, dofs, vals)
    ctx.set_state(None, None)
    tc = 0
    for _ in range(2):
        ctx.run_central_difference(dt, tc, stage, stage, store=False); tc += stage
    barrier()
    w0 = time.perf_counter()
    dev = 0.0
    for _ in range(n_st):
        _, _, _, st = ctx.run_central_difference(dt, tc, stage, stage, store=False); tc += stage
        dev += st["seconds_device"]
    barrier()
    wall = max_over_ranks(time.perf_counter() - w0)
    dev = max_over_ranks(dev)
    halo = max_over_ranks(st.get("seconds_halo", 0.0))
    total_dof = sum_over_ranks(float(n_owned))
    u_norm = sum_over_ranks(float(np.abs(ctx.get_state()[0][dom.owned_eq]).sum()))
    ctx.close()
    return {"workload": f"hexa8 box {T}^3 elements ({int(total_dof)} DOF) over {world} z-slabs, explicit central difference, fixed total size",
            "scaling": "strong", "dof_total": total_dof, "dof_per_gpu_max": max_over_ranks(float(n_owned)), "nnz_rank0": nnz,
            "dof_timesteps_per_s": total_dof * n_st * stage / wall, "ms_per_time_step": 1e3 * wall / (n_st * stage),
            "device_ms_per_time_step": 1e3 * dev / (n_st * stage), "halo_seconds_per_step": halo / stage,
            "assembly_seconds": t_asm, "host_mesh_seconds": t_mesh, "checksum_abs_u": u_norm}


def run_newmark_multi(args, rank, world, local_rank, dist, barrier, max_over_ranks, sum_over_ranks):
    """Implicit path on several GPUs (weak scaling): hexa8 box of --nm-size^3 elements per GPU cut into z-slabs, Newmark + PCG
    (FSAI built per rank on its own rows, projection, true-residual check; one halo exchange per product, one all-reduce
    per pair of dots) at rtol 1e-12 -- the tolerance the N-rank parity check of this run uses."""
    from scatter_b200 import _lib, boxmesh, partition, system_matrix
    s = args.nm_size
    dom = partition.slab_partition(s, s, s, rank, world, H, "hexa8")
    model = dom.model
    ne = len(model.elem)
    E = boxmesh.lognormal_young(ne, E_MEAN, E_STD, seed=77 + rank)
    mx = system_matrix.GenerateMatrix(model.number_eq, 2, device=local_rank)
    ctx = mx.ctx
    uid = [_lib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.dist_init(rank, world, uid[0])
    ctx.set_mesh("hexa8", model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), model.number_eq, dom.active)
    ctx.set_materials(E, np.full(ne, NU), np.full(ne, RHO))
    ctx.build_pattern()
    ctx.assemble(2, _lib.ASM_K | _lib.ASM_M_FULL)
    mx.damping_Rayleigh(DAMPING)
    ctx.set_halo(dom.neighbor_rank, dom.send_ptr, dom.send_idx, dom.recv_ptr, dom.recv_idx)
    n_owned = int(len(dom.owned_eq))
    nst, dt = 10, 5e-4
    total = nst + 8
    ptr = np.zeros(total + 1, dtype=np.int64); dofs = np.zeros(0, dtype=np.int64); vals = np.zeros(0)
    if rank == world // 2:
        d = int(dom.owned_eq[len(dom.owned_eq) // 2])
        ptr = np.arange(total + 1, dtype=np.int64); dofs = np.full(total, d, dtype=np.int64)
        vals = -1000.0 * np.minimum(1.0, np.arange(total) / 4.0)
    ctx.set_load_schedule(ptr, dofs, vals)
    ctx.set_state(None, None)
    ctx.run_newmark(dt, 0, 2, 1, rtol=1e-12, store=False)
    barrier()
    w0 = time.perf_counter()
    _, _, _, st = ctx.run_newmark(dt, 2, nst, 1, rtol=1e-12, store=False)
    barrier()
    wall = max_over_ranks(time.perf_counter() - w0)
    total_dof = sum_over_ranks(float(n_owned))
    pre = ctx.precond_info()
    ctx.close()
    return {"workload": f"hexa8 box {s}x{s}x{s * world} elements ({int(total_dof)} DOF) over {world} z-slabs, Newmark + PCG (rtol 1e-12), dt {dt}",
            "scaling": "weak", "dof_total": total_dof, "dof_timesteps_per_s": total_dof * nst / wall, "ms_per_time_step": 1e3 * wall / nst,
            "pcg_iterations_per_step": st["pcg_iterations"] / nst, "last_residual": st["last_residual"],
            "fsai_entries_rank0": pre["fsai_nnz"], "projection_vectors": pre["projection_vectors"]}


def run_secondary_newmark(args, local_rank):
    """BASELINE.json config 4: structured hexa20 box, ~10 M DOF, Newmark (beta=1/4, gamma=1/2) with PCG on one B200, at the
    PCG tolerance whose parity is checked right here on a 12^3 box against the oracle's direct solve."""
    from scatter_b200 import _lib, boxmesh, system_matrix
    rtol = args.rtol20
    dt = 5e-4
    # --- parity at the benchmarked tolerance and with the PCG driver the big box uses (stream-ordered, not the cooperative kernel)
    parity = None
    try:
        import scipy.sparse as sp
        oracle = _oracle()
        sp_ = 12
        pm = boxmesh.box_model(sp_, sp_, sp_, H, "hexa20", hexa20_order=args.order20); pm.connectivities()
        pne = len(pm.elem)
        pE = boxmesh.lognormal_young(pne, E_MEAN, E_STD, seed=20)
        Ko, Mo = oracle.assemble_global(oracle.model_from_readmesh(pm), pE, np.full(pne, NU), np.full(pne, RHO), 2)
        c0, c1 = oracle.rayleigh_coefficients(DAMPING)
        pn = pm.number_eq
        pd = int(pm.eq_nb_dof[boxmesh.top_centre_node(sp_, sp_, sp_, model=pm, h=H) - 1, 1])
        nst = 20

        def force(t):
            f = np.zeros(pn); f[pd] = -1000.0 * min(1.0, t / 4.0)
            return f
        Uo, Vo, _, _ = oracle.newmark(Mo, Mo * c0 + Ko * c1, Ko, force, np.arange(nst + 1) * dt, 5)
        pmx = system_matrix.GenerateMatrix(pn, 2, device=local_rank)
        pctx = pmx.ctx
        pctx.set_option("small_pcg", 0)
        pctx.set_mesh("hexa20", pm.nodes[:, 1:], pm.node_rows(), pm.equation_table_int(), pn, None)
        pctx.set_materials(pE, np.full(pne, NU), np.full(pne, RHO))
        pctx.build_pattern()
        pctx.assemble(2, _lib.ASM_K | _lib.ASM_M_FULL)
        pmx.damping_Rayleigh(DAMPING)
        kd = float(np.abs(pctx.get_values(_lib.MAT_K) - sp.csr_matrix(Ko).data).max() / np.abs(Ko.data).max())
        pctx.set_load_schedule(np.arange(nst + 2, dtype=np.int64), np.full(nst + 1, pd, dtype=np.int64), -1000.0 * np.minimum(1.0, np.arange(nst + 1) / 4.0))
        pctx.set_state(None, None)
        u, v, _, pst = pctx.run_newmark(dt, 0, nst, 5, rtol=rtol)
        rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
        parity = {"box": f"{sp_}^3 hexa20 ({pn} DOF)", "k_rel": kd, "nm_rel_l2": max(rel(u, Uo), rel(v, Vo)), "nm_steps": nst, "pcg_rtol": rtol,
                  "pcg_iterations_per_step": pst["pcg_iterations"] / nst, "passed": bool(kd <= 1e-12 and max(rel(u, Uo), rel(v, Vo)) <= 1e-8)}
        pctx.close()
    except Exception as exc:
        parity = {"error": repr(exc)}
    s = args.size20
    t0 = time.perf_counter()
    model = boxmesh.box_model(s, s, s, H, "hexa20", hexa20_order=args.order20)
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
    nsteps = args.steps20
    total = nsteps * 3 + 8
    d = int(model.eq_nb_dof[boxmesh.top_centre_node(s, s, s, model=model, h=H) - 1, 1])
    ramp = np.ones(total); ramp[:5] = np.linspace(0, 1, 5)
    ctx.set_load_schedule(np.arange(total + 1, dtype=np.int64), np.full(total, d, dtype=np.int64), -1000.0 * ramp)
    ctx.set_state(None, None)
    _, _, _, st0 = ctx.run_newmark(dt, 0, 2, 1, rtol=rtol, store=False)        # warm-up (also builds Khat and its FSAI factor)
    _, _, _, st = ctx.run_newmark(dt, 2, nsteps, 1, rtol=rtol, store=False)
    its = st["pcg_iterations"] / max(nsteps, 1)
    pre = ctx.precond_info()
    # bytes per step: rhs two-matrix SpMV (values of M and K, CSR columns once) + per PCG iteration one SpMV (node-blocked
    # index: one column list per node), the two FSAI factor products (8 bytes per entry of G and of G^T, row pointers) and
    # 16 vector passes (SpMV x/y, fused update, preconditioner in/out, direction update)
    ps = ctx.pattern_stats()
    idx_bytes = ps["node_col_entries"] * 4 + ps["n_nodes"] * 24 if ps["node_blocked"] else nnz * 4 + n * 8
    rhs_bytes = nnz * 20 + n * 8 * 12
    fsai_bytes = 2 * (pre["fsai_nnz"] * 8 + n * 8) if pre["fsai_nnz"] else 0
    it_bytes = nnz * 8 + idx_bytes + fsai_bytes + n * 8 * (16 if pre["fsai_nnz"] else 14)
    step_bytes = rhs_bytes + its * it_bytes
    sec = st["seconds_device"] / nsteps
    # second line: the loosest tolerance that still keeps a 12^3 hexa20 history within 1e-8 of the direct solve over 1000
    # steps (4.3e-9 at 1e-11, 3.2e-8 at 1e-10: profiles/r2_hexa20_rtol_probe_12cube_1000steps.txt); same matrix and factor
    loose = None
    if rtol < 1e-11:
        try:
            ctx.set_state(None, None)
            ctx.run_newmark(dt, 0, 2, 1, rtol=1e-11, store=False)
            _, _, _, stl = ctx.run_newmark(dt, 2, nsteps, 1, rtol=1e-11, store=False)
            loose = {"pcg_rtol": 1e-11, "ms_per_time_step": 1e3 * stl["seconds_device"] / nsteps,
                     "dof_timesteps_per_s": n * nsteps / stl["seconds_device"], "pcg_iterations_per_step": stl["pcg_iterations"] / nsteps}
        except Exception as exc:
            loose = {"error": repr(exc)}
    peak, _ = measured_peak()
    out = {"workload": f"hexa20 soil box {s}^3 elements ({args.order20} node numbering), Newmark + PCG (rtol {rtol:g}), dt {dt}", "dof": n, "nnz": nnz,
           "node_numbering": args.order20, "column_dictionary_patterns": ps.get("dict_patterns"),
           "dof_timesteps_per_s": n / sec, "ms_per_time_step": 1e3 * sec, "pcg_iterations_per_step": its, "pcg_rtol": rtol,
           "pcg_stagnations": st.get("pcg_stagnations", 0),
           "preconditioner": ({"kind": "FSAI (G^T G, FP32 factor on a filtered lower pattern) + projection onto previous solutions",
                               "fsai_entries": pre["fsai_nnz"], "fsai_entries_per_matrix_entry": pre["fsai_nnz"] / nnz,
                               "fsai_setup_seconds": st0.get("fsai_setup_seconds", 0.0), "projection_vectors": pre["projection_vectors"]}
                              if pre["fsai_nnz"] else {"kind": "Jacobi"}),
           "roofline": {"bound": "hbm", "achieved": step_bytes / sec / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": step_bytes / sec / 1e9 / peak, "bytes_per_step": step_bytes},
           "assembly": {"seconds": t_asm, "elements_per_s": ne / t_asm, "pattern_seconds": t_pat, "host_mesh_seconds": t_mesh},
           "last_residual": st["last_residual"], "parity_check": parity, "at_loosest_parity_tolerance": loose}
    ctx.close()
    return out


# FP64 operations the assembly spends per hexa8 element (order 2): per Gauss point 121 for J, J^-1, detJ -- once per element
# (k_elem_records; k_assemble_blk of round 1 repeated it in each of the ~4 blocks that see the element) -- and 84 in each
# of the 16 pair lanes of k_assemble_tma, plus the 8 x 8 x 9 ordered additions of the row gather
ASM_FMA_PER_HEXA8 = 8 * (121 + 16 * 84) + 576

# dram__bytes_read.sum + dram__bytes_write.sum of one fused central-difference launch (k_spmv_node<2,..> with the column
# dictionary) from the committed ncu capture, by box size
TRAFFIC = {255: 34492233000 + 788973056}      # profiles/r2_v7_k_spmv_node_ng2_255cube.txt (1 GPU; lagged damping: two more vector passes)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=255, help="elements per box edge per GPU")
    ap.add_argument("--stage", type=int, default=100, help="time steps per bench step = output interval of the e2e arm "
                                                             "(the reference's run scripts store every 100th step, run_scatter_rose_2D.py:19)")
    ap.add_argument("--cpu-size", type=int, default=12, help="box edge of the bounded CPU sample (the reference assembles ~350 elements/s)")
    ap.add_argument("--cpu-steps", type=int, default=200, help="Newmark steps of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--parity", type=int, default=1, help="run the oracle parity check of this rank count (outside the timed region)")
    ap.add_argument("--e2e-sweep", type=int, default=1, help="e2e at output intervals 10 and 1 and with an output selection (N = 1)")
    ap.add_argument("--scatter-e2e", type=int, default=1, help="time the whole scatter(...) call at --size (N = 1 only)")
    ap.add_argument("--secondary", type=int, default=1, help="also run the hexa20 Newmark/PCG workload (N = 1 only)")
    ap.add_argument("--config5", type=int, default=1, help="N > 1: also run BASELINE config 5 as named (--total^3 box, fixed size)")
    ap.add_argument("--nm-size", type=int, default=128, help="N > 1: hexa8 elements per box edge per GPU of the Newmark weak-scaling block (0: skip)")
    ap.add_argument("--total", type=int, default=405, help="box edge of config 5 (405^3 elements = 200.8 M DOF)")
    ap.add_argument("--random-field", type=int, default=1, help="also time the random-field sampler on this rank's elements")
    ap.add_argument("--size20", type=int, default=94, help="hexa20 box edge (elements) of the secondary workload")
    ap.add_argument("--steps20", type=int, default=30)
    ap.add_argument("--order20", default="interleaved", choices=["interleaved", "grouped"],
                    help="node numbering of the synthetic hexa20 box: cell by cell (translation invariant, like the hexa8 lattice) or "
                         "corner lattice followed by the three edge lattices (round 1)")
    ap.add_argument("--rtol20", type=float, default=1e-12, help="PCG tolerance of the hexa20 workload (history error 4.7e-10 after 1000 steps on the 12^3 probe, profiles/r2_hexa20_rtol_probe_12cube_1000steps.txt; the solver classes default to 1e-14)")
    ap.add_argument("--option", action="append", default=[], metavar="NAME=VALUE",
                    help="library option for A/B runs (sc_set_option), e.g. halo_overlap=0; the defaults are the product path")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.option:
            from scatter_b200 import _lib
            for kv in args.option:
                k, v = kv.split("=")
                _lib.DEFAULT_OPTIONS[k] = int(v)
        run_ours(args)


if __name__ == "__main__":
    main()
