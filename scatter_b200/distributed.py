"""`scatter(...)` on several GPUs of one node -- one process per GPU (launch with torchrun), no counterpart in the reference.

    torchrun --nproc-per-node 4 --master-addr 127.0.0.1 run_case.py      # run_case.py calls scatter_distributed(...)

Every rank reads the mesh and numbers the equations exactly like the serial run (cheap, whole-array numpy), takes its
part of a node-based domain decomposition (`partition.owner_by_rcb` / `owner_by_slabs` + `partition_model`: all elements
touching an owned node, ghost nodes for the rest), assembles and integrates its own rows on its GPU with one NCCL halo
exchange per matrix-vector product, and sends the stored rows of its own dofs to rank 0, which writes the same
`data.pickle` / VTK files as the serial entry point.  Host-side plumbing (NCCL id, result gather) goes through a gloo
group of `torch.distributed`; the data path between GPUs is the library's own NCCL communicator (`sc_dist_init`).

Tests: the pieces (partitioning, halo plan, schedule localisation, result gather) run in the world_size 2/3 gloo tests on
CPU; the entry point as a whole is compared with the single-domain oracle on two GPUs in
`tests/test_gpu_multirank.py::test_scatter_distributed_matches_serial` (Newmark and central difference, absorbing faces).
"""
from __future__ import annotations

import os
import tempfile
import types

import numpy as np

from . import _lib, export_results, partition
from .scatter import Pipeline, Solver, _SOLVER_CLASSES


def gather_histories(dom: partition.LocalDomain, fields, n_global_eq: int, group=None, dst: int = 0):
    """Collect (n_out, n_local_eq) histories of every rank into (n_out, n_global_eq) arrays on rank `dst` (others get None).
    Only the columns a rank owns travel; every global equation is owned by exactly one rank."""
    import torch.distributed as dist
    mine = [np.ascontiguousarray(f[:, dom.owned_eq]) if f is not None else None for f in fields]
    payload = (np.asarray(dom.global_eq_of_owned, dtype=np.int64), mine)
    rank = dist.get_rank(group)
    bucket = [None] * dist.get_world_size(group) if rank == dst else None
    dist.gather_object(payload, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    out = []
    for k, f in enumerate(fields):
        if f is None:
            out.append(None)
            continue
        g = np.full((f.shape[0], n_global_eq), np.nan)
        for geq, parts in bucket:
            g[:, geq] = parts[k]
        if np.isnan(g).any():
            raise RuntimeError("result gather left equations without an owner (inconsistent domain decomposition)")
        out.append(g)
    return out


def scatter_distributed(mesh_file: str, outfile_folder: str, materials: dict, boundaries: dict, inp_settings: dict, loading: dict,
                        time_step: float = 0.1, solver: Solver = Solver.NEWMARK_EXPLICIT, random_props=False,
                        partitioner: str = "rcb"):
    """Same arguments and outputs as `scatter_b200.scatter` (rank 0 returns the `export_results.Write` object, the other
    ranks None).  RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT come from the launcher."""
    import torch
    import torch.distributed as dist
    from . import validator
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    validator.ValidateLoad.validate(loading)
    if world == 1:
        from .scatter import scatter
        return scatter(mesh_file, outfile_folder, materials, boundaries, inp_settings, loading, time_step, solver, random_props,
                       device=local_rank)
    if solver == Solver.STATIC:
        raise NotImplementedError("the static solver is not domain-decomposed")
    torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        dist.init_process_group("gloo")
    host = None if dist.get_backend() == "gloo" else dist.new_group(backend="gloo")

    run = Pipeline(materials, boundaries, inp_settings, loading, time_step, solver, random_props, local_rank)
    # the mesh, its numbering and the random field are computed identically on every rank (deterministic kernels)
    run.mesh(mesh_file).random_field(outfile_folder if rank == 0 else tempfile.mkdtemp(prefix="scatter_b200_rf_"))
    model = run.model
    owner = partition.owner_by_rcb(model, world) if partitioner == "rcb" else partition.owner_by_slabs(model, world)
    dom = partition.partition_model(model, owner, rank)
    loc = dom.model

    # matrices of the rank's rows
    from . import system_matrix
    mx = system_matrix.GenerateMatrix(loc.number_eq, inp_settings["int_order"], device=local_rank)
    uid = [_lib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0, group=host)
    mx.ctx.dist_init(rank, world, uid[0])
    explicit = solver == Solver.CENTRAL_DIFFERENCE
    mx.want_full_mass, mx.want_lumped_mass = not explicit, explicit
    mx.generate_stiffness_and_mass(loc, run.materials, active=dom.active)
    mx.absorbing_boundaries(loc, run.materials, inp_settings["absorbing_BC"], inp_settings["absorbing_BC_stiff"], owned_rows=dom.owned_eq)
    mx.damping_Rayleigh(inp_settings["damping"])
    mx.ctx.set_halo(dom.neighbor_rank, dom.send_ptr, dom.send_idx, dom.recv_ptr, dom.recv_idx)

    # time axis, solver object of the local system, loads compiled on the global numbering and restricted to owned dofs
    total = loading["time"]
    time = np.linspace(0, total, int(np.ceil(total / time_step) + 1))
    num = _SOLVER_CLASSES[solver]()
    num.output_interval = inp_settings.get("output_interval", 1)
    num.initialise(loc.number_eq, time)
    num.bind(mx)
    from . import force_external
    force = force_external.Force()
    top = model.get_top_surface() if loading["type"] == "moving_at_plane" else []
    force.initialise_load(loading, time, model, num, top_surface_elements=top)
    num.load_schedule = partition.localise_schedule(dom, *force.compile_schedule())
    num.update(0)
    num.calculate(None, None, None, None, 0, len(time) - 1)

    fields = gather_histories(dom, (num.u, num.v, num.a), model.number_eq, group=host)
    dist.barrier(group=host)
    if rank != 0:
        return None
    glob = types.SimpleNamespace(u=fields[0], v=fields[1], a=fields[2], output_time=num.output_time, time=num.time)
    res = export_results.Write(outfile_folder, model, run.materials, glob)
    res.pickle(write=inp_settings["pickle"], nodes=inp_settings["pickle_nodes"])
    res.vtk(write=inp_settings["VTK"], binary=inp_settings["VTK_binary"], output_interval=1)
    return res
