"""Two-GPU parity (needs >= 2 CUDA devices; run with `gpurun --gpus 2`): domain-decomposed assembly + time loops with
NCCL halo exchange against the single-domain oracle."""
import os
import socket
import sys

import numpy as np
import pytest

import cases
from conftest import ROOT, load_oracle

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mesh_path, out_dir, transport):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scatter_b200 import _lib, mesher, partition, system_matrix
    if transport == "nccl":                    # default: halo values stored straight into the neighbour's HBM (CUDA IPC windows)
        _lib.DEFAULT_OPTIONS["peer_halo"] = 0
    m = mesher.ReadMesh(mesh_path)
    m.read_gmsh(); m.read_bc(cases.BC_CUBE); m.mapping(); m.connectivities()
    owner = partition.owner_by_slabs(m, world, axis=2)
    dom = partition.partition_model(m, owner, rank)
    loc = dom.model
    uid = [_lib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    mx = system_matrix.GenerateMatrix(loc.number_eq, 2, device=rank)
    ctx = mx.ctx
    ctx.dist_init(rank, world, uid[0])
    ne = len(loc.elem)
    mx.generate_stiffness_and_mass(loc, None, elem_props=(np.full(ne, 10e6), np.full(ne, 0.2), np.full(ne, 1500.0)), active=dom.active)
    mx.damping_Rayleigh([1, 0.01, 30, 0.01])
    ctx.set_halo(dom.neighbor_rank, dom.send_ptr, dom.send_idx, dom.recv_ptr, dom.recv_idx)
    # halo exchange hook: owned entries = global eq number, ghosts must arrive from the owner
    leq = loc.equation_table_int()
    geq_all = m.equation_table_int()[dom.global_nodes]
    x = np.full(loc.number_eq, -1.0)
    x[dom.owned_eq] = dom.global_eq_of_owned
    y = ctx.halo_exchange(x)
    free = leq >= 0
    assert np.array_equal(y[leq[free]], geq_all[free].astype(float)), "halo exchange delivered wrong ghost values"
    # the rows a rank owns equal the corresponding rows of the single-domain matrix (values and global columns)
    rowptr, col = mx.pattern()
    kv = ctx.get_values(0)
    l2g = np.full(loc.number_eq, -1, dtype=np.int64)
    l2g[leq[free]] = geq_all[free]
    out_rows = {}
    for le, ge in zip(dom.owned_eq[::7], dom.global_eq_of_owned[::7]):
        sl = slice(rowptr[le], rowptr[le + 1])
        out_rows[int(ge)] = (l2g[col[sl]], kv[sl])
    np.savez(os.path.join(out_dir, f"rows{rank}.npz"), ge=np.array(list(out_rows.keys())),
             cols=np.concatenate([v[0] for v in out_rows.values()]), vals=np.concatenate([v[1] for v in out_rows.values()]),
             ptr=np.cumsum([0] + [len(v[0]) for v in out_rows.values()]))
    # load on global node 8, owner rank only
    nt = 61
    grow = int(np.where(m.nodes[:, 0] == 8)[0][0])
    ptr = np.zeros(nt + 1, dtype=np.int64); dofs = np.zeros(0, dtype=np.int64); vals = np.zeros(0)
    if owner[grow] == rank:
        lrow = int(np.where(dom.global_nodes == grow)[0][0])
        ramp = np.ones(nt); ramp[:5] = np.linspace(0, 1, 5)
        ptr = np.arange(nt + 1, dtype=np.int64); dofs = np.full(nt, leq[lrow, 1]); vals = -1000.0 * ramp
    ctx.set_load_schedule(ptr, dofs, vals)
    out = {}
    ctx.set_state(None, None)
    u, v, a, st = ctx.run_central_difference(2e-4, 0, nt - 1, 10)
    out["cd_u"], out["cd_v"] = u[:, dom.owned_eq], v[:, dom.owned_eq]
    ctx.set_state(None, None)
    u, v, a, st = ctx.run_newmark(5e-3, 0, 20, 5)
    out["nm_u"], out["nm_v"], out["nm_a"] = u[:, dom.owned_eq], v[:, dom.owned_eq], a[:, dom.owned_eq]
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), geq=dom.global_eq_of_owned, **out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["peer_memory", "nccl"])
def test_two_gpu_time_loops_match_single_domain_oracle(transport, golden_meshes, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    oracle = load_oracle()
    world = 2
    mesh_path = golden_meshes["cube.msh"]
    mp.spawn(_worker, args=(world, _free_port(), mesh_path, str(tmp_path), transport), nprocs=world, join=True)
    from scatter_b200 import mesher
    m = mesher.ReadMesh(mesh_path)
    m.read_gmsh(); m.read_bc(cases.BC_CUBE); m.mapping(); m.connectivities()
    om = oracle.model_from_readmesh(m)
    ne = len(m.elem)
    K, M = oracle.assemble_global(om, np.full(ne, 10e6), np.full(ne, 0.2), np.full(ne, 1500.0), 2)
    c0, c1 = oracle.rayleigh_coefficients([1, 0.01, 30, 0.01])
    C = M * c0 + K * c1
    n = m.number_eq
    d = int(m.eq_nb_dof[int(np.where(m.nodes[:, 0] == 8)[0][0]), 1])

    def force(t):
        f = np.zeros(n)
        f[d] = -1000.0 * (min(t, 4) / 4.0)
        return f
    Ucd, Vcd, _, _ = oracle.central_difference(M, C, K, force, np.arange(61) * 2e-4, 10, c1=c1)
    Unm, Vnm, Anm, _ = oracle.newmark(M, C, K, force, np.arange(21) * 5e-3, 5)
    got = {k: np.full(ref.shape, np.nan) for k, ref in (("cd_u", Ucd), ("cd_v", Vcd), ("nm_u", Unm), ("nm_v", Vnm), ("nm_a", Anm))}
    for r in range(world):
        z = np.load(os.path.join(tmp_path, f"rank{r}.npz"))
        for k in got:
            got[k][:, z["geq"]] = z[k]
    Kc = K.tocsr()
    for r in range(world):
        z = np.load(os.path.join(tmp_path, f"rows{r}.npz"))
        for i, ge in enumerate(z["ge"]):
            sl = slice(z["ptr"][i], z["ptr"][i + 1])
            ref_sl = slice(Kc.indptr[ge], Kc.indptr[ge + 1])
            assert np.array_equal(z["cols"][sl], Kc.indices[ref_sl])
            assert np.abs(z["vals"][sl] - Kc.data[ref_sl]).max() <= 1e-12 * np.abs(Kc.data).max()
    for k, ref in (("cd_u", Ucd), ("cd_v", Vcd), ("nm_u", Unm), ("nm_v", Vnm), ("nm_a", Anm)):
        assert not np.isnan(got[k]).any()
        err = np.linalg.norm(got[k] - ref) / np.linalg.norm(ref)
        assert err <= (1e-8 if k != "nm_a" else 1e-7), (k, err)


def _dist_entry_worker(rank, world, port, mesh_path, out_dir, solver_name):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from scatter_b200 import Solver
    from scatter_b200.distributed import scatter_distributed
    sett = cases.settings(damping=[1, 0.01, 30, 0.01], output_interval=5, VTK=True, VTK_binary=False)
    load = {"force": [0, -1000, 0], "node": [8], "time": 0.05 if solver_name == "newmark" else 0.02, "type": "heaviside", "ini_steps": 5}
    dt = 1e-3 if solver_name == "newmark" else 2e-4
    res = scatter_distributed(mesh_path, os.path.join(out_dir, f"out_{solver_name}"), cases.materials(), cases.BC_CUBE_ABS, sett, load,
                              time_step=dt, solver=Solver.NEWMARK_EXPLICIT if solver_name == "newmark" else Solver.CENTRAL_DIFFERENCE)
    if rank == 0:
        np.savez(os.path.join(out_dir, f"dist_{solver_name}.npz"), u=res.dis, v=res.vel, a=res.acc, t=res.time)
    else:
        assert res is None
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("solver_name", ["newmark", "cd"])
def test_scatter_distributed_matches_serial(solver_name, golden_meshes, tmp_path):
    """The multi-GPU entry point (RCB partition, absorbing faces on owned rows, localised loads, gathered histories, rank-0
    export) against the single-domain oracle run of the same case."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    oracle = load_oracle()
    world = 2
    mesh_path = golden_meshes["cube.msh"]
    mp.spawn(_dist_entry_worker, args=(world, _free_port(), mesh_path, str(tmp_path), solver_name), nprocs=world, join=True)
    sett = cases.settings(damping=[1, 0.01, 30, 0.01], output_interval=5)
    load = {"force": [0, -1000, 0], "node": [8], "time": 0.05 if solver_name == "newmark" else 0.02, "type": "heaviside", "ini_steps": 5}
    dt = 1e-3 if solver_name == "newmark" else 2e-4
    _, _, (U, V, A, tt) = oracle.run_case(mesh_path, cases.materials(), cases.BC_CUBE_ABS, sett, load, dt,
                                          solver="newmark" if solver_name == "newmark" else "cd")
    z = np.load(os.path.join(tmp_path, f"dist_{solver_name}.npz"))
    assert z["u"].shape == U.shape and np.abs(U).max() > 0
    for got, ref, tol in ((z["u"], U, 1e-8), (z["v"], V, 1e-8), (z["a"], A, 1e-7)):
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= tol
    assert os.path.isfile(os.path.join(tmp_path, f"out_{solver_name}", "data.pickle"))
    assert os.path.isfile(os.path.join(tmp_path, f"out_{solver_name}", "VTK", "data_1.vtk"))
