"""Multi-rank host logic on CPU (gloo, world_size 2): domain decomposition + halo plan.

Each rank partitions the reference's cube mesh, assembles its local rows with the oracle, and advances the explicit
central-difference recurrence with numpy, exchanging ghost values through torch.distributed (gloo) exactly along the
halo plan the CUDA path hands to NCCL.  The gathered result must equal the single-domain oracle run."""
import os
import socket
import sys

import numpy as np
import pytest

import cases
from conftest import GOLDEN, ROOT, load_oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _global_model(mesh_path):
    from scatter_b200 import mesher
    m = mesher.ReadMesh(mesh_path)
    m.read_gmsh(); m.read_bc(cases.BC_CUBE); m.mapping(); m.connectivities()
    return m


def _worker(rank, world, port, mesh_path, out_dir):
    import torch.distributed as dist
    import torch
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = load_oracle()
    from scatter_b200 import partition
    m = _global_model(mesh_path)
    owner = partition.owner_by_slabs(m, world, axis=2)
    dom = partition.partition_model(m, owner, rank)
    loc = dom.model
    om = oracle.model_from_readmesh(loc)
    ne = len(loc.elem)
    E, nu, rho = np.full(ne, 10e6), np.full(ne, 0.2), np.full(ne, 1500.0)
    K, M = oracle.assemble_global(om, E, nu, rho, 2)
    leq = loc.equation_table_int()
    act_rows = np.zeros(loc.number_eq, dtype=bool)
    act_rows[dom.owned_eq] = True
    ml = oracle.lump_rows(M)
    dt = 2e-4
    a0 = 1 / dt ** 2
    inv_d = np.where(act_rows, 1 / (a0 * np.where(ml != 0, ml, 1.0)), 0.0)
    n = loc.number_eq
    u = np.zeros(n); up = np.zeros(n)
    # load on global node 8 (owner rank only)
    f = np.zeros(n)
    grow = int(np.where(m.nodes[:, 0] == 8)[0][0])
    if owner[grow] == rank:
        lrow = int(np.where(dom.global_nodes == grow)[0][0])
        f[leq[lrow, 1]] = -1000.0

    def halo(x):
        reqs, bufs = [], []
        for k, nb in enumerate(dom.neighbor_rank):
            snd = torch.from_numpy(np.ascontiguousarray(x[dom.send_idx[dom.send_ptr[k]:dom.send_ptr[k + 1]]]))
            rcv = torch.zeros(int(dom.recv_ptr[k + 1] - dom.recv_ptr[k]), dtype=torch.float64)
            reqs.append(dist.isend(snd, int(nb))); reqs.append(dist.irecv(rcv, int(nb)))
            bufs.append((k, rcv, snd))
        for r in reqs:
            r.wait()
        for k, rcv, _ in bufs:
            x[dom.recv_idx[dom.recv_ptr[k]:dom.recv_ptr[k + 1]]] = rcv.numpy()

    for _ in range(40):
        un = np.where(act_rows, inv_d * (f - K @ u) + 2 * u - up, 0.0)
        halo(un)
        up, u = u, un
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), u=u[dom.owned_eq], geq=dom.global_eq_of_owned,
             n_ghost=int((dom.active == 0).sum()), n_send=len(dom.send_idx), n_recv=len(dom.recv_idx))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_decomposition_reproduces_single_domain(golden_meshes, tmp_path):
    import torch.multiprocessing as mp
    oracle = load_oracle()
    world = 2
    mesh_path = golden_meshes["cube.msh"]
    mp.spawn(_worker, args=(world, _free_port(), mesh_path, str(tmp_path)), nprocs=world, join=True)
    # single-domain reference
    m = _global_model(mesh_path)
    om = oracle.model_from_readmesh(m)
    ne = len(m.elem)
    K, M = oracle.assemble_global(om, np.full(ne, 10e6), np.full(ne, 0.2), np.full(ne, 1500.0), 2)
    ml = oracle.lump_rows(M)
    dt = 2e-4
    inv_d = 1 / (ml / dt ** 2)
    n = m.number_eq
    u = np.zeros(n); up = np.zeros(n); f = np.zeros(n)
    f[int(m.eq_nb_dof[int(np.where(m.nodes[:, 0] == 8)[0][0]), 1])] = -1000.0
    for _ in range(40):
        un = inv_d * (f - K @ u) + 2 * u - up
        up, u = u, un
    got = np.full(n, np.nan)
    for r in range(world):
        d = np.load(os.path.join(tmp_path, f"rank{r}.npz"))
        assert d["n_ghost"] > 0 and d["n_send"] > 0 and d["n_recv"] > 0
        assert np.isnan(got[d["geq"]]).all()            # every dof owned exactly once
        got[d["geq"]] = d["u"]
    assert not np.isnan(got).any()
    assert np.abs(got - u).max() <= 1e-12 * np.abs(u).max()


def test_slab_partition_matches_generic_partition():
    """The direct slab builder (no global mesh) gives the same local meshes / halo plan as the generic partitioner."""
    from scatter_b200 import boxmesh, partition
    nx, ny, nzp, world = 3, 4, 2, 3
    g = boxmesh.box_model(nx, ny, nzp * world, 0.5, "hexa8")
    g.connectivities()
    npl = (nx + 1) * (ny + 1)
    plane = np.arange(len(g.nodes)) // npl
    owner = np.minimum(plane // nzp, world - 1).astype(np.int32)
    for r in range(world):
        a = partition.slab_partition(nx, ny, nzp, r, world, 0.5)
        b = partition.partition_model(g, owner, r)
        assert np.array_equal(a.model.nodes[:, 1:], b.model.nodes[:, 1:])
        assert np.array_equal(a.model.elem, b.model.elem)
        assert np.array_equal(a.active, b.active)
        assert np.array_equal(a.model.equation_table_int(), b.model.equation_table_int())
        assert np.array_equal(a.neighbor_rank, b.neighbor_rank)
        assert np.array_equal(a.send_idx, b.send_idx) and np.array_equal(a.recv_idx, b.recv_idx)
        assert np.array_equal(a.send_ptr, b.send_ptr) and np.array_equal(a.recv_ptr, b.recv_ptr)
