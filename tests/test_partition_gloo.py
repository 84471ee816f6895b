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


def _worker(rank, world, port, mesh_path, out_dir, partitioner="slabs"):
    import torch.distributed as dist
    import torch
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = load_oracle()
    from scatter_b200 import partition
    m = _global_model(mesh_path)
    owner = partition.owner_by_slabs(m, world, axis=2) if partitioner == "slabs" else partition.owner_by_rcb(m, world)
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
    # result gather of the multi-GPU entry point: a history whose value encodes (output row, global equation)
    from scatter_b200 import distributed
    hist = np.full((3, n), -7.0)
    hist[:, dom.owned_eq] = np.arange(3)[:, None] * 1e6 + dom.global_eq_of_owned[None, :]
    got = distributed.gather_histories(dom, (hist, None, 2 * hist), m.number_eq)
    if rank == 0:
        want = np.arange(3)[:, None] * 1e6 + np.arange(m.number_eq)[None, :]
        assert got[1] is None and np.array_equal(got[0], want) and np.array_equal(got[2], 2 * want)
    else:
        assert got is None
    # the load schedule of a moving load, restricted to the rank's dofs
    from scatter_b200 import force_external
    time = np.linspace(0, 0.3, 61)
    load = {"force": [0, -1000, 0], "node": 8, "time": 0.3, "type": "moving", "speed": 10, "ini_steps": 20}
    F = force_external.Force(); F.initialise_load(load, time, m, None)
    gp, gd, gv = F.compile_schedule()
    lp, ld, lv = partition.localise_schedule(dom, gp, gd, gv)
    l2g = np.full(n, -1, dtype=np.int64); l2g[dom.owned_eq] = dom.global_eq_of_owned
    assert len(lp) == len(gp) and (l2g[ld] >= 0).all()
    np.savez(os.path.join(out_dir, f"sched{rank}.npz"), ptr=lp, dof=l2g[ld], val=lv)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,partitioner", [(2, "slabs"), (3, "rcb")])
def test_two_rank_decomposition_reproduces_single_domain(world, partitioner, golden_meshes, tmp_path):
    import torch.multiprocessing as mp
    oracle = load_oracle()
    mesh_path = golden_meshes["cube.msh"]
    mp.spawn(_worker, args=(world, _free_port(), mesh_path, str(tmp_path), partitioner), nprocs=world, join=True)
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
    # every entry of the global load schedule is applied by exactly one rank, in its time step
    from scatter_b200 import force_external
    time = np.linspace(0, 0.3, 61)
    load = {"force": [0, -1000, 0], "node": 8, "time": 0.3, "type": "moving", "speed": 10, "ini_steps": 20}
    F = force_external.Force(); F.initialise_load(load, time, m, None)
    gp, gd, gv = F.compile_schedule()
    parts = [np.load(os.path.join(tmp_path, f"sched{r}.npz")) for r in range(world)]
    assert len(gd) > 0
    for t in range(len(gp) - 1):
        want = sorted(zip(gd[gp[t]:gp[t + 1]].tolist(), gv[gp[t]:gp[t + 1]].tolist()))
        have = sorted(sum([list(zip(p["dof"][p["ptr"][t]:p["ptr"][t + 1]].tolist(), p["val"][p["ptr"][t]:p["ptr"][t + 1]].tolist()))
                           for p in parts], []))
        assert want == have


def test_rcb_is_balanced_and_deterministic(golden_meshes):
    from scatter_b200 import partition
    m = _global_model(golden_meshes["cube.msh"])
    for world in (1, 2, 3, 5, 8):
        owner = partition.owner_by_rcb(m, world)
        counts = np.bincount(owner, minlength=world)
        assert counts.sum() == len(m.nodes) and counts.max() - counts.min() <= world and counts.min() > 0
        assert np.array_equal(owner, partition.owner_by_rcb(m, world))
    # compact parts: 8 ranks on a cube -> every part spans about half of each axis
    owner = partition.owner_by_rcb(m, 8)
    ext = m.nodes[:, 1:].max(axis=0) - m.nodes[:, 1:].min(axis=0)
    for r in range(8):
        p = m.nodes[owner == r, 1:]
        assert ((p.max(axis=0) - p.min(axis=0)) <= 0.62 * ext).all()


def test_absorbing_entries_of_a_partition_cover_the_global_ones(golden_meshes, oracle):
    """Domain-decomposed absorbing boundaries: the rows a rank owns get exactly the global entries (ghost faces included)."""
    from scatter_b200 import mesher, partition, system_matrix
    m = mesher.ReadMesh(golden_meshes["cube.msh"])
    m.read_gmsh(); m.read_bc(cases.BC_CUBE_ABS); m.mapping(); m.connectivities()
    ne = len(m.elem)
    props = (np.full(ne, 30e6), np.full(ne, 0.2), np.full(ne, 1500.0))
    cg, kg = system_matrix.absorbing_entries(m, *props, 2, [1, 1], 1e3)
    assert len(cg) > 0
    world = 3
    owner = partition.owner_by_rcb(m, world)
    seen = {}
    for r in range(world):
        dom = partition.partition_model(m, owner, r)
        nl = len(dom.model.elem)
        cl, kl = system_matrix.absorbing_entries(dom.model, np.full(nl, 30e6), np.full(nl, 0.2), np.full(nl, 1500.0), 2, [1, 1], 1e3)
        leq = dom.model.equation_table_int()
        geq = m.equation_table_int()[dom.global_nodes]
        l2g = np.full(dom.model.number_eq, -1, dtype=np.int64)
        l2g[leq[leq >= 0]] = geq[leq >= 0]
        owned = set(dom.owned_eq.tolist())
        plan = system_matrix.absorbing_plan(dom.model)                                  # what the device path receives
        if plan is None:
            assert not cl
        else:
            plan.restrict_rows(dom.owned_eq)
            assert set(zip(plan.rows.tolist(), plan.cols.tolist())) == {k for k in cl if k[0] in owned}
            assert plan.grp_ptr[0] == 0 and plan.grp_ptr[-1] == len(plan.grp_entry) and (np.diff(plan.grp_ptr) > 0).all()
        for (i, j), v in cl.items():
            if i in owned:
                key = (int(l2g[i]), int(l2g[j]))
                assert key not in seen
                seen[key] = (v, kl[(i, j)])
    assert set(seen) == set(cg)
    for key, (cv, kv) in seen.items():
        assert abs(cv - cg[key]) <= 1e-13 * abs(cg[key]) and abs(kv - kg[key]) <= 1e-13 * abs(kg[key])


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
