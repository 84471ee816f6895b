"""Domain decomposition for multi-GPU runs (one process per GPU).  The reference is serial; this is
the build's own design.

Ownership is by *node*: every node (hence every matrix row) belongs to exactly one rank.  A rank holds all elements
touching one of its nodes ("ghost elements" included), so it assembles complete rows for its own dofs with no
communication at assembly time.  Nodes of those elements owned by somebody else are ghost nodes: they get local
equation numbers (so that the local CSR can address them as columns) but their rows stay empty (`active = 0` in
`sc_set_mesh`) and their vector entries are refreshed by a point-to-point halo exchange before every SpMV.

Local numbering keeps the global (node, dof) order, so local equation numbers are monotone -- the property the pattern
builder needs -- and restricted to owned rows the local matrix equals the corresponding rows of the global one.

`partition_model` works on any `ReadMesh`-shaped model given an owner array (tests, unstructured meshes);
`slab_partition` builds the z-slab of a structured box directly, without ever materialising the global mesh.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import boxmesh
from .mesher import ReadMesh


@dataclass
class LocalDomain:
    model: ReadMesh                 # local mesh with local equation numbering (ghost dofs included)
    active: np.ndarray              # (n_local_nodes,) uint8, 1 = owned
    global_nodes: np.ndarray        # (n_local_nodes,) global node rows (or -1 if unknown)
    neighbor_rank: np.ndarray       # (n_nb,) int32
    send_ptr: np.ndarray            # (n_nb+1,) int64
    send_idx: np.ndarray            # local equation numbers to send
    recv_ptr: np.ndarray
    recv_idx: np.ndarray            # local equation numbers to receive into
    owned_eq: np.ndarray            # local equation numbers of owned dofs (ascending)
    global_eq_of_owned: np.ndarray  # matching global equation numbers (or empty)
    n_global_eq: int = 0


def owner_by_slabs(model: ReadMesh, world: int, axis: int = 2) -> np.ndarray:
    """Owner rank of every node: equal-count slabs along a coordinate axis."""
    x = model.nodes[:, 1 + axis]
    order = np.argsort(x, kind="stable")
    owner = np.empty(len(x), dtype=np.int32)
    bounds = np.linspace(0, len(x), world + 1).astype(np.int64)
    for r in range(world):
        owner[order[bounds[r]:bounds[r + 1]]] = r
    # keep nodes with identical coordinate on one rank (clean planar interfaces)
    for r in range(1, world):
        cut = x[order[bounds[r]]] if bounds[r] < len(x) else np.inf
        same = np.isclose(x, cut)
        owner[same & (owner == r - 1)] = r
    return owner


def owner_by_rcb(model: ReadMesh, world: int) -> np.ndarray:
    """Owner rank of every node by recursive coordinate bisection (any world size, unstructured meshes): a group of ranks
    is split in two halves of floor / ceil size, the nodes in proportion, along the longest axis of the group's bounding
    box; ties are broken by node row, so the result is deterministic and every rank gets within one node of its share."""
    xyz = model.nodes[:, 1:4]
    owner = np.empty(len(xyz), dtype=np.int32)
    stack = [(np.arange(len(xyz)), 0, world)]
    while stack:
        idx, r0, nr = stack.pop()
        if nr == 1:
            owner[idx] = r0
            continue
        left = nr // 2
        pts = xyz[idx]
        axis = int(np.argmax(pts.max(axis=0) - pts.min(axis=0))) if len(idx) else 0
        order = np.argsort(pts[:, axis], kind="stable")
        cut = (len(idx) * left) // nr
        stack.append((idx[order[:cut]], r0, left))
        stack.append((idx[order[cut:]], r0 + left, nr - left))
    return owner


def localise_schedule(dom: "LocalDomain", step_ptr, dof, val):
    """Restrict a global load schedule (CSR over time steps of (global equation, value), `Force.compile_schedule`) to the
    dofs `dom` owns, renumbered to its local equations -- every load entry is applied by exactly one rank."""
    step_ptr = np.asarray(step_ptr, dtype=np.int64); dof = np.asarray(dof, dtype=np.int64); val = np.asarray(val, dtype=float)
    g2l = -np.ones(max(int(dom.n_global_eq), int(dof.max()) + 1 if len(dof) else 0), dtype=np.int64)
    g2l[dom.global_eq_of_owned] = dom.owned_eq
    loc = g2l[dof]
    keep = loc >= 0
    step = np.repeat(np.arange(len(step_ptr) - 1), np.diff(step_ptr))
    counts = np.bincount(step[keep], minlength=len(step_ptr) - 1)
    return np.concatenate([[0], np.cumsum(counts)]).astype(np.int64), loc[keep], val[keep]


def partition_model(model: ReadMesh, owner: np.ndarray, rank: int) -> LocalDomain:
    """Local domain of `rank` for a global model whose BC / equation numbering are already set."""
    rows = model.node_rows()
    mine = owner == rank
    elem_sel = np.where(mine[rows].any(axis=1))[0]
    local_nodes = np.unique(rows[elem_sel])                       # ascending global node rows
    g2l = -np.ones(len(model.nodes), dtype=np.int64)
    g2l[local_nodes] = np.arange(len(local_nodes))
    nodes = model.nodes[local_nodes].copy()
    nodes[:, 0] = np.arange(1, len(local_nodes) + 1)
    elem = g2l[rows[elem_sel]] + 1
    loc = ReadMesh.from_arrays(nodes, elem, np.asarray(model.materials_index)[elem_sel], model.materials, model.element_type)
    loc.BC = np.asarray(model.BC)[local_nodes]
    loc.BC_dir = np.asarray(model.BC_dir)[local_nodes]
    loc.mapping()
    loc.connectivities()
    active = mine[local_nodes].astype(np.uint8)
    geq = model.equation_table_int()[local_nodes]                 # global eq of local dofs (-1 fixed)
    leq = loc.equation_table_int()
    own_l = owner[local_nodes]
    # halo plan.  I receive the free dofs of my ghost nodes from their owners; I send the free dofs of those of my
    # nodes that are ghosts on rank s.  Node n (owned by me) is a ghost on s iff it shares an element with a node of s.
    nb_ranks = sorted(set(int(s) for s in np.unique(own_l) if s != rank))
    # which ranks touch each of my nodes: via the local elements (they contain every element touching my nodes)
    touch = {s: np.zeros(len(local_nodes), dtype=bool) for s in nb_ranks}
    lrows = elem - 1
    own_e = own_l[lrows]                                          # (ne_local, nne)
    for s in nb_ranks:
        has_s = (own_e == s).any(axis=1)
        touch[s][np.unique(lrows[has_s])] = True
    send_ptr, recv_ptr, send_idx, recv_idx = [0], [0], [], []
    for s in nb_ranks:
        snd_nodes = np.where(touch[s] & (own_l == rank))[0]       # ascending global order on both sides
        rcv_nodes = np.where(own_l == s)[0]
        sd = leq[snd_nodes].ravel(); sd = sd[sd >= 0]
        rd = leq[rcv_nodes].ravel(); rd = rd[rd >= 0]
        send_idx.append(sd); recv_idx.append(rd)
        send_ptr.append(send_ptr[-1] + len(sd)); recv_ptr.append(recv_ptr[-1] + len(rd))
    owned_mask = (leq >= 0) & (active[:, None] == 1)
    return LocalDomain(model=loc, active=active, global_nodes=local_nodes, neighbor_rank=np.array(nb_ranks, dtype=np.int32),
                       send_ptr=np.array(send_ptr, dtype=np.int64), send_idx=np.concatenate(send_idx) if send_idx else np.zeros(0, np.int64),
                       recv_ptr=np.array(recv_ptr, dtype=np.int64), recv_idx=np.concatenate(recv_idx) if recv_idx else np.zeros(0, np.int64),
                       owned_eq=leq[owned_mask], global_eq_of_owned=geq[owned_mask], n_global_eq=int(model.number_eq))


def slab_planes(nz: int, rank: int, world: int):
    """Owned node planes [p0, p1) of `rank` when the nz + 1 node planes of a box are cut into `world` slabs of (nearly) equal
    thickness: plane bounds at round(r * nz / world); the last rank also owns the final plane."""
    b0, b1 = (rank * nz) // world, ((rank + 1) * nz) // world
    return b0, b1 + (1 if rank == world - 1 else 0)


def slab_partition(nx: int, ny: int, nz_per_rank, rank: int, world: int, h: float = 0.5, element_type: str = "hexa8",
                   bottom: str = "111", nz_total: int | None = None) -> LocalDomain:
    """z-slab of the global box nx x ny x nz, nz = nz_per_rank*world (weak scaling) or `nz_total` (a fixed box; the layer
    count need not divide by the rank count): rank r owns the node planes `slab_planes(nz, r, world)` and holds one ghost
    element layer towards each neighbour (hexa8 only)."""
    if element_type != "hexa8":
        raise NotImplementedError("slab_partition builds hexa8 boxes; use partition_model for other element types")
    nz = int(nz_total) if nz_total is not None else nz_per_rank * world
    if nz < world:
        raise ValueError("fewer element layers than ranks")
    p0, p1 = slab_planes(nz, rank, world)                                  # owned planes [p0, p1)
    z0, z1 = max(p0 - 1, 0), min(p1, nz)                                   # local element layers [z0, z1)
    nodes, elem = boxmesh.box_arrays(nx, ny, nz, h, element_type, z_range=(z0, z1))
    m = ReadMesh.from_arrays(nodes, elem, np.ones(len(elem), dtype=np.int64), [[3.0, 1, "solid"]], element_type)
    m.read_bc(boxmesh.box_boundaries(nx, ny, nz, h, bottom))
    m.mapping()
    npl = (nx + 1) * (ny + 1)                                              # nodes per plane
    plane = np.arange(len(nodes)) // npl + z0                              # global plane of every local node
    active = ((plane >= p0) & (plane < p1)).astype(np.uint8)
    leq = m.equation_table_int()

    def plane_dofs(p):
        d = leq[(p - z0) * npl:(p - z0 + 1) * npl].ravel()
        return d[d >= 0]

    nb, send, recv = [], [], []
    if rank > 0:
        nb.append(rank - 1); send.append(plane_dofs(p0)); recv.append(plane_dofs(p0 - 1))
    if rank < world - 1:
        nb.append(rank + 1); send.append(plane_dofs(p1 - 1)); recv.append(plane_dofs(p1))
    sp = np.concatenate([[0], np.cumsum([len(s) for s in send])]).astype(np.int64)
    rp = np.concatenate([[0], np.cumsum([len(s) for s in recv])]).astype(np.int64)
    owned_mask = (leq >= 0) & (active[:, None] == 1)
    return LocalDomain(model=m, active=active, global_nodes=-np.ones(len(nodes), dtype=np.int64),
                       neighbor_rank=np.array(nb, dtype=np.int32), send_ptr=sp,
                       send_idx=np.concatenate(send) if send else np.zeros(0, np.int64), recv_ptr=rp,
                       recv_idx=np.concatenate(recv) if recv else np.zeros(0, np.int64), owned_eq=leq[owned_mask],
                       global_eq_of_owned=np.zeros(0, np.int64))
