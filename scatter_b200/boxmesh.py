"""Synthetic structured soil boxes (hexa8 / hexa20) -- the benchmark inputs of BASELINE.json configs 4 and 5.

Node id = 1 + i + (nx+1) (j + (ny+1) k) with x fastest; hexa8 connectivity in gmsh order
(`scatter/element_types.py:10-21`), hexa20 mid-edge nodes appended after the corner lattice in gmsh edge order
(`element_types.py:119-131`).  y is the vertical axis as in the reference (`scatter/scatter.py:41-44`).

`box_model` returns a `ReadMesh`-shaped object built from arrays (no file); `write_box_msh` emits the same mesh as
gmsh 2.2 ASCII for small sizes so that the file reader sees the identical model.  `z_range` restricts the box to a slab
of element layers [k0, k1) -- used by the domain decomposition so that no rank ever materialises the global mesh.
"""
from __future__ import annotations

import numpy as np

from . import gmsh_io
from .mesher import ReadMesh

_HEX20_EDGES = [(0, 1), (0, 3), (0, 4), (1, 2), (1, 5), (2, 3), (2, 6), (3, 7), (4, 5), (4, 7), (5, 6), (6, 7)]
_CORNER_OFF = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])


def box_arrays(nx: int, ny: int, nz: int, h: float = 0.5, element_type: str = "hexa8", z_range=None, hexa20_order: str = "grouped"):
    """-> nodes (Nn,4) [id,x,y,z], elem (Ne,nne) 1-based ids, with ids local to the slab `z_range` = (k0, k1).

    Everything is generated in (k, j, i) C order, which is the id order (x fastest), so no scatter is needed.
    `hexa20_order`: "grouped" numbers the corner lattice first and the x-, y-, z-edge nodes after it (three more lattices with
    their own strides); "interleaved" numbers cell by cell -- corner, then the x-, y-, z-edge node the cell owns -- which
    makes the numbering translation invariant in the interior (what the hexa8 lattice is by construction): neighbours
    stay close in memory and the relative column lists of the node-blocked SpMV repeat (column dictionary)."""
    k0, k1 = (0, nz) if z_range is None else z_range
    nzl = k1 - k0
    NX, NY, NZ = nx + 1, ny + 1, nzl + 1

    def lattice(nk, nj, ni, off=(0.0, 0.0, 0.0)):
        """coordinates of an (nk, nj, ni) lattice, x fastest; `off` shifts it by a fraction of h along (x, y, z)"""
        out = np.empty((nk, nj, ni, 3))
        out[..., 0] = (np.arange(ni) + off[0]) * h
        out[..., 1] = ((np.arange(nj) + off[1]) * h)[:, None]
        out[..., 2] = ((np.arange(nk) + k0 + off[2]) * h)[:, None, None]
        return out.reshape(-1, 3)

    n_corner = NX * NY * NZ
    ne = nx * ny * nzl
    off8 = np.array([(dk * NY + dj) * NX + di for di, dj, dk in _CORNER_OFF], dtype=np.int64)
    if element_type == "hexa8":
        # first corner of element (ek, ej, ei), in id order, by broadcasting (no index grids), and the 1-based ids in the
        # same pass; coordinates written straight into the node table
        first = ((np.arange(nzl, dtype=np.int64) * NY)[:, None, None] + np.arange(ny, dtype=np.int64)[None, :, None]) * NX \
            + np.arange(nx, dtype=np.int64)[None, None, :]
        elem = first.reshape(-1, 1) + (off8 + 1)[None, :]
        nodes = np.empty((n_corner, 4))
        nodes[:, 0] = np.arange(1, n_corner + 1)
        grid = nodes.reshape(NZ, NY, NX, 4)
        grid[..., 1] = np.arange(NX) * h
        grid[..., 2] = (np.arange(NY) * h)[:, None]
        grid[..., 3] = ((np.arange(NZ) + k0) * h)[:, None, None]
        return nodes, elem
    if element_type != "hexa20":
        raise ValueError("box meshes are hexa8 or hexa20")
    # element (ek, ej, ei) in id order; its first corner on the node lattice
    ek, ej, ei = (a.ravel() for a in np.meshgrid(np.arange(nzl), np.arange(ny), np.arange(nx), indexing="ij"))
    first = (ek * NY + ej) * NX + ei
    corner = first[:, None] + off8[None, :]              # one contiguous pass (column-wise fills are strided)
    # mid-edge nodes: x-edges, then y-edges, then z-edges, each group x fastest
    nxe, nye = nx * NY * NZ, NX * ny * NZ
    base = {0: n_corner, 1: n_corner + nxe, 2: n_corner + nxe + nye}

    def edge_id(axis, i, j, k):
        if axis == 0:
            return base[0] + (k * NY + j) * nx + i
        if axis == 1:
            return base[1] + (k * ny + j) * NX + i
        return base[2] + (k * NY + j) * NX + i

    n_total = n_corner + nxe + nye + NX * NY * nzl
    nodes = np.empty((n_total, 4))
    nodes[:, 0] = np.arange(1, n_total + 1)
    nodes[:n_corner, 1:] = lattice(NZ, NY, NX)
    nodes[base[0]:base[1], 1:] = lattice(NZ, NY, nx, (0.5, 0.0, 0.0))
    nodes[base[1]:base[2], 1:] = lattice(NZ, ny, NX, (0.0, 0.5, 0.0))
    nodes[base[2]:, 1:] = lattice(nzl, NY, NX, (0.0, 0.0, 0.5))
    elem = np.empty((ne, 20), dtype=np.int64)
    elem[:, :8] = corner
    for m, (ca, cb) in enumerate(_HEX20_EDGES):
        oa, ob = _CORNER_OFF[ca], _CORNER_OFF[cb]
        axis = int(np.nonzero(ob - oa)[0][0])
        lo = np.minimum(oa, ob)
        elem[:, 8 + m] = edge_id(axis, ei + lo[0], ej + lo[1], ek + lo[2])
    if hexa20_order == "interleaved":
        # sort key: owner cell (x fastest) * 4 + slot (0 corner, 1..3 edge along x, y, z); the edge lattices are in cell order already
        key = np.empty(n_total, dtype=np.int64)
        kk, jj, ii = (a.ravel() for a in np.meshgrid(np.arange(NZ), np.arange(NY), np.arange(NX), indexing="ij"))
        key[:n_corner] = ((kk * NY + jj) * NX + ii) * 4
        kk, jj, ii = (a.ravel() for a in np.meshgrid(np.arange(NZ), np.arange(NY), np.arange(nx), indexing="ij"))
        key[base[0]:base[1]] = ((kk * NY + jj) * NX + ii) * 4 + 1
        kk, jj, ii = (a.ravel() for a in np.meshgrid(np.arange(NZ), np.arange(ny), np.arange(NX), indexing="ij"))
        key[base[1]:base[2]] = ((kk * NY + jj) * NX + ii) * 4 + 2
        kk, jj, ii = (a.ravel() for a in np.meshgrid(np.arange(nzl), np.arange(NY), np.arange(NX), indexing="ij"))
        key[base[2]:] = ((kk * NY + jj) * NX + ii) * 4 + 3
        order = np.argsort(key, kind="stable")           # new position -> old row
        new_of_old = np.empty(n_total, dtype=np.int64)
        new_of_old[order] = np.arange(n_total)
        nodes = nodes[order]
        nodes[:, 0] = np.arange(1, n_total + 1)
        elem = new_of_old[elem]
    elif hexa20_order != "grouped":
        raise ValueError("hexa20_order is 'grouped' or 'interleaved'")
    return nodes, elem + 1


def box_boundaries(nx: int, ny: int, nz: int, h: float = 0.5, bottom: str = "111") -> dict:
    """Bottom (y=0) `bottom`, x-sides roller "100", z-sides roller "001", top free (integration_test.py:513-518)."""
    X, Y, Z = nx * h, ny * h, nz * h
    return {"bottom": [bottom, [[0, 0, 0], [X, 0, 0], [0, 0, Z], [X, 0, Z]]],
            "left": ["100", [[0, 0, 0], [0, 0, Z], [0, Y, 0], [0, Y, Z]]],
            "right": ["100", [[X, 0, 0], [X, 0, Z], [X, Y, 0], [X, Y, Z]]],
            "front": ["001", [[0, 0, 0], [X, 0, 0], [0, Y, 0], [X, Y, 0]]],
            "back": ["001", [[0, 0, Z], [X, 0, Z], [0, Y, Z], [X, Y, Z]]]}


def box_model(nx: int, ny: int, nz: int, h: float = 0.5, element_type: str = "hexa8", bottom: str = "111",
              z_range=None, hexa20_order: str = "grouped") -> ReadMesh:
    nodes, elem = box_arrays(nx, ny, nz, h, element_type, z_range, hexa20_order)
    m = ReadMesh.from_arrays(nodes, elem, np.ones(len(elem), dtype=np.int64), [[3.0, 1, "solid"]], element_type)
    bc = box_boundaries(nx, ny, nz, h, bottom)
    m.read_bc(bc)
    m.mapping()
    m.prepared_bc = bc              # `Pipeline.mesh` skips its own read_bc / mapping when handed the same boundaries
    return m


def top_centre_node(nx: int, ny: int, nz: int, model=None, h: float = 0.5) -> int:
    """1-based id of the corner-lattice node at the centre of the free top surface (y = ny*h).  With a `model` whose node
    numbering is not the grouped lattice order (hexa20_order="interleaved") the node is located by its coordinates."""
    if model is not None:
        target = np.array([(nx // 2) * h, ny * h, (nz // 2) * h])
        hit = np.flatnonzero(np.abs(model.nodes[:, 1:] - target).max(axis=1) < 1e-9 * max(h, 1.0))
        return int(model.nodes[hit[0], 0])
    return 1 + nx // 2 + (nx + 1) * (ny + (ny + 1) * (nz // 2))


def lognormal_young(n_elem: int, mean: float = 30e6, std: float = 1e6, seed: int = 26021981) -> np.ndarray:
    """Per-element Young's modulus, lognormal with the given mean / standard deviation (random_fields.py:73-75)."""
    rng = np.random.default_rng(seed)
    sig2 = np.log(1.0 + (std / mean) ** 2)
    mu = np.log(mean) - 0.5 * sig2
    return np.exp(mu + np.sqrt(sig2) * rng.standard_normal(n_elem))


def write_box_msh(path: str, nx: int, ny: int, nz: int, h: float = 0.5, element_type: str = "hexa8"):
    nodes, elem = box_arrays(nx, ny, nz, h, element_type)
    gmsh_io.write_msh(path, nodes, elem, np.ones(len(elem), dtype=int), [[3, 1, "solid"]], element_type)
