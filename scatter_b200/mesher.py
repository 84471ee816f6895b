"""Mesh model feeding the hot path: gmsh 2.2 reader, boundary conditions, equation numbering.

Host-side mirror of the reference's `scatter/mesher.py` `ReadMesh` (same attribute names, same numbering, same error
strings) written with whole-array numpy operations: the reference walks nodes and elements in Python and matches node
ids with `np.where` per lookup (O(N^2), `mesher.py:323`), which cannot produce the 10^7..10^8-dof inputs the GPU path
is built for.  Semantics restated from:

* `read_gmsh`        mesher.py:55-228, utils.py:62-90 (section parser), element codes mesher.py:180-219
* `read_bc`          mesher.py:230-274, utils.py:32-59 (plane through three points, un-normalised normal)
* `mapping`          mesher.py:276-310 (file order, dof x,y,(z); 0 free, 1 fixed -> NaN, 2 absorbing)
* `connectivities`   mesher.py:312-326
* `get_mesh_edges`   mesher.py:328-391 (hexa8 only), `get_top_surface` mesher.py:393-424
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import gmsh_io

_LOWER = {"hexa8": ("quad4", 4), "hexa20": ("quad8", 8), "tetra4": ("tri3", 3), "tetra10": ("tri6", 6)}
_HEX8_SURFACES = np.array([[0, 1, 2, 3], [0, 1, 4, 5], [4, 5, 6, 7], [2, 3, 6, 7], [0, 3, 4, 7], [1, 2, 5, 6]])


class ReadMesh:
    def __init__(self, file_name: str | None) -> None:
        if file_name is not None:
            if os.path.isfile(file_name):
                self.file_name = file_name
            else:
                sys.exit("ERROR: Mesh file does not exit.")
            if os.path.splitext(self.file_name)[-1] != ".msh":
                sys.exit("ERROR: Mesh file is not a valid file")
        else:
            self.file_name = None
        self.nodes = []
        self.elem = []
        self.rose_elem = []
        self.rose_nodes = []
        self.nb_nodes_elem = []
        self.materials = []
        self.BC = []
        self.BC_dir = []
        self.number_eq = []
        self.eq_nb_dof = []
        self.eq_nb_dof_rose_nodes = []
        self.rose_eq_nb = []
        # the reference's string / per-element tables (type_BC, eq_nb_elem, type_BC_elem, boundary_elem) are derived data that
        # cost 1-10 GB on a 50 M-dof mesh: they are built on first access (the hot path itself consumes BC and eq_nb_dof)
        self._type_BC = self._eq_nb_elem = self._type_BC_elem = self._boundary_elem = None
        self._connected = False
        self.element_type = []
        self.lower_element_type = []
        self.nb_nodes_lower_elem = []
        self.materials_index = []
        self.dimension = 3
        self._node_rows = None          # (Ne, nne) 0-based rows into self.nodes

    # ---- lazily derived attributes of the reference's data model ---------------------------------------------------
    @property
    def type_BC(self):
        """"Normal" / "Fixed" / "Absorb" per (node, dof) -- mesher.py:293-305."""
        if self._type_BC is None:
            if len(self.BC) == 0 or len(self.eq_nb_dof) == 0:
                return []
            bc = np.asarray(self.BC)
            t = np.full(bc.shape, "Normal")
            t[bc == 1] = "Fixed"
            t[bc == 2] = "Absorb"
            self._type_BC = t
        return self._type_BC

    @type_BC.setter
    def type_BC(self, value):
        self._type_BC = value

    @property
    def eq_nb_elem(self):
        """Equation numbers per element, node-major (NaN = fixed) -- mesher.py:312-326."""
        if self._eq_nb_elem is None:
            if not self._connected:
                return []
            rows = self.node_rows()
            self._eq_nb_elem = self.eq_nb_dof[rows].reshape(rows.shape[0], -1)
        return self._eq_nb_elem

    @eq_nb_elem.setter
    def eq_nb_elem(self, value):
        self._eq_nb_elem = value

    @property
    def type_BC_elem(self):
        if self._type_BC_elem is None:
            if not self._connected:
                return []
            rows = self.node_rows()
            self._type_BC_elem = self.type_BC[rows].reshape(rows.shape[0], -1)
        return self._type_BC_elem

    @type_BC_elem.setter
    def type_BC_elem(self, value):
        self._type_BC_elem = value

    @property
    def boundary_elem(self):
        """Boundary faces of a hexa8 mesh (mesher.py:328-391); computed by `get_mesh_edges`, on first access at the latest."""
        if self._boundary_elem is None:
            if self.element_type != "hexa8":
                return []
            self._compute_mesh_edges()
        return self._boundary_elem

    @boundary_elem.setter
    def boundary_elem(self, value):
        self._boundary_elem = value

    # ------------------------------------------------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, nodes, elem, materials_index, materials, element_type):
        """Build the model directly from arrays (synthetic meshes too large to round-trip through a .msh file)."""
        m = cls(None)
        m._set(np.asarray(nodes, dtype=float), np.asarray(elem), np.asarray(materials_index), list(materials), element_type)
        return m

    def _set(self, nodes, elem, tags, names, element_type):
        self.element_type = element_type
        self.dimension = 2 if element_type in ("tri3", "tri6", "quad4", "quad8") else 3
        low = _LOWER.get(element_type, ([], []))
        self.lower_element_type, self.nb_nodes_lower_elem = low
        self.nodes = nodes
        self.elem = elem
        self.materials_index = tags
        self.materials = names
        self.nb_nodes_elem = elem.shape[1]

    def read_gmsh(self) -> None:
        msh = gmsh_io.read_msh(self.file_name)
        names = msh["physical_names"]
        rose_tag = next((n[1] for n in names if n[2] == "rose"), None)
        blocks = msh["elements"]          # list of (gmsh_type, tags(int array n x ntags), nodes(int array n x k)) in file order
        geo = []
        for gtype, phys, conn in blocks:
            if rose_tag is not None:
                keep = phys != rose_tag
                if (~keep).any():
                    self.rose_elem = conn[~keep] if len(self.rose_elem) == 0 else np.vstack([self.rose_elem, conn[~keep]])
                if not keep.any():
                    continue
                phys, conn = phys[keep], conn[keep]
            geo.append((gtype, phys, conn))
        if len(self.rose_elem):
            self.rose_nodes = np.unique(self.rose_elem)
        codes = {g[0] for g in geo}
        if not all(c in gmsh_io.GMSH_TO_TYPE for c in codes):
            sys.exit("ERROR: Element type not supported")
        if len(codes) != 1:
            sys.exit("ERROR: Element type not supported")
        code = next(iter(codes))
        elem = np.vstack([g[2] for g in geo])
        tags = np.concatenate([g[1] for g in geo])
        self._set(msh["nodes"], elem, tags, names, gmsh_io.GMSH_TO_TYPE[code])

    # ------------------------------------------------------------------------------------------------------------
    def read_bc(self, bc: dict) -> None:
        nn, dim = len(self.nodes), self.dimension
        self.BC = np.zeros((nn, dim), dtype=int)
        self.BC_dir = np.zeros((nn, dim), dtype=int)
        xyz = self.nodes[:, 1:]
        for boundary in bc:
            typ, pts = bc[boundary][0], bc[boundary][1]
            if dim == 3:
                p1, p2, p3 = (np.array(p, dtype=float) for p in pts[:3])
                cp = np.cross(p3 - p1, p2 - p1)
                direction = np.abs(cp / np.linalg.norm(cp))
                # plane equation c + sum_k x_k cp_k, terms added in the order k = 0, 1, 2 (c + a == a + c bit for bit, so the
                # first product doubles as the accumulator); axis-aligned planes touch one coordinate only
                c = -np.dot(cp, p3)
                ks = [k for k in range(3) if cp[k] != 0.0]
                if ks:
                    # chunk by chunk, so that the temporaries of the four passes stay in cache (17 M nodes: 0.2 -> 0.1 s per plane)
                    pieces = []
                    for lo in range(0, nn, 1 << 20):
                        blk = xyz[lo:lo + (1 << 20)]
                        r = blk[:, ks[0]] * cp[ks[0]]
                        r += c
                        for k in ks[1:]:
                            r += blk[:, k] * cp[k]
                        np.abs(r, out=r)
                        hit = np.flatnonzero(r <= 1.0e-5)
                        if len(hit):
                            pieces.append(hit + lo)
                    idx = np.concatenate(pieces) if pieces else np.zeros(0, dtype=np.intp)
                    residual = None
                else:
                    residual = np.full(nn, c)
            elif dim == 2:
                p1, p2 = np.array(pts[0], dtype=float), np.array(pts[1], dtype=float)
                vector = p2 - p1
                direction = np.array([-vector[1], vector[0]])
                residual = (np.linalg.norm(p1[None, :] - xyz, axis=1) + np.linalg.norm(p2[None, :] - xyz, axis=1)
                            - np.linalg.norm(p1 - p2))
            else:
                sys.exit(f"ERROR: dimension: {dim}, is  not supported")
            if residual is not None:
                np.abs(residual, out=residual)
                idx = np.flatnonzero(residual <= 1.0e-5)   # == np.isclose(residual, 0.0, atol=1e-5), without its temporaries
            for j, val in enumerate(typ):
                if j >= dim:
                    break
                self.BC[idx, j] = np.maximum(self.BC[idx, j], int(val))
                self.BC_dir[idx, j] = np.maximum(self.BC_dir[idx, j], abs(int(direction[j])))

    def mapping(self) -> None:
        bc = np.asarray(self.BC)
        if bc.size and (bc.min() < 0 or bc.max() > 2):
            bad = (bc < 0) | (bc > 2)
            sys.exit("Error in the boundary condition definition. \n"
                     f"{bc[bad][0]} is not a valid boundary condition.")
        fixed = (bc == 1).ravel()
        numbers = np.cumsum(~fixed, dtype=np.int64)        # 1-based equation number of every free dof
        self.number_eq = int(numbers[-1]) if numbers.size else 0
        eq = numbers.astype(float)
        eq -= 1.0
        np.putmask(eq, fixed, np.nan)
        self.eq_nb_dof = eq.reshape(bc.shape)
        self._type_BC = self._eq_nb_elem = self._type_BC_elem = None

    def node_rows(self) -> np.ndarray:
        """(Ne, nne) 0-based row of every element node in `self.nodes` (ids need not be contiguous)."""
        if self._node_rows is None:
            ids = self.nodes[:, 0].astype(np.int64)
            if np.array_equal(ids, np.arange(1, len(ids) + 1)):
                rows = np.asarray(self.elem, dtype=np.int64) - 1
            else:
                order = np.argsort(ids, kind="stable")
                pos = np.searchsorted(ids[order], self.elem)
                rows = order[pos]
            self._node_rows = rows
        return self._node_rows

    def connectivities(self) -> None:
        """mesher.py:312-326: per-element equation / BC-type tables (`eq_nb_elem`, `type_BC_elem`) -- available from now on,
        materialised on first access."""
        self.node_rows()
        self._connected = True
        self._eq_nb_elem = self._type_BC_elem = None

    def equation_table_int(self) -> np.ndarray:
        """eq_nb_dof as int64 with -1 for fixed dofs (the C ABI's representation)."""
        return np.where(np.isnan(self.eq_nb_dof), -1, self.eq_nb_dof).astype(np.int64)

    # ------------------------------------------------------------------------------------------------------------
    def get_mesh_edges(self):
        """Boundary faces of a hexa8 mesh (node ids), unique rows -- mesher.py:328-391.  The faces are only consumed by
        `get_top_surface` (moving_at_plane loads): they are computed when `boundary_elem` is first read."""
        self._boundary_elem = None

    def _compute_mesh_edges(self):
        rows = self.node_rows()
        nn = len(self.nodes)
        count = np.bincount(rows.ravel(), minlength=nn)
        is_bnd = count < 8
        nb = is_bnd[rows]                              # (Ne, 8)
        nsum = nb.sum(axis=1)
        faces = []
        one = np.where(nsum == 4)[0]                   # exactly one boundary face: the boundary nodes in local order
        if len(one):
            loc = np.argsort(~nb[one], axis=1, kind="stable")[:, :4]
            faces.append(np.take_along_axis(self.elem[one], loc, axis=1))
        many = np.where(nsum > 4)[0]
        for s in _HEX8_SURFACES:
            ok = many[nb[many][:, s].all(axis=1)] if len(many) else many
            if len(ok):
                faces.append(self.elem[ok][:, s])
        if faces:
            self.boundary_elem = np.unique(np.vstack(faces), axis=0)
        else:
            self.boundary_elem = np.zeros((0, 4), dtype=int)

    def get_top_surface(self):
        """mesher.py:393-424"""
        epsilon = 1e-10
        coords = self.nodes[:, 1:]
        c = coords[self.boundary_elem - 1, :].mean(axis=1)
        min_x, max_x = c[:, 0].min(), c[:, 0].max()
        min_y = c[:, 1].min()
        min_z, max_z = c[:, 2].min(), c[:, 2].max()
        sel = ((c[:, 0] > min_x + epsilon) & (c[:, 0] < max_x - epsilon) & (c[:, 1] > min_y + epsilon)
               & (c[:, 2] > min_z + epsilon) & (c[:, 2] < max_z - epsilon))
        return self.boundary_elem[sel]

    def rose_connectivities(self, rose_model):
        raise NotImplementedError("ROSE train-track coupling is outside the scope of the B200 hot path (SURVEY.md 2, #10)")
