"""Result export -- same output layout as the reference's `scatter/export_results.py:10-213`.

* `Write.data` = {"time", "nodes", "position", "displacement"|"velocity"|"acceleration": {str(node): {"x","y"[,"z"]}}}
  with zeros for fixed dofs (`export_results.py:68-103`).  The three per-node dictionaries are lazy mappings
  (`NodalHistories`): an entry is built when it is read, the reference builds O(Nn*dim) Python objects eagerly.
* `pickle()` writes `<folder>/data.pickle` (all nodes or a subset, `:105-141`).
* `vtk()` writes legacy-VTK unstructured grids `<folder>/VTK/data_<k>.vtk` (`:143-213`).  The reference delegates the
  file format to the un-vendored `vtk_tools` package; the writer below reproduces the byte layout of the reference's
  golden files (`integration_tests/results_mean/VTK/data_*.vtk`): POINTS float / CELLS / CELL_TYPES / POINT_DATA with
  VECTORS displacement, velocity, boundary_conditions (double) / CELL_DATA with SCALARS material_index and
  material_prop_<key> (double, LOOKUP_TABLE default).
"""
from __future__ import annotations

import os
import pickle
from collections.abc import Mapping

import numpy as np

_VTK_PERM = {"hexa8": list(range(8)), "hexa20": list(range(20)), "quad4": list(range(4)), "tri3": list(range(3)),
             "tri6": list(range(6)), "tetra4": list(range(4)), "tetra10": [0, 1, 2, 3, 4, 5, 6, 7, 9, 8],
             "quad8": list(range(8))}
_VTK_CELL = {"hexa8": 12, "hexa20": 25, "quad4": 9, "quad8": 23, "tri3": 5, "tri6": 22, "tetra4": 10, "tetra10": 24}


class NodalHistories(Mapping):
    """`data["displacement"]` & co. of the reference's result dictionary -- {str(node id): {"x": history, "y": ..., ["z": ...]}}
    (export_results.py:74-102) -- as a read-only mapping over the (n_out, n_eq) result array: an entry is built when it is
    asked for, so a 10^7-node run does not create 10^8 Python objects it never reads.  `to_dict()` gives the plain dict."""

    def __init__(self, write, field):
        self._w, self._field = write, field

    def __len__(self):
        return len(self._w.nodes)

    def __iter__(self):
        return (str(int(n)) for n in self._w.nodes)

    def __contains__(self, key):
        try:
            self._w.node_row(int(key))
            return True
        except (ValueError, TypeError):
            return False

    def __getitem__(self, key):
        try:
            row = self._w.node_row(int(key))
        except (ValueError, TypeError):
            raise KeyError(key) from None
        return self._w.node_entry(self._field, row)

    def to_dict(self, rows=None) -> dict:
        w = self._w
        rows = range(len(w.nodes)) if rows is None else rows
        return {str(int(w.nodes[r])): w.node_entry(self._field, r) for r in rows}


class Write:
    def __init__(self, output_folder: str, model: object, materials: dict, numerical: object) -> None:
        if not os.path.isdir(output_folder):
            os.makedirs(output_folder)
        self.element_type = model.element_type
        self.idx_vtk = _VTK_PERM[model.element_type]
        self.output_folder = output_folder
        self.nodes = model.nodes[:, 0].astype(int)
        self.eq_nb_dof = model.eq_nb_dof
        self.coordinates = model.nodes[:, 1:]
        self._model_elem = model.elem
        self._elements = None
        self.time = numerical.output_time
        self.dis = numerical.u
        self.vel = numerical.v
        self.acc = numerical.a
        # the solver may have kept only some equations (`output_dofs`): column of every equation in the stored rows, or -1
        sel = getattr(numerical, "output_dofs", None)
        self._col_of = None
        if sel is not None:
            self._col_of = -np.ones(int(model.number_eq), dtype=np.int64)
            self._col_of[np.asarray(sel, dtype=np.int64)] = np.arange(len(sel))
        self.mat = model.materials
        self.mat_idx = model.materials_index
        self.materials = materials
        self.bc = model.BC
        self.n_dim = model.dimension
        self._data = None
        self._row_lookup = None
        self._zeros = None

    @property
    def elements(self):
        """0-based element table in VTK node order (export_results.py:51) -- only the VTK writer needs it."""
        if self._elements is None:
            self._elements = np.asarray(self._model_elem)[:, self.idx_vtk] - 1
        return self._elements

    # ------------------------------------------------------------------------------------------------------------
    def node_row(self, node_id: int) -> int:
        """Row of a node id (first occurrence, like `list.index`); raises ValueError for an unknown id."""
        ids = self.nodes
        if self._row_lookup is None:
            if len(ids) and ids[0] == 1 and ids[-1] == len(ids) and np.array_equal(ids, np.arange(1, len(ids) + 1)):
                self._row_lookup = "identity"
            else:
                order = np.argsort(ids, kind="stable")
                self._row_lookup = (order, ids[order])
        if isinstance(self._row_lookup, str):
            if not 1 <= node_id <= len(ids):
                raise ValueError(f"{node_id} is not in list")
            return node_id - 1
        order, sorted_ids = self._row_lookup
        k = int(np.searchsorted(sorted_ids, node_id, side="left"))
        if k >= len(sorted_ids) or sorted_ids[k] != node_id:
            raise ValueError(f"{node_id} is not in list")
        return int(order[k])

    def node_entry(self, field: np.ndarray, row: int) -> dict:
        """{"x": history, ...} of one node: zeros for fixed dofs (export_results.py:89-102)."""
        if self._zeros is None:
            self._zeros = np.zeros(len(self.time))
        out = {}
        for j, lab in enumerate(("x", "y", "z")[:self.n_dim]):
            dof = self.eq_nb_dof[row][j]
            if np.isnan(dof):
                out[lab] = self._zeros
                continue
            col = int(dof) if self._col_of is None else int(self._col_of[int(dof)])
            if col < 0:
                raise KeyError(f"the history of node row {row} was not stored (output_dofs does not include equation {int(dof)})")
            out[lab] = field[:, col]
        return out

    @property
    def data(self) -> dict:
        if self._data is None:
            self._data = {}
            self.parse_data()
        return self._data

    def nodal_field(self, field: np.ndarray, rows=None) -> np.ndarray:
        """(n_out, Nn, dim) array of a result field with zeros at fixed dofs; `rows`: only these output rows."""
        eq = self.eq_nb_dof
        free = ~np.isnan(eq)
        cols = eq[free].astype(np.int64)
        if self._col_of is not None:
            cols = self._col_of[cols]
            if (cols < 0).any():
                raise KeyError("whole-mesh fields need full output rows (output_dofs is set)")
        src = field if rows is None else field[rows]
        out = np.zeros((src.shape[0],) + eq.shape)
        out[:, free] = src[:, cols]
        return out

    def parse_data(self) -> None:
        """The reference's result dictionary (export_results.py:68-103), with the three per-node dictionaries as lazy
        mappings (`NodalHistories`)."""
        self._data.update({"time": self.time, "nodes": self.nodes.tolist(), "position": self.coordinates,
                           "displacement": NodalHistories(self, self.dis), "velocity": NodalHistories(self, self.vel),
                           "acceleration": NodalHistories(self, self.acc)})

    def pickle(self, name="data", write=True, nodes="all") -> None:
        """`<folder>/<name>.pickle` with plain dictionaries, all nodes or a subset (export_results.py:105-141).  The subset
        path touches only the requested nodes."""
        if not write:
            return
        fields = (("displacement", self.dis), ("velocity", self.vel), ("acceleration", self.acc))
        if isinstance(nodes, str) and nodes == "all":
            data = {"time": self.time, "nodes": self.nodes.tolist(), "position": self.coordinates}
            for key, field in fields:
                data[key] = NodalHistories(self, field).to_dict()
        else:
            idx = [self.node_row(int(i)) for i in nodes]
            data = {"time": self.time, "nodes": nodes, "position": [self.coordinates[i] for i in idx]}
            for key, field in fields:
                data[key] = {str(n): self.node_entry(field, i) for n, i in zip(nodes, idx)}
        with open(os.path.join(self.output_folder, f"{name}.pickle"), "wb") as f:
            pickle.dump(data, f)

    # ------------------------------------------------------------------------------------------------------------
    def vtk(self, name="data", binary=True, write=True, output_interval=1) -> None:
        if not write:
            return
        nb_elements = len(self.elements)
        list_props = list(set([tuple(i.keys()) for i in self.materials.values()]))[0]
        tag_to_name = {int(m[1]): m[2] for m in self.mat}
        material = np.asarray(self.mat_idx, dtype=float)
        tags, inv = np.unique(np.asarray(self.mat_idx).astype(np.int64), return_inverse=True)
        material_prop = np.zeros((nb_elements, len(list_props)))
        for j, m in enumerate(list_props):
            material_prop[:, j] = np.array([self.materials[tag_to_name[int(t)]][m] for t in tags], dtype=float)[inv]
        if self.n_dim == 2:
            bc = np.zeros((self.bc.shape[0], 3))
            bc[:, :2] = self.bc
        else:
            bc = self.bc
        folder = os.path.join(self.output_folder, "VTK")
        os.makedirs(folder, exist_ok=True)
        for output_t in range(int(len(self.time) / output_interval)):
            t = int(output_t * output_interval)
            d3 = np.zeros((len(self.nodes), 3)); v3 = np.zeros((len(self.nodes), 3))
            d3[:, :self.n_dim] = self.nodal_field(self.dis, [t])[0]        # one frame at a time: no (n_out, Nn, dim) copy
            v3[:, :self.n_dim] = self.nodal_field(self.vel, [t])[0]
            w = LegacyVtk(os.path.join(folder, f"{name}_{output_t}.vtk"), f"{name}_{output_t}", binary)
            w.mesh(self.coordinates, self.elements, self.element_type)
            w.point_vectors([("displacement", d3), ("velocity", v3), ("boundary_conditions", bc)])
            w.cell_scalars([("material_index", material)] + [(f"material_prop_{m}", material_prop[:, j]) for j, m in enumerate(list_props)])
            w.save()


class LegacyVtk:
    """Minimal legacy-VTK (DataFile Version 2.0) unstructured-grid writer, ASCII or big-endian binary."""

    def __init__(self, path: str, title: str, binary: bool):
        self.path, self.title, self.binary = path, title, binary
        self.chunks = []

    def _text(self, s: str):
        self.chunks.append(s.encode())

    def _rows(self, a: np.ndarray, dtype: str):
        if self.binary:
            self.chunks.append(np.ascontiguousarray(a).astype(dtype).tobytes())
            self.chunks.append(b"\n")
        else:
            a = np.asarray(a)
            if a.ndim == 1:
                a = a[:, None]
            if np.issubdtype(a.dtype, np.integer):
                self._text("\n".join(" ".join(str(int(x)) for x in row) for row in a) + "\n")
            else:
                self._text("\n".join(" ".join(repr(float(x)) for x in row) for row in a) + "\n")

    def mesh(self, points, cells, element_type):
        n, ne, nne = len(points), len(cells), cells.shape[1]
        self._text(f"# vtk DataFile Version 2.0\n{self.title}\n{'BINARY' if self.binary else 'ASCII'}\nDATASET UNSTRUCTURED_GRID\n")
        self._text(f"POINTS {n} float\n")
        self._rows(points, ">f4")
        self._text(f"CELLS {ne} {ne * (nne + 1)}\n")
        self._rows(np.column_stack([np.full(ne, nne, dtype=np.int64), cells.astype(np.int64)]), ">i4")
        self._text(f"CELL_TYPES {ne}\n")
        self._rows(np.full(ne, _VTK_CELL[element_type], dtype=np.int64), ">i4")
        self._n, self._ne = n, ne

    def point_vectors(self, fields):
        self._text(f"POINT_DATA {self._n}\n")
        for name, arr in fields:
            self._text(f"VECTORS {name} double\n")
            self._rows(arr, ">f8")

    def cell_scalars(self, fields):
        self._text(f"CELL_DATA {self._ne}\n")
        for name, arr in fields:
            self._text(f"SCALARS {name} double\nLOOKUP_TABLE default\n")
            self._rows(np.asarray(arr, dtype=float), ">f8")

    def save(self):
        with open(self.path, "wb") as f:
            for c in self.chunks:
                f.write(c)
