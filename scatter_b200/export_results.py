"""Result export -- same output layout as the reference's `scatter/export_results.py:10-213`.

* `Write.data` = {"time", "nodes", "position", "displacement"|"velocity"|"acceleration": {str(node): {"x","y"[,"z"]}}}
  with zeros for fixed dofs (`export_results.py:68-103`); built lazily (the reference builds O(Nn*dim) Python objects
  eagerly, which is unusable beyond ~1e6 nodes).
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
from collections import defaultdict

import numpy as np

_VTK_PERM = {"hexa8": list(range(8)), "hexa20": list(range(20)), "quad4": list(range(4)), "tri3": list(range(3)),
             "tri6": list(range(6)), "tetra4": list(range(4)), "tetra10": [0, 1, 2, 3, 4, 5, 6, 7, 9, 8],
             "quad8": list(range(8))}
_VTK_CELL = {"hexa8": 12, "hexa20": 25, "quad4": 9, "quad8": 23, "tri3": 5, "tri6": 22, "tetra4": 10, "tetra10": 24}


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
        self.elements = model.elem[:, self.idx_vtk] - 1
        self.time = numerical.output_time
        self.dis = numerical.u
        self.vel = numerical.v
        self.acc = numerical.a
        self.mat = model.materials
        self.mat_idx = model.materials_index
        self.materials = materials
        self.bc = model.BC
        self.n_dim = model.dimension
        self._data = None

    # ------------------------------------------------------------------------------------------------------------
    @property
    def data(self) -> dict:
        if self._data is None:
            self._data = {}
            self.parse_data()
        return self._data

    def nodal_field(self, field: np.ndarray) -> np.ndarray:
        """(n_out, Nn, dim) array of a result field with zeros at fixed dofs."""
        eq = self.eq_nb_dof
        free = ~np.isnan(eq)
        out = np.zeros((field.shape[0],) + eq.shape)
        out[:, free] = field[:, eq[free].astype(np.int64)]
        return out

    def parse_data(self) -> None:
        labels = ["x", "y", "z"][:self.n_dim]
        d = {"time": self.time, "nodes": list(map(int, self.nodes)), "position": self.coordinates,
             "displacement": defaultdict(dict), "velocity": defaultdict(dict), "acceleration": defaultdict(dict)}
        zeros = np.zeros(len(self.time))
        for name, field in (("displacement", self.dis), ("velocity", self.vel), ("acceleration", self.acc)):
            target = d[name]
            for i, nid in enumerate(self.nodes):
                key = str(int(nid))
                for j, lab in enumerate(labels):
                    dof = self.eq_nb_dof[i][j]
                    target[key][lab] = zeros if np.isnan(dof) else field[:, int(dof)]
        self._data.update(d)

    def pickle(self, name="data", write=True, nodes="all") -> None:
        if not write:
            return
        if nodes != "all":
            idx = [self.data["nodes"].index(int(i)) for i in nodes]
            data = {"time": self.data["time"], "nodes": nodes, "position": [self.data["position"][i] for i in idx],
                    "displacement": defaultdict(dict), "velocity": defaultdict(dict), "acceleration": defaultdict(dict)}
            for n in nodes:
                for key in ("displacement", "velocity", "acceleration"):
                    data[key].update({str(n): self.data[key][str(n)]})
        else:
            data = self.data
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
        dis = self.nodal_field(self.dis)
        vel = self.nodal_field(self.vel)
        folder = os.path.join(self.output_folder, "VTK")
        os.makedirs(folder, exist_ok=True)
        for output_t in range(int(len(self.time) / output_interval)):
            t = int(output_t * output_interval)
            d3 = np.zeros((len(self.nodes), 3)); v3 = np.zeros((len(self.nodes), 3))
            d3[:, :self.n_dim] = dis[t]
            v3[:, :self.n_dim] = vel[t]
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
