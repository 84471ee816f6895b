"""External loads -- host mirror of the reference's `scatter/force_external.py` `Force` class.

The reference recomputes a dense force vector in a Python callback every time step (`scatter.py:151`,
`force_external.py:55-74`).  Here the same loads are *compiled once* into a per-step sparse schedule
(step_ptr, dof, value) that lives on the device (`sc_set_load_schedule`), which removes the per-step host round trip.
`update_load_at_t(t)` is kept (same name, same dense result) for code that drives a solver by hand.

Restated from: pulse `force_external.py:76-85, 250-264`; heaviside `:87-97, 266-279`; moving `:99-149, 320-345`
(including the `x * l` interpolation quirk at `:335-336`); moving_at_plane `:151-212, 281-318` (point-in-polygon done
with a convex-polygon test instead of shapely).
"""
from __future__ import annotations

import sys

import numpy as np


class Force:
    def __init__(self):
        self.force_vector = []
        self.factor = None
        self.contact_nodes = None
        self.step_factors = None

    # ------------------------------------------------------------------------------------------------------------
    def initialise_load(self, load_set, time, model, solver, **kwargs):
        self.nb_equations = model.number_eq
        self.factor = load_set.get("force")
        self.contact_nodes = load_set.get("node")
        self.eq_nb_dof = model.eq_nb_dof
        self.model_nodes = model.nodes
        self.time = time
        self.steps = load_set["ini_steps"]
        self.loading_type = load_set["type"]
        self.solver = solver
        self._ids = model.nodes[:, 0].astype(np.int64)
        self._row_of = None
        if self.loading_type in ("pulse", "heaviside", "moving", "moving_at_plane"):
            if len(self.time) <= self.steps:
                sys.exit("Error: Number of loading steps smaller than " + str(self.steps))
        if self.loading_type == "pulse":
            s = self.steps
            self.step_factors = np.append(np.linspace(0, 1, int((s - 1) / 2), endpoint=False), np.linspace(1, 0, int((s + 1) / 2)))
        elif self.loading_type in ("heaviside", "moving", "moving_at_plane"):
            self.step_factors = np.ones(len(self.time))
            self.step_factors[:self.steps] = np.linspace(0, 1, self.steps)
            if self.loading_type == "moving":
                self._init_moving(load_set["speed"])
            elif self.loading_type == "moving_at_plane":
                self._init_moving_at_plane(kwargs["top_surface_elements"], load_set["speed"], load_set["direction"],
                                           load_set["start_coord"])
        elif self.loading_type == "rose":
            raise NotImplementedError("ROSE train-track coupling is outside the scope of the B200 hot path")
        else:
            sys.exit(f'Error: Load type {load_set["type"]} not supported')
        self.force_vector = self.update_load_at_t(0)

    def _node_row(self, node_id: int) -> int:
        """Row of a node id in the node table; first occurrence wins, like the reference's `list(...).index(n)`
        (force_external.py:262,275,341)."""
        ids = self._ids
        if self._row_of is None:
            if len(ids) and ids[0] == 1 and ids[-1] == len(ids) and np.array_equal(ids, np.arange(1, len(ids) + 1)):
                self._row_of = "identity"
            else:
                order = np.argsort(ids, kind="stable")
                self._row_of = (order, ids[order])
        if isinstance(self._row_of, str):
            row = int(node_id) - 1
            if not 0 <= row < len(ids):
                raise ValueError(f"{node_id} is not in list")
            return row
        order, sorted_ids = self._row_of
        k = int(np.searchsorted(sorted_ids, int(node_id), side="left"))
        if k >= len(sorted_ids) or sorted_ids[k] != int(node_id):
            raise ValueError(f"{node_id} is not in list")
        return int(order[k])

    def _init_moving(self, load_speed):
        nd = self.model_nodes
        idx = np.where(nd[:, 0] == self.contact_nodes)[0][0]
        lst = np.where((nd[:, 1] == nd[idx, 1]) & (nd[:, 2] == nd[idx, 2]))[0]
        dist = np.sqrt((nd[lst, 3] - nd[idx, 3]) ** 2)
        self.idx_list = lst[np.argsort(dist)]
        self.node_distances = np.sort(dist)
        speed = np.ones(len(self.time)) * load_speed
        speed[:self.steps] = 0
        self.load_distances = speed * (self.time - self.time[self.steps])

    def _init_moving_at_plane(self, xz_plane_elements, load_speed, load_direction, start_coord):
        from scipy.spatial import cKDTree
        self._plane_elems = np.asarray(xz_plane_elements)
        coords = self.model_nodes[self._plane_elems - 1, 1:][:, :, [0, 2]]        # (nf, 4, 2)
        dt = np.diff(self.time[self.steps:])
        distance = np.zeros(len(self.time))
        distance[self.steps:] = np.append(0, np.cumsum(load_speed * dt))
        if np.isclose(load_direction[0], 0):
            angle = 0.5 * np.pi if load_direction[1] > 0 else -0.5 * np.pi
        else:
            angle = np.arctan(load_direction[1] / load_direction[0])
        self.position = np.array([np.cos(angle) * distance + start_coord[0], np.sin(angle) * distance + start_coord[1]])
        # convex hull ordering of each quad (counter-clockwise about its centroid) for the inside test
        cen = coords.mean(axis=1)
        ang = np.arctan2(coords[:, :, 1] - cen[:, None, 1], coords[:, :, 0] - cen[:, None, 0])
        hull = np.take_along_axis(coords, np.argsort(ang, axis=1)[:, :, None], axis=1)
        tree = cKDTree(cen)
        self.active_elements = []
        for p in self.position.T:
            _, near = tree.query(p, min(10, len(cen)))
            found = None
            for k in np.atleast_1d(near):
                poly = hull[k]
                edge = np.roll(poly, -1, axis=0) - poly
                rel = p[None, :] - poly
                cross = edge[:, 0] * rel[:, 1] - edge[:, 1] * rel[:, 0]
                if (cross > 0).all():                 # strictly inside, like shapely's `contains`
                    found = self._plane_elems[k]
                    break
            self.active_elements.append(found)

    # ------------------------------------------------------------------------------------------------------------
    def _entries(self, t: int):
        """Sparse form of the force at time index t: (dofs, values); later entries overwrite earlier ones."""
        dofs, vals = [], []
        kind = self.loading_type

        def put(node_id, scale):
            row = self._node_row(node_id)
            for i, eq in enumerate(self.eq_nb_dof[row]):
                if not np.isnan(eq):
                    dofs.append(int(eq)); vals.append(float(self.factor[i]) * scale)

        if kind == "pulse":
            if t < self.steps - 1:
                for n in self.contact_nodes:
                    put(n, self.step_factors[t])
        elif kind == "heaviside":
            for n in self.contact_nodes:
                put(n, self.step_factors[t])
        elif kind == "moving":
            if not self.load_distances[t] >= np.max(self.node_distances):
                k = np.where(self.node_distances <= self.load_distances[t])[0][-1]
                nd = self.model_nodes
                pair = [int(nd[self.idx_list[k], 0]), int(nd[self.idx_list[k + 1], 0])]
                x = self.load_distances[t] - nd[self.idx_list[k], 3] + nd[self.idx_list[0], 3]
                l = self.node_distances[k + 1] - self.node_distances[k]
                shp = [1 - x / l, x * l]                # sic: the reference multiplies by l (force_external.py:335-336)
                for j, n in enumerate(pair):
                    put(n, shp[j] * self.step_factors[t])
        elif kind == "moving_at_plane":
            el = self.active_elements[t]
            xz = self.model_nodes[el - 1][:, [1, 3]]
            dist = xz - self.position[:, t]
            xd, zd = np.abs(dist[:, 0]), np.abs(dist[:, 1])
            xw = (xd < 1e-10) * 1 if np.any(xd < 1e-10) else 1 / xd
            zw = (zd < 1e-10) * 1 if np.any(zd < 1e-10) else 1 / zd
            w = xw * zw
            w = w / w.sum()
            point_load = np.array(self.factor) * self.step_factors[t]
            nodal = w[:, None].dot(point_load[None, :])
            act = self.eq_nb_dof[el - 1]
            ok = ~np.isnan(act)
            dofs.extend(act[ok].astype(int).tolist()); vals.extend(nodal[ok].tolist())
        # duplicates: the reference assigns (does not add), so the last write wins
        if len(dofs) != len(set(dofs)):
            last = {}
            for d, v in zip(dofs, vals):
                last[d] = v
            dofs, vals = list(last.keys()), list(last.values())
        return np.array(dofs, dtype=np.int64), np.array(vals, dtype=np.float64)

    def update_load_at_t(self, t, **kwargs):
        f = np.zeros(self.nb_equations)
        d, v = self._entries(int(t))
        f[d] = v
        self.force_vector = f
        return self.force_vector

    def compile_schedule(self):
        """(step_ptr, dof, val) for every time index -- the device-resident form of the callback."""
        n = len(self.time)
        if self.loading_type in ("pulse", "heaviside"):
            d0, _ = self._entries_template()
            ptr = [0]
            dofs, vals = [], []
            for t in range(n):
                if self.loading_type == "pulse" and not t < self.steps - 1:
                    ptr.append(ptr[-1]); continue
                dofs.append(d0[0]); vals.append(d0[1] * self.step_factors[t]); ptr.append(ptr[-1] + len(d0[0]))
            dofs = np.concatenate(dofs) if dofs else np.zeros(0, dtype=np.int64)
            vals = np.concatenate(vals) if vals else np.zeros(0)
            return np.array(ptr, dtype=np.int64), dofs, vals
        ptr = np.zeros(n + 1, dtype=np.int64)
        dofs, vals = [], []
        for t in range(n):
            d, v = self._entries(t)
            dofs.append(d); vals.append(v); ptr[t + 1] = ptr[t] + len(d)
        return ptr, np.concatenate(dofs), np.concatenate(vals)

    def _entries_template(self):
        """(dofs, unit values) of a fixed-node load; value(t) = unit * step_factor(t) (bitwise the same product as
        `float(factor[i]) * step_factors[t]`)."""
        dofs, vals = [], []
        for n in self.contact_nodes:
            row = self._node_row(n)
            for i, eq in enumerate(self.eq_nb_dof[row]):
                if not np.isnan(eq):
                    dofs.append(int(eq)); vals.append(float(self.factor[i]))
        last = {}
        for d, v in zip(dofs, vals):
            last[d] = v
        return (np.array(list(last.keys()), dtype=np.int64), np.array(list(last.values()), dtype=np.float64)), None
