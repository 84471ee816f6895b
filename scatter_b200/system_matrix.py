"""`GenerateMatrix` -- the reference's assembly seam (`scatter/system_matrix.py:10-376`) backed by the CUDA library.

Same constructor and method names as the reference:

    matrix = GenerateMatrix(model.number_eq, inp_settings["int_order"])
    matrix.generate_stiffness_and_mass(model, materials)
    matrix.absorbing_boundaries(model, materials, inp_settings["absorbing_BC"], inp_settings["absorbing_BC_stiff"])
    matrix.damping_Rayleigh(inp_settings["damping"])
    matrix.K, matrix.M, matrix.C           # scipy.sparse, fetched from the device on access

All matrix values live on the GPU (`matrix.ctx`); the time loop consumes them there.  `.K/.M/.C` copy them back as
scipy CSR for inspection and parity tests.  Differences to the reference that are visible through this seam:

* the matrices are CSR (the reference hands out `lil` for M and for K before `absorbing_boundaries`);
* `.K/.M/.C` carry the *structural* pattern -- every free (i,k) pair of every element, explicit zeros kept -- which is
  what `coo_matrix(...).tolil()` holds at system_matrix.py:120-121.  The reference's later scipy binops
  (system_matrix.py:375-376, :198) silently drop entries that are exactly 0.0; `pruned()` reproduces that view.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import _lib


def resolve_element_properties(data, material: dict):
    """Per-element (E, nu, rho): physical tag -> material name -> dict (system_matrix.py:52,64-71)."""
    names = {int(m[1]): m[2] for m in data.materials}
    tags, inv = np.unique(np.asarray(data.materials_index).astype(np.int64), return_inverse=True)
    E = np.array([material[names[int(t)]]["Young"] for t in tags], dtype=float)[inv]
    nu = np.array([material[names[int(t)]]["poisson"] for t in tags], dtype=float)[inv]
    rho = np.array([material[names[int(t)]]["density"] for t in tags], dtype=float)[inv]
    return E, nu, rho


def rayleigh_coefficients(damp):
    """system_matrix.py:183-196: 1/2 [[1/w1, w1],[1/w2, w2]] [c0, c1]^T = [d1, d2]^T"""
    f1, d1, f2, d2 = damp
    if f1 == f2:
        raise SystemExit('Frequencies for the Rayleigh damping are the same.')
    w1, w2 = 2 * np.pi * f1, 2 * np.pi * f2
    A = 0.5 * np.array([[1 / w1, w1], [1 / w2, w2]])
    c = np.linalg.solve(A, np.array([d1, d2], dtype=float))
    return float(c[0]), float(c[1])


# ---- absorbing boundary faces (host, O(surface)) -----------------------------------------------------------------
def _same_slope(pts) -> bool:
    def slope(p, q):
        return (q[1] - p[1]) / (q[0] - p[0]) if q[0] != p[0] else float("inf")
    first = slope(pts[0], pts[1])
    for k in range(2, len(pts)):
        if slope(pts[k - 1], pts[k]) != first:
            return False
    return True


def gmsh_face_order(points: np.ndarray) -> np.ndarray:
    """Order the nodes of a planar face like the reference does (utils.py:141-175): counter-clockwise by angle about the
    lowest point, corner nodes first, remaining (mid-side) nodes afterwards in set order."""
    pts = [np.asarray(p) for p in points]
    lowest = pts[int(np.argmin([p[1] for p in pts]))]
    ang = [np.arctan2(p[1] - lowest[1], p[0] - lowest[0]) for p in pts]
    ordered = [pts[k] for k in sorted(range(len(pts)), key=lambda k: ang[k])]
    corners = [ordered[0]]
    start = 0
    for i in range(len(ordered) - 1):
        if not _same_slope(ordered[start:i + 2]):
            corners.append(ordered[i])
            start = i
    rest = set(map(tuple, ordered)).symmetric_difference(set(map(tuple, corners)))
    corners.extend([list(t) for t in rest])
    return np.array(corners)


def _absorbing_faces(data):
    """(element, direction, node rows[nl]) of every absorbing face, in the reference's element / direction order
    (system_matrix.py:274-316): a face exists where exactly `nb_nodes_lower_elem` nodes of an element absorb in a direction."""
    # "Absorb" <=> BC code 2 (mesher.py:293-305); the integer table avoids the (Nn, dim) string array on large meshes
    absorb = (np.asarray(data.BC) == 2) if len(data.BC) else (np.asarray(data.type_BC) == "Absorb")   # (Nn, dim)
    rows = data.node_rows()
    nl, dim = data.nb_nodes_lower_elem, data.dimension
    touched = np.where(absorb.any(axis=1)[rows].any(axis=1))[0]
    out = []
    # 2-D element types have no face element (`nb_nodes_lower_elem == []`, mesher.py:187-198): the reference's test
    # `len(...) == data.nb_nodes_lower_elem` is then never true, so absorbing codes on 2-D meshes are silently ignored and
    # its "not implemented for 2D" exit (system_matrix.py:324-326) is unreachable
    if len(touched) == 0 or not isinstance(nl, (int, np.integer)):
        return out
    fl = absorb[rows[touched]]                                     # (nt, nne, dim)
    cnt = fl.sum(axis=1)                                           # (nt, dim)
    for t, d in zip(*np.where(cnt == nl)):                         # row-major: element order, then direction
        e = touched[t]
        out.append((e, d, rows[e][fl[t, :, d]]))
    return out


class AbsorbingPlan:
    """Index plan of the absorbing faces (all integer / ordering work of system_matrix.py:274-358; no arithmetic).

    faces_nodes (nf, nl) node rows in the reference's face-node order (utils.py:141-175), elem (nf,), direction (nf,),
    i1 (nf, nl) sorted equation numbers of the face dofs -- paired positionally with the face-node order, as the reference
    does -- perp (nf, nl) 1 where the dof is perpendicular to its boundary (BC_dir == 1: compression wave, else shear),
    and the grouping of the nf*nl*nl entries (face, a, b) -> key (i1[a], i1[b]) in face order: rows/cols (sorted unique
    keys), grp_ptr / grp_entry (entry ids of every key, ascending = the order in which the reference accumulates)."""

    def __init__(self, face_type, nodes, elem, direction, i1, perp, n_eq):
        self.face_type = face_type
        self.nodes = np.ascontiguousarray(nodes, dtype=np.int32)
        self.elem = np.ascontiguousarray(elem, dtype=np.int32)
        self.direction = np.ascontiguousarray(direction, dtype=np.int32)
        self.i1 = np.ascontiguousarray(i1, dtype=np.int64)
        self.perp = np.ascontiguousarray(perp, dtype=np.uint8)
        nf, nl = self.i1.shape
        r = np.repeat(self.i1[:, :, None], nl, axis=2).ravel()
        c = np.repeat(self.i1[:, None, :], nl, axis=1).ravel()
        key = r * int(n_eq) + c
        order = np.argsort(key, kind="stable")                    # entries of one key stay in face order
        uniq, first = np.unique(key[order], return_index=True)
        self.rows = (uniq // int(n_eq)).astype(np.int64)
        self.cols = (uniq % int(n_eq)).astype(np.int64)
        self.grp_ptr = np.append(first, len(key)).astype(np.int64)
        self.grp_entry = order.astype(np.int64)

    def restrict_rows(self, owned_rows):
        """Keep the keys whose row is in `owned_rows` (domain-decomposed runs: ghost rows belong to another rank)."""
        keep = np.isin(self.rows, np.asarray(owned_rows))
        lens = np.diff(self.grp_ptr)[keep]
        sel = np.concatenate([np.arange(a, b) for a, b in zip(self.grp_ptr[:-1][keep], self.grp_ptr[1:][keep])]) if keep.any() \
            else np.zeros(0, dtype=np.int64)
        self.grp_entry = self.grp_entry[sel]
        self.grp_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        self.rows, self.cols = self.rows[keep], self.cols[keep]
        return self


def absorbing_plan(data):
    """-> AbsorbingPlan or None when the model has no absorbing face."""
    faces = _absorbing_faces(data)
    if not faces:
        return None
    if data.dimension == 2:
        raise SystemExit("Absorbing boundaries not implemented for 2D yet")      # system_matrix.py:324-326
    eq = data.eq_nb_dof
    el = np.array([f[0] for f in faces]); dr = np.array([f[1] for f in faces]); nodes = np.array([f[2] for f in faces])
    nf = len(faces)
    if data.lower_element_type == "quad4":
        # whole-array version of utils.py:141-175 for four-node faces: counter-clockwise by angle about the first lowest point
        keep = np.array([[1, 2], [0, 2], [0, 1]])[dr]              # in-plane coordinate columns after dropping `d`
        xy = np.take_along_axis(data.nodes[nodes, 1:], keep[:, None, :], axis=2)     # (nf, 4, 2)
        low = np.argmin(xy[:, :, 1], axis=1)
        ref = xy[np.arange(nf), low]
        ang = np.arctan2(xy[:, :, 1] - ref[:, None, 1], xy[:, :, 0] - ref[:, None, 0])
        ordered = np.take_along_axis(nodes, np.argsort(ang, axis=1, kind="stable"), axis=1)
    else:
        ordered = np.empty_like(nodes)
        for k, (e, d, nd) in enumerate(faces):
            pts = np.delete(data.nodes[nd, 1:], d, axis=1)
            srt = gmsh_face_order(pts)
            perm = [int(np.where((pts == q).all(axis=1))[0][0]) for q in srt]
            ordered[k] = nd[perm]
    i1 = np.sort(eq[nodes, dr[:, None]], axis=1).astype(np.int64)  # (nf, nl): sorted, not in face-node order (reference quirk)
    eqi = np.where(np.isnan(eq), -1, eq).astype(np.int64)
    dir_of_eq = np.zeros(data.number_eq, dtype=np.int64)
    dir_of_eq[eqi[eqi >= 0]] = np.asarray(data.BC_dir)[eqi >= 0]
    return AbsorbingPlan(data.lower_element_type, ordered, el, dr, i1, dir_of_eq[i1] == 1, data.number_eq)


def absorbing_entries(data, E, nu, rho, order: int, viscous, stiff: float):
    """Host evaluation of an `absorbing_plan` -> ({(row, col): C_abs value}, {(row, col): K_abs value / stiff}).

    CPU tests use it to pin the index plan against the oracle; the product path (`GenerateMatrix.absorbing_boundaries`)
    hands the same plan to the device (`sc_add_absorbing_faces`), which does this arithmetic in `k_abs_faces`."""
    plan = absorbing_plan(data)
    if plan is None:
        return {}, {}
    el, dr = plan.elem, plan.direction
    nf, nl = plan.i1.shape
    Ec = E[el] * (1 - nu[el]) / ((1 + nu[el]) * (1 - 2 * nu[el]))
    G = E[el] / (2 * (1 + nu[el]))
    vp, vs = np.sqrt(Ec / rho[el]), np.sqrt(G / rho[el])
    keep = np.array([[1, 2], [0, 2], [0, 1]])[dr]
    xy = np.take_along_axis(data.nodes[plan.nodes, 1:], keep[:, None, :], axis=2)   # (nf, nl, 2) in face-node order
    N, dN, w = _lib.shape_table(plan.face_type, order)
    J = np.einsum("gad,fak->fgdk", dN, xy)
    det = J[..., 0, 0] * J[..., 1, 1] - J[..., 0, 1] * J[..., 1, 0]
    S = np.einsum("ga,gb,fg->fab", N, N, det * w[None, :])
    perp = plan.perp.astype(bool)
    fct = np.where(perp, (viscous[0] * rho[el] * vp)[:, None], (viscous[1] * rho[el] * vs)[:, None])
    fct2 = np.where(perp, Ec[:, None], G[:, None])
    cv = (S * fct[:, None, :]).ravel()
    kv = (np.abs(S) * fct2[:, None, :]).ravel()
    grp = np.repeat(np.arange(len(plan.rows)), np.diff(plan.grp_ptr))
    csum = np.zeros(len(plan.rows)); ksum = np.zeros(len(plan.rows))
    np.add.at(csum, grp, cv[plan.grp_entry])                      # sequential: face order inside every key
    np.add.at(ksum, grp, kv[plan.grp_entry])
    keys = list(zip(plan.rows.tolist(), plan.cols.tolist()))
    return dict(zip(keys, csum)), dict(zip(keys, ksum / stiff))


class GenerateMatrix:
    def __init__(self, nb_equations: int, order: int, device: int = 0, ctx: "_lib.Context | None" = None) -> None:
        self.nb_equations = int(nb_equations)
        self.order = order
        self.ctx = ctx if ctx is not None else _lib.Context(device)
        self.c0 = self.c1 = 0.0
        self._pattern = None
        self._props = None
        self.assembly_seconds = None
        self.want_full_mass = True
        self.want_lumped_mass = True

    # ---- reference interface --------------------------------------------------------------------------------
    def generate_stiffness_and_mass(self, data, material: dict, elem_props=None, active=None) -> None:
        """system_matrix.py:35-121 -- element integration + assembly, on the device."""
        E, nu, rho = resolve_element_properties(data, material) if elem_props is None else elem_props
        self._props = (np.asarray(E, float), np.asarray(nu, float), np.asarray(rho, float))
        import time
        t0 = time.perf_counter()
        rows = data.node_rows()
        eq = data.equation_table_int()
        t1 = time.perf_counter()
        self.ctx.set_mesh(data.element_type, data.nodes[:, 1:], rows, eq, data.number_eq, active)
        self.ctx.set_materials(*self._props)
        t2 = time.perf_counter()
        self.ctx.build_pattern()
        t3 = time.perf_counter()
        flags = _lib.ASM_K | (_lib.ASM_M_FULL if self.want_full_mass else 0) | (_lib.ASM_M_LUMPED if self.want_lumped_mass else 0)
        self.assembly_seconds = self.ctx.assemble(self.order, flags)
        self._pattern = None
        # host wall-clock split of this call (bench.py reports it for the whole-pipeline run)
        self.timings = {"host_tables_seconds": t1 - t0, "h2d_seconds": t2 - t1, "pattern_seconds": t3 - t2,
                        "assembly_kernel_seconds": self.assembly_seconds, "assembly_call_seconds": time.perf_counter() - t3}

    def absorbing_boundaries(self, data, material: dict, parameters_viscous: list, parameters_stiff: float, owned_rows=None) -> None:
        """system_matrix.py:256-376 -- Lysmer-Kuhlemeyer dashpots into C, springs into K.  The host only plans (which faces,
        node order, equation numbers); face matrices, wave speeds and the ordered accumulation run on the device.
        `owned_rows` (domain-decomposed runs): local equations whose rows this rank assembles."""
        plan = absorbing_plan(data)
        if plan is None:
            return
        if owned_rows is not None:
            plan.restrict_rows(owned_rows)
        if len(plan.rows):
            self.ctx.add_absorbing_faces(plan, self.order, parameters_viscous[0], parameters_viscous[1], parameters_stiff)

    def damping_Rayleigh(self, damp) -> None:
        """system_matrix.py:166-198 -- C = C + c0 M + c1 K (applied on the fly on the device)."""
        self.c0, self.c1 = rayleigh_coefficients(damp)
        self.ctx.set_rayleigh(self.c0, self.c1)

    # ---- matrices back on the host --------------------------------------------------------------------------
    def pattern(self):
        if self._pattern is None:
            self._pattern = self.ctx.get_pattern()
        return self._pattern

    def _csr(self, which):
        rowptr, col = self.pattern()
        n = self.nb_equations
        return sp.csr_matrix((self.ctx.get_values(which), col, rowptr), shape=(n, n))

    @property
    def K(self):
        return self._csr(_lib.MAT_K)

    @property
    def M(self):
        return self._csr(_lib.MAT_M)

    @property
    def C(self):
        return self._csr(_lib.MAT_C)

    def pruned(self, which: str):
        """The matrix as the reference's scipy binops leave it: exact zeros dropped."""
        m = getattr(self, which).copy()
        m.eliminate_zeros()
        return m

    def lumped_mass(self):
        return self.ctx.get_lumped_mass()
