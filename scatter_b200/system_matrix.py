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


def face_mass(face_type: str, order: int, xy: np.ndarray) -> np.ndarray:
    """S[a,b] = sum_g N_a N_b detJ w over a 2-D face element (discretisation.py:419-433 with unit density)."""
    N, dN, w = _lib.shape_table(face_type, order)
    J = np.einsum("gad,ak->gdk", dN, xy[:, :2])
    det = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
    return np.einsum("ga,gb,g->ab", N, N, det * w)


def _absorbing_faces(data):
    """(element, direction, node rows[nl]) of every absorbing face, in the reference's element / direction order
    (system_matrix.py:274-316): a face exists where exactly `nb_nodes_lower_elem` nodes of an element absorb in a direction."""
    type_bc = np.asarray(data.type_BC)
    absorb = type_bc == "Absorb"                                   # (Nn, dim)
    rows = data.node_rows()
    nl, dim = data.nb_nodes_lower_elem, data.dimension
    touched = np.where(absorb.any(axis=1)[rows].any(axis=1))[0]
    out = []
    if len(touched) == 0:
        return out
    fl = absorb[rows[touched]]                                     # (nt, nne, dim)
    cnt = fl.sum(axis=1)                                           # (nt, dim)
    for t, d in zip(*np.where(cnt == nl)):                         # row-major: element order, then direction
        e = touched[t]
        out.append((e, d, rows[e][fl[t, :, d]]))
    return out


def absorbing_entries(data, E, nu, rho, order: int, viscous, stiff: float):
    """COO entries of C_abs and K_abs/stiff, duplicates summed in face order (system_matrix.py:256-376).

    quad4 faces (hexa8 meshes) are processed with whole-array numpy so that large boxes stay cheap; other face types use
    the per-face path."""
    faces = _absorbing_faces(data)
    cdict, kdict = {}, {}
    if not faces:
        return cdict, kdict
    dim, nl = data.dimension, data.nb_nodes_lower_elem
    if dim == 2:
        raise SystemExit("Absorbing boundaries not implemented for 2D yet")
    eq = data.eq_nb_dof
    el = np.array([f[0] for f in faces]); dr = np.array([f[1] for f in faces]); nodes = np.array([f[2] for f in faces])
    nf = len(faces)
    Ec = E[el] * (1 - nu[el]) / ((1 + nu[el]) * (1 - 2 * nu[el]))
    G = E[el] / (2 * (1 + nu[el]))
    vp, vs = np.sqrt(Ec / rho[el]), np.sqrt(G / rho[el])
    # unit face matrices in the reference's face-node order
    S = np.empty((nf, nl, nl))
    if data.lower_element_type == "quad4":
        keep = np.array([[1, 2], [0, 2], [0, 1]])[dr]              # in-plane coordinate columns after dropping `d`
        xy = np.take_along_axis(data.nodes[nodes, 1:], keep[:, None, :], axis=2)     # (nf, 4, 2)
        low = np.argmin(xy[:, :, 1], axis=1)                        # first lowest point
        ref = xy[np.arange(nf), low]
        ang = np.arctan2(xy[:, :, 1] - ref[:, None, 1], xy[:, :, 0] - ref[:, None, 0])
        srt = np.argsort(ang, axis=1, kind="stable")
        xy = np.take_along_axis(xy, srt[:, :, None], axis=1)        # counter-clockwise from the lowest point
        N, dN, w = _lib.shape_table("quad4", order)
        J = np.einsum("gad,fak->fgdk", dN, xy)
        det = J[..., 0, 0] * J[..., 1, 1] - J[..., 0, 1] * J[..., 1, 0]
        S = np.einsum("ga,gb,fg->fab", N, N, det * w[None, :])
    else:
        for k, (e, d, nd) in enumerate(faces):
            S[k] = face_mass(data.lower_element_type, order, gmsh_face_order(np.delete(data.nodes[nd, 1:], d, axis=1)))
    # equation numbers of the face dofs, sorted (the reference pairs them with the face-node order as is)
    i1 = np.sort(eq[nodes, dr[:, None]], axis=1).astype(np.int64)  # (nf, nl)
    eqi = np.where(np.isnan(eq), -1, eq).astype(np.int64)
    dir_of_eq = np.zeros(data.number_eq, dtype=np.int64)
    dir_of_eq[eqi[eqi >= 0]] = np.asarray(data.BC_dir)[eqi >= 0]
    perp = dir_of_eq[i1] == 1
    fct = np.where(perp, (viscous[0] * rho[el] * vp)[:, None], (viscous[1] * rho[el] * vs)[:, None])
    fct2 = np.where(perp, Ec[:, None], G[:, None])
    r = np.repeat(i1[:, :, None], nl, axis=2).ravel()
    c = np.repeat(i1[:, None, :], nl, axis=1).ravel()
    cv = (S * fct[:, None, :]).ravel()
    kv = (np.abs(S) * fct2[:, None, :]).ravel()
    # sum duplicates in face order (np.add.at is sequential) on the unique key set
    key = r * data.number_eq + c
    uniq, inv = np.unique(key, return_inverse=True)
    csum = np.zeros(len(uniq)); ksum = np.zeros(len(uniq))
    np.add.at(csum, inv, cv)
    np.add.at(ksum, inv, kv)
    keys = [(int(u // data.number_eq), int(u % data.number_eq)) for u in uniq]
    cdict = dict(zip(keys, csum))
    kdict = dict(zip(keys, ksum / stiff))
    return cdict, kdict


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
        rows = data.node_rows()
        eq = data.equation_table_int()
        self.ctx.set_mesh(data.element_type, data.nodes[:, 1:], rows, eq, data.number_eq, active)
        self.ctx.set_materials(*self._props)
        self.ctx.build_pattern()
        flags = _lib.ASM_K | (_lib.ASM_M_FULL if self.want_full_mass else 0) | (_lib.ASM_M_LUMPED if self.want_lumped_mass else 0)
        self.assembly_seconds = self.ctx.assemble(self.order, flags)
        self._pattern = None

    def absorbing_boundaries(self, data, material: dict, parameters_viscous: list, parameters_stiff: float, owned_rows=None) -> None:
        """system_matrix.py:256-376 -- Lysmer-Kuhlemeyer dashpots into C, springs into K.  `owned_rows` (domain-decomposed
        runs): local equation numbers whose rows this rank assembles; entries of ghost rows are left to their owner."""
        E, nu, rho = self._props if self._props is not None else resolve_element_properties(data, material)
        cdict, kdict = absorbing_entries(data, E, nu, rho, self.order, parameters_viscous, parameters_stiff)
        if cdict:
            keys = np.array(list(cdict.keys()), dtype=np.int64)
            cv, kv = np.array(list(cdict.values())), np.array([kdict[tuple(k)] for k in keys])
            if owned_rows is not None:
                keep = np.isin(keys[:, 0], np.asarray(owned_rows))
                keys, cv, kv = keys[keep], cv[keep], kv[keep]
            if len(keys):
                self.ctx.add_entries(_lib.MAT_C, keys[:, 0], keys[:, 1], cv)
                self.ctx.add_entries(_lib.MAT_K, keys[:, 0], keys[:, 1], kv)

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
