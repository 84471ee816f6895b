"""TEST INFRASTRUCTURE -- CPU oracle (numpy/scipy restatement) of SCATTER's hot path.

This file is the parity yard-stick for the CUDA path.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product (`scatter_b200/`) never does.

What it restates (file:line relative to the reference tree, PlatypusBytes/scatter):

* shape functions N, dN/dxi for tri3/tri6/quad4/quad8/tetra4/tetra10/hexa8/hexa20 -- `scatter/element_types.py:52-104`
  (hexa8), `:155-254` (hexa20), `:303-333` (quad4), `:382-424` (quad8, incl. its non-serendipity corner functions),
  `:475-501` (tri3), `:552-587` (tri6), `:646-683` (tetra4), `:742-802` (tetra10)
* Gauss tables -- `scatter/discretisation.py:436-497`; point loop order u outer, w inner -- `:32-38`, `:303-306`
* Jacobian, signed detJ, global derivatives -- `scatter/discretisation.py:107-129` (3D), `:337-351` (2D)
* B / N matrices, Ke = sum B^T D B detJ w, Me = rho sum N^T N detJ w -- `scatter/discretisation.py:139-222`, `:353-417`
* isotropic elastic D (3D / plane strain) -- `scatter/material_models.py:5-43`
* global assembly on the structural pattern (all free (i,k) pairs of every element, explicit zeros kept) --
  `scatter/system_matrix.py:35-121`
* Lysmer-Kuhlemeyer absorbing boundary -- `scatter/system_matrix.py:256-376`, `scatter/utils.py:141-196`
* Rayleigh damping coefficients -- `scatter/system_matrix.py:166-198`
* mesh model: BC planes / equation numbering -- `scatter/mesher.py:230-326`, `scatter/utils.py:5-60`
* loads (pulse / heaviside / moving) -- `scatter/force_external.py:76-149, 250-345`
* time integration: the reference calls the un-vendored package PuggleSolvers==1.0.1 (`scatter/scatter.py:120-159`);
  its default solver is restated as the classical incremental Newmark (beta=1/4, gamma=1/2), which reproduces every
  golden history the reference ships (see `oracle/VALIDATION.md`).  The central-difference scheme is this
  repository's own documented scheme (no reference fixture exists): PARITY UNPINNED for that integrator.

Parity status: PINNED for assembly (against the imported reference, `oracle/validate_against_reference.py`) and
for Newmark (against the reference's golden result files, copied in compact form to `tests/golden/`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# --------------------------------------------------------------------------------------------------------------
# element catalogue
# --------------------------------------------------------------------------------------------------------------
ELEMENT_INFO = {
    # name: (nne, dim, integration family, gmsh code)
    "tri3": (3, 2, "tri", 2),
    "tri6": (6, 2, "tri", 9),
    "quad4": (4, 2, "quad", 3),
    "quad8": (8, 2, "quad", 16),
    "tetra4": (4, 3, "tetra", 4),
    "tetra10": (10, 3, "tetra", 11),
    "hexa8": (8, 3, "quad", 5),
    "hexa20": (20, 3, "quad", 17),
}

# natural coordinates of the hexa8 corners (gmsh order)
_HEX_CORNER = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                        [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=float)
# hexa20 mid-edge nodes (gmsh order 8..19): the two corners each one sits between
_HEX20_EDGES = [(0, 1), (0, 3), (0, 4), (1, 2), (1, 5), (2, 3), (2, 6), (3, 7), (4, 5), (4, 7), (5, 6), (6, 7)]
_QUAD_CORNER = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], dtype=float)


def shape_functions(elem_type: str, xi) -> tuple[np.ndarray, np.ndarray]:
    """N (nne,) and dN/dxi (nne, dim) at natural coordinate `xi`."""
    xi = np.asarray(xi, dtype=float)
    if elem_type == "hexa8":
        c = _HEX_CORNER
        f = 1.0 + c * xi  # (8,3)
        N = f[:, 0] * f[:, 1] * f[:, 2] / 8.0
        dN = np.stack([c[:, 0] * f[:, 1] * f[:, 2], f[:, 0] * c[:, 1] * f[:, 2], f[:, 0] * f[:, 1] * c[:, 2]], axis=1) / 8.0
        return N, dN
    if elem_type == "hexa20":
        N = np.zeros(20)
        dN = np.zeros((20, 3))
        c = _HEX_CORNER
        f = 1.0 + c * xi
        s = (c * xi).sum(axis=1) - 2.0
        N[:8] = f[:, 0] * f[:, 1] * f[:, 2] * s / 8.0
        for d in range(3):
            o1, o2 = (d + 1) % 3, (d + 2) % 3
            dN[:8, d] = c[:, d] * f[:, o1] * f[:, o2] * (s + f[:, d]) / 8.0
        for k, (a, b) in enumerate(_HEX20_EDGES):
            mid = 0.5 * (c[a] + c[b])          # one component is 0: the edge direction
            e = int(np.where(mid == 0.0)[0][0])
            o1, o2 = (e + 1) % 3, (e + 2) % 3
            g1, g2 = 1.0 + mid[o1] * xi[o1], 1.0 + mid[o2] * xi[o2]
            q = 1.0 - xi[e] ** 2
            N[8 + k] = q * g1 * g2 / 4.0
            dN[8 + k, e] = -2.0 * xi[e] * g1 * g2 / 4.0
            dN[8 + k, o1] = q * mid[o1] * g2 / 4.0
            dN[8 + k, o2] = q * g1 * mid[o2] / 4.0
        return N, dN
    if elem_type == "quad4":
        c = _QUAD_CORNER
        f = 1.0 + c * xi
        N = f[:, 0] * f[:, 1] / 4.0
        dN = np.stack([c[:, 0] * f[:, 1], f[:, 0] * c[:, 1]], axis=1) / 4.0
        return N, dN
    if elem_type == "quad8":
        # NOTE: the reference keeps the plain bilinear corner functions (no serendipity correction) and uses a factor
        # 1/2 on the mid-side functions (element_types.py:396-423); replicated verbatim for parity.
        u, v = xi
        c = _QUAD_CORNER
        f = 1.0 + c * xi
        N = np.zeros(8)
        dN = np.zeros((8, 2))
        N[:4] = f[:, 0] * f[:, 1] / 4.0
        dN[:4] = np.stack([c[:, 0] * f[:, 1], f[:, 0] * c[:, 1]], axis=1) / 4.0
        N[4] = 0.5 * (1 - u * u) * (1 - v)
        N[5] = 0.5 * (1 + u) * (1 - v * v)
        N[6] = 0.5 * (1 - u * u) * (1 + v)
        N[7] = 0.5 * (1 - u) * (1 - v * v)
        dN[4] = [-u * (1 - v), -0.5 * (1 - u * u)]
        dN[5] = [0.5 * (1 - v * v), -v * (1 + u)]
        dN[6] = [-u * (1 + v), 0.5 * (1 - u * u)]
        dN[7] = [-0.5 * (1 - v * v), -v * (1 - u)]
        return N, dN
    if elem_type == "tri3":
        u, v = xi
        return np.array([1 - u - v, u, v]), np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])
    if elem_type == "tri6":
        u, v = xi
        L = 1 - u - v
        N = np.array([(2 * L - 1) * L, (2 * u - 1) * u, (2 * v - 1) * v, 4 * L * u, 4 * u * v, 4 * L * v])
        dN = np.array([[1 - 4 * L, 1 - 4 * L], [4 * u - 1, 0.0], [0.0, 4 * v - 1],
                       [4 * L - 4 * u, -4 * u], [4 * v, 4 * u], [-4 * v, 4 * L - 4 * v]])
        return N, dN
    if elem_type == "tetra4":
        u, v, w = xi
        return (np.array([1 - u - v - w, u, v, w]),
                np.array([[-1.0, -1.0, -1.0], [1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]]))
    if elem_type == "tetra10":
        u, v, w = xi
        x = 1 - u - v - w
        N = np.array([(2 * x - 1) * x, (2 * u - 1) * u, (2 * v - 1) * v, (2 * w - 1) * w,
                      4 * u * x, 4 * u * v, 4 * v * x, 4 * w * x, 4 * v * w, 4 * u * w])
        d0 = 1 - 4 * x
        dN = np.array([[d0, d0, d0], [4 * u - 1, 0, 0], [0, 4 * v - 1, 0], [0, 0, 4 * w - 1],
                       [4 * x - 4 * u, -4 * u, -4 * u], [4 * v, 4 * u, 0], [-4 * v, 4 * x - 4 * v, -4 * v],
                       [-4 * w, -4 * w, 4 * x - 4 * w], [0, 4 * w, 4 * v], [4 * w, 0, 4 * u]], dtype=float)
        return N, dN
    raise ValueError(f"element type {elem_type} not supported")


def gauss_rule(family: str, n: int) -> tuple[np.ndarray, np.ndarray]:
    """1-D rule for 'quad', full rule (dim, npts) for 'tri' / 'tetra' -- discretisation.py:436-497."""
    if family == "quad":
        if n == 1:
            return np.array([0.0]), np.array([2.0])
        if n == 2:
            a = math.sqrt(1.0 / 3.0)
            return np.array([-a, a]), np.array([1.0, 1.0])
        if n == 3:
            a = math.sqrt(3.0 / 5.0)
            return np.array([-a, 0.0, a]), np.array([5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0])
    elif family == "tri":
        if n == 1:
            return np.array([[1 / 3], [1 / 3]]), np.array([1 / 2])
        if n == 2:
            return np.array([[1 / 6, 2 / 3, 1 / 6], [1 / 6, 1 / 6, 2 / 3]]), np.array([1 / 6, 1 / 6, 1 / 6])
        if n == 3:
            return (np.array([[1 / 3, 1 / 5, 3 / 5, 1 / 5], [1 / 3, 1 / 5, 1 / 5, 3 / 5]]),
                    np.array([-27 / 96, 25 / 96, 25 / 96, 25 / 96]))
    elif family == "tetra":
        if n == 1:
            return np.array([[1 / 4], [1 / 4], [1 / 4]]), np.array([1 / 6])
        if n == 2:
            a = 1 / 4 - 1 / 20 * math.sqrt(5)
            b = 1 / 4 + 3 / 20 * math.sqrt(5)
            return np.array([[a, a, a, b], [a, a, b, a], [a, b, a, a]]), np.full(4, 1 / 24)
    raise SystemExit(f"ERROR: integration order not supported for type {family}")


def integration_points(elem_type: str, order: int) -> tuple[np.ndarray, np.ndarray]:
    """All Gauss points (ngp, dim) and weights (ngp,) in the reference's loop order."""
    nne, dim, family, _ = ELEMENT_INFO[elem_type]
    if family == "quad":
        x, w = gauss_rule("quad", order)
        pts, wts = [], []
        if dim == 3:
            for i in range(order):
                for j in range(order):
                    for k in range(order):
                        pts.append([x[i], x[j], x[k]])
                        wts.append(np.prod([w[i], w[j], w[k]]))
        else:
            for i in range(order):
                for j in range(order):
                    pts.append([x[i], x[j]])
                    wts.append(np.prod([w[i], w[j]]))
        return np.array(pts), np.array(wts)
    x, w = gauss_rule(family, order)
    return np.ascontiguousarray(x.T), np.asarray(w, dtype=float)


def reference_tables(elem_type: str, order: int):
    """(N[ngp,nne], dN[ngp,nne,dim], w[ngp]) -- element independent."""
    pts, w = integration_points(elem_type, order)
    Ns, dNs = zip(*(shape_functions(elem_type, p) for p in pts))
    return np.array(Ns), np.array(dNs), w


def elasticity_matrix(E: float, nu: float, dim: int) -> np.ndarray:
    """material_models.py:5-43"""
    if dim == 3:
        D = np.zeros((6, 6))
        D[:3, :3] = nu
        D[np.arange(3), np.arange(3)] = 1.0 - nu
        D[np.arange(3, 6), np.arange(3, 6)] = (1.0 - 2.0 * nu) / 2
        return D * (E / ((1.0 + nu) * (1.0 - 2.0 * nu)))
    if dim == 2:
        D = np.zeros((3, 3))
        D[:2, :2] = nu
        D[0, 0] = D[1, 1] = 1.0 - nu
        D[2, 2] = (1.0 - 2.0 * nu) / 2
        return D * E / ((1.0 + nu) * (1.0 - 2.0 * nu))
    raise SystemExit(f"ERROR dimension: {dim} is not supported")


def element_matrices(elem_type: str, order: int, xyz: np.ndarray, E, nu, rho, chunk: int = 4096):
    """Ke, Me for a batch of elements.

    xyz: (Ne, nne, 3) nodal coordinates (2-D elements use the first two columns, discretisation.py:329).
    E, nu, rho: scalars or (Ne,) arrays.  Returns Ke, Me of shape (Ne, nne*dim, nne*dim).
    """
    nne, dim, _, _ = ELEMENT_INFO[elem_type]
    xyz = np.asarray(xyz, dtype=float)
    ne = xyz.shape[0]
    E = np.broadcast_to(np.asarray(E, dtype=float), (ne,))
    nu = np.broadcast_to(np.asarray(nu, dtype=float), (ne,))
    rho = np.broadcast_to(np.asarray(rho, dtype=float), (ne,))
    Nref, dNref, w = reference_tables(elem_type, order)
    ngp = len(w)
    nd = nne * dim
    nstrain = 6 if dim == 3 else 3
    Ke = np.empty((ne, nd, nd))
    Me = np.empty((ne, nd, nd))
    # N matrix per gauss point (dim, nd)
    Nmat = np.zeros((ngp, dim, nd))
    for d in range(dim):
        Nmat[:, d, d::dim] = Nref
    NtN = np.einsum("gkj,gkl->gjl", Nmat, Nmat)
    for s in range(0, ne, chunk):
        x = xyz[s:s + chunk, :, :dim]
        m = x.shape[0]
        J = np.einsum("gad,eak->egdk", dNref, x)            # J[d,k] = sum_a dN[a,d] x[a,k]  == dN^T . xyz
        detJ = np.linalg.det(J)                               # signed, no abs (discretisation.py:126)
        invJT = np.linalg.inv(np.swapaxes(J, -1, -2))          # inv(J^T)
        dNg = np.einsum("gad,egdk->egak", dNref, invJT)       # dN . inv(J^T)
        B = np.zeros((m, ngp, nstrain, nd))
        if dim == 3:
            B[:, :, 0, 0::3] = dNg[..., 0]
            B[:, :, 1, 1::3] = dNg[..., 1]
            B[:, :, 2, 2::3] = dNg[..., 2]
            B[:, :, 3, 0::3] = dNg[..., 1]
            B[:, :, 3, 1::3] = dNg[..., 0]
            B[:, :, 4, 1::3] = dNg[..., 2]
            B[:, :, 4, 2::3] = dNg[..., 1]
            B[:, :, 5, 0::3] = dNg[..., 2]
            B[:, :, 5, 2::3] = dNg[..., 0]
        else:
            B[:, :, 0, 0::2] = dNg[..., 0]
            B[:, :, 1, 1::2] = dNg[..., 1]
            B[:, :, 2, 0::2] = dNg[..., 1]
            B[:, :, 2, 1::2] = dNg[..., 0]
        D = np.stack([elasticity_matrix(E[s + i], nu[s + i], dim) for i in range(m)])
        wj = detJ * w[None, :]
        BtD = np.einsum("egkj,ekl->egjl", B, D)
        Ke[s:s + m] = np.einsum("egjk,egkl,eg->ejl", BtD, B, wj)
        Me[s:s + m] = np.einsum("gjl,eg->ejl", NtN, wj) * rho[s:s + m, None, None]
    return Ke, Me


# --------------------------------------------------------------------------------------------------------------
# mesh model (ReadMesh-shaped)
# --------------------------------------------------------------------------------------------------------------
@dataclass
class Model:
    """The attributes of `mesher.ReadMesh` the hot path consumes (mesher.py:31-52)."""
    nodes: np.ndarray                 # (Nn, 4) [id, x, y, z]
    elem: np.ndarray                  # (Ne, nne) 1-based node ids
    materials_index: np.ndarray       # (Ne,) physical tag
    materials: list                   # [[dim, tag, name], ...]
    element_type: str
    dimension: int
    BC: np.ndarray = None
    BC_dir: np.ndarray = None
    eq_nb_dof: np.ndarray = None      # (Nn, dim) float with NaN for fixed dofs
    type_BC: np.ndarray = None
    number_eq: int = 0
    eq_nb_elem: np.ndarray = None     # (Ne, nne*dim) float with NaN
    lower_element_type: str = ""
    nb_nodes_lower_elem: int = 0
    nb_nodes_elem: int = 0
    extra: dict = field(default_factory=dict)


_LOWER = {"hexa8": ("quad4", 4), "hexa20": ("quad8", 8), "tetra4": ("tri3", 3), "tetra10": ("tri6", 6)}
_GMSH_TO_TYPE = {2: "tri3", 9: "tri6", 3: "quad4", 5: "hexa8", 17: "hexa20", 4: "tetra4", 11: "tetra10"}


def parse_gmsh(path: str) -> Model:
    """gmsh 2.2 ASCII reader following mesher.py:137-228 / utils.py:62-90 (plain loops)."""
    with open(path) as f:
        lines = f.read().splitlines()

    def section(a, b):
        i0 = next(i for i, l in enumerate(lines) if l.startswith(a))
        i1 = next(i for i, l in enumerate(lines) if l.startswith(b))
        return [l.split() for l in lines[i0 + 2:i1]]

    nodes = np.array([[float(t) for t in row] for row in section("$Nodes", "$EndNodes")])
    names = [[float(r[0]), int(float(r[1])), r[2].replace('"', "")] for r in section("$PhysicalNames", "$EndPhysicalNames")]
    elems = [[int(float(t)) for t in row] for row in section("$Elements", "$EndElements")]
    rose_tag = next((n[1] for n in names if n[2] == "rose"), None)
    geo = [e for e in elems if e[3] != rose_tag]
    codes = {e[1] for e in geo}
    if len(codes) != 1 or next(iter(codes)) not in _GMSH_TO_TYPE:
        raise SystemExit("ERROR: Element type not supported")
    etype = _GMSH_TO_TYPE[next(iter(codes))]
    nne, dim, _, _ = ELEMENT_INFO[etype]
    geo = np.array(geo)
    low = _LOWER.get(etype, ([], []))
    return Model(nodes=nodes, elem=geo[:, 5:], materials_index=geo[:, 3], materials=names, element_type=etype,
                 dimension=dim, lower_element_type=low[0], nb_nodes_lower_elem=low[1], nb_nodes_elem=nne)


def apply_boundary_conditions(model: Model, bc: dict) -> None:
    """mesher.py:230-274 -- plane (3D) / segment (2D) membership with atol 1e-5; max code wins."""
    nn, dim = len(model.nodes), model.dimension
    BC = np.zeros((nn, dim), dtype=int)
    BCd = np.zeros((nn, dim), dtype=int)
    xyz = model.nodes[:, 1:]
    for name in bc:
        typ, pts = bc[name][0], bc[name][1]
        if dim == 3:
            p1, p2, p3 = (np.array(p, dtype=float) for p in pts[:3])
            cp = np.cross(p3 - p1, p2 - p1)
            direction = np.abs(cp / np.linalg.norm(cp))
            resid = xyz[:, 0] * cp[0] + xyz[:, 1] * cp[1] + xyz[:, 2] * cp[2] - np.dot(cp, p3)
        else:
            p1, p2 = np.array(pts[0], dtype=float), np.array(pts[1], dtype=float)
            vec = p2 - p1
            direction = np.array([-vec[1], vec[0]])
            resid = (np.linalg.norm(p1[None, :] - xyz, axis=1) + np.linalg.norm(p2[None, :] - xyz, axis=1)
                     - np.linalg.norm(p1 - p2))
        for idx in np.where(np.isclose(resid, 0.0, atol=1e-5))[0]:
            for j, val in enumerate(typ):
                BC[idx, j] = max(BC[idx, j], int(val))
                BCd[idx, j] = max(BCd[idx, j], abs(int(direction[j])))
    model.BC, model.BC_dir = BC, BCd


def number_equations(model: Model) -> None:
    """mesher.py:276-326 -- node-file order, dof x,y,(z), NaN for fixed dofs."""
    nn, dim = len(model.nodes), model.dimension
    eq = np.zeros((nn, dim))
    typ = np.full((nn, dim), "Normal")
    k = 0
    for i in range(nn):
        for j in range(dim):
            c = model.BC[i, j]
            if c == 0:
                eq[i, j] = k; k += 1
            elif c == 1:
                eq[i, j] = np.nan; typ[i, j] = "Fixed"
            elif c == 2:
                eq[i, j] = k; typ[i, j] = "Absorb"; k += 1
            else:
                raise SystemExit(f"Error in the boundary condition definition. \n{c} is not a valid boundary condition.")
    model.eq_nb_dof, model.type_BC, model.number_eq = eq, typ, k
    row_of_id = {int(n): i for i, n in enumerate(model.nodes[:, 0])}
    rows = np.array([[row_of_id[int(n)] for n in el] for el in model.elem])
    model.extra["node_rows"] = rows
    model.eq_nb_elem = eq[rows].reshape(len(model.elem), -1)


def build_model(path: str, bc: dict) -> Model:
    m = parse_gmsh(path)
    apply_boundary_conditions(m, bc)
    number_equations(m)
    return m


def element_properties(model: Model, materials: dict):
    """Per-element (E, nu, rho) via physical tag -> name -> dict (system_matrix.py:52,64-71)."""
    tag_to_name = {int(m[1]): m[2] for m in model.materials}
    E = np.array([materials[tag_to_name[int(t)]]["Young"] for t in model.materials_index], dtype=float)
    nu = np.array([materials[tag_to_name[int(t)]]["poisson"] for t in model.materials_index], dtype=float)
    rho = np.array([materials[tag_to_name[int(t)]]["density"] for t in model.materials_index], dtype=float)
    return E, nu, rho


# --------------------------------------------------------------------------------------------------------------
# global assembly
# --------------------------------------------------------------------------------------------------------------
def structural_pattern(eq_elem: np.ndarray, n_eq: int) -> sp.csr_matrix:
    """CSR (int8 ones) of all free (i,k) pairs of every element; sorted columns -- system_matrix.py:90-121."""
    ne, nd = eq_elem.shape
    free = ~np.isnan(eq_elem)
    eqi = np.where(free, eq_elem, -1).astype(np.int64)
    r = np.repeat(eqi[:, :, None], nd, axis=2).ravel()
    c = np.repeat(eqi[:, None, :], nd, axis=1).ravel()
    ok = (r >= 0) & (c >= 0)
    key = np.unique(r[ok] * n_eq + c[ok])
    rows, cols = key // n_eq, key % n_eq
    indptr = np.zeros(n_eq + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr)
    return sp.csr_matrix((np.ones(len(cols), dtype=np.int8), cols.astype(np.int32), indptr), shape=(n_eq, n_eq))


def assemble_global(model: Model, E, nu, rho, order: int):
    """K, M as CSR on the structural pattern (explicit zeros kept), element contributions summed in element order."""
    nne, dim, _, _ = ELEMENT_INFO[model.element_type]
    rows = model.extra["node_rows"]
    xyz = model.nodes[:, 1:][rows]
    Ke, Me = element_matrices(model.element_type, order, xyz, E, nu, rho)
    n_eq = model.number_eq
    pat = structural_pattern(model.eq_nb_elem, n_eq)
    indptr, indices = pat.indptr.astype(np.int64), pat.indices.astype(np.int64)
    eq = model.eq_nb_elem
    free = ~np.isnan(eq)
    eqi = np.where(free, eq, 0).astype(np.int64)
    nd = nne * dim
    kv = np.zeros(len(indices))
    mv = np.zeros(len(indices))
    # slot lookup through a sorted global key array; np.add.at accumulates in array (= element) order
    keys = np.repeat(np.arange(n_eq, dtype=np.int64), np.diff(indptr)) * n_eq + indices
    for s in range(0, len(eq), 2048):
        e = eqi[s:s + 2048]
        f = free[s:s + 2048]
        r = np.repeat(e[:, :, None], nd, axis=2)
        c = np.repeat(e[:, None, :], nd, axis=1)
        ok = (f[:, :, None] & f[:, None, :]).ravel()
        slot = np.searchsorted(keys, (r * n_eq + c).ravel()[ok])
        np.add.at(kv, slot, Ke[s:s + 2048].ravel()[ok])
        np.add.at(mv, slot, Me[s:s + 2048].ravel()[ok])
    K = sp.csr_matrix((kv, pat.indices, pat.indptr), shape=(n_eq, n_eq))
    M = sp.csr_matrix((mv, pat.indices, pat.indptr), shape=(n_eq, n_eq))
    return K, M


def rayleigh_coefficients(damp) -> tuple[float, float]:
    """system_matrix.py:183-196"""
    f1, d1, f2, d2 = damp
    if f1 == f2:
        raise SystemExit("Frequencies for the Rayleigh damping are the same.")
    A = 0.5 * np.array([[1 / (2 * np.pi * f1), 2 * np.pi * f1], [1 / (2 * np.pi * f2), 2 * np.pi * f2]])
    c = np.linalg.solve(A, np.array([d1, d2], dtype=float))
    return float(c[0]), float(c[1])


# ---- absorbing boundaries (plain loops; small cases only) -----------------------------------------------------
def _collinear(pts) -> bool:
    def slope(p, q):
        return (q[1] - p[1]) / (q[0] - p[0]) if q[0] != p[0] else float("inf")
    s0 = slope(pts[0], pts[1])
    return all(slope(pts[i - 1], pts[i]) == s0 for i in range(2, len(pts)))


def clockwise_sort(points: np.ndarray) -> np.ndarray:
    """utils.py:141-175 -- angle sort about the lowest point, corners first then mid-side nodes."""
    ref = min(points, key=lambda p: p[1])
    srt = sorted(points, key=lambda p: np.arctan2(p[1] - ref[1], p[0] - ref[0]))
    corner = [srt[0]]
    base = 0
    for i in range(len(srt) - 1):
        if not _collinear(srt[base:i + 2]):
            corner.append(srt[i])
            base = i
    s1 = set(map(tuple, srt))
    s2 = set(map(tuple, corner))
    corner.extend(list(map(list, s1.symmetric_difference(s2))))
    return np.array(corner)


def face_unit_matrix(face_type: str, order: int, xy: np.ndarray) -> np.ndarray:
    """sum N^T N detJ w of a 2-D face element with 2 dof/node (discretisation.py:419-433)."""
    nne = ELEMENT_INFO[face_type][0]
    Nref, dNref, w = reference_tables(face_type, order)
    out = np.zeros((2 * nne, 2 * nne))
    for g in range(len(w)):
        J = dNref[g].T @ xy[:, :2]
        Nm = np.zeros((2, 2 * nne))
        Nm[0, 0::2] = Nref[g]
        Nm[1, 1::2] = Nref[g]
        out = out + Nm.T @ Nm * np.linalg.det(J) * w[g]
    return out


def absorbing_matrices(model: Model, E, nu, rho, order: int, viscous, stiff: float):
    """C_abs and K_abs (already divided by `stiff`) as CSR -- system_matrix.py:256-376, loop for loop."""
    n_eq, dim, nl = model.number_eq, model.dimension, model.nb_nodes_lower_elem
    Cabs = sp.lil_matrix((n_eq, n_eq))
    Kabs = sp.lil_matrix((n_eq, n_eq))
    rows = model.extra["node_rows"]
    for e in range(len(model.elem)):
        Ec = E[e] * (1 - nu[e]) / ((1 + nu[e]) * (1 - 2 * nu[e]))
        G = E[e] / (2 * (1 + nu[e]))
        vp, vs = np.sqrt(Ec / rho[e]), np.sqrt(G / rho[e])
        bc_type, xyz_abs, eq_nb = [], [], []
        for r in rows[e]:
            if "Absorb" in model.type_BC[r]:
                bc_type.append(model.type_BC[r])
                xyz_abs.append(model.nodes[r, 1:])
                eq_nb.append(model.eq_nb_dof[r])
        if not bc_type:
            continue
        bc_type = np.array(bc_type)
        for d in range(dim):
            sel = np.where(bc_type[:, d] == "Absorb")[0]
            if len(sel) != nl:
                continue
            if dim == 2:
                raise SystemExit("Absorbing boundaries not implemented for 2D yet")
            xy = clockwise_sort(np.delete(np.array(xyz_abs)[sel, :], d, axis=1))
            unit = face_unit_matrix(model.lower_element_type, order, xy)
            ext = np.copy(unit)
            for i in range(nl):
                ext = np.insert(ext, 2 + i * dim, np.zeros(ext.shape[0]), axis=1)
            for i in range(nl):
                new_row = np.copy(ext[2 + i * dim - 1, :])
                new_row = np.concatenate(([new_row[-1]], new_row[:-1]))
                ext = np.insert(ext, 2 + i * dim, new_row, axis=0)
            i1 = np.sort(np.array(eq_nb)[sel, d]).astype(int)
            i2 = np.linspace(d, d + (nl - 1) * dim, nl, dtype=int)
            fct = np.ones(len(i1)) * viscous[1] * rho[e] * vs
            fct2 = np.ones(len(i1)) * G
            for i, val in enumerate(i1):
                j = np.where(model.eq_nb_dof == val)
                if model.BC_dir[j[0], j[1]] == 1:
                    fct[i] = viscous[0] * rho[e] * vp
                    fct2[i] = Ec
            blk = ext[i2.reshape(-1, 1), i2]
            Cabs[i1.reshape(-1, 1), i1] = Cabs[i1.reshape(-1, 1), i1] + blk * fct
            Kabs[i1.reshape(-1, 1), i1] += np.abs(blk) * fct2
    return Cabs.tocsr(), (Kabs / stiff).tocsr()


def system_matrices(model: Model, materials: dict, settings: dict, elem_props=None):
    """K, M, C exactly as `scatter.scatter` builds them (scatter.py:103-114).  Returns CSR K, M, C plus (c0, c1)."""
    E, nu, rho = element_properties(model, materials) if elem_props is None else elem_props
    order = settings["int_order"]
    K, M = assemble_global(model, E, nu, rho, order)
    has_abs = model.type_BC is not None and (model.type_BC == "Absorb").any()
    n = model.number_eq
    if has_abs:
        Cabs, Kabs = absorbing_matrices(model, E, nu, rho, order, settings["absorbing_BC"], settings["absorbing_BC_stiff"])
    else:
        Cabs, Kabs = sp.csr_matrix((n, n)), sp.csr_matrix((n, n))
    c0, c1 = rayleigh_coefficients(settings["damping"])
    Kf = (K + Kabs).tocsr()
    C = (Cabs + (M * c0 + Kf * c1)).tocsr()
    return Kf, M, C, (c0, c1)


# --------------------------------------------------------------------------------------------------------------
# loads
# --------------------------------------------------------------------------------------------------------------
class LoadSchedule:
    """External force vector per time index (force_external.py:55-74); pulse / heaviside / moving."""

    def __init__(self, model: Model, loading: dict, time: np.ndarray):
        self.model, self.time = model, time
        self.kind = loading["type"]
        self.steps = loading.get("ini_steps", 5)
        self.factor = loading["force"]
        self.nodes = loading["node"]
        self.n_eq = model.number_eq
        if len(time) <= self.steps:
            raise SystemExit("Error: Number of loading steps smaller than " + str(self.steps))
        s = self.steps
        if self.kind == "pulse":
            self.sf = np.append(np.linspace(0, 1, int((s - 1) / 2), endpoint=False), np.linspace(1, 0, int((s + 1) / 2)))
        elif self.kind in ("heaviside", "moving"):
            self.sf = np.ones(len(time))
            self.sf[:s] = np.linspace(0, 1, s)
        else:
            raise SystemExit(f'Error: Load type {self.kind} not supported')
        self.ids = list(model.nodes[:, 0].astype(int))
        if self.kind == "moving":
            nd = model.nodes
            idx = np.where(nd[:, 0] == self.nodes)[0][0]
            lst = np.where((nd[:, 1] == nd[idx, 1]) & (nd[:, 2] == nd[idx, 2]))[0]
            dist = np.array([np.sqrt((nd[i, 3] - nd[idx, 3]) ** 2) for i in lst])
            self.idx_list = lst[np.argsort(dist)]
            self.node_dist = np.sort(dist)
            speed = np.ones(len(time)) * loading["speed"]
            speed[:s] = 0
            self.load_dist = speed * (time - time[s])

    def _put(self, f, node_id, scale):
        row = self.ids.index(node_id)
        for i, eq in enumerate(self.model.eq_nb_dof[row]):
            if not np.isnan(eq):
                f[int(eq)] = float(self.factor[i]) * scale

    def __call__(self, t: int) -> np.ndarray:
        f = np.zeros(self.n_eq)
        if self.kind == "pulse":
            if t < self.steps - 1:
                for n in self.nodes:
                    self._put(f, n, self.sf[t])
        elif self.kind == "heaviside":
            for n in self.nodes:
                self._put(f, n, self.sf[t])
        else:  # moving (force_external.py:320-345, including the `x * l` quirk)
            if self.load_dist[t] >= np.max(self.node_dist):
                return f
            k = np.where(self.node_dist <= self.load_dist[t])[0][-1]
            nd = self.model.nodes
            pair = [int(nd[self.idx_list[k], 0]), int(nd[self.idx_list[k + 1], 0])]
            x = self.load_dist[t] - nd[self.idx_list[k], 3] + nd[self.idx_list[0], 3]
            l = self.node_dist[k + 1] - self.node_dist[k]
            shp = [1 - x / l, x * l]
            for j, n in enumerate(pair):
                self._put(f, n, shp[j] * self.sf[t])
        return f


# --------------------------------------------------------------------------------------------------------------
# time integration
# --------------------------------------------------------------------------------------------------------------
def newmark(M, C, K, force, time: np.ndarray, output_interval: int = 1, beta: float = 0.25, gamma: float = 0.5):
    """Incremental Newmark.  `force(t_index) -> ndarray(n_eq)`.  Returns u, v, a, output_time."""
    M, C, K = sp.csc_matrix(M), sp.csc_matrix(C), sp.csc_matrix(K)
    n = M.shape[0]
    nt = len(time)
    dt = (time[-1] - time[0]) / (nt - 1)
    out_idx = np.arange(0, nt, output_interval)
    U = np.zeros((len(out_idx), n)); V = np.zeros_like(U); A = np.zeros_like(U)
    u = np.zeros(n); v = np.zeros(n)
    f_prev = force(0)
    a = spla.splu(M).solve(f_prev - C @ v - K @ u)
    A[0] = a
    Khat = (K + C * (gamma / (beta * dt)) + M * (1.0 / (beta * dt * dt))).tocsc()
    lu = spla.splu(Khat)
    row = 1
    for t in range(1, nt):
        f = force(t)
        rhs = (f - f_prev) + M @ (v / (beta * dt) + a / (2 * beta)) + C @ ((gamma / beta) * v + dt * (gamma / (2 * beta) - 1) * a)
        du = lu.solve(rhs)
        dv = gamma / (beta * dt) * du - (gamma / beta) * v + dt * (1 - gamma / (2 * beta)) * a
        da = du / (beta * dt * dt) - v / (beta * dt) - a / (2 * beta)
        u = u + du; v = v + dv; a = a + da
        f_prev = f
        if t % output_interval == 0:
            U[row], V[row], A[row] = u, v, a
            row += 1
    return U, V, A, time[out_idx]


def lump_rows(M) -> np.ndarray:
    return np.asarray(sp.csr_matrix(M).sum(axis=1)).ravel()


def central_difference(M, C, K, force, time: np.ndarray, output_interval: int = 1, c1: float = 0.0):
    """Explicit central difference with row-sum lumped M (Bathe, Table 9.1).  PARITY UNPINNED: the reference only forwards
    `Solver.CENTRAL_DIFFERENCE` to the un-vendored PuggleSolvers class (scatter/scatter.py:124-125) and ships no fixture.

    The damping C = C_abs + c0 M + c1 K (system_matrix.py:198) is split.  Row sums of K vanish away from supports (a rigid
    translation produces no force), so lumping c1 K would silently drop the stiffness-proportional damping; instead
    `c1` names that part, which acts on the lagged velocity (u(t) - u(t-dt))/dt and stays a full matrix, while the rest,
    c_d = rowsum(C - c1 K) = c0 m + rowsum(C_abs), is diagonal and centred in time:

        a0 = 1/dt^2, a1 = 1/(2dt), g = c1/dt
        u(-dt) = u0 - dt v0 + dt^2/2 acc0,   acc0 = (F0 - K (u0 + c1 v0) - c_d v0) / m
        (a0 m + a1 c_d) u(t+dt) = F(t) - K [(1+g) u(t) - g u(t-dt)] + 2 a0 m u(t) - (a0 m - a1 c_d) u(t-dt)
        v(t) = a1 (u(t+dt) - u(t-dt)),   a(t) = a0 (u(t+dt) - 2 u(t) + u(t-dt))
    """
    K = sp.csr_matrix(K)
    m = lump_rows(M)
    c = lump_rows(C) - c1 * lump_rows(K)
    n = len(m)
    nt = len(time)
    dt = (time[-1] - time[0]) / (nt - 1)
    a0, a1 = 1.0 / dt ** 2, 1.0 / (2 * dt)
    a2 = 2 * a0
    g = c1 / dt
    inv_d = 1.0 / (a0 * m + a1 * c)
    out_idx = np.arange(0, nt, output_interval)
    U = np.zeros((len(out_idx), n)); V = np.zeros_like(U); A = np.zeros_like(U)
    u = np.zeros(n)
    acc0 = (force(0) - K @ u) / m
    u_prev = u + 0.5 * dt * dt * acc0
    row = 0
    for t in range(nt):
        w = (1.0 + g) * u - g * u_prev if g != 0.0 else u
        u_next = inv_d * (force(t) - K @ w + a2 * m * u - (a0 * m - a1 * c) * u_prev)
        if t % output_interval == 0:
            U[row] = u
            V[row] = a1 * (u_next - u_prev)
            A[row] = a0 * (u_next - 2 * u + u_prev)
            row += 1
        u_prev, u = u, u_next
    return U, V, A, time[out_idx]


def time_array(total_time: float, time_step: float) -> np.ndarray:
    """scatter.py:117"""
    return np.linspace(0, total_time, int(np.ceil(total_time / time_step) + 1))


def run_case(mesh_file: str, materials: dict, bc: dict, settings: dict, loading: dict, time_step: float,
             solver: str = "newmark", elem_props=None):
    """End-to-end oracle run: the CPU path of `scatter.scatter` (scatter.py:36-171) without file output."""
    loading = dict(loading)
    loading.setdefault("ini_steps", 5)
    model = build_model(mesh_file, bc)
    K, M, C, _ = system_matrices(model, materials, settings, elem_props)
    time = time_array(loading["time"], time_step)
    force = LoadSchedule(model, loading, time)
    oi = settings.get("output_interval", 1)
    if solver == "newmark":
        U, V, A, t_out = newmark(M, C, K, force, time, oi)
    else:
        U, V, A, t_out = central_difference(M, C, K, force, time, oi, c1=rayleigh_coefficients(settings["damping"])[1])
    return model, (K, M, C), (U, V, A, t_out)


def model_from_readmesh(m) -> Model:
    """Oracle `Model` view of a product `ReadMesh` object (tests: lets the oracle assemble partitioned / synthetic meshes)."""
    om = Model(nodes=np.asarray(m.nodes), elem=np.asarray(m.elem), materials_index=np.asarray(m.materials_index),
               materials=m.materials, element_type=m.element_type, dimension=m.dimension, BC=m.BC, BC_dir=m.BC_dir,
               eq_nb_dof=m.eq_nb_dof, type_BC=m.type_BC, number_eq=m.number_eq, eq_nb_elem=m.eq_nb_elem,
               lower_element_type=m.lower_element_type, nb_nodes_lower_elem=m.nb_nodes_lower_elem,
               nb_nodes_elem=m.nb_nodes_elem)
    om.extra["node_rows"] = m.node_rows()
    return om


def bathe(M, C, K, force, time: np.ndarray, output_interval: int = 1):
    """Bathe composite scheme (trapezoidal rule over dt/2, then 3-point backward Euler).  PARITY UNPINNED: the reference
    only forwards `Solver.BATHE` to the un-vendored PuggleSolvers class and ships no fixture for it; this is the textbook
    scheme (Bathe & Baig 2005) with the half-step force taken as the mean of the two step forces."""
    M, C, K = sp.csc_matrix(M), sp.csc_matrix(C), sp.csc_matrix(K)
    n = M.shape[0]
    nt = len(time)
    dt = (time[-1] - time[0]) / (nt - 1)
    out_idx = np.arange(0, nt, output_interval)
    U = np.zeros((len(out_idx), n)); V = np.zeros_like(U); A = np.zeros_like(U)
    u = np.zeros(n); v = np.zeros(n)
    a = spla.splu(M).solve(force(0) - C @ v - K @ u)
    A[0] = a
    lu1 = spla.splu((K + C * (4 / dt) + M * (16 / dt ** 2)).tocsc())
    lu2 = spla.splu((K + C * (3 / dt) + M * (9 / dt ** 2)).tocsc())
    row = 1
    for t in range(1, nt):
        f_half = 0.5 * (force(t - 1) + force(t))
        u1 = lu1.solve(f_half + M @ (16 * u / dt ** 2 + 8 * v / dt + a) + C @ (4 * u / dt + v))
        v1 = 4 * (u1 - u) / dt - v
        u2 = lu2.solve(force(t) + M @ (12 * u1 / dt ** 2 - 3 * u / dt ** 2 + 4 * v1 / dt - v / dt) + C @ (4 * u1 / dt - u / dt))
        v2 = (u - 4 * u1 + 3 * u2) / dt
        a2 = (v - 4 * v1 + 3 * v2) / dt
        u, v, a = u2, v2, a2
        if t % output_interval == 0:
            U[row], V[row], A[row] = u, v, a
            row += 1
    return U, V, A, time[out_idx]


def static(K, force, time: np.ndarray, output_interval: int = 1):
    """K u(t) = F(t) for every time index.  PARITY UNPINNED (no reference fixture for Solver.STATIC)."""
    lu = spla.splu(sp.csc_matrix(K))
    nt = len(time)
    out_idx = np.arange(0, nt, output_interval)
    U = np.zeros((len(out_idx), K.shape[0]))
    for row, t in enumerate(out_idx):
        U[row] = lu.solve(force(int(t)))
    return U, time[out_idx]


# ---- boundary faces / top surface / moving load on the top plane (plain loops; small cases only) -----------------
_HEX8_FACES = [[0, 1, 2, 3], [0, 1, 4, 5], [4, 5, 6, 7], [2, 3, 6, 7], [0, 3, 4, 7], [1, 2, 5, 6]]


def boundary_faces_hexa8(model: Model) -> np.ndarray:
    """mesher.py:328-391 -- faces (node ids) whose nodes all touch fewer than 8 elements, unique rows."""
    ids = model.nodes[:, 0].astype(int)
    is_bnd = np.array([np.count_nonzero(model.elem == n) < 8 for n in ids])
    bnd_ids = set(ids[is_bnd].tolist())
    faces = []
    for el in model.elem:
        flags = np.array([int(n) in bnd_ids for n in el])
        if not flags.any():
            continue
        if flags.sum() > 4:
            for f in _HEX8_FACES:
                if flags[f].all():
                    faces.append(el[f])
        elif flags.sum() == 4:
            faces.append(el[flags])
    return np.unique(np.array(faces), axis=0)


def top_surface_faces(model: Model, faces: np.ndarray) -> np.ndarray:
    """mesher.py:393-424 -- boundary faces whose centroid is strictly inside the x/z extent and above the bottom."""
    eps = 1e-10
    cen = np.array([model.nodes[f - 1, 1:].mean(axis=0) for f in faces])
    keep = [(c[0] > cen[:, 0].min() + eps) and (c[0] < cen[:, 0].max() - eps) and (c[1] > cen[:, 1].min() + eps)
            and (c[2] > cen[:, 2].min() + eps) and (c[2] < cen[:, 2].max() - eps) for c in cen]
    return faces[np.array(keep)]


def _inside_convex(poly, pt) -> bool:
    """strict interior test by ray casting on the angularly sorted polygon"""
    c = poly.mean(axis=0)
    order = np.argsort(np.arctan2(poly[:, 1] - c[1], poly[:, 0] - c[0]))
    p = poly[order]
    inside = False
    n = len(p)
    for i in range(n):
        x1, y1 = p[i]
        x2, y2 = p[(i + 1) % n]
        if (y1 > pt[1]) != (y2 > pt[1]):
            xi = x1 + (pt[1] - y1) * (x2 - x1) / (y2 - y1)
            if xi > pt[0]:
                inside = not inside
    return inside


class MovingAtPlaneLoad:
    """force_external.py:151-212, 281-318 -- point load travelling over the top surface, bilinear-like nodal weights."""

    def __init__(self, model: Model, loading: dict, time: np.ndarray, top_faces: np.ndarray):
        self.model, self.time = model, time
        s = loading.get("ini_steps", 5)
        self.sf = np.ones(len(time)); self.sf[:s] = np.linspace(0, 1, s)
        self.factor = loading["force"]
        dist = np.zeros(len(time))
        dist[s:] = np.append(0, np.cumsum(loading["speed"] * np.diff(time[s:])))
        d = loading["direction"]
        ang = (0.5 * np.pi if d[1] > 0 else -0.5 * np.pi) if np.isclose(d[0], 0) else np.arctan(d[1] / d[0])
        self.pos = np.array([np.cos(ang) * dist + loading["start_coord"][0], np.sin(ang) * dist + loading["start_coord"][1]])
        self.active = []
        for p in self.pos.T:
            hit = None
            for f in top_faces:
                if _inside_convex(model.nodes[f - 1][:, [1, 3]], p):
                    hit = f
                    break
            self.active.append(hit)

    def __call__(self, t: int) -> np.ndarray:
        f = np.zeros(self.model.number_eq)
        el = self.active[t]
        xz = self.model.nodes[el - 1][:, [1, 3]]
        dx, dz = np.abs(xz[:, 0] - self.pos[0, t]), np.abs(xz[:, 1] - self.pos[1, t])
        wx = (dx < 1e-10) * 1 if np.any(dx < 1e-10) else 1 / dx
        wz = (dz < 1e-10) * 1 if np.any(dz < 1e-10) else 1 / dz
        w = wx * wz
        w = w / w.sum()
        load = np.array(self.factor) * self.sf[t]
        eq = self.model.eq_nb_dof[el - 1]
        for a in range(len(el)):
            for d in range(3):
                if not np.isnan(eq[a, d]):
                    f[int(eq[a, d])] = w[a] * load[d]
        return f


# ---------------------------------------------------------------------------------------------------------------------
# Random field by the randomisation method (scatter/random_fields.py:59-104 calls gstools SRF; gstools==1.7.0 is an
# un-vendored dependency -- **parity unpinned** against it.  Restated from its published algorithm (Hesse et al. 2014,
# gstools.field.generator.RandMeth): field(x) = mean + sqrt(var/N) sum_j (z1_j cos(k_j.x) + z2_j sin(k_j.x)).
def srf_field(pos: np.ndarray, k: np.ndarray, z1: np.ndarray, z2: np.ndarray, scale: float, mean: float, lognormal: bool) -> np.ndarray:
    """Sequential sum over the modes (the order gstools' `summate` and the CUDA kernel use)."""
    pos = np.asarray(pos, dtype=float)
    acc = np.zeros(len(pos))
    for j in range(len(z1)):
        ph = (k[j, 0] * pos[:, 0] + k[j, 1] * pos[:, 1]) + k[j, 2] * pos[:, 2]
        acc = acc + (z1[j] * np.cos(ph) + z2[j] * np.sin(ph))
    f = mean + scale * acc
    return np.exp(f) if lognormal else f
