"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the reference's golden vectors.

Tolerances (BASELINE.json north_star): DOF numbering and CSR pattern bit-exact; K/M/C entries <= 1e-12 relative to the
matrix max-norm; displacement / velocity histories <= 1e-8 relative L2.
"""
import hashlib
import os

import numpy as np
import pytest
import scipy.sparse as sp

import cases
from conftest import max_rel, probe_vector, rel_l2

pytestmark = pytest.mark.gpu

TOL_MAT = 1e-12
TOL_HIST = 1e-8


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build(path, bc, materials, settings, explicit=False, elem_props=None):
    from scatter_b200 import mesher, system_matrix
    m = mesher.ReadMesh(path)
    m.read_gmsh(); m.read_bc(bc); m.mapping(); m.connectivities(); m.get_mesh_edges()
    mx = system_matrix.GenerateMatrix(m.number_eq, settings["int_order"])
    mx.want_full_mass = True
    mx.want_lumped_mass = True
    mx.generate_stiffness_and_mass(m, materials, elem_props=elem_props)
    mx.absorbing_boundaries(m, materials, settings["absorbing_BC"], settings["absorbing_BC_stiff"])
    mx.damping_Rayleigh(settings["damping"])
    return m, mx


@pytest.mark.parametrize("case", list(cases.MATRIX_CASES))
def test_assembly_matches_oracle_and_reference(case, golden_meshes, golden_matrices, oracle):
    fn, bc = cases.MATRIX_CASES[case]
    G = golden_matrices
    mat, sett = cases.case_materials(case), cases.settings()
    m, mx = build(golden_meshes[fn], bc, mat, sett)
    # numbering (bit-exact with the reference)
    assert m.number_eq == int(G[case + "__n_eq"])
    assert np.array_equal(m.equation_table_int(), G[case + "__eq_nb_dof"])
    # pattern (bit-exact with the reference's structural pattern)
    rowptr, col = mx.pattern()
    assert rowptr.dtype == np.int64 and col.dtype == np.int32
    assert len(col) == int(G[case + "__nnz"])
    assert sha(rowptr) + sha(col) == str(G[case + "__pattern_sha"])
    # values against the oracle: structural K (before absorbing springs are visible only through Kf) and M
    om = oracle.build_model(golden_meshes[fn], bc)
    Kf, Mo, Co, _ = oracle.system_matrices(om, mat, sett)
    n = m.number_eq
    K = mx.K; M = mx.M; C = mx.C
    Ko_s = sp.csr_matrix(Kf); Ko_s.sort_indices()
    dK = abs(K - Kf).max() / abs(Kf).max()
    dM = abs(M - Mo).max() / abs(Mo).max()
    dC = abs(C - Co).max() / abs(Co).max()
    assert dK <= TOL_MAT and dM <= TOL_MAT and dC <= TOL_MAT, (dK, dM, dC)
    # values against the reference itself (probes / samples written by oracle/make_golden.py)
    x = probe_vector(n)
    assert max_rel(K @ x, G[case + "__Kfx"]) <= 1e-11
    assert max_rel(M @ x, G[case + "__Mx"]) <= 1e-11
    assert max_rel(C @ x, G[case + "__Cfx"]) <= 1e-11
    if (case + "__Mdata") in G.files:
        assert np.abs(M.data - G[case + "__Mdata"]).max() <= TOL_MAT * float(G[case + "__Mmax"])
    else:
        idx = G[case + "__sample_idx"]
        assert np.abs(M.data[idx] - G[case + "__Msample"]).max() <= TOL_MAT * float(G[case + "__Mmax"])
    # device SpMV against scipy on the same values
    y = mx.ctx.spmv(0, x)
    assert max_rel(y, K @ x) <= 1e-14
    # lumped mass = row sums of the consistent mass
    ml = mx.lumped_mass()
    assert max_rel(ml, np.asarray(M.sum(axis=1)).ravel()) <= 1e-13


@pytest.mark.parametrize("etype,order", [("hexa8", 1), ("hexa8", 3), ("hexa20", 3), ("quad4", 1), ("quad4", 3), ("tri3", 1),
                                         ("tri3", 3), ("tri6", 3), ("tetra4", 1), ("tetra10", 1), ("hexa20", 1), ("tri6", 1)])
def test_assembly_other_orders(etype, order, golden_meshes, oracle):
    fn, bc = {"hexa8": cases.MATRIX_CASES["column"], "hexa20": cases.MATRIX_CASES["column_high_order"],
              "quad4": cases.MATRIX_CASES["column_2D"], "tri3": cases.MATRIX_CASES["column_2D_tri3"],
              "tri6": cases.MATRIX_CASES["column_2D_tri6"], "tetra4": cases.MATRIX_CASES["column_3D_tetra4"],
              "tetra10": cases.MATRIX_CASES["column_3D_tetra10"]}[etype]
    mat, sett = cases.materials(), cases.settings(int_order=order)
    m, mx = build(golden_meshes[fn], bc, mat, sett)
    om = oracle.build_model(golden_meshes[fn], bc)
    Kf, Mo, Co, _ = oracle.system_matrices(om, mat, sett)
    assert abs(mx.K - Kf).max() / abs(Kf).max() <= TOL_MAT
    assert abs(mx.M - Mo).max() / max(abs(Mo).max(), 1e-300) <= TOL_MAT


def test_quad8_elements(oracle, tmp_path):
    """quad8 is unreachable through the reference's mesh reader (mesher.py:181) but is one of the eight element types
    of the path (face element of hexa20): assemble a small distorted quad8 patch through the C ABI directly."""
    from scatter_b200 import _lib
    rng = np.random.default_rng(3)
    nx, ny = 3, 2
    # corner lattice + mid-side nodes
    ids = {}
    pts = []

    def node(x, y):
        key = (round(x * 2), round(y * 2))
        if key not in ids:
            ids[key] = len(pts)
            pts.append([x + 0.05 * rng.uniform(-1, 1), y + 0.05 * rng.uniform(-1, 1), 0.0])
        return ids[key]

    conn = []
    for j in range(ny):
        for i in range(nx):
            c = [node(i, j), node(i + 1, j), node(i + 1, j + 1), node(i, j + 1),
                 node(i + 0.5, j), node(i + 1, j + 0.5), node(i + 0.5, j + 1), node(i, j + 0.5)]
            conn.append(c)
    xyz = np.array(pts); conn = np.array(conn, dtype=np.int32)
    nn = len(xyz)
    bcode = np.zeros((nn, 2), dtype=int)
    bcode[xyz[:, 1] < 0.2, 1] = 1
    bcode[xyz[:, 0] < 0.2, 0] = 1
    free = bcode != 1
    eq = np.where(free.ravel(), np.cumsum(free.ravel()) - 1, -1).reshape(nn, 2)
    n_eq = int(free.sum())
    E = 30e6 * rng.uniform(0.5, 2, len(conn)); nu = rng.uniform(0.1, 0.35, len(conn)); rho = 1500 * rng.uniform(0.8, 1.2, len(conn))
    for order in (2, 3):
        ctx = _lib.Context(0)
        ctx.set_mesh("quad8", xyz, conn, eq, n_eq)
        ctx.set_materials(E, nu, rho)
        ctx.build_pattern()
        ctx.assemble(order, _lib.ASM_K | _lib.ASM_M_FULL)
        rowptr, col = ctx.get_pattern()
        K = sp.csr_matrix((ctx.get_values(0), col, rowptr), shape=(n_eq, n_eq))
        M = sp.csr_matrix((ctx.get_values(1), col, rowptr), shape=(n_eq, n_eq))
        Ke, Me = oracle.element_matrices("quad8", order, xyz[conn], E, nu, rho)
        Kd = np.zeros((n_eq, n_eq)); Md = np.zeros((n_eq, n_eq))
        for e in range(len(conn)):
            g = eq[conn[e]].ravel()
            ok = g >= 0
            Kd[np.ix_(g[ok], g[ok])] += Ke[e][np.ix_(ok, ok)]
            Md[np.ix_(g[ok], g[ok])] += Me[e][np.ix_(ok, ok)]
        assert np.abs(K.toarray() - Kd).max() <= TOL_MAT * np.abs(Kd).max()
        assert np.abs(M.toarray() - Md).max() <= TOL_MAT * np.abs(Md).max()
        ctx.close()


def _dense_from_elements(oracle, etype, order, xyz, conn, eq, n_eq, E, nu, rho):
    Ke, Me = oracle.element_matrices(etype, order, xyz[conn], E, nu, rho)
    Kd = np.zeros((n_eq, n_eq)); Md = np.zeros((n_eq, n_eq))
    for e in range(len(conn)):
        g = eq[conn[e]].ravel()
        ok = g >= 0
        Kd[np.ix_(g[ok], g[ok])] += Ke[e][np.ix_(ok, ok)]
        Md[np.ix_(g[ok], g[ok])] += Me[e][np.ix_(ok, ok)]
    return Kd, Md


def _gpu_dense(etype, order, xyz, conn, eq, n_eq, E, nu, rho):
    from scatter_b200 import _lib
    ctx = _lib.Context(0)
    ctx.set_mesh(etype, xyz, conn, eq, n_eq)
    ctx.set_materials(E, nu, rho)
    ctx.build_pattern()
    ctx.assemble(order, _lib.ASM_K | _lib.ASM_M_FULL | _lib.ASM_M_LUMPED)
    rowptr, col = ctx.get_pattern()
    K = sp.csr_matrix((ctx.get_values(0), col, rowptr), shape=(n_eq, n_eq)).toarray()
    M = sp.csr_matrix((ctx.get_values(1), col, rowptr), shape=(n_eq, n_eq)).toarray()
    Ml = ctx.get_lumped_mass()
    ctx.close()
    return K, M, Ml


@pytest.mark.parametrize("etype,mesh,orders", [("tri3", "column_2D_tri3.msh", (1, 2)), ("tri6", "column_2D_tri6.msh", (2, 3)),
                                               ("quad4", "column_2D.msh", (1, 2, 3)), ("tetra4", "column_3D_tetra4.msh", (1, 2)),
                                               ("tetra10", "column_3D_tetra10.msh", (1, 2)), ("hexa8", "cube.msh", (1, 2, 3)),
                                               ("hexa20", "column_high_order.msh", (1, 2, 3))])
def test_single_distorted_element(etype, mesh, orders, golden_meshes, oracle):
    """Smallest possible mesh (one element, every dof free) with perturbed nodes: the Jacobian differs at every Gauss point
    and the assembled matrices are Ke / Me themselves."""
    from scatter_b200 import mesher
    m = mesher.ReadMesh(golden_meshes[mesh]); m.read_gmsh()
    rng = np.random.default_rng(11)
    rows = m.node_rows()[0]
    xyz = m.nodes[rows, 1:].copy()
    dim = m.dimension
    size = np.ptp(xyz[:, :dim], axis=0).max()
    xyz[:, :dim] += 0.06 * size * rng.uniform(-1, 1, (len(rows), dim))
    nne = len(rows)
    conn = np.arange(nne, dtype=np.int32)[None, :]
    eq = np.arange(nne * dim).reshape(nne, dim)
    n_eq = nne * dim
    E, nu, rho = np.array([12.5e6]), np.array([0.31]), np.array([1830.0])
    for order in orders:
        K, M, Ml = _gpu_dense(etype, order, xyz, conn, eq, n_eq, E, nu, rho)
        Kd, Md = _dense_from_elements(oracle, etype, order, xyz, conn, eq, n_eq, E, nu, rho)
        assert np.abs(K - Kd).max() <= TOL_MAT * np.abs(Kd).max(), (etype, order)
        assert np.abs(M - Md).max() <= TOL_MAT * np.abs(Md).max(), (etype, order)
        assert np.abs(Ml - Md.sum(axis=1)).max() <= TOL_MAT * np.abs(Md).max(), (etype, order)
        assert np.abs(K - K.T).max() <= 1e-13 * np.abs(K).max()


@pytest.mark.parametrize("nt", [150, 300])
def test_high_valence_node_uses_the_warp_per_node_kernel(nt, oracle):
    """A fan of 150 / 300 triangles around one node: more (node, element) pairs than a block of the pair kernels holds, so the
    assembly falls back to the warp-per-node kernel; neighbour lists of 151 nodes exercise the long-list paths of the
    pattern builder as well."""
    rng = np.random.default_rng(5)
    ang = np.linspace(0, 2 * np.pi, nt, endpoint=False)
    r = 1.0 + 0.2 * rng.uniform(-1, 1, nt)
    xyz = np.zeros((nt + 1, 3))
    xyz[1:, 0] = r * np.cos(ang); xyz[1:, 1] = r * np.sin(ang)
    conn = np.array([[0, 1 + i, 1 + (i + 1) % nt] for i in range(nt)], dtype=np.int32)
    free = np.ones((nt + 1, 2), dtype=bool)
    free[1, :] = False; free[40, 1] = False
    eq = np.where(free.ravel(), np.cumsum(free.ravel()) - 1, -1).reshape(nt + 1, 2)
    n_eq = int(free.sum())
    E = 30e6 * rng.uniform(0.5, 2, nt); nu = rng.uniform(0.1, 0.35, nt); rho = 1500 * rng.uniform(0.8, 1.2, nt)
    K, M, Ml = _gpu_dense("tri3", 2, xyz, conn, eq, n_eq, E, nu, rho)
    Kd, Md = _dense_from_elements(oracle, "tri3", 2, xyz, conn, eq, n_eq, E, nu, rho)
    assert np.abs(K - Kd).max() <= TOL_MAT * np.abs(Kd).max()
    assert np.abs(M - Md).max() <= TOL_MAT * np.abs(Md).max()
    assert np.abs(Ml - Md.sum(axis=1)).max() <= TOL_MAT * np.abs(Md).max()


def test_zero_load_and_zero_steps(golden_meshes):
    """Degenerate runs: no load at all (PCG sees a zero right-hand side), zero time steps, free vibration from an initial
    velocity only."""
    m, mx = build(golden_meshes["cube.msh"], cases.BC_CUBE, cases.materials(), cases.settings(damping=[1, 0.01, 30, 0.01]))
    ctx = mx.ctx
    n = m.number_eq
    ctx.set_load_schedule(np.zeros(12, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0))
    ctx.set_state(None, None)
    u, v, a, st = ctx.run_newmark(1e-3, 0, 10, 5)
    assert u.shape == (3, n) and not u.any() and not v.any() and not a.any()
    u, v, a, st = ctx.run_central_difference(2e-4, 0, 10, 5)
    assert u.shape == (3, n) and not u.any() and not v.any()
    u, v, a, st = ctx.run_newmark(1e-3, 0, 0, 1)
    assert u.shape == (1, n) and not u.any()
    v0 = np.zeros(n); v0[n // 2] = 1e-3
    ctx.set_state(None, v0)
    u, v, a, st = ctx.run_newmark(1e-3, 0, 10, 5)
    assert np.isfinite(u).all() and np.abs(u[-1]).max() > 0 and np.array_equal(v[0], v0)
    ctx.close()


def test_assembly_is_bit_reproducible(golden_meshes):
    fn, bc = cases.MATRIX_CASES["cube"]
    _, a = build(golden_meshes[fn], bc, cases.materials(), cases.settings())
    _, b = build(golden_meshes[fn], bc, cases.materials(), cases.settings())
    assert sha(a.ctx.get_values(0)) == sha(b.ctx.get_values(0))
    assert sha(a.ctx.get_values(1)) == sha(b.ctx.get_values(1))
    k1 = a.ctx.get_values(0)
    a.ctx.assemble(2, 3)
    assert sha(k1) != "" and np.array_equal(k1, a.ctx.get_values(0))


@pytest.mark.parametrize("case", ["cube", "column_3D_tetra4", "column_2D", "column_2D_tri6", "column_high_order", "column_3D_tetra10",
                                  "rose_2D_side"])
def test_assembly_kernel_generations_agree(case, golden_meshes, monkeypatch):
    """The persistent TMA-fed kernel (default: k_elem_records + k_assemble_tma), k_assemble_blk (Jacobian set-up inside every
    block) and the warp-per-node k_assemble sum the same contributions in the same order.  The first two do the same
    arithmetic per pair -- identical bits; the last differs in how a single element contribution is rounded (material law
    per Gauss point vs once)."""
    if case not in cases.MATRIX_CASES:
        pytest.skip("case not in the fixture set")
    fn, bc = cases.MATRIX_CASES[case]
    vals = {}
    from scatter_b200 import _lib
    for name, opt, value in (("tma", None, None), ("blk", "assembly_records", 0), ("generic", "generic_assembly", 1)):
        if opt:
            monkeypatch.setitem(_lib.DEFAULT_OPTIONS, opt, value)
        _, mx = build(golden_meshes[fn], bc, cases.case_materials(case), cases.settings())
        vals[name] = (mx.ctx.get_values(0), mx.ctx.get_values(1), mx.ctx.get_lumped_mass())
        mx.ctx.close()
        if opt:
            monkeypatch.delitem(_lib.DEFAULT_OPTIONS, opt)
    for other in ("blk",):
        for a, b in zip(vals["tma"], vals[other]):
            assert a.shape == b.shape and np.array_equal(a, b), other
    for a, b in zip(vals["tma"], vals["generic"]):
        assert a.shape == b.shape
        assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()


def test_box_mesh_random_field(oracle, tmp_path):
    """Synthetic structured boxes (the benchmark generator) with per-element random Young's modulus."""
    from scatter_b200 import boxmesh, system_matrix
    for et, n in (("hexa8", 6), ("hexa20", 3)):
        path = os.path.join(tmp_path, f"box_{et}.msh")
        boxmesh.write_box_msh(path, n, n + 1, n - 1, 0.5, et)
        bc = boxmesh.box_boundaries(n, n + 1, n - 1, 0.5)
        model = boxmesh.box_model(n, n + 1, n - 1, 0.5, et)
        model.connectivities()
        om = oracle.build_model(path, bc)
        assert om.number_eq == model.number_eq
        assert np.array_equal(np.nan_to_num(om.eq_nb_dof, nan=-1), np.nan_to_num(model.eq_nb_dof, nan=-1))
        ne = len(model.elem)
        E = boxmesh.lognormal_young(ne); nu = np.full(ne, 0.2); rho = np.full(ne, 1500.0)
        mx = system_matrix.GenerateMatrix(model.number_eq, 2)
        mx.generate_stiffness_and_mass(model, None, elem_props=(E, nu, rho))
        Ko, Mo = oracle.assemble_global(om, E, nu, rho, 2)
        rowptr, col = mx.pattern()
        assert np.array_equal(rowptr, Ko.indptr) and np.array_equal(col, Ko.indices)
        assert np.abs(mx.ctx.get_values(0) - Ko.data).max() <= TOL_MAT * np.abs(Ko.data).max()
        assert np.abs(mx.ctx.get_values(1) - Mo.data).max() <= TOL_MAT * np.abs(Mo.data).max()


def test_mid_size_box_vs_oracle(oracle):
    """32^3 hexa8 box (105 k dofs) of the benchmark generator: big enough for interior column-dictionary nodes, z-face tiles
    with explicit lists, several tiles per CTA and both consumer groups of the node-blocked SpMV.  Pattern bit for bit, K and
    M against the oracle, and a fused central-difference history (lagged stiffness-proportional damping) against the
    oracle's (a sparse LU of 100 k 3-D equations for the Newmark oracle takes > 10 min: the implicit solver is checked on
    smaller boxes)."""
    from scatter_b200 import _lib, boxmesh, system_matrix
    s, h = 32, 0.5
    model = boxmesh.box_model(s, s, s, h, "hexa8"); model.connectivities()
    ne, n = len(model.elem), model.number_eq
    E = boxmesh.lognormal_young(ne, 30e6, 1e6, seed=11); nu = np.full(ne, 0.2); rho = np.full(ne, 1500.0)
    Ko, Mo = oracle.assemble_global(oracle.model_from_readmesh(model), E, nu, rho, 2)
    Ko = sp.csr_matrix(Ko); Mo = sp.csr_matrix(Mo)
    mx = system_matrix.GenerateMatrix(n, 2)
    ctx = mx.ctx
    ctx.set_mesh("hexa8", model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), n, None)
    ctx.set_materials(E, nu, rho)
    ctx.build_pattern()
    ctx.assemble(2, _lib.ASM_K | _lib.ASM_M_FULL | _lib.ASM_M_LUMPED)
    rowptr, col = ctx.get_pattern()
    assert np.array_equal(rowptr, Ko.indptr) and np.array_equal(col, Ko.indices)
    assert np.abs(ctx.get_values(_lib.MAT_K) - Ko.data).max() <= TOL_MAT * np.abs(Ko.data).max()
    assert np.abs(ctx.get_values(_lib.MAT_M) - Mo.data).max() <= TOL_MAT * np.abs(Mo.data).max()
    assert ctx.pattern_stats()["dict_patterns"] > 0
    damping = [1, 0.01, 30, 0.01]
    mx.damping_Rayleigh(damping)
    c0, c1 = oracle.rayleigh_coefficients(damping)
    d = int(model.eq_nb_dof[boxmesh.top_centre_node(s, s, s) - 1, 1])
    nt = 42

    def force(t):
        f = np.zeros(n); f[d] = -1000.0 * min(1.0, t / 4.0)
        return f
    ctx.set_load_schedule(np.arange(nt + 1, dtype=np.int64), np.full(nt, d, dtype=np.int64), -1000.0 * np.minimum(1.0, np.arange(nt) / 4.0))
    # explicit: dt at 0.3 of the CFL estimate of the stiffest element
    dt = 0.3 * h / np.sqrt(E.max() * (1 - 0.2) / ((1 + 0.2) * (1 - 0.4)) / 1500.0)
    Uc, Vc, _, _ = oracle.central_difference(Mo, Mo * c0 + Ko * c1, Ko, force, np.arange(41) * dt, 10, c1=c1)
    ctx.set_state(None, None)
    u, v, _, _ = ctx.run_central_difference(dt, 0, 40, 10)
    assert rel_l2(u, Uc) <= TOL_HIST and rel_l2(v, Vc) <= TOL_HIST
    ctx.close()


@pytest.mark.parametrize("case", ["cube", "cube_abs", "rose_2D_side", "column_3D_tetra4", "box"])
def test_column_dictionary_is_bitwise_neutral(case, golden_meshes, monkeypatch, tmp_path):
    """node_dict.cu replaces the explicit column lists of nodes with a frequent relative list by a dictionary id: SpMV, the
    fused central-difference step and a Newmark stage must give bit-identical results with and without it."""
    from scatter_b200 import _lib, boxmesh, solvers, system_matrix
    out = {}
    for mode in ("dict", "explicit"):
        if mode == "explicit":
            monkeypatch.setitem(_lib.DEFAULT_OPTIONS, "column_dictionary", 0)
        if case == "box":
            model = boxmesh.box_model(14, 9, 11, 0.5, "hexa8")
            model.connectivities()
            ne = len(model.elem)
            mx = system_matrix.GenerateMatrix(model.number_eq, 2)
            mx.generate_stiffness_and_mass(model, None, elem_props=(boxmesh.lognormal_young(ne), np.full(ne, 0.2), np.full(ne, 1500.0)))
            mx.damping_Rayleigh([1, 0.01, 30, 0.01])
            m = model
        else:
            fn, bc = cases.MATRIX_CASES[case]
            m, mx = build(golden_meshes[fn], bc, cases.case_materials(case), cases.settings(damping=[1, 0.01, 30, 0.01]))
        ctx = mx.ctx
        st = ctx.pattern_stats()
        n = m.number_eq
        x = probe_vector(n)
        y = ctx.spmv(0, x)
        ptr = np.arange(41, dtype=np.int64)
        ctx.set_load_schedule(ptr, np.full(40, n // 2, dtype=np.int64), np.full(40, -1000.0))
        ctx.set_state(None, None)
        ucd, vcd, _, _ = ctx.run_central_difference(1e-5, 0, 30, 10)
        ctx.set_state(None, None)
        unm, _, anm, _ = ctx.run_newmark(1e-3, 0, 6, 3)
        out[mode] = (st, y, ucd, vcd, unm, anm)
        ctx.close()
        if mode == "explicit":
            monkeypatch.delitem(_lib.DEFAULT_OPTIONS, "column_dictionary")
    sd, se = out["dict"][0], out["explicit"][0]
    assert se["dict_patterns"] == 0 and sd["nnz"] == se["nnz"]
    if case in ("cube", "cube_abs", "box"):
        assert sd["node_blocked"] == 1 and sd["dict_patterns"] > 0 and sd["node_col_entries"] < se["node_col_entries"]
    for a, b in zip(out["dict"][1:], out["explicit"][1:]):
        assert np.isfinite(a).all() and np.abs(a).max() > 0
        assert np.array_equal(a, b)


@pytest.mark.parametrize("model_name", ["Gaussian", "Exponential", "Matern"])
def test_random_field_kernel_vs_oracle(model_name, oracle):
    """k_srf (randomisation method) against the sequential CPU restatement on the same modes; anisotropic 3-D and 2-D."""
    from scatter_b200 import _lib, random_fields
    rng = np.random.default_rng(3)
    ctx = _lib.Context(0)
    for dim, n in ((3, 3001), (2, 517), (3, 1)):
        pos = rng.uniform(-40, 60, (n, 3))
        if dim == 2:
            pos[:, 2] = 0.0
        sf = random_fields.SpectralField(model_name, dim, var=0.01, mean=17.2, len_scale=[15.0, 1.5, 3.0], angles=0.3 if dim == 2 else 0.0,
                                         seed=26021981)
        for logn in (False, True):
            ref = oracle.srf_field(sf.isometrize(pos), sf.k, sf.z1, sf.z2, np.sqrt(sf.var / sf.mode_no), sf.mean, logn)
            got = sf(pos, lognormal=logn, ctx=ctx)
            assert got.shape == ref.shape and np.isfinite(got).all()
            assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max(), (model_name, dim, logn, np.abs(got - ref).max())
            assert np.array_equal(got, sf(pos, lognormal=logn, ctx=ctx))                 # reproducible
    out, sec = ctx.srf_sample(np.zeros((0, 3)), sf.k, sf.z1, sf.z2, 1.0, 0.0, False)      # no points: no launch
    assert out.shape == (0,)
    ctx.close()


def test_scatter_with_random_field(golden_meshes, oracle, tmp_path):
    """BASELINE config 2/3 shape: multi-material embankment mesh, lognormal random Young's modulus in one soil layer
    (run_scatter_rose_2D.py:62-76 without the train).  The field itself is pinned in test_random_field_kernel_vs_oracle;
    here the whole entry point is checked against the oracle run with the materials the field produced."""
    from scatter_b200 import scatter
    case = "rose_2D_side"
    fn, bc = cases.MATRIX_CASES[case]
    mats = cases.case_materials(case)
    sett = cases.settings(damping=[1, 0.005, 20, 0.005], VTK=True, VTK_binary=False, output_interval=5)
    load = {"force": [0, -1e4, 0], "node": [4], "time": 0.1, "type": "heaviside", "ini_steps": 5}
    out = os.path.join(tmp_path, "rf2d")
    res = scatter(golden_meshes[fn], out, mats, bc, sett, dict(load), time_step=1e-3, random_props=cases.rf_properties(case, "Gaussian"))
    young = np.array([v["Young"] for k, v in mats.items() if k.startswith("material_")])
    # one realisation over a domain of a few correlation lengths: its mean sits within a couple of field standard deviations (6 %)
    assert len(young) == 602 and abs(young.mean() / 500e5 - 1) < 0.15 and 0.2 < young.std() / 3e6 < 3.0
    assert os.path.isfile(os.path.join(out, "rf_props.txt"))
    # oracle with the same per-element materials (tags 0..N-1 = field elements, the others shifted)
    om = oracle.build_model(golden_meshes[fn], bc)
    tag_soil = [m[1] for m in om.materials if m[2] == "soil1"][0]
    base = cases.case_materials(case)
    E, nu, rho = oracle.element_properties(om, base)
    E[np.asarray(om.materials_index) == tag_soil] = young
    _, _, (U, V, A, _) = oracle.run_case(golden_meshes[fn], base, bc, sett, load, 1e-3, elem_props=(E, nu, rho))
    assert np.abs(U).max() > 0
    assert rel_l2(res.dis, U) <= TOL_HIST and rel_l2(res.vel, V) <= TOL_HIST
    with open(os.path.join(out, "VTK", "data_2.vtk")) as f:
        txt = f.read().splitlines()
    i = txt.index("SCALARS material_prop_Young double")
    vtk_young = np.array([float(t) for t in txt[i + 2:i + 2 + len(om.elem)]])
    assert np.abs(vtk_young - E).max() <= 1e-9 * E.max()


# ---- time histories ----------------------------------------------------------------------------------------------
def run_history(name, golden_meshes, solver="newmark"):
    from scatter_b200 import force_external, solvers
    c = cases.history_case(name)
    load = dict(c["loading"]); load.setdefault("ini_steps", 5)
    m, mx = build(golden_meshes[c["mesh"]], c["bc"], c["materials"], c["settings"])
    time = np.linspace(0, load["time"], int(np.ceil(load["time"] / c["time_step"]) + 1))
    num = solvers.NewmarkExplicit() if solver == "newmark" else solvers.CentralDifferenceSolver()
    num.output_interval = c["settings"].get("output_interval", 1)
    num.initialise(m.number_eq, time)
    num.bind(mx)
    F = force_external.Force()
    F.initialise_load(load, time, m, num, top_surface_elements=[])
    num.update_rhs_at_time_step_func = F.update_load_at_t
    num.update(0)
    num.calculate(None, None, None, F.force_vector, 0, len(time) - 1)
    return m, mx, num


def test_newmark_hexa8_pulse_vs_reference_golden(golden_meshes, golden_histories):
    H = golden_histories
    m, mx, num = run_history("hexa8_pulse", golden_meshes)
    eq = m.eq_nb_dof
    uy = np.zeros((num.u.shape[0], len(eq))); vy = np.zeros_like(uy)
    free = ~np.isnan(eq[:, 1])
    uy[:, free] = num.u[:, eq[free, 1].astype(int)]
    vy[:, free] = num.v[:, eq[free, 1].astype(int)]
    steps = H["hexa8_pulse__steps"]
    assert rel_l2(uy[steps], H["hexa8_pulse__uy"]) <= TOL_HIST
    assert rel_l2(vy[steps], H["hexa8_pulse__vy"]) <= TOL_HIST
    sel = H["hexa8_pulse__nodes_full"]
    assert rel_l2(uy[:, sel], H["hexa8_pulse__uy_full"]) <= TOL_HIST
    assert rel_l2(vy[:, sel], H["hexa8_pulse__vy_full"]) <= TOL_HIST
    assert num.stats[0]["pcg_iterations"] > 0


@pytest.mark.parametrize("pcg_path", ["cooperative", "graph", "eager", "graph_jacobi", "graph_fsai_no_projection", "eager_jacobi_projection"])
def test_newmark_quad4_heaviside_vs_reference_golden(pcg_path, golden_meshes, golden_histories, monkeypatch):
    # the three PCG drivers (one cooperative kernel for small systems; stream-ordered iterations replayed from a CUDA
    # graph, or launched one by one as on multi-GPU runs) must all reproduce the reference history -- the stream-ordered
    # ones with the FSAI preconditioner and the projection onto previous solutions (the defaults), and with either switched off
    from scatter_b200 import _lib
    if pcg_path != "cooperative":
        monkeypatch.setitem(_lib.DEFAULT_OPTIONS, "small_pcg", 0)
    if pcg_path.startswith("eager"):
        monkeypatch.setitem(_lib.DEFAULT_OPTIONS, "pcg_graph", 0)
    if "jacobi" in pcg_path:
        monkeypatch.setitem(_lib.DEFAULT_OPTIONS, "fsai", 0)
    if pcg_path in ("graph_jacobi", "graph_fsai_no_projection"):
        monkeypatch.setitem(_lib.DEFAULT_OPTIONS, "pcg_projection", 0)
    H = golden_histories
    m, mx, num = run_history("quad4_heaviside", golden_meshes)
    ids = list(m.nodes[:, 0].astype(int))
    eq = m.eq_nb_dof
    for name, arr in (("displacement", num.u), ("velocity", num.v), ("acceleration", num.a)):
        gold = H["quad4_heaviside__" + name]          # (nn, 2, nt)
        mine = np.zeros_like(gold)
        for k, nid in enumerate(H["quad4_heaviside__nodes"]):
            i = ids.index(int(nid))
            for d in range(2):
                if not np.isnan(eq[i, d]):
                    mine[k, d] = arr[:, int(eq[i, d])]
        assert rel_l2(mine, gold) <= (TOL_HIST if name != "acceleration" else 1e-7), name
    assert np.allclose(num.output_time, H["quad4_heaviside__time"])


@pytest.mark.parametrize("etype", ["tri3", "tri6", "tetra4", "tetra10"])
@pytest.mark.parametrize("stream_pcg", [False, True])
def test_newmark_benchmark_set_2_vs_reference_golden(etype, stream_pcg, golden_meshes, golden_histories, monkeypatch):
    if stream_pcg:        # FSAI-preconditioned, projected PCG (what large systems run) instead of the cooperative kernel
        from scatter_b200 import _lib
        monkeypatch.setitem(_lib.DEFAULT_OPTIONS, "small_pcg", 0)
    H = golden_histories
    m, mx, num = run_history(etype, golden_meshes)
    assert num.u.shape[0] == 201
    assert rel_l2(num.u[:, 0], H[etype + "__uy"][0::10]) <= TOL_HIST
    assert rel_l2(num.v[:, 0], H[etype + "__vy"][0::10]) <= TOL_HIST
    # the reference's own assertion (test_benchmark_set_2.py:93-94): 6 decimals
    np.testing.assert_array_almost_equal(num.u[:, 0], H[etype + "__uy"][0::10])
    np.testing.assert_array_almost_equal(num.v[:, 0], H[etype + "__vy"][0::10])


@pytest.mark.parametrize("stream_pcg", [False, True])
def test_newmark_absorbing_and_hexa20_vs_oracle(stream_pcg, golden_meshes, oracle, monkeypatch):
    """Cases whose reference goldens are missing blobs: compare with the (pinned) oracle instead.  stream_pcg: the FSAI +
    projection driver of large systems on the effective matrix with absorbing dashpots / springs and on hexa20."""
    from scatter_b200 import _lib, force_external, solvers
    if stream_pcg:
        monkeypatch.setitem(_lib.DEFAULT_OPTIONS, "small_pcg", 0)
    for mesh, bc, kind in (("column.msh", cases.BC_COLUMN_ABS, "heaviside"), ("column_high_order.msh", cases.BC_COLUMN, "pulse")):
        load = {"force": [0, -1000, 0], "node": [3, 4, 7, 8], "time": 0.1, "type": kind, "ini_steps": 5}
        sett = cases.settings()
        model, mats, (U, V, A, tt) = oracle.run_case(golden_meshes[mesh], cases.materials(), bc, sett, load, 0.5e-3)
        m, mx = build(golden_meshes[mesh], bc, cases.materials(), sett)
        time = oracle.time_array(load["time"], 0.5e-3)
        num = solvers.NewmarkExplicit(); num.initialise(m.number_eq, time); num.bind(mx)
        F = force_external.Force(); F.initialise_load(load, time, m, num)
        num.update_rhs_at_time_step_func = F.update_load_at_t
        num.update(0); num.calculate(None, None, None, F.force_vector, 0, len(time) - 1)
        assert rel_l2(num.u, U) <= TOL_HIST and rel_l2(num.v, V) <= TOL_HIST and rel_l2(num.a, A) <= 1e-7


def test_newmark_moving_load_vs_oracle(golden_meshes, oracle):
    """integration_test.py:493-545 (moving load on cube.msh); the reference's golden pickle is a missing blob, so the
    history is compared with the pinned oracle (first 60 of the 201 steps)."""
    from scatter_b200 import force_external, solvers
    mat = cases.materials(); mat["solid"]["Young"] = 10e6
    sett = cases.settings(damping=[1, 0.0, 30, 0.0])
    load = {"force": [0, -1000, 0], "node": 8, "time": 0.3, "type": "moving", "speed": 10, "ini_steps": 20}
    model, mats, (U, V, A, tt) = oracle.run_case(golden_meshes["cube.msh"], mat, cases.BC_CUBE, sett, load, 0.5e-2)
    m, mx = build(golden_meshes["cube.msh"], cases.BC_CUBE, mat, sett)
    time = oracle.time_array(load["time"], 0.5e-2)
    num = solvers.NewmarkExplicit(); num.initialise(m.number_eq, time); num.bind(mx)
    F = force_external.Force(); F.initialise_load(load, time, m, num)
    num.update_rhs_at_time_step_func = F.update_load_at_t
    num.update(0); num.calculate(None, None, None, F.force_vector, 0, len(time) - 1)
    assert np.abs(U).max() > 0
    assert rel_l2(num.u, U) <= TOL_HIST and rel_l2(num.v, V) <= TOL_HIST and rel_l2(num.a, A) <= 1e-7


def test_central_difference_vs_oracle(golden_meshes, oracle):
    from scatter_b200 import force_external, solvers
    mesh, bc = "cube.msh", cases.BC_CUBE
    mat = cases.materials()
    sett = cases.settings(damping=[1, 0.01, 30, 0.01])
    dt = 2e-4     # h = 1 m, vp ~ 150 m/s: well inside the stability limit
    load = {"force": [0, -1000, 0], "node": [8], "time": 0.05, "type": "heaviside", "ini_steps": 5}
    model, (K, M, C), (U, V, A, tt) = oracle.run_case(golden_meshes[mesh], mat, bc, dict(sett, output_interval=5), load, dt, solver="cd")
    m, mx = build(golden_meshes[mesh], bc, mat, sett)
    time = oracle.time_array(load["time"], dt)
    num = solvers.CentralDifferenceSolver(); num.output_interval = 5
    num.initialise(m.number_eq, time); num.bind(mx)
    F = force_external.Force(); F.initialise_load(load, time, m, num)
    num.update_rhs_at_time_step_func = F.update_load_at_t
    num.update(0)
    # two stages: exercises the restart hook calculate(t0, t1)
    half = (len(time) - 1) // 2 // 5 * 5
    num.calculate(None, None, None, F.force_vector, 0, half)
    num.calculate(None, None, None, F.force_vector, half, len(time) - 1)
    errs = (rel_l2(num.u, U), rel_l2(num.v, V), rel_l2(num.a, A))
    rows = [rel_l2(num.u[k], U[k]) for k in range(1, len(U))]
    assert errs[0] <= TOL_HIST and errs[1] <= TOL_HIST and errs[2] <= 1e-7, (errs, rows[:3], rows[24:28], rows[-2:])


def test_central_difference_reports_divergence(golden_meshes):
    """A time step far above the stability limit must raise, not hand back NaN/Inf histories."""
    from scatter_b200 import force_external, solvers
    from scatter_b200._lib import ScatterB200Error
    mesh, bc = "cube.msh", cases.BC_CUBE
    mat, sett = cases.materials(), cases.settings(damping=[1, 0.01, 30, 0.01])
    dt = 0.5
    load = {"force": [0, -1000, 0], "node": [8], "time": 200.0, "type": "heaviside", "ini_steps": 5}
    m, mx = build(golden_meshes[mesh], bc, mat, sett)
    time = np.linspace(0, load["time"], int(np.ceil(load["time"] / dt) + 1))
    num = solvers.CentralDifferenceSolver(); num.output_interval = 50
    num.initialise(m.number_eq, time); num.bind(mx)
    F = force_external.Force(); F.initialise_load(load, time, m, num)
    num.update_rhs_at_time_step_func = F.update_load_at_t
    num.update(0)
    with pytest.raises(ScatterB200Error, match="diverged"):
        num.calculate(None, None, None, F.force_vector, 0, len(time) - 1)


@pytest.mark.parametrize("stream_pcg", [False, True])
def test_rows_without_entries_do_not_disturb_pcg(stream_pcg, golden_meshes, oracle, monkeypatch):
    """Rank 0's sub-domain of a two-way decomposition, run alone on one GPU with no halo exchange: ghost nodes own equation
    numbers but empty rows, so the system reduces to the owned block with the ghost dofs held at zero.  Regression test:
    the SpMV must write q = 0 on the empty rows, otherwise stale values leak into the PCG residual norms.  With the
    stream-ordered driver the FSAI factor leaves the ghost rows / columns out (block preconditioner per rank)."""
    from scatter_b200 import _lib, mesher, partition, system_matrix
    if stream_pcg:
        monkeypatch.setitem(_lib.DEFAULT_OPTIONS, "small_pcg", 0)
    m = mesher.ReadMesh(golden_meshes["cube.msh"])
    m.read_gmsh(); m.read_bc(cases.BC_CUBE); m.mapping(); m.connectivities()
    owner = partition.owner_by_slabs(m, 2, axis=2)
    dom = partition.partition_model(m, owner, 0)
    loc = dom.model
    ne = len(loc.elem)
    mx = system_matrix.GenerateMatrix(loc.number_eq, 2)
    mx.generate_stiffness_and_mass(loc, None, elem_props=(np.full(ne, 10e6), np.full(ne, 0.2), np.full(ne, 1500.0)), active=dom.active)
    mx.damping_Rayleigh([1, 0.01, 30, 0.01])
    ctx = mx.ctx
    own = np.asarray(dom.owned_eq)
    ghost = np.setdiff1d(np.arange(loc.number_eq), own)
    assert len(ghost) > 0
    nt = 21
    d = int(own[len(own) // 2])
    ramp = np.ones(nt); ramp[:5] = np.linspace(0, 1, 5)
    ctx.set_load_schedule(np.arange(nt + 1, dtype=np.int64), np.full(nt, d, dtype=np.int64), -1000.0 * ramp)
    # dirty the scratch vectors on the ghost rows first (an explicit run with a non-zero initial velocity leaves
    # non-zero accelerations there and shares its scratch with the PCG), then the implicit solve from rest
    ctx.set_state(None, np.ones(loc.number_eq))
    ctx.run_central_difference(2e-4, 0, 10, 5)
    ctx.set_state(None, None)
    u, v, a, st = ctx.run_newmark(5e-3, 0, nt - 1, 5)
    assert np.all(u[:, ghost] == 0.0)
    # oracle: the owned block of the single-domain matrices
    om = oracle.model_from_readmesh(m)
    neg = len(m.elem)
    K, M = oracle.assemble_global(om, np.full(neg, 10e6), np.full(neg, 0.2), np.full(neg, 1500.0), 2)
    c0, c1 = oracle.rayleigh_coefficients([1, 0.01, 30, 0.01])
    g = np.asarray(dom.global_eq_of_owned)
    Koo = sp.csr_matrix(K)[g][:, g]; Moo = sp.csr_matrix(M)[g][:, g]
    Coo = c0 * Moo + c1 * Koo
    k = int(np.where(own == d)[0][0])

    def force(t):
        f = np.zeros(len(g)); f[k] = -1000.0 * ramp[min(t, nt - 1)]
        return f
    U, V, A = oracle.newmark(Moo, Coo, Koo, force, np.arange(nt) * 5e-3, output_interval=5)[:3]
    assert rel_l2(u[:, own], U) <= TOL_HIST and rel_l2(v[:, own], V) <= TOL_HIST
    ctx.close()


@pytest.mark.parametrize("stream_pcg", [False, True])
def test_bathe_and_static_vs_oracle(stream_pcg, golden_meshes, oracle, monkeypatch):
    """Solver.BATHE / Solver.STATIC have no reference fixture: compare the device loops with the oracle's textbook
    restatement, and Bathe with the pinned Newmark oracle (both second order)."""
    from scatter_b200 import _lib, force_external, solvers
    if stream_pcg:        # FSAI-preconditioned stream-ordered PCG (two factors for Bathe's two effective matrices)
        monkeypatch.setitem(_lib.DEFAULT_OPTIONS, "small_pcg", 0)
    mesh, bc = "column.msh", cases.BC_COLUMN
    mat, sett = cases.materials(), cases.settings()
    load = {"force": [0, -1000, 0], "node": [3, 4, 7, 8], "time": 0.05, "type": "heaviside", "ini_steps": 20}
    dt = 5e-4
    om = oracle.build_model(golden_meshes[mesh], bc)
    K, M, C, _ = oracle.system_matrices(om, mat, sett)
    time = oracle.time_array(load["time"], dt)
    force = oracle.LoadSchedule(om, load, time)
    Ub, Vb, Ab, _ = oracle.bathe(M, C, K, force, time, 2)
    Un, Vn, An, _ = oracle.newmark(M, C, K, force, time, 2)
    Us, _ = oracle.static(K, force, time[:12], 1)
    assert rel_l2(Ub, Un) < 2e-2          # same physics, different second-order schemes
    for cls, ref in ((solvers.BatheSolver, (Ub, Vb, Ab)), (solvers.StaticSolver, (Us,))):
        m, mx = build(golden_meshes[mesh], bc, mat, sett)
        num = cls()
        static = cls is solvers.StaticSolver
        tt = time[:12] if static else time
        num.output_interval = 1 if static else 2
        num.initialise(m.number_eq, tt); num.bind(mx)
        F = force_external.Force(); F.initialise_load(load, tt if not static else time, m, num)
        num.update_rhs_at_time_step_func = F.update_load_at_t
        if static:
            num.calculate(None, F.force_vector, 0, len(tt) - 1)
            assert rel_l2(num.u, ref[0]) <= 1e-8
        else:
            num.update(0)
            num.calculate(None, None, None, F.force_vector, 0, len(tt) - 1)
            assert rel_l2(num.u, ref[0]) <= TOL_HIST and rel_l2(num.v, ref[1]) <= TOL_HIST and rel_l2(num.a, ref[2]) <= 1e-7


def test_thin_slab_with_tiles_of_ghost_nodes(monkeypatch):
    """Rank 1 of an 8-slab partition of a 24^3 box, run alone: three owned element layers between two ghost planes, so whole
    tiles of the node-blocked SpMV hold ghost nodes only (their stage carries descriptors and nothing else).  Regression:
    the PCG product read its dot-product operand from the unloaded stage and 0 * stale shared memory turned the solve into
    NaN (first seen in the 8-GPU parity check of bench.py).  The row-wise kernels are the reference."""
    from scatter_b200 import _lib, partition, system_matrix
    hist = {}
    for name, opts in (("node", {}), ("rowwise", {"node_spmv": 0, "tma_spmv": 0})):
        monkeypatch.setitem(_lib.DEFAULT_OPTIONS, "small_pcg", 0)
        for k, v in opts.items():
            monkeypatch.setitem(_lib.DEFAULT_OPTIONS, k, v)
        dom = partition.slab_partition(24, 24, 3, 1, 8, 0.5, "hexa8")
        model = dom.model
        ne = len(model.elem)
        mx = system_matrix.GenerateMatrix(model.number_eq, 2)
        ctx = mx.ctx
        ctx.set_mesh("hexa8", model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), model.number_eq, dom.active)
        ctx.set_materials(np.full(ne, 30e6), np.full(ne, 0.2), np.full(ne, 1500.0))
        ctx.build_pattern(); ctx.assemble(2, _lib.ASM_K | _lib.ASM_M_FULL)
        mx.damping_Rayleigh([1, 0.01, 30, 0.01])
        nt = 12
        d = int(dom.owned_eq[len(dom.owned_eq) // 2])
        ctx.set_load_schedule(np.arange(nt + 1, dtype=np.int64), np.full(nt, d, dtype=np.int64), -1000.0 * np.minimum(1.0, np.arange(nt) / 4.0))
        ctx.set_state(None, None)
        u, v, a, st = ctx.run_newmark(5e-4, 0, 10, 5, rtol=1e-13)
        assert np.isfinite(u).all() and np.abs(u).max() > 0
        hist[name] = u
        ctx.close()
    assert rel_l2(hist["node"], hist["rowwise"]) <= 1e-9


def test_hexa20_box_stream_pcg_vs_oracle(oracle, monkeypatch):
    """10^3 hexa20 box (12 k dofs) through the kernels a large box runs: node-blocked SpMV with two consumer groups on a ring
    with an odd number of stages (regression: a group running ahead took the other group's completed phase for its own),
    FSAI + projection PCG.  Products against scipy, 40 Newmark steps against the oracle's direct solve."""
    from scatter_b200 import _lib, boxmesh, system_matrix
    monkeypatch.setitem(_lib.DEFAULT_OPTIONS, "small_pcg", 0)
    s, h, dt, nst = 10, 0.5, 5e-4, 40
    pm = boxmesh.box_model(s, s, s, h, "hexa20"); pm.connectivities()
    ne, n = len(pm.elem), pm.number_eq
    E = boxmesh.lognormal_young(ne, 30e6, 1e6, seed=3)
    Ko, Mo = oracle.assemble_global(oracle.model_from_readmesh(pm), E, np.full(ne, 0.2), np.full(ne, 1500.0), 2)
    c0, c1 = oracle.rayleigh_coefficients([1, 0.01, 30, 0.01])
    d = int(pm.eq_nb_dof[boxmesh.top_centre_node(s, s, s) - 1, 1])

    def force(t):
        f = np.zeros(n); f[d] = -1000.0 * min(1.0, t / 4.0)
        return f
    Uo, Vo, _, _ = oracle.newmark(Mo, Mo * c0 + Ko * c1, Ko, force, np.arange(nst + 1) * dt, 5)
    mx = system_matrix.GenerateMatrix(n, 2)
    ctx = mx.ctx
    ctx.set_mesh("hexa20", pm.nodes[:, 1:], pm.node_rows(), pm.equation_table_int(), n, None)
    ctx.set_materials(E, np.full(ne, 0.2), np.full(ne, 1500.0))
    ctx.build_pattern(); ctx.assemble(2, _lib.ASM_K | _lib.ASM_M_FULL)
    mx.damping_Rayleigh([1, 0.01, 30, 0.01])
    Ks = sp.csr_matrix(Ko)
    x = np.sin(0.37 * np.arange(n) + 0.11)
    ref = Ks @ x
    for _ in range(30):                                    # the race was timing dependent
        assert np.abs(ctx.spmv(_lib.MAT_K, x) - ref).max() <= 1e-12 * np.abs(ref).max()
    ctx.set_load_schedule(np.arange(nst + 2, dtype=np.int64), np.full(nst + 1, d, dtype=np.int64), -1000.0 * np.minimum(1.0, np.arange(nst + 1) / 4.0))
    ctx.set_state(None, None)
    u, v, _, st = ctx.run_newmark(dt, 0, nst, 5, rtol=1e-13)
    assert rel_l2(u, Uo) <= TOL_HIST and rel_l2(v, Vo) <= TOL_HIST
    info = ctx.precond_info()
    assert info["fsai_nnz"] > n and info["projection_vectors"] > 0
    assert st["pcg_iterations"] / nst < 150               # Jacobi needs ~330 per step on this matrix
    ctx.close()


def test_scatter_newmark_implicit_equals_explicit(golden_meshes, tmp_path):
    """`Solver.NEWMARK_IMPLICIT` (scatter.py:118-121 picks NewmarkImplicitForce) is the total-force form of the same
    recurrence: for the linear systems of this path it must give the history of the default solver, through the whole
    `scatter(...)` call."""
    from scatter_b200 import scatter, Solver
    c = cases.history_case("quad4_heaviside")
    res = {}
    for name, sv in (("explicit", Solver.NEWMARK_EXPLICIT), ("implicit", Solver.NEWMARK_IMPLICIT)):
        out = os.path.join(tmp_path, name)
        res[name] = scatter(golden_meshes[c["mesh"]], out, c["materials"], c["bc"], dict(c["settings"]), dict(c["loading"]),
                            time_step=c["time_step"], solver=sv)
    assert np.abs(res["explicit"].dis).max() > 0
    for f in ("dis", "vel", "acc"):
        assert np.array_equal(getattr(res["explicit"], f), getattr(res["implicit"], f)), f


def test_scatter_entry_point_writes_reference_layout(golden_meshes, golden_histories, tmp_path):
    from scatter_b200 import scatter
    c = cases.history_case("quad4_heaviside")
    sett = dict(c["settings"], VTK=True, VTK_binary=False)
    out = os.path.join(tmp_path, "res2d")
    res = scatter(golden_meshes[c["mesh"]], out, c["materials"], c["bc"], sett, dict(c["loading"]), time_step=c["time_step"])
    import pickle
    with open(os.path.join(out, "data.pickle"), "rb") as f:
        data = pickle.load(f)
    H = golden_histories
    assert list(data["nodes"]) == list(H["quad4_heaviside__nodes"])
    for k, nid in enumerate(H["quad4_heaviside__nodes"]):
        for d, lab in enumerate("xy"):
            np.testing.assert_almost_equal(data["displacement"][str(int(nid))][lab], H["quad4_heaviside__displacement"][k, d], decimal=5)
    # VTK file: same lines as the reference's golden file (numbers to 5 decimals, integration_test.py:304-319)
    with open(os.path.join(out, "VTK", "data_3.vtk")) as f:
        mine = f.read().splitlines()
    gold = str(H["quad4_heaviside__vtk_step3"]).splitlines()
    assert len(mine) == len(gold)
    for a, b in zip(mine, gold):
        ta, tb = a.split(), b.split()
        try:
            fb = [float(t) for t in tb]
        except ValueError:
            assert a == b
            continue
        np.testing.assert_almost_equal([float(t) for t in ta], fb, decimal=5)
    assert res.dis.shape == (201, res.eq_nb_dof[~np.isnan(res.eq_nb_dof)].size)
