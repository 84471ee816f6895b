"""GPU tests of the boundary seams added in round 2: caller-supplied matrices (`sc_set_csr`), output selections
(`sc_set_output_dofs`), the always-stored last step, the re-run protocol, and the explicit scheme against the analytical
column solution.  Everything goes through the C ABI; the oracle is only the checker."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import cases
from conftest import rel_l2
from test_gpu_parity import build

pytestmark = pytest.mark.gpu
TOL_HIST = 1e-8


def _oracle_system(oracle, golden_meshes, case="cube", damping=(1, 0.01, 30, 0.01), bc=None):
    fn, bc0 = cases.MATRIX_CASES[case]
    om = oracle.build_model(golden_meshes[fn], bc or bc0)
    K, M, C, _ = oracle.system_matrices(om, cases.case_materials(case), cases.settings(damping=list(damping)))
    return om, K, M, C


def _point_load(n, dof, nt, scale=-1000.0):
    def force(t):
        f = np.zeros(n)
        f[dof] = scale * min(t, 4) / 4.0
        return f
    ptr = np.arange(nt + 1, dtype=np.int64)
    return force, (ptr, np.full(nt, dof, dtype=np.int64), np.array([scale * min(t, 4) / 4.0 for t in range(nt)]))


@pytest.mark.parametrize("case", ["cube", "cube_abs", "column_2D_tri6"])
def test_unbound_solver_integrates_caller_matrices(case, oracle, golden_meshes):
    """Seam B2 as the reference uses it (scatter.py:159): `calculate(matrix.M, matrix.C, matrix.K, F, t0, t1)` with scipy
    matrices this library did not assemble -- here the oracle's (= the reference's, oracle/VALIDATION.md), in the sparse
    formats the reference hands over (M lil, C and K csr with exact zeros pruned)."""
    from scatter_b200 import solvers
    om, K, M, C = _oracle_system(oracle, golden_meshes, case)
    n = om.number_eq
    Kc = sp.csr_matrix(K); Kc.eliminate_zeros()
    Cc = sp.csr_matrix(C); Cc.eliminate_zeros()
    Ml = sp.lil_matrix(M)
    nt = 41
    time = np.arange(nt) * 1e-3
    force, sched = _point_load(n, n // 2, nt)
    U, V, A, _ = oracle.newmark(M, C, K, force, time, 5)
    num = solvers.NewmarkExplicit(); num.output_interval = 5
    num.initialise(n, time)
    num.update_rhs_at_time_step_func = force                 # plain callback: evaluated once per step on the host
    num.update(0)
    num.calculate(Ml, Cc, Kc, force(0), 0, 20)               # two stages through the restart protocol
    num.update(20)
    num.calculate(Ml, Cc, Kc, force(0), 20, nt - 1)
    assert np.abs(U).max() > 0
    assert rel_l2(num.u, U) <= TOL_HIST and rel_l2(num.v, V) <= TOL_HIST and rel_l2(num.a, A) <= 1e-7
    # Bathe and the static solver on the same uploaded matrices
    Ub = oracle.bathe(M, C, K, force, time[:11], 5)[0]
    nb = solvers.BatheSolver(); nb.output_interval = 5
    nb.initialise(n, time[:11]); nb.update_rhs_at_time_step_func = force
    nb.update(0); nb.calculate(Ml, Cc, Kc, force(0), 0, 10)
    assert rel_l2(nb.u, Ub) <= TOL_HIST
    Us = oracle.static(K, force, time[:7], 3)[0]
    ns = solvers.StaticSolver(); ns.output_interval = 3
    ns.initialise(n, time[:7]); ns.update_rhs_at_time_step_func = force
    ns.calculate(Kc, force(0), 0, 6)
    assert rel_l2(ns.u, Us) <= 1e-7
    # explicit scheme: with caller matrices all of C is lumped by row sums (no Rayleigh split is known).  Not for tri6: the
    # row sums of a quadratic triangle's mass matrix vanish at the corner nodes, no explicit scheme runs on that
    if case == "column_2D_tri6":
        return
    dt = 1e-4
    t2 = np.arange(61) * dt
    Ucd = oracle.central_difference(M, C, K, force, t2, 10, c1=0.0)[0]
    nc = solvers.CentralDifferenceSolver(); nc.output_interval = 10
    nc.initialise(n, t2); nc.update_rhs_at_time_step_func = force
    nc.update(0); nc.calculate(Ml, Cc, Kc, force(0), 0, 60)
    assert rel_l2(nc.u, Ucd) <= TOL_HIST


def test_output_selection_and_last_row(oracle, golden_meshes):
    """`sc_set_output_dofs`: rows hold only the selected equations (any order); an output interval that does not divide the
    number of steps still stores the final state; both for the implicit and the explicit loop."""
    from scatter_b200 import solvers
    m, mx = build(golden_meshes["cube.msh"], cases.BC_CUBE, cases.materials(), cases.settings(damping=[1, 0.01, 30, 0.01]))
    om, K, M, C = _oracle_system(oracle, golden_meshes, "cube")
    c1 = oracle.rayleigh_coefficients([1, 0.01, 30, 0.01])[1]
    n = m.number_eq
    sel = np.array([n - 1, 5, n // 2, 17, 0], dtype=np.int64)
    for kind, dt, nt in (("newmark", 1e-3, 24), ("cd", 2e-4, 47)):
        time = np.arange(nt) * dt
        force, sched = _point_load(n, n // 2, nt)
        if kind == "newmark":
            U, V, A, _ = oracle.newmark(M, C, K, force, time, 1)
        else:
            U, V, A, _ = oracle.central_difference(M, C, K, force, time, 1, c1=c1)
        idx = list(range(0, nt, 5)) + ([nt - 1] if (nt - 1) % 5 else [])
        for output_dofs in (None, sel):
            num = solvers.NewmarkExplicit() if kind == "newmark" else solvers.CentralDifferenceSolver()
            num.output_interval = 5
            num.output_dofs = output_dofs
            num.initialise(n, time); num.bind(mx)
            num.load_schedule = sched
            num.update(0)
            half = 10 if kind == "newmark" else 23            # second stage starts off the output grid
            num.calculate(None, None, None, None, 0, half)
            num.calculate(None, None, None, None, half, nt - 1)
            cols = slice(None) if output_dofs is None else sel
            assert list(num.output_time_indices) == idx and num.u.shape == (len(idx), n if output_dofs is None else len(sel))
            assert rel_l2(num.u, U[idx][:, cols]) <= TOL_HIST and rel_l2(num.v, V[idx][:, cols]) <= TOL_HIST
            assert rel_l2(num.a, A[idx][:, cols]) <= 1e-7
    mx.ctx.set_output_dofs(None)
    mx.ctx.close()


def test_second_calculate_restarts_from_u0(oracle, golden_meshes):
    """a second `calculate(0, n)` on the same solver must reproduce the first history (start from u0 / v0), not
    continue from the end state on the device; a stage that continues the previous one must not re-upload."""
    from scatter_b200 import solvers
    m, mx = build(golden_meshes["cube.msh"], cases.BC_CUBE, cases.materials(), cases.settings(damping=[1, 0.01, 30, 0.01]))
    n = m.number_eq
    for cls, dt in ((solvers.NewmarkExplicit, 1e-3), (solvers.CentralDifferenceSolver, 2e-4)):
        nt = 21
        time = np.arange(nt) * dt
        _, sched = _point_load(n, n // 2, nt)
        num = cls(); num.output_interval = 4
        num.initialise(n, time); num.bind(mx); num.load_schedule = sched
        num.update(0); num.calculate(None, None, None, None, 0, nt - 1)
        first = (num.u.copy(), num.v.copy(), num.a.copy())
        assert np.abs(first[0]).max() > 0
        num.calculate(None, None, None, None, 0, nt - 1)          # no update(): u0 / v0 still are the initial rows
        assert np.array_equal(num.u, first[0]) and np.array_equal(num.v, first[1]) and np.array_equal(num.a, first[2])
        num.update(0); num.calculate(None, None, None, None, 0, 8)
        e = mx.ctx.state_epoch
        num.calculate(None, None, None, None, 8, nt - 1)
        assert mx.ctx.state_epoch == e + 1                        # only the run, no set_state
        assert np.array_equal(num.u, first[0]) and np.array_equal(num.v, first[1])
    mx.ctx.close()


def test_central_difference_on_device_vs_analytical_column(golden_meshes):
    """The explicit loop on the reference's hexa8 column against the closed-form wave solution (Churchill; the reference's
    integration_tests/analytical_solutions/analytical_wave_prop.py:50-81)."""
    from scatter_b200 import force_external, solvers
    c = cases.history_case("hexa8_pulse")
    mat = c["materials"]
    load = dict(c["loading"], time=0.3, type="heaviside", ini_steps=2)
    m, mx = build(golden_meshes[c["mesh"]], c["bc"], mat, dict(c["settings"], damping=[1, 0.0, 30, 0.0]))
    dt = 5e-5
    time = np.linspace(0, load["time"], int(np.ceil(load["time"] / dt) + 1))
    num = solvers.CentralDifferenceSolver(); num.output_interval = 10
    num.initialise(m.number_eq, time); num.bind(mx)
    F = force_external.Force(); F.initialise_load(load, time, m, num, top_surface_elements=[])
    num.update_rhs_at_time_step_func = F.update_load_at_t
    num.update(0); num.calculate(None, None, None, F.force_vector, 0, len(time) - 1)
    E, nu, rho = mat["solid"]["Young"], mat["solid"]["poisson"], mat["solid"]["density"]
    L, Kb = 20.0, E * (1 - nu) / ((1 + nu) * (1 - 2 * nu))
    p0 = -1000.0 * len(load["node"]) / (0.1 * 0.1)
    cw = np.sqrt(Kb / rho)
    k = np.arange(1, 400)[:, None]
    lam = (2 * k - 1) * np.pi / (2 * L)
    tt = num.output_time
    u_top = p0 / Kb * (L + 8 * L / np.pi ** 2 * ((-1.0) ** k / (2 * k - 1) ** 2 * np.sin(lam * L) * np.cos(lam * cw * tt[None, :])).sum(axis=0))
    top = int(m.eq_nb_dof[int(np.where(m.nodes[:, 0] == load["node"][0])[0][0]), 1])
    assert rel_l2(num.u[:, top], u_top) <= 0.01
    mx.ctx.close()


def test_mid_size_box_against_oracle(oracle):
    """Structured-box generator + interior column dictionary + explicitly listed z-face tiles together, at a size where the
    oracle still finishes in seconds: pattern bit-exact, K / lumped M <= 1e-12, explicit history
    <= 1e-8, all against the single-domain numpy restatement."""
    from scatter_b200 import _lib, boxmesh, system_matrix
    s = 28
    model = boxmesh.box_model(s, s, s, 0.5, "hexa8")
    model.connectivities()
    ne, n = len(model.elem), model.number_eq
    E = boxmesh.lognormal_young(ne)
    om = oracle.model_from_readmesh(model)
    Ko, Mo = oracle.assemble_global(om, E, np.full(ne, 0.2), np.full(ne, 1500.0), 2)
    Ko = sp.csr_matrix(Ko); Mo = sp.csr_matrix(Mo)
    mx = system_matrix.GenerateMatrix(n, 2)
    mx.want_full_mass, mx.want_lumped_mass = False, True
    mx.generate_stiffness_and_mass(model, None, elem_props=(E, np.full(ne, 0.2), np.full(ne, 1500.0)))
    mx.damping_Rayleigh([1, 0.01, 30, 0.01])
    ctx = mx.ctx
    rowptr, col = mx.pattern()
    assert np.array_equal(rowptr, Ko.indptr) and np.array_equal(col, Ko.indices)
    assert ctx.pattern_stats()["dict_patterns"] > 0
    assert np.abs(ctx.get_values(_lib.MAT_K) - Ko.data).max() <= 1e-12 * np.abs(Ko.data).max()
    ml = oracle.lump_rows(Mo)
    assert np.abs(ctx.get_lumped_mass() - ml).max() <= 1e-12 * ml.max()
    c0, c1 = oracle.rayleigh_coefficients([1, 0.01, 30, 0.01])
    nt = 41
    dof = int(model.eq_nb_dof[boxmesh.top_centre_node(s, s, s) - 1, 1])
    force, sched = _point_load(n, dof, nt)
    dt = 0.3 * 0.5 / np.sqrt(36e6 * 0.8 / (1.2 * 0.6) / 1500.0)
    U, V, A, _ = oracle.central_difference(sp.diags(ml), sp.diags(ml) * c0 + Ko * c1, Ko, force, np.arange(nt) * dt, 10, c1=c1)
    ctx.set_load_schedule(*sched)
    ctx.set_state(None, None)
    u, v, a, st = ctx.run_central_difference(dt, 0, nt - 1, 10)
    assert np.abs(U).max() > 0 and rel_l2(u, U) <= TOL_HIST and rel_l2(v, V) <= TOL_HIST
    ctx.close()


def test_reference_random_field_case_on_device(golden_meshes, tmp_path):
    """The reference's random-field test (integration_test.py:376-431) end to end on the GPU: gstools-compatible modes
    (`random_fields.gstools_modes`), field evaluation in `k_srf`, assembly, Newmark -- against the reference's own golden
    `results_rf_2d/data.pickle` (tests/golden/history_rf_2d.npz)."""
    from scatter_b200.scatter import scatter
    from test_pipeline_cpu import RF_2D_PROPS, check_rf_2d_golden
    sett = cases.settings(damping=[1, 0.005, 20, 0.005])
    load = {"force": [0, -1e6, 0], "node": [3, 4, 25], "time": 1.0, "type": "heaviside"}
    res = scatter(golden_meshes["column_2D.msh"], str(tmp_path), cases.materials(), cases.BC_2D, sett, load, time_step=5e-3,
                  random_props=dict(RF_2D_PROPS))
    check_rf_2d_golden(res)


def test_layered_box_with_moving_plane_load(oracle, tmp_path):
    """Substitute for BASELINE config 3 (`mesh/embankment_rose.msh` is a missing blob and its load needs the un-vendored ROSE
    package): a layered 3-D box with the three materials of run_scatter_rose_3D.py:35-43 (embankment over
    soil1 over soil2), absorbing bottom, and the in-tree moving load on the top surface (`moving_at_plane`,
    force_external.py:151-212,281-318; settings of integration_test.py:551-602), through `scatter(...)` against the oracle."""
    from scatter_b200 import boxmesh, gmsh_io
    from scatter_b200.scatter import scatter
    nx, ny, nz, h = 8, 6, 7, 1.0
    nodes, elem = boxmesh.box_arrays(nx, ny, nz, h, "hexa8")
    layer = (np.arange(len(elem)) // nx) % ny                       # element layer along y (x fastest, then y, then z)
    tags = np.where(layer >= 4, 1, np.where(layer >= 2, 2, 3))
    phys = [[3, 1, "embankment"], [3, 2, "soil1"], [3, 3, "soil2"]]
    path = os.path.join(tmp_path, "layers.msh")
    gmsh_io.write_msh(path, nodes, elem, tags, phys, "hexa8")
    mats = {"embankment": {"density": 2000, "Young": 100e6, "poisson": 0.2}, "soil1": {"density": 1700, "Young": 40e6, "poisson": 0.2},
            "soil2": {"density": 2000, "Young": 10e6, "poisson": 0.2}}
    bc = boxmesh.box_boundaries(nx, ny, nz, h, bottom="020")
    sett = cases.settings(damping=[1, 0.01, 30, 0.01], pickle_nodes=[int(boxmesh.top_centre_node(nx, ny, nz))], output_interval=5)
    load = {"force": [0, -1000, 0], "start_coord": [1.5, 1.5], "time": 0.1, "type": "moving_at_plane", "direction": [1, 0.5],
            "speed": 30, "ini_steps": 10}
    dt = 1e-3
    res = scatter(path, os.path.join(tmp_path, "out"), mats, bc, sett, dict(load), time_step=dt)
    om = oracle.build_model(path, bc)
    K, M, C, _ = oracle.system_matrices(om, mats, sett)
    time = oracle.time_array(load["time"], dt)
    top = oracle.top_surface_faces(om, oracle.boundary_faces_hexa8(om))
    force = oracle.MovingAtPlaneLoad(om, dict(load), time, top)
    U, V, A, tt = oracle.newmark(M, C, K, force, time, 5)
    assert len(set(tags)) == 3 and np.abs(U).max() > 0 and (np.asarray(om.type_BC) == "Absorb").any()
    assert rel_l2(res.dis, U) <= TOL_HIST and rel_l2(res.vel, V) <= TOL_HIST and np.allclose(res.time, tt)
    # the load really travels: the loaded dofs at the first and last loaded step differ
    f0, f1 = force(12), force(len(time) - 1)
    assert set(np.nonzero(f0)[0]) != set(np.nonzero(f1)[0])


@pytest.mark.parametrize("order20,vertex_first", [("grouped", 1), ("interleaved", 1), ("interleaved", 0)])
def test_hexa20_box_implicit_vs_oracle(order20, vertex_first, oracle, monkeypatch):
    """BASELINE config 4 at a size the oracle's direct solve finishes in seconds: hexa20 box of the benchmark generator in both
    node numberings (lattice by lattice / cell by cell), pattern bit-exact, K and consistent M <= 1e-12, and a Newmark history
    through the driver large systems use (FSAI with either elimination order + projection + true-residual check, rtol 1e-12)
    <= 1e-8 against the oracle's `splu` recurrence."""
    from scatter_b200 import _lib, boxmesh, system_matrix
    monkeypatch.setitem(_lib.DEFAULT_OPTIONS, "small_pcg", 0)
    monkeypatch.setitem(_lib.DEFAULT_OPTIONS, "fsai_vertex_first", vertex_first)
    s, h = 8, 0.5
    model = boxmesh.box_model(s, s, s, h, "hexa20", hexa20_order=order20)
    model.connectivities()
    ne, n = len(model.elem), model.number_eq
    E = boxmesh.lognormal_young(ne, 30e6, 1e6, seed=5); nu = np.full(ne, 0.2); rho = np.full(ne, 1500.0)
    Ko, Mo = oracle.assemble_global(oracle.model_from_readmesh(model), E, nu, rho, 2)
    Ko = sp.csr_matrix(Ko); Mo = sp.csr_matrix(Mo)
    mx = system_matrix.GenerateMatrix(n, 2)
    mx.generate_stiffness_and_mass(model, None, elem_props=(E, nu, rho))
    damping = [1, 0.01, 30, 0.01]
    mx.damping_Rayleigh(damping)
    ctx = mx.ctx
    rowptr, col = mx.pattern()
    assert np.array_equal(rowptr, Ko.indptr) and np.array_equal(col, Ko.indices)
    assert np.abs(ctx.get_values(_lib.MAT_K) - Ko.data).max() <= 1e-12 * np.abs(Ko.data).max()
    assert np.abs(ctx.get_values(_lib.MAT_M) - Mo.data).max() <= 1e-12 * np.abs(Mo.data).max()
    c0, c1 = oracle.rayleigh_coefficients(damping)
    nt = 21
    dof = int(model.eq_nb_dof[boxmesh.top_centre_node(s, s, s, model=model, h=h) - 1, 1])
    force, sched = _point_load(n, dof, nt)
    dt = 5e-4
    U, V, A, _ = oracle.newmark(Mo, Mo * c0 + Ko * c1, Ko, force, np.arange(nt) * dt, 5)
    ctx.set_load_schedule(*sched)
    ctx.set_state(None, None)
    u, v, a, st = ctx.run_newmark(dt, 0, nt - 1, 5, rtol=1e-12)
    assert np.abs(U).max() > 0 and rel_l2(u, U) <= TOL_HIST and rel_l2(v, V) <= TOL_HIST
    info = ctx.precond_info()
    assert info["fsai_nnz"] > n and st["pcg_iterations"] > 0
    ctx.close()
