"""The Python layer of `scatter(...)` end to end on CPU: mesh -> random field -> matrices -> solver -> loads -> integrate ->
export, for every solver of the `Solver` enum.

There is no GPU here, so `_lib.Context` (the ctypes binding of the CUDA library) is replaced by a stand-in built on the
oracle -- TEST INFRASTRUCTURE, the product never does this -- which receives exactly the calls the real context gets
(same methods, same argument layout) and checks them.  What this pins every round, without a GPU: the dict schemas, the
stage protocol (`update` / `calculate`), the load schedule handed to the device, output-row bookkeeping, pickle / VTK
files and the returned object.  The CUDA path itself is pinned by the `-m gpu` tests.
"""
import os
import pickle

import numpy as np
import pytest
import scipy.sparse as sp

import cases
from conftest import load_oracle, rel_l2


class OracleContext:
    """Stand-in for scatter_b200._lib.Context (methods used by GenerateMatrix / the solver classes)."""

    def __init__(self, device=0):
        self.oracle = load_oracle()
        self.device, self.n_eq, self.nnz, self.closed = device, 0, 0, False
        self.c0 = self.c1 = 0.0
        self.Cabs = None
        self.calls = []
        self.state_epoch = 0
        self.final_output_step = -1
        self.output_dofs = None

    def set_final_output_step(self, step):
        self.final_output_step = -1 if step is None else int(step)

    def set_output_dofs(self, dofs=None):
        assert dofs is None

    # ---- mesh / matrices
    def set_mesh(self, elem_type, xyz, conn, eq, n_eq, active=None):
        assert xyz.shape[1] == 3 and conn.dtype.kind == "i" and eq.shape == (xyz.shape[0], 2 if elem_type in ("tri3", "tri6", "quad4", "quad8") else 3)
        assert active is None
        self.elem_type, self.xyz, self.conn, self.eq, self.n_eq = elem_type, np.asarray(xyz, float), np.asarray(conn), np.asarray(eq), int(n_eq)
        self.n_elem, self.n_nodes = len(conn), len(xyz)
        self.calls.append("set_mesh")

    def set_materials(self, E, nu, rho):
        self.E, self.nu, self.rho = (np.broadcast_to(np.asarray(a, float), (self.n_elem,)).copy() for a in (E, nu, rho))
        self.calls.append("set_materials")

    def _model(self):
        o = self.oracle
        dim = self.eq.shape[1]
        nodes = np.c_[np.arange(1, self.n_nodes + 1), self.xyz]
        eqf = np.where(self.eq < 0, np.nan, self.eq.astype(float))
        m = o.Model(nodes=nodes, elem=self.conn + 1, materials_index=np.ones(self.n_elem, int), materials=[[float(dim), 1, "m"]],
                    element_type=self.elem_type, dimension=dim, BC=None, BC_dir=None, eq_nb_dof=eqf, type_BC=None, number_eq=self.n_eq,
                    eq_nb_elem=eqf[self.conn].reshape(self.n_elem, -1), nb_nodes_elem=self.conn.shape[1])
        m.extra["node_rows"] = self.conn
        return m

    def build_pattern(self):
        self.calls.append("build_pattern")
        return 0

    def assemble(self, order, flags):
        K, M = self.oracle.assemble_global(self._model(), self.E, self.nu, self.rho, order)
        self.K, self.M, self.flags = K.tocsr(), M.tocsr(), flags
        self.nnz = self.K.nnz
        self.calls.append("assemble")
        return 0.0

    def add_absorbing_faces(self, plan, order, p0, p1, stiff):
        # evaluate the plan like k_abs_faces / k_abs_reduce do
        from scatter_b200 import _lib
        N, dN, w = _lib.shape_table(plan.face_type, order)
        nf, nl = plan.i1.shape
        keep = np.array([[1, 2], [0, 2], [0, 1]])[plan.direction]
        xy = np.take_along_axis(self.xyz[plan.nodes], keep[:, None, :], axis=2)
        J = np.einsum("gad,fak->fgdk", dN, xy)
        det = J[..., 0, 0] * J[..., 1, 1] - J[..., 0, 1] * J[..., 1, 0]
        S = np.einsum("ga,gb,fg->fab", N, N, det * w[None, :])
        E, nu, rho = self.E[plan.elem], self.nu[plan.elem], self.rho[plan.elem]
        Ec = E * (1 - nu) / ((1 + nu) * (1 - 2 * nu)); G = E / (2 * (1 + nu))
        perp = plan.perp.astype(bool)
        fct = np.where(perp, (p0 * rho * np.sqrt(Ec / rho))[:, None], (p1 * rho * np.sqrt(G / rho))[:, None])
        fct2 = np.where(perp, Ec[:, None], G[:, None])
        cv, kv = (S * fct[:, None, :]).ravel(), (np.abs(S) * fct2[:, None, :]).ravel()
        grp = np.repeat(np.arange(len(plan.rows)), np.diff(plan.grp_ptr))
        cs = np.zeros(len(plan.rows)); ks = np.zeros(len(plan.rows))
        np.add.at(cs, grp, cv[plan.grp_entry]); np.add.at(ks, grp, kv[plan.grp_entry])
        n = self.n_eq
        self.Cabs = sp.csr_matrix((cs, (plan.rows, plan.cols)), shape=(n, n))
        self.K = (self.K + sp.csr_matrix((ks / stiff, (plan.rows, plan.cols)), shape=(n, n))).tocsr()
        self.calls.append("add_absorbing_faces")

    def set_rayleigh(self, c0, c1):
        self.c0, self.c1 = c0, c1

    def _C(self):
        C = self.M * self.c0 + self.K * self.c1
        return (C + self.Cabs).tocsr() if self.Cabs is not None else C.tocsr()

    # ---- loads / state / time loops
    def set_load_schedule(self, ptr, dof, val):
        ptr, dof, val = np.asarray(ptr), np.asarray(dof), np.asarray(val)
        assert ptr[0] == 0 and ptr[-1] == len(dof) == len(val) and (np.diff(ptr) >= 0).all()
        assert len(dof) == 0 or (dof.min() >= 0 and dof.max() < self.n_eq)
        self.sched = (ptr, dof, val)

    def _force(self, t):
        ptr, dof, val = self.sched
        f = np.zeros(self.n_eq)
        if 0 <= t < len(ptr) - 1:
            f[dof[ptr[t]:ptr[t + 1]]] = val[ptr[t]:ptr[t + 1]]
        return f

    def set_state(self, u=None, v=None):
        self.state = (None if u is None else np.array(u), None if v is None else np.array(v))
        self.state_epoch += 1

    def _store(self, res, outs):
        for src, dst in zip(res, outs):
            if dst is not None:
                assert dst.shape == src.shape, (dst.shape, src.shape)
                dst[...] = src

    def run_newmark(self, dt, t_start, n_steps, oi=1, beta=0.25, gamma=0.5, rtol=1e-14, maxit=20000, u_out=None, v_out=None, a_out=None, store=True):
        assert self.flags & 2                            # full mass assembled
        # a stage is the tail of the run from rest (the device continues from the state of the previous stage)
        full = self.oracle.newmark(self.M, self._C(), self.K, self._force, np.arange(t_start + n_steps + 1) * dt, 1, beta, gamma)[:3]
        u0, v0 = getattr(self, "state", (None, None))
        if t_start > 0 and u0 is not None:               # restart hook: the uploaded state is the stored row of step t_start
            assert np.allclose(u0, full[0][t_start], rtol=0, atol=1e-12 * np.abs(full[0]).max())
            assert np.allclose(v0, full[1][t_start], rtol=0, atol=1e-12 * np.abs(full[1]).max())
        self.state = (None, None)                        # consumed: a following stage without set_state continues on the device
        self.state_epoch += 1
        steps = [t for t in range(t_start, t_start + n_steps + 1) if t % oi == 0 or t == self.final_output_step]
        self._store(tuple(f[steps] for f in full), tuple(None if o is None else o[:len(steps)] for o in (u_out, v_out, a_out)))
        return u_out, v_out, a_out, {"pcg_iterations": 0}

    def run_central_difference(self, dt, t_start, n_steps, oi=1, u_out=None, v_out=None, a_out=None, store=True):
        assert t_start == 0 and self.flags & 4           # lumped mass assembled
        U, V, A, _ = self.oracle.central_difference(self.M, self._C(), self.K, self._force, np.arange(n_steps + 1) * dt, oi, c1=self.c1)
        self._store((U, V, A), (u_out, v_out, a_out))
        return u_out, v_out, a_out, {}

    def run_bathe(self, dt, t_start, n_steps, oi=1, rtol=1e-14, maxit=20000, u_out=None, v_out=None, a_out=None):
        assert t_start == 0
        U, V, A, _ = self.oracle.bathe(self.M, self._C(), self.K, self._force, np.arange(n_steps + 1) * dt, oi)
        self._store((U, V, A), (u_out, v_out, a_out))
        return u_out, v_out, a_out, {}

    def run_static(self, t_start, n_steps, oi=1, rtol=1e-12, maxit=100000, u_out=None):
        assert t_start == 0
        U, _ = self.oracle.static(self.K, self._force, np.arange(n_steps + 1, dtype=float), oi)
        self._store((U,), (u_out,))
        return u_out, {}

    def get_pattern(self):
        return self.K.indptr.astype(np.int64), self.K.indices.astype(np.int32)

    def get_values(self, which):
        return {0: self.K, 1: self.M, 2: self._C()}[which].data

    def close(self):
        self.closed = True


@pytest.fixture
def oracle_device(monkeypatch):
    from scatter_b200 import _lib, random_fields
    made = []

    def factory(device=0):
        made.append(OracleContext(device))
        return made[-1]
    monkeypatch.setattr(_lib, "Context", factory)

    def cpu_field(self, pos, lognormal=False, device=0, ctx=None):
        o = load_oracle()
        return o.srf_field(self.isometrize(pos), self.k, self.z1, self.z2, np.sqrt(self.var / self.mode_no), self.mean, lognormal)
    monkeypatch.setattr(random_fields.SpectralField, "__call__", cpu_field)
    return made


SOLVER_OF = {"NEWMARK_EXPLICIT": "newmark", "NEWMARK_IMPLICIT": "newmark", "CENTRAL_DIFFERENCE": "cd", "BATHE": "bathe", "STATIC": "static"}


@pytest.mark.parametrize("solver_name", list(SOLVER_OF))
def test_scatter_entry_point_every_solver(solver_name, oracle_device, golden_meshes, oracle, tmp_path):
    from scatter_b200 import scatter, Solver
    mesh, bc = golden_meshes["cube.msh"], cases.BC_CUBE_ABS                       # absorbing bottom and left face
    dt = 2e-4 if solver_name == "CENTRAL_DIFFERENCE" else 2e-3
    load = {"force": [0, -1000, 0], "node": [8, 9], "time": 20 * dt, "type": "heaviside"}   # ini_steps defaults to 5
    sett = cases.settings(damping=[1, 0.01, 30, 0.01], output_interval=4, VTK=True, VTK_binary=(solver_name == "BATHE"), pickle_nodes=[8, 3])
    mats = cases.materials()
    res = scatter(mesh, str(tmp_path), mats, bc, sett, load, time_step=dt, solver=getattr(Solver, solver_name))
    assert load["ini_steps"] == 5                                                  # validator default, written into the caller's dict
    ctx = oracle_device[-1]
    assert ctx.calls[:4] == ["set_mesh", "set_materials", "build_pattern", "assemble"] and "add_absorbing_faces" in ctx.calls
    # the oracle's own end-to-end run of the same case
    om = oracle.build_model(mesh, bc)
    K, M, C, _ = oracle.system_matrices(om, cases.materials(), sett)
    time = oracle.time_array(load["time"], dt)
    force = oracle.LoadSchedule(om, dict(load, ini_steps=5), time)
    kind = SOLVER_OF[solver_name]
    if kind == "static":
        U, tt = oracle.static(K, force, time, 4)
        V = A = np.zeros_like(U)
    else:
        kw = {"c1": oracle.rayleigh_coefficients(sett["damping"])[1]} if kind == "cd" else {}
        U, V, A, tt = {"newmark": oracle.newmark, "cd": oracle.central_difference, "bathe": oracle.bathe}[kind](M, C, K, force, time, 4, **kw)
    assert res.dis.shape == U.shape == (6, om.number_eq) and np.abs(U).max() > 0
    assert rel_l2(res.dis, U) <= 1e-10 and np.allclose(res.time, tt)
    if kind != "static":
        assert rel_l2(res.vel, V) <= 1e-10 and rel_l2(res.acc, A) <= 1e-10
    # files: pickle of the two requested nodes, one VTK file per stored row
    with open(os.path.join(tmp_path, "data.pickle"), "rb") as f:
        data = pickle.load(f)
    assert data["nodes"] == [8, 3] and set(data["displacement"]) == {"8", "3"} and len(data["position"]) == 2
    row = int(np.where(om.nodes[:, 0] == 8)[0][0])
    assert np.array_equal(data["displacement"]["8"]["y"], res.dis[:, int(om.eq_nb_dof[row, 1])])
    vtk = sorted(os.listdir(os.path.join(tmp_path, "VTK")))
    assert vtk == sorted(f"data_{k}.vtk" for k in range(6))
    head = open(os.path.join(tmp_path, "VTK", "data_2.vtk"), "rb").read(200).decode("latin1").splitlines()
    assert head[0] == "# vtk DataFile Version 2.0" and head[2] == ("BINARY" if solver_name == "BATHE" else "ASCII")


def test_scatter_entry_point_with_random_field_cpu(oracle_device, golden_meshes, oracle, tmp_path):
    from scatter_b200 import scatter
    case = "embankment_rose2D"
    fn, bc = cases.MATRIX_CASES[case]
    mats = cases.case_materials(case)
    sett = cases.settings(damping=[1, 0.005, 20, 0.005])
    load = {"force": [0, -1e4, 0], "node": [4], "time": 0.02, "type": "pulse", "ini_steps": 7}
    res = scatter(golden_meshes[fn], str(tmp_path), mats, bc, sett, load, time_step=1e-3, random_props=cases.rf_properties(case, "Exponential"))
    om = oracle.build_model(golden_meshes[fn], bc)
    tag = [m[1] for m in om.materials if m[2] == "soil1"][0]
    n_rf = int((np.asarray(om.materials_index) == tag).sum())
    young = np.array([mats[f"material_{k + 1}"]["Young"] for k in range(n_rf)])
    assert len([k for k in mats if k.startswith("material_")]) == n_rf and young.std() > 0
    base = cases.case_materials(case)
    E, nu, rho = oracle.element_properties(om, base)
    E[np.asarray(om.materials_index) == tag] = young
    _, _, (U, V, A, _) = oracle.run_case(golden_meshes[fn], base, bc, sett, load, 1e-3, elem_props=(E, nu, rho))
    assert np.abs(U).max() > 0 and rel_l2(res.dis, U) <= 1e-10 and rel_l2(res.vel, V) <= 1e-10
    assert os.path.isfile(os.path.join(tmp_path, "rf_props.txt")) and os.path.isfile(os.path.join(tmp_path, "data.pickle"))
    assert res.dis.shape[0] == len(oracle.time_array(0.02, 1e-3))


RF_2D_PROPS = {"number_realisations": 1, "element_size": 1, "theta": 5, "seed_number": -26021981, "material": "solid",
               "key_material": "Young", "std_value": 3e6, "aniso_x": 2 / 5, "aniso_z": 1 / 5, "model_name": "Exponential"}


def check_rf_2d_golden(res, decimal_tol=1e-8):
    """The reference's own assertion for this case (integration_test.py:423-431: every array of the result dictionary
    against results_rf_2d/data.pickle, 5 decimals), tightened to a relative L2 tolerance."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "history_rf_2d.npz"))
    data = res.data
    assert [int(n) for n in data["nodes"]] == [int(n) for n in G["nodes"]] and np.allclose(data["time"], G["time"])
    for name in ("displacement", "velocity", "acceleration"):
        mine = np.array([[data[name][str(int(n))][lab] for lab in "xy"] for n in G["nodes"]])
        assert rel_l2(mine, G[name]) <= (decimal_tol if name != "acceleration" else 100 * decimal_tol), name
        np.testing.assert_almost_equal(mine, G[name], decimal=5)


def test_reference_random_field_case_reproduces_its_golden(oracle_device, golden_meshes, tmp_path):
    """Test1DWavePropagation_2D.test_2 of the reference (integration_test.py:376-431): quad4 column whose Young's modulus is a
    gstools Exponential random field.  Pins `random_fields.gstools_modes` (the restated RandMeth seed -> modes path): the
    golden history only comes out if every element receives the modulus gstools 1.7.0 gave it."""
    from scatter_b200.scatter import scatter
    sett = cases.settings(damping=[1, 0.005, 20, 0.005])
    load = {"force": [0, -1e6, 0], "node": [3, 4, 25], "time": 1.0, "type": "heaviside"}
    res = scatter(golden_meshes["column_2D.msh"], str(tmp_path), cases.materials(), cases.BC_2D, sett, load, time_step=5e-3,
                  random_props=dict(RF_2D_PROPS))
    check_rf_2d_golden(res)


@pytest.mark.parametrize("with_update", [False, True])
def test_solver_stages_fill_the_right_output_rows(with_update, oracle_device, golden_meshes, oracle):
    """Stage protocol of scatter.py:153-159 / rose_utils.py: `calculate(t0, t1)` called stage by stage (optionally with the
    restart hook `update(t0)` in between) must fill the same rows as one run over the whole time axis."""
    from scatter_b200 import force_external, mesher, solvers, system_matrix
    mesh, bc = golden_meshes["column_2D.msh"], cases.BC_2D
    m = mesher.ReadMesh(mesh)
    m.read_gmsh(); m.read_bc(bc); m.mapping(); m.connectivities()
    mx = system_matrix.GenerateMatrix(m.number_eq, 2)
    mx.generate_stiffness_and_mass(m, cases.materials())
    mx.absorbing_boundaries(m, cases.materials(), [1, 1], 1e3)
    mx.damping_Rayleigh([1, 0.005, 20, 0.005])
    load = {"force": [0, -1e6, 0], "node": [3, 4, 25], "time": 0.135, "type": "heaviside", "ini_steps": 5}
    time = oracle.time_array(load["time"], 5e-3)                                   # 28 time indices
    num = solvers.NewmarkExplicit(); num.output_interval = 3
    num.initialise(m.number_eq, time); num.bind(mx)
    F = force_external.Force(); F.initialise_load(load, time, m, num)
    num.update_rhs_at_time_step_func = F.update_load_at_t
    num.update(0)
    for t0, t1 in ((0, 7), (7, 9), (9, 20), (20, 27)):                              # stage ends on and off the output grid
        if with_update and t0 % 3 == 0 and t0 > 0:
            num.update(t0)
        num.calculate(None, None, None, F.force_vector, t0, t1)
    om = oracle.build_model(mesh, bc)
    K, M, C, _ = oracle.system_matrices(om, cases.materials(), cases.settings(damping=[1, 0.005, 20, 0.005]))
    U, V, A, tt = oracle.newmark(M, C, K, oracle.LoadSchedule(om, load, time), time, 3)
    assert num.u.shape == U.shape == (10, m.number_eq) and np.allclose(num.output_time, tt)
    assert rel_l2(num.u, U) <= 1e-10 and rel_l2(num.v, V) <= 1e-10 and rel_l2(num.a, A) <= 1e-10


def test_last_step_is_stored_and_rerun_restarts_from_u0(oracle_device, golden_meshes, oracle):
    """(i) An output interval that does not divide the number of steps still keeps the final state (extra last row);
    (ii) calling `calculate(0, n)` a second time starts again from u0 / v0 (the reference protocol), not from the state the
    first run left on the device: the initial state is uploaded again and the histories are identical."""
    from scatter_b200 import force_external, mesher, solvers, system_matrix
    mesh, bc = golden_meshes["column_2D.msh"], cases.BC_2D
    m = mesher.ReadMesh(mesh)
    m.read_gmsh(); m.read_bc(bc); m.mapping(); m.connectivities()
    mx = system_matrix.GenerateMatrix(m.number_eq, 2)
    mx.generate_stiffness_and_mass(m, cases.materials())
    mx.absorbing_boundaries(m, cases.materials(), [1, 1], 1e3)
    mx.damping_Rayleigh([1, 0.005, 20, 0.005])
    load = {"force": [0, -1e6, 0], "node": [3, 4, 25], "time": 0.14, "type": "heaviside", "ini_steps": 5}
    time = np.linspace(0, 0.14, 29)                                                # 29 time indices, 28 steps
    num = solvers.NewmarkExplicit(); num.output_interval = 3
    num.initialise(m.number_eq, time); num.bind(mx)
    assert list(num.output_time_indices) == [0, 3, 6, 9, 12, 15, 18, 21, 24, 27, 28] and num.u.shape[0] == 11
    F = force_external.Force(); F.initialise_load(load, time, m, num)
    num.update_rhs_at_time_step_func = F.update_load_at_t
    num.update(0)
    num.calculate(None, None, None, F.force_vector, 0, 28)
    om = oracle.build_model(mesh, bc)
    K, M, C, _ = oracle.system_matrices(om, cases.materials(), cases.settings(damping=[1, 0.005, 20, 0.005]))
    U = oracle.newmark(M, C, K, oracle.LoadSchedule(om, load, time), time, 1)[0]
    assert rel_l2(num.u, U[num.output_time_indices]) <= 1e-10
    first = num.u.copy()
    ctx = mx.ctx
    e0 = ctx.state_epoch
    num.update(0)
    num.calculate(None, None, None, F.force_vector, 0, 28)
    assert ctx.state_epoch == e0 + 2                     # set_state + run: the stale device state was not reused
    assert np.array_equal(num.u, first)
    # a stage that continues exactly where the previous one ended does not upload anything
    num2 = solvers.NewmarkExplicit(); num2.output_interval = 3
    num2.initialise(m.number_eq, time); num2.bind(mx)
    num2.update_rhs_at_time_step_func = F.update_load_at_t
    num2.update(0)
    num2.calculate(None, None, None, F.force_vector, 0, 10)
    e1 = ctx.state_epoch
    num2.calculate(None, None, None, F.force_vector, 10, 28)
    assert ctx.state_epoch == e1 + 1 and rel_l2(num2.u, first) <= 1e-12


def test_output_row_count():
    from scatter_b200 import _lib
    rows = _lib.Context.n_output_rows
    assert rows(7, 2, 3) == 1 and rows(10, 1, 3) == 0 and rows(0, 27, 3) == 10 and rows(0, 0, 1) == 1 and rows(5, 0, 5) == 1
    # the last step of the time axis is always stored (sc_set_final_output_step)
    assert rows(0, 28, 3, 28) == 11 and rows(0, 27, 3, 27) == 10 and rows(9, 10, 3, 28) == 4 and rows(27, 1, 3, 28) == 2
