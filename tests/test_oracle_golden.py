"""The CPU oracle (oracle/fem_np.py) against the golden vectors produced by the unmodified reference
(tests/golden/*.npz, written by oracle/make_golden.py) and the reference's own golden result files."""
import hashlib

import numpy as np
import pytest

import cases
from conftest import max_rel, probe_vector, rel_l2


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_shape_functions_and_element_matrices_vs_reference(oracle, golden_elements):
    G = golden_elements
    keys = sorted({k.rsplit("__", 1)[0] for k in G.files})
    assert len(keys) == 22          # 8 element types x supported orders
    for key in keys:
        et, o = key.split("__")
        order = int(o[1:])
        N, dN, w = oracle.reference_tables(et, order)
        assert np.abs(N - G[key + "__N"]).max() <= 1e-14
        assert np.abs(dN - G[key + "__dN"]).max() <= 1e-14
        assert np.abs(w - G[key + "__W"]).max() <= 1e-15
        E, nu, rho = G[key + "__props"]
        Ke, Me = oracle.element_matrices(et, order, G[key + "__xyz"][None], E, nu, rho)
        assert max_rel(Ke[0], G[key + "__Ke"]) <= 1e-13
        assert max_rel(Me[0], G[key + "__Me"]) <= 1e-13


@pytest.mark.parametrize("case", list(cases.MATRIX_CASES))
def test_oracle_assembly_vs_reference(case, golden_meshes, golden_matrices, oracle):
    fn, bc = cases.MATRIX_CASES[case]
    G = golden_matrices
    om = oracle.build_model(golden_meshes[fn], bc)
    assert om.number_eq == int(G[case + "__n_eq"])
    assert np.array_equal(np.nan_to_num(om.eq_nb_dof, nan=-1).astype(np.int64), G[case + "__eq_nb_dof"])
    assert np.array_equal(om.BC, G[case + "__BC"]) and np.array_equal(om.BC_dir, G[case + "__BC_dir"])
    mat = cases.case_materials(case)
    E, nu, rho = oracle.element_properties(om, mat)
    K, M = oracle.assemble_global(om, E, nu, rho, 2)
    assert sha(K.indptr.astype(np.int64)) + sha(K.indices.astype(np.int32)) == str(G[case + "__pattern_sha"])
    x = probe_vector(om.number_eq)
    assert max_rel(K @ x, G[case + "__Kx"]) <= 1e-12
    assert max_rel(M @ x, G[case + "__Mx"]) <= 1e-12
    Kf, Mf, Cf, _ = oracle.system_matrices(om, mat, cases.settings())
    assert max_rel(Kf @ x, G[case + "__Kfx"]) <= 1e-12
    assert max_rel(Cf @ x, G[case + "__Cfx"]) <= 1e-12
    if (case + "__Kdata") in G.files:
        assert np.abs(K.data - G[case + "__Kdata"]).max() <= 1e-13 * float(G[case + "__Kmax"])
        assert np.abs(M.data - G[case + "__Mdata"]).max() <= 1e-13 * float(G[case + "__Mmax"])
    else:
        idx = G[case + "__sample_idx"]
        assert np.abs(K.data[idx] - G[case + "__Ksample"]).max() <= 1e-13 * float(G[case + "__Kmax"])


def _run(oracle, golden_meshes, name):
    c = cases.history_case(name)
    return oracle.run_case(golden_meshes[c["mesh"]], c["materials"], c["bc"], c["settings"], c["loading"], c["time_step"])


def test_oracle_newmark_hexa8_vs_reference_golden(oracle, golden_meshes, golden_histories):
    H = golden_histories
    model, _, (U, V, A, tt) = _run(oracle, golden_meshes, "hexa8_pulse")
    eq = model.eq_nb_dof
    free = ~np.isnan(eq[:, 1])
    uy = np.zeros((len(tt), len(eq))); vy = np.zeros_like(uy)
    uy[:, free] = U[:, eq[free, 1].astype(int)]; vy[:, free] = V[:, eq[free, 1].astype(int)]
    st = H["hexa8_pulse__steps"]
    assert rel_l2(uy[st], H["hexa8_pulse__uy"]) <= 1e-10
    assert rel_l2(vy[st], H["hexa8_pulse__vy"]) <= 1e-10
    sel = H["hexa8_pulse__nodes_full"]
    assert rel_l2(uy[:, sel], H["hexa8_pulse__uy_full"]) <= 1e-10


def test_oracle_newmark_quad4_vs_reference_golden(oracle, golden_meshes, golden_histories):
    H = golden_histories
    model, _, (U, V, A, tt) = _run(oracle, golden_meshes, "quad4_heaviside")
    ids = list(model.nodes[:, 0].astype(int))
    eq = model.eq_nb_dof
    for name, arr in (("displacement", U), ("velocity", V), ("acceleration", A)):
        gold = H["quad4_heaviside__" + name]
        mine = np.zeros_like(gold)
        for k, nid in enumerate(H["quad4_heaviside__nodes"]):
            i = ids.index(int(nid))
            for d in range(2):
                if not np.isnan(eq[i, d]):
                    mine[k, d] = arr[:, int(eq[i, d])]
        assert rel_l2(mine, gold) <= 1e-10


@pytest.mark.parametrize("etype", ["tri3", "tri6", "tetra4", "tetra10"])
def test_oracle_newmark_benchmark_set_2(etype, oracle, golden_meshes, golden_histories):
    H = golden_histories
    model, _, (U, V, A, tt) = _run(oracle, golden_meshes, etype)
    assert rel_l2(U[:, 0], H[etype + "__uy"][0::10]) <= 1e-10
    assert rel_l2(V[:, 0], H[etype + "__vy"][0::10]) <= 1e-10


def test_oracle_central_difference_converges_to_newmark(oracle, golden_meshes):
    """The central-difference scheme has no reference fixture (parity unpinned): cross-check it against the pinned
    Newmark oracle on the hexa8 column at a small time step (both are second-order accurate)."""
    c = cases.history_case("hexa8_pulse")
    # smooth load (long ramp): the two mass discretisations only agree on wave lengths the mesh resolves
    load = dict(c["loading"], time=0.03, type="heaviside", ini_steps=500)
    sett = dict(c["settings"], damping=[1, 0.0, 30, 0.0])
    dt = 2e-5
    _, _, (U1, V1, A1, t1) = oracle.run_case(golden_meshes[c["mesh"]], c["materials"], c["bc"], sett, load, dt, solver="newmark")
    _, (K, M, C), (U2, V2, A2, t2) = oracle.run_case(golden_meshes[c["mesh"]], c["materials"], c["bc"], sett, load, dt, solver="cd")
    # lumped vs consistent mass differ at O(h^2): agreement to a few percent of the peak is what the schemes allow
    assert np.abs(U1 - U2).max() <= 0.02 * np.abs(U1).max()


def _lumped_newmark_reference(oracle, golden_meshes, damping, dt, t_end, ini_steps_time=0.01):
    """Column case integrated two ways on the same lumped-mass system: Newmark with the *full* Rayleigh matrix
    C = c0 diag(m) + c1 K (what the reference builds, system_matrix.py:198, on a lumped mass) and the explicit scheme."""
    import scipy.sparse as sp
    c = cases.history_case("hexa8_pulse")
    model = oracle.build_model(golden_meshes[c["mesh"]], c["bc"])
    sett = dict(c["settings"], damping=damping)
    K, M, C, _ = oracle.system_matrices(model, c["materials"], sett)
    c0, c1 = oracle.rayleigh_coefficients(damping)
    m = oracle.lump_rows(M)
    Ml = sp.diags(m).tocsr()
    Cl = (Ml * c0 + sp.csr_matrix(K) * c1).tocsr()
    time = oracle.time_array(t_end, dt)
    load = dict(c["loading"], time=t_end, type="heaviside", ini_steps=max(int(round(ini_steps_time / dt)), 2))
    force = oracle.LoadSchedule(model, load, time)
    U1 = oracle.newmark(Ml, Cl, K, force, time)[0]
    U2 = oracle.central_difference(Ml, Cl, K, force, time, 1, c1=c1)[0]
    U3 = oracle.central_difference(Ml, Ml * c0, K, force, time, 1, c1=0.0)[0]     # stiffness-proportional part dropped
    return U1, U2, U3


def test_oracle_central_difference_keeps_stiffness_proportional_damping(oracle, golden_meshes):
    """Central difference vs Newmark on the same lumped-mass system WITH Rayleigh damping [1, 0.01, 30, 0.01] (the bench
    setting).  The lagged c1 K term makes the scheme first-order consistent in the damping force, so the gap closes with
    dt; a scheme that drops c1 K (row-sum lumping of C: rowsum(K) = 0 away from supports) keeps a dt-independent gap."""
    damping = [1, 0.05, 30, 0.05]
    errs, errs_dropped = [], []
    for dt in (4e-5, 2e-5):
        U1, U2, U3 = _lumped_newmark_reference(oracle, golden_meshes, damping, dt, 0.03)
        errs.append(rel_l2(U2, U1)); errs_dropped.append(rel_l2(U3, U1))
    assert errs[0] <= 2e-3 and errs[1] <= 0.6 * errs[0], errs            # converges (between first and second order)
    assert errs_dropped[1] >= 10 * errs[1] and errs_dropped[1] >= 0.8 * errs_dropped[0], (errs, errs_dropped)


def test_oracle_central_difference_vs_analytical_column(oracle, golden_meshes):
    """1-D wave in a finite column under a suddenly applied pressure (Churchill; the reference's
    integration_tests/analytical_solutions/analytical_wave_prop.py:50-81, only ever plotted there): top displacement
    u(L, t) = p0/K (L + 8L/pi^2 sum_k (-1)^k/(2k-1)^2 sin(lam_k L) cos(lam_k c t)), lam_k = (2k-1) pi / (2L)."""
    c = cases.history_case("hexa8_pulse")                # 0.1 x 20 x 0.1 m column, 200 hexa8 over the height, roller sides
    mat, load = c["materials"], dict(c["loading"], time=0.3, type="heaviside", ini_steps=2)
    sett = dict(c["settings"], damping=[1, 0.0, 30, 0.0], output_interval=1)
    dt = 5e-5
    model, _, (U, V, A, tt) = oracle.run_case(golden_meshes[c["mesh"]], mat, c["bc"], sett, load, dt, solver="cd")
    E, nu, rho = mat["solid"]["Young"], mat["solid"]["poisson"], mat["solid"]["density"]
    L, Kb = 20.0, E * (1 - nu) / ((1 + nu) * (1 - 2 * nu))                 # laterally confined column: constrained modulus
    p0 = -1000.0 * len(load["node"]) / (0.1 * 0.1)
    cw = np.sqrt(Kb / rho)
    k = np.arange(1, 400)[:, None]
    lam = (2 * k - 1) * np.pi / (2 * L)
    u_top = p0 / Kb * (L + 8 * L / np.pi ** 2 * ((-1.0) ** k / (2 * k - 1) ** 2 * np.sin(lam * L) * np.cos(lam * cw * tt[None, :])).sum(axis=0))
    top = int(model.eq_nb_dof[int(np.where(model.nodes[:, 0] == load["node"][0])[0][0]), 1])
    assert rel_l2(U[:, top], u_top) <= 0.01 and np.abs(U[:, top] - u_top).max() <= 0.01 * np.abs(u_top).max()
