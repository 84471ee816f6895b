"""TEST INFRASTRUCTURE (container-only) -- pin `oracle/fem_np.py` against the imported, unmodified reference.

Run in the build container (needs /root/reference):  python oracle/validate_against_reference.py
Writes oracle/VALIDATION.md with the measured discrepancies.
"""
import os
import sys
import pickle
import time
import warnings

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402
import fem_np as orc  # noqa: E402

warnings.filterwarnings("ignore")
ref = ref_import.load_reference()
IT = "/root/reference/integration_tests"
lines = []


def log(s):
    print(s)
    lines.append(s)


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


# ---- 1. shape functions ---------------------------------------------------------------------------------------
CLS = {"hexa8": "HexEight", "hexa20": "HexTwenty", "quad4": "QuadFour", "quad8": "QuadEight", "tri3": "TriThree",
       "tri6": "TriSix", "tetra4": "TetraFour", "tetra10": "TetraTen"}
rng = np.random.default_rng(1)
log("## shape functions (max abs diff over 20 random points)")
for et, cn in CLS.items():
    dim = orc.ELEMENT_INFO[et][1]
    worst = 0.0
    for _ in range(20):
        xi = rng.uniform(-1, 1, dim)
        obj = getattr(ref.element_types, cn)()
        obj.shape_functions(list(xi))
        N, dN = orc.shape_functions(et, xi)
        worst = max(worst, np.abs(np.asarray(obj.N).ravel() - N).max(), np.abs(obj.dN - dN).max())
    log(f"- {et}: {worst:.2e}")

# ---- 2. gauss tables + element matrices -----------------------------------------------------------------------
log("## Ke / Me on randomly distorted elements (relative to max entry)")
UNIT = {
    "hexa8": orc._HEX_CORNER * 0.5, "quad4": np.c_[orc._QUAD_CORNER * 0.5, np.zeros(4)],
    "tri3": np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0.]]), "tetra4": np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.]]),
}
c = orc._HEX_CORNER * 0.5
UNIT["hexa20"] = np.vstack([c] + [0.5 * (c[a] + c[b]) for a, b in orc._HEX20_EDGES])
q = UNIT["quad4"]
UNIT["quad8"] = np.vstack([q, 0.5 * (q[0] + q[1]), 0.5 * (q[1] + q[2]), 0.5 * (q[2] + q[3]), 0.5 * (q[3] + q[0])])
t = UNIT["tri3"]
UNIT["tri6"] = np.vstack([t, 0.5 * (t[0] + t[1]), 0.5 * (t[1] + t[2]), 0.5 * (t[0] + t[2])])
t = UNIT["tetra4"]
UNIT["tetra10"] = np.vstack([t, 0.5 * (t[0] + t[1]), 0.5 * (t[1] + t[2]), 0.5 * (t[0] + t[2]), 0.5 * (t[0] + t[3]),
                             0.5 * (t[2] + t[3]), 0.5 * (t[1] + t[3])])
for et in CLS:
    nne, dim, fam, _ = orc.ELEMENT_INFO[et]
    orders = [1, 2, 3] if fam in ("quad", "tri") else [1, 2]
    for order in orders:
        worst_k = worst_m = 0.0
        for _ in range(3):
            xyz = UNIT[et] + rng.uniform(-0.08, 0.08, UNIT[et].shape)
            if dim == 2:
                xyz[:, 2] = 0.0
            E, nu, rho = 30e6 * rng.uniform(0.5, 2), rng.uniform(0.0, 0.4), 1500 * rng.uniform(0.5, 2)
            el = (ref.discretisation.VolumeElement if dim == 3 else ref.discretisation.SurfaceElement)(et, order)
            el.generate(xyz)
            D = ref.material_models.stiffness_elasticity(E, nu, dim)
            Kr, Mr = el.compute_stiffness(D), el.compute_mass(rho)
            Ko, Mo = orc.element_matrices(et, order, xyz[None], E, nu, rho)
            worst_k = max(worst_k, rel(Ko[0], Kr)); worst_m = max(worst_m, rel(Mo[0], Mr))
        log(f"- {et} order {order}: Ke {worst_k:.2e}  Me {worst_m:.2e}")

# ---- 3. mesh model + global matrices --------------------------------------------------------------------------
log("## mesher + assembled K/M/C on the reference's test meshes")
x, y, z = 0.1, 20, -0.1
BC_COL = {"bottom": ["010", [[0, 0, 0], [x, 0, 0], [0, 0, z], [x, 0, z]]],
          "left": ["100", [[0, 0, 0], [0, 0, z], [0, y, 0], [0, y, z]]],
          "right": ["100", [[x, 0, 0], [x, 0, z], [x, y, 0], [x, y, z]]],
          "front": ["001", [[0, 0, 0], [z, 0, 0], [0, y, 0], [x, y, 0]]],
          "back": ["001", [[0, 0, z], [x, 0, z], [0, y, z], [x, y, z]]]}
BC_COL_ABS = dict(BC_COL, bottom=["020", BC_COL["bottom"][1]])
x, y, z = 10, 10, -10
BC_CUBE = {"bottom": ["010", [[0, 0, 0], [x, 0, 0], [0, 0, z], [x, 0, z]]],
           "left": ["100", [[0, 0, 0], [0, 0, z], [0, y, 0], [0, y, z]]],
           "right": ["100", [[x, 0, 0], [x, 0, z], [x, y, 0], [x, y, z]]],
           "front": ["001", [[0, 0, 0], [z, 0, 0], [0, y, 0], [x, y, 0]]],
           "back": ["001", [[0, 0, z], [x, 0, z], [0, y, z], [x, y, z]]]}
BC_CUBE_ABS = dict(BC_CUBE, bottom=["020", BC_CUBE["bottom"][1]], left=["200", BC_CUBE["left"][1]])
x, y, z = 1, 10, -1
BC_B2_3D = {"bottom": ["010", [[0, 0, 0], [x, 0, 0], [0, 0, z], [x, 0, z]]],
            "left": ["100", [[0, 0, 0], [0, 0, z], [0, y, 0], [0, y, z]]],
            "right": ["100", [[x, 0, 0], [x, 0, z], [x, y, 0], [x, y, z]]],
            "front": ["001", [[0, 0, 0], [x, 0, 0], [0, y, 0], [x, y, 0]]],
            "back": ["001", [[0, 0, z], [x, 0, z], [0, y, z], [x, y, z]]]}
BC_2D = {"bottom": ["01", [[0, 0, 0], [x, 0, 0]]], "left": ["10", [[0, 0, 0], [0, y, 0]]], "right": ["10", [[x, 0, 0], [x, y, 0]]]}
MAT = {"solid": {"density": 1500, "Young": 30e6, "poisson": 0.2}, "bottom": {"density": 1200, "Young": 300e6, "poisson": 0.25}}
SETT = {"int_order": 2, "damping": [1, 0.001, 30, 0.001], "absorbing_BC": [1, 1], "absorbing_BC_stiff": 1e3}
CASES = [("column.msh", BC_COL), ("column.msh", BC_COL_ABS), ("column_high_order.msh", BC_COL),
         ("column_high_order.msh", BC_COL_ABS), ("cube.msh", BC_CUBE), ("cube.msh", BC_CUBE_ABS), ("column_2D.msh", BC_2D),
         ("column_2D_tri3.msh", BC_2D), ("column_2D_tri6.msh", BC_2D), ("column_3D_tetra4.msh", BC_B2_3D),
         ("column_3D_tetra10.msh", BC_B2_3D)]


def reference_system(mesh, bc, mat, sett):
    m = ref.mesher.ReadMesh(mesh)
    m.read_gmsh(); m.read_bc(bc); m.mapping(); m.connectivities()
    mx = ref.system_matrix.GenerateMatrix(m.number_eq, sett["int_order"])
    t0 = time.time()
    mx.generate_stiffness_and_mass(m, mat)
    t_asm = time.time() - t0
    Ks = sp.csr_matrix(mx.K); Ms = sp.csr_matrix(mx.M)   # structural (before the pruning binops)
    Ks.sort_indices(); Ms.sort_indices()
    mx.absorbing_boundaries(m, mat, sett["absorbing_BC"], sett["absorbing_BC_stiff"])
    mx.damping_Rayleigh(sett["damping"])
    return m, mx, Ks, Ms, t_asm


# meshes of the reference's run scripts (mesh/*.msh, run_scatter_rose_2D.py:22-41): multi-material tri3 / quad4
def _bc2(xx, y0, y1):
    return {"bottom": ["11", [[0, y0, 0], [xx, y0, 0]]], "left": ["10", [[0, y0, 0], [0, y1, 0]]], "right": ["10", [[xx, y0, 0], [xx, y1, 0]]]}


MAT_EMB = {"embankment": {"density": 2000, "Young": 100e6, "poisson": 0.2}, "soil1": {"density": 1700, "Young": 500e5, "poisson": 0.2},
           "soil2": {"density": 2000, "Young": 200e5, "poisson": 0.2}}
RUN_CASES = [("rose_2D_side.msh", _bc2(90, -3, 0.5), MAT_EMB), ("embankment_rose2D.msh", _bc2(10, -5, 0.5), MAT_EMB),
             ("box2d.msh", _bc2(120, 0, 1.8), MAT)]
ALL_CASES = [(os.path.join(IT, "mesh", mesh), bc, MAT) for mesh, bc in CASES] + \
            [(os.path.join("/root/reference/mesh", mesh), bc, mat) for mesh, bc, mat in RUN_CASES]
for path, bc, MAT_CASE in ALL_CASES:
    mesh = os.path.basename(path)
    m, mx, Ks, Ms, t_asm = reference_system(path, bc, MAT_CASE, SETT)
    om = orc.build_model(path, bc)
    assert om.number_eq == m.number_eq and om.element_type == m.element_type
    assert np.array_equal(np.nan_to_num(om.eq_nb_dof, nan=-1), np.nan_to_num(m.eq_nb_dof, nan=-1))
    assert np.array_equal(np.nan_to_num(om.eq_nb_elem, nan=-1), np.nan_to_num(m.eq_nb_elem, nan=-1))
    assert np.array_equal(om.BC, m.BC) and np.array_equal(om.BC_dir, m.BC_dir)
    assert np.array_equal(om.type_BC, m.type_BC)
    E, nu, rho = orc.element_properties(om, MAT_CASE)
    Ko, Mo = orc.assemble_global(om, E, nu, rho, 2)
    pat_ok = (np.array_equal(Ko.indptr, Ks.indptr) and np.array_equal(Ko.indices, Ks.indices)
              and np.array_equal(Mo.indptr, Ms.indptr) and np.array_equal(Mo.indices, Ms.indices))
    assert pat_ok, mesh
    ek, em = rel(Ko.data, Ks.data), rel(Mo.data, Ms.data)
    Kf, Mf, Cf, _ = orc.system_matrices(om, MAT_CASE, SETT)
    ekf = abs(Kf - sp.csr_matrix(mx.K)).max() / abs(mx.K).max()
    ecf = abs(Cf - sp.csr_matrix(mx.C)).max() / abs(sp.csr_matrix(mx.C)).max()
    nabs = int((om.type_BC == "Absorb").sum())
    log(f"- {mesh} ({m.element_type}, {len(m.elem)} el, {m.number_eq} eq, absorbing dofs {nabs}): numbering+BC exact, "
        f"structural pattern bit-exact (nnz {Ks.nnz}), K {ek:.2e}, M {em:.2e}, final K {ekf:.2e}, final C {ecf:.2e}; "
        f"reference assembly {len(m.elem) / t_asm:.0f} elem/s")

# ---- 4. golden histories --------------------------------------------------------------------------------------
log("## Newmark restatement vs the reference's golden result files (relative L2 over the whole history)")


def l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b))


def read_vtk_vectors(path, nn):
    with open(path) as f:
        L = f.read().splitlines()
    iu = L.index("VECTORS displacement double"); iv = L.index("VECTORS velocity double")
    u = np.array([[float(t) for t in l.split()] for l in L[iu + 1:iu + 1 + nn]])
    v = np.array([[float(t) for t in l.split()] for l in L[iv + 1:iv + 1 + nn]])
    return u, v


# hexa8 column, pulse (integration_test.py:67-102) vs results_mean/VTK
load = {"force": [0, -1000, 0], "node": [3, 4, 7, 8], "time": 0.5, "type": "pulse"}
model, mats, (U, V, A, tt) = orc.run_case(os.path.join(IT, "mesh/column.msh"), MAT, BC_COL, SETT, load, 0.5e-3)
eq = model.eq_nb_dof
free = ~np.isnan(eq)
gu = np.zeros((len(tt),) + eq.shape); gv = np.zeros_like(gu)
for k in range(len(tt)):
    gu[k], gv[k] = read_vtk_vectors(os.path.join(IT, f"results_mean/VTK/data_{k}.vtk"), len(eq))
ou = np.zeros_like(gu); ov = np.zeros_like(gv)
ou[:, free] = U[:, eq[free].astype(int)]; ov[:, free] = V[:, eq[free].astype(int)]
log(f"- hexa8 column pulse (1001 steps x 804 nodes, VTK goldens): disp {l2(ou, gu):.2e} vel {l2(ov, gv):.2e}")

# quad4 column heaviside (integration_test.py:336-369) vs results_mean_2d/data.pickle
sett2 = dict(SETT, damping=[1, 0.005, 20, 0.005])
load = {"force": [0, -1e6, 0], "node": [3, 4, 25], "time": 1.0, "type": "heaviside"}
model, mats, (U, V, A, tt) = orc.run_case(os.path.join(IT, "mesh/column_2D.msh"), MAT, BC_2D, sett2, load, 5e-3)
with open(os.path.join(IT, "results_mean_2d/data.pickle"), "rb") as f:
    gold = pickle.load(f)
eq = model.eq_nb_dof
errs = []
for name, arr in (("displacement", U), ("velocity", V), ("acceleration", A)):
    g = np.zeros((len(tt),) + eq.shape); o = np.zeros_like(g)
    for i, nid in enumerate(model.nodes[:, 0].astype(int)):
        for d, lab in enumerate("xy"):
            g[:, i, d] = gold[name][str(nid)][lab]
            if not np.isnan(eq[i, d]):
                o[:, i, d] = arr[:, int(eq[i, d])]
    errs.append(l2(o, g))
log(f"- quad4 column heaviside (201 steps x 63 nodes): disp {errs[0]:.2e} vel {errs[1]:.2e} acc {errs[2]:.2e}")

# benchmark set 2 (test_benchmark_set_2.py:32-94)
B2 = [("tri3", [3, 4, 25], 2), ("tri6", [3, 4, 47, 48, 49], 2), ("tetra4", [3, 4, 7, 8, 29, 69, 91, 92, 132], 3),
      ("tetra10", [3, 4, 7, 8, 51, 52, 53, 135, 136, 137, 183, 184, 185, 186, 187, 188, 432, 433, 434, 435, 436, 437,
                   438, 439, 440], 3)]
mat0 = {"solid": {"density": 1500, "Young": 30e6, "poisson": 0.0}}
for et, nodes, nd in B2:
    load = {"force": [0, 1000 / len(nodes), 0], "node": nodes, "time": 1, "type": "heaviside"}
    settb = dict(SETT, output_interval=10)
    model, mats, (U, V, A, tt) = orc.run_case(os.path.join(IT, f"mesh/column_{nd}D_{et}.msh"), mat0,
                                              BC_2D if nd == 2 else BC_B2_3D, settb, load, 5e-4)
    with open(os.path.join(IT, f"test_data/column_{nd}D_{et}.pickle"), "rb") as f:
        gold = pickle.load(f)
    gu = np.asarray(gold["displacement"]["3"]["y"])[0::10]; gv = np.asarray(gold["velocity"]["3"]["y"])[0::10]
    log(f"- {et} column heaviside (2001 steps, output_interval 10, top node): disp {l2(U[:, 0], gu):.2e} vel {l2(V[:, 0], gv):.2e}")

with open(os.path.join(HERE, "VALIDATION.md"), "w") as f:
    f.write("# Oracle validation (generated by oracle/validate_against_reference.py in the build container)\n\n"
            "Reference = unmodified /root/reference (PlatypusBytes/scatter) imported with stub modules for the\n"
            "uninstalled third-party packages; numpy %s.\n\n" % np.__version__)
    f.write("\n".join(lines) + "\n")
