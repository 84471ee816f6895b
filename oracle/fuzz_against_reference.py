"""TEST INFRASTRUCTURE (container-only) -- fuzz the host-side mirror of the reference interface against the unmodified
reference imported from /root/reference (stub modules for the uninstalled third-party packages, see ref_import.py).

    python oracle/fuzz_against_reference.py          # appends its report to oracle/VALIDATION_FUZZ.md

Compared bit for bit on randomly generated inputs:
  * boundary-condition planes / lines -> BC codes, BC directions, equation numbering, element equation tables
    (`scatter/mesher.py:230-326` vs `scatter_b200/mesher.py`)
  * pulse / heaviside / moving loads -> dense force vector of every time index, and the compiled device schedule
    (`scatter/force_external.py:55-345` vs `scatter_b200/force_external.py`)
  * result export -> `Write.data`, `data.pickle` for all nodes and for a node subset
    (`scatter/export_results.py:48-141` vs `scatter_b200/export_results.py`)
  * boundary faces / top surface of hexa8 meshes (`scatter/mesher.py:328-424`)
  * absorbing codes on a 2-D mesh (ignored by the reference: `nb_nodes_lower_elem == []`)
"""
import importlib
import os
import pickle
import sys
import tempfile
import types
import warnings

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)
import cases  # noqa: E402
import ref_import  # noqa: E402

warnings.filterwarnings("ignore")
ref = ref_import.load_reference()
from scatter_b200 import export_results, force_external, mesher  # noqa: E402

IT = "/root/reference/integration_tests/mesh/"
RUN = "/root/reference/mesh/"
lines = []


def log(s):
    print(s)
    lines.append(s)


def both(fn, bc):
    out = []
    for mod in (ref.mesher, mesher):
        m = mod.ReadMesh(fn)
        m.read_gmsh(); m.read_bc(bc); m.mapping(); m.connectivities()
        out.append(m)
    return out


def nan(a):
    return np.nan_to_num(np.asarray(a, dtype=float), nan=-1)


# ---- 1. boundary conditions ------------------------------------------------------------------------------------------
rng = np.random.default_rng(0)


def random_bc(dim, lo, hi, n):
    bc = {}
    for k in range(n):
        code = "".join(str(rng.integers(0, 3)) for _ in range(dim))
        ax = int(rng.integers(0, dim))
        val = rng.choice([lo[ax], hi[ax], 0.5 * (lo[ax] + hi[ax])])
        if dim == 3:
            o = [a for a in range(3) if a != ax]
            pts = []
            for u, v in ((0, 0), (1, 0), (0, 1), (1, 1)):
                p = [0, 0, 0]
                p[ax] = val; p[o[0]] = (lo, hi)[u][o[0]]; p[o[1]] = (lo, hi)[v][o[1]]
                pts.append(p)
            if rng.random() < 0.3:
                rng.shuffle(pts)
        else:
            p0, p1 = [0, 0, 0], [0, 0, 0]
            p0[ax] = p1[ax] = val; p0[1 - ax] = lo[1 - ax]; p1[1 - ax] = hi[1 - ax]
            pts = [p0, p1]
        bc[f"b{k}"] = [code, pts]
    return bc


tot = bad = 0
for fn, dim in ((IT + "cube.msh", 3), (IT + "column_2D.msh", 2), (RUN + "box2d.msh", 2), (IT + "column_3D_tetra4.msh", 3)):
    m0 = ref.mesher.ReadMesh(fn); m0.read_gmsh()
    lo, hi = m0.nodes[:, 1:].min(0), m0.nodes[:, 1:].max(0)
    for _ in range(25 if dim == 3 else 40):
        bc = random_bc(dim, lo, hi, int(rng.integers(1, 5)))
        res, errs = [], []
        for mod in (ref.mesher, mesher):
            try:
                m = mod.ReadMesh(fn); m.read_gmsh(); m.read_bc(bc); m.mapping(); m.connectivities()
                res.append(m); errs.append(None)
            except BaseException as e:      # noqa: BLE001  (the reference uses sys.exit)
                res.append(None); errs.append(repr(e))
        tot += 1
        if errs[0] or errs[1]:
            bad += (errs[0] is None) != (errs[1] is None)
            continue
        a, b = res
        bad += not (np.array_equal(a.BC, b.BC) and np.array_equal(a.BC_dir, b.BC_dir) and a.number_eq == b.number_eq
                    and np.array_equal(nan(a.eq_nb_dof), nan(b.eq_nb_dof)) and np.array_equal(nan(a.eq_nb_elem), nan(b.eq_nb_elem))
                    and np.array_equal(np.asarray(a.type_BC), np.asarray(b.type_BC)))
log(f"- boundary conditions / numbering: {tot} random BC sets on 4 meshes, {bad} differences")

# ---- 2. loads ----------------------------------------------------------------------------------------------------------
rfe = importlib.import_module("scatter.force_external")
rng = np.random.default_rng(1)
a, b = both(IT + "cube.msh", cases.BC_CUBE)
ids = a.nodes[:, 0].astype(int)
top = ids[np.isclose(a.nodes[:, 2], a.nodes[:, 2].max())]
tot = bad = 0
for _ in range(60):
    kind = rng.choice(["pulse", "heaviside", "moving"])
    nt = int(rng.integers(8, 60)); T = float(rng.uniform(0.05, 0.5)); time = np.linspace(0, T, nt)
    ini = int(rng.integers(2, 7))
    if kind == "moving":
        load = {"force": [0, -1000.0, 0], "node": int(rng.choice(top)), "time": T, "type": "moving", "speed": float(rng.uniform(1, 80)), "ini_steps": ini}
    else:
        load = {"force": list(rng.normal(size=3) * 1000), "node": [int(x) for x in rng.choice(ids, int(rng.integers(1, 5)), replace=False)],
                "time": T, "type": kind, "ini_steps": ini}
    va = vb = vc = None
    ea = eb = None
    try:
        Fa = rfe.Force(); Fa.initialise_load(dict(load), time, a, types.SimpleNamespace(), top_surface_elements=[])
        va = np.array([np.array(Fa.update_load_at_t(t)).copy() for t in range(nt)])
    except BaseException as e:              # noqa: BLE001
        ea = repr(e)
    try:
        Fb = force_external.Force(); Fb.initialise_load(dict(load), time, b, types.SimpleNamespace(), top_surface_elements=[])
        vb = np.array([Fb.update_load_at_t(t).copy() for t in range(nt)])
        ptr, dof, val = Fb.compile_schedule()
        vc = np.zeros_like(vb)
        for t in range(nt):
            vc[t, dof[ptr[t]:ptr[t + 1]]] = val[ptr[t]:ptr[t + 1]]
    except BaseException as e:              # noqa: BLE001
        eb = repr(e)
    tot += 1
    if ea or eb:
        bad += (ea is None) != (eb is None)
        continue
    bad += not (np.array_equal(va, vb) and np.array_equal(vb, vc))
log(f"- loads (pulse / heaviside / moving; dense vector per step and compiled schedule): {tot} random load cases, {bad} differences")

# ---- 3. exporter -------------------------------------------------------------------------------------------------------
rex = importlib.import_module("scatter.export_results")
rng = np.random.default_rng(2)
bad = tot = 0
for fn, bc in ((IT + "cube.msh", cases.BC_CUBE), (IT + "column_2D.msh", cases.BC_2D), (IT + "column_3D_tetra10.msh", cases.BC_B2_3D)):
    a, b = both(fn, bc)
    nt, n = 7, a.number_eq
    num = types.SimpleNamespace(u=rng.normal(size=(nt, n)), v=rng.normal(size=(nt, n)), a=rng.normal(size=(nt, n)),
                                output_time=np.linspace(0, 1, nt), time=np.linspace(0, 1, nt))
    da, db = tempfile.mkdtemp(), tempfile.mkdtemp()
    Wa, Wb = rex.Write(da, a, cases.materials(), num), export_results.Write(db, b, cases.materials(), num)
    sub = [int(x) for x in rng.choice(a.nodes[:, 0].astype(int), 5, replace=False)]
    for nodes in ("all", sub):
        Wa.pickle(write=True, nodes=nodes); Wb.pickle(write=True, nodes=nodes)
        A = pickle.load(open(os.path.join(da, "data.pickle"), "rb")); B = pickle.load(open(os.path.join(db, "data.pickle"), "rb"))
        ok = (set(A) == set(B) and list(A["nodes"]) == list(B["nodes"]) and np.array_equal(np.asarray(A["time"]), np.asarray(B["time"]))
              and np.array_equal(np.asarray(A["position"]), np.asarray(B["position"])) and type(A["position"]) is type(B["position"])
              and type(A["nodes"]) is type(B["nodes"]))
        for key in ("displacement", "velocity", "acceleration"):
            ok = ok and set(A[key]) == set(B[key])
            for nid in A[key]:
                ok = ok and set(A[key][nid]) == set(B[key][nid]) and all(np.array_equal(A[key][nid][l], B[key][nid][l]) for l in A[key][nid])
        tot += 1
        bad += not ok
log(f"- exporter (`Write.data`, data.pickle for all nodes / a subset, container types included): {tot} cases, {bad} differences")

# ---- 4. boundary faces / top surface -------------------------------------------------------------------------------------
bad = 0
for fn, bc in ((IT + "cube.msh", cases.BC_CUBE), (IT + "column.msh", cases.BC_COLUMN)):
    a, b = both(fn, bc)
    a.get_mesh_edges(); b.get_mesh_edges()
    bad += not (np.array_equal(a.boundary_elem, b.boundary_elem) and np.array_equal(a.get_top_surface(), b.get_top_surface()))
log(f"- hexa8 boundary faces and top surface (cube.msh, column.msh): {bad} differences")

# ---- 5. absorbing codes on a 2-D mesh -----------------------------------------------------------------------------------
bc = dict(cases.BC_2D); bc["bottom"] = ["02", bc["bottom"][1]]
a, _ = both(IT + "column_2D.msh", bc)
mx = ref.system_matrix.GenerateMatrix(a.number_eq, 2)
mx.generate_stiffness_and_mass(a, cases.materials())
K0 = sp.csr_matrix(mx.K).copy()
mx.absorbing_boundaries(a, cases.materials(), [1, 1], 1e3)
log(f"- absorbing codes on a 2-D mesh: the reference runs through (its 2-D exit is unreachable), |K - K0| = "
    f"{abs(sp.csr_matrix(mx.K) - K0).max():.1f}, C has {sp.csr_matrix(mx.C).nnz} entries -> ignored; scatter_b200 does the same")

# ---- 6. oracle assembly on perturbed, two-material meshes ----------------------------------------------------------------
import fem_np as orc  # noqa: E402

rng = np.random.default_rng(3)
worst = 0.0
n_cases = 0
MAT2 = {"solid": {"density": 1500, "Young": 30e6, "poisson": 0.2}, "bottom": {"density": 2100, "Young": 170e6, "poisson": 0.33}}
for fn, bc in ((IT + "cube.msh", cases.BC_CUBE_ABS), (IT + "column_3D_tetra10.msh", cases.BC_B2_3D), (IT + "column_2D_tri6.msh", cases.BC_2D),
               (IT + "column_2D.msh", cases.BC_2D), (IT + "column_high_order.msh", cases.BC_COLUMN_ABS), (RUN + "embankment_rose2D.msh", cases.MATRIX_CASES["embankment_rose2D"][1])):
    m = ref.mesher.ReadMesh(fn); m.read_gmsh()
    corner = {"hexa8": 8, "hexa20": 8, "tetra10": 4, "tetra4": 4, "tri6": 3, "tri3": 3, "quad4": 4}[m.element_type]
    # move the corner nodes a little (mid-side nodes follow their edge, boundary planes stay planes: interior nodes only)
    lo, hi = m.nodes[:, 1:].min(0), m.nodes[:, 1:].max(0)
    interior = np.all((m.nodes[:, 1:1 + m.dimension] > lo[:m.dimension] + 1e-9) & (m.nodes[:, 1:1 + m.dimension] < hi[:m.dimension] - 1e-9), axis=1)
    if m.element_type in ("hexa8", "tri3", "quad4"):
        h = np.min(np.linalg.norm(m.nodes[m.elem[:, 0] - 1, 1:] - m.nodes[m.elem[:, 1] - 1, 1:], axis=1))
        m.nodes[interior, 1:1 + m.dimension] += rng.uniform(-0.15, 0.15, (int(interior.sum()), m.dimension)) * h
    if "embankment" in fn:
        mat = cases.materials_embankment()
    else:
        mat = MAT2
        m.materials = [[float(m.dimension), 1, "solid"], [float(m.dimension), 2, "bottom"]]
        m.materials_index = rng.integers(1, 3, len(m.elem))
    m.read_bc(bc); m.mapping(); m.connectivities()
    mx = ref.system_matrix.GenerateMatrix(m.number_eq, 2)
    mx.generate_stiffness_and_mass(m, mat)
    Ks, Ms = sp.csr_matrix(mx.K), sp.csr_matrix(mx.M)
    mx.absorbing_boundaries(m, mat, [1, 1], 1e3)
    mx.damping_Rayleigh([1, 0.01, 30, 0.01])
    om = orc.Model(nodes=m.nodes, elem=m.elem, materials_index=m.materials_index, materials=m.materials, element_type=m.element_type,
                   dimension=m.dimension, BC=m.BC, BC_dir=m.BC_dir, eq_nb_dof=m.eq_nb_dof, type_BC=m.type_BC, number_eq=m.number_eq,
                   eq_nb_elem=m.eq_nb_elem, nb_nodes_elem=m.elem.shape[1])
    om.lower_element_type, om.nb_nodes_lower_elem = m.lower_element_type, m.nb_nodes_lower_elem
    ids = m.nodes[:, 0].astype(np.int64)
    om.extra["node_rows"] = np.searchsorted(ids, m.elem) if not np.array_equal(ids, np.arange(1, len(ids) + 1)) else m.elem - 1
    E, nu, rho = orc.element_properties(om, mat)
    Ko, Mo = orc.assemble_global(om, E, nu, rho, 2)
    Kf, Mf, Cf, _ = orc.system_matrices(om, mat, {"int_order": 2, "damping": [1, 0.01, 30, 0.01], "absorbing_BC": [1, 1], "absorbing_BC_stiff": 1e3})
    errs = [abs(Ko - Ks).max() / abs(Ks).max(), abs(Mo - Ms).max() / abs(Ms).max(), abs(Kf - sp.csr_matrix(mx.K)).max() / abs(Ks).max(),
            abs(Cf - sp.csr_matrix(mx.C)).max() / abs(sp.csr_matrix(mx.C)).max()]
    worst = max(worst, max(errs))
    n_cases += 1
log(f"- oracle assembly (K, M, final K and C incl. absorbing + Rayleigh) on {n_cases} perturbed / randomly two-material meshes "
    f"(hexa8, hexa20, tetra10, tri6, quad4, tri3): worst difference {worst:.1e} of the matrix max-norm")

with open(os.path.join(HERE, "VALIDATION_FUZZ.md"), "w") as f:
    f.write("# Host-logic fuzzing against the unmodified reference (generated by oracle/fuzz_against_reference.py in the build container)\n\n")
    f.write("\n".join(lines) + "\n")
