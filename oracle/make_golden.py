"""TEST INFRASTRUCTURE (container-only) -- generate tests/golden/*.npz from the unmodified reference.

Run in the build container:  python oracle/make_golden.py
Produces
  tests/golden/meshes.npz     the reference's test meshes as arrays (the .msh text itself is not copied; tests
                              re-emit gmsh 2.2 ASCII from these arrays with the repo's own writer)
  tests/golden/matrices.npz   per MATRIX_CASE: equation numbering, BC codes, structural CSR pattern, K/M values
                              (full for small cases, sha256 + probes always), final K / C probes -- all computed by
                              the imported reference code (system_matrix.GenerateMatrix)
  tests/golden/histories.npz  compact copies of the reference's golden result files (u/v/a histories)
"""
import hashlib
import os
import pickle
import sys
import warnings

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_import  # noqa: E402
import cases  # noqa: E402

warnings.filterwarnings("ignore")
ref = ref_import.load_reference()
IT = "/root/reference/integration_tests"
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def probe_vector(n):
    i = np.arange(n, dtype=np.float64)
    return np.sin(0.37 * i + 0.11) + 0.25 * np.cos(1.3 * i)


# ---- meshes ---------------------------------------------------------------------------------------------------
mesh_out = {}
RUN_MESH = "/root/reference/mesh"                 # meshes of the reference's run scripts
EXTRA = ["rose_2D_side.msh", "embankment_rose2D.msh", "box2d.msh"]


def mesh_path(fn):
    return os.path.join(RUN_MESH if fn in EXTRA else os.path.join(IT, "mesh"), fn)


for fn in sorted(os.listdir(os.path.join(IT, "mesh"))) + EXTRA:
    m = ref.mesher.ReadMesh(mesh_path(fn))
    m.read_gmsh()
    key = fn[:-4]
    gm = {"tri3": 2, "tri6": 9, "quad4": 3, "hexa8": 5, "hexa20": 17, "tetra4": 4, "tetra10": 11}[m.element_type]
    mesh_out[key + "__nodes"] = np.asarray(m.nodes, dtype=np.float64)
    mesh_out[key + "__elem"] = np.asarray(m.elem, dtype=np.int64)
    mesh_out[key + "__tags"] = np.asarray(m.materials_index, dtype=np.int64)
    mesh_out[key + "__gmsh_type"] = np.int64(gm)
    mesh_out[key + "__phys"] = np.array([f"{int(p[0])}|{int(p[1])}|{p[2]}" for p in m.materials])
np.savez_compressed(os.path.join(OUT, "meshes.npz"), **mesh_out)

# ---- matrices -------------------------------------------------------------------------------------------------
mat_out = {}
MAT = cases.materials()
SETT = cases.settings()
for key, (fn, bc) in cases.MATRIX_CASES.items():
    m = ref.mesher.ReadMesh(mesh_path(fn))
    m.read_gmsh(); m.read_bc(bc); m.mapping(); m.connectivities()
    MAT = cases.case_materials(key)
    mx = ref.system_matrix.GenerateMatrix(m.number_eq, 2)
    mx.generate_stiffness_and_mass(m, MAT)
    Ks = sp.csr_matrix(mx.K); Ms = sp.csr_matrix(mx.M)
    Ks.sort_indices(); Ms.sort_indices()
    assert np.array_equal(Ks.indices, Ms.indices) and np.array_equal(Ks.indptr, Ms.indptr)
    mx.absorbing_boundaries(m, MAT, SETT["absorbing_BC"], SETT["absorbing_BC_stiff"])
    mx.damping_Rayleigh(SETT["damping"])
    Kf = sp.csr_matrix(mx.K); Cf = sp.csr_matrix(mx.C)
    x = probe_vector(m.number_eq)
    p = key + "__"
    mat_out[p + "n_eq"] = np.int64(m.number_eq)
    mat_out[p + "eq_nb_dof"] = np.nan_to_num(m.eq_nb_dof, nan=-1).astype(np.int64)
    mat_out[p + "BC"] = m.BC.astype(np.int8)
    mat_out[p + "BC_dir"] = m.BC_dir.astype(np.int8)
    mat_out[p + "pattern_sha"] = np.array(sha(Ks.indptr.astype(np.int64)) + sha(Ks.indices.astype(np.int32)))
    mat_out[p + "nnz"] = np.int64(Ks.nnz)
    mat_out[p + "Kx"] = Ks @ x
    mat_out[p + "Mx"] = Ms @ x
    mat_out[p + "Kfx"] = Kf @ x
    mat_out[p + "Cfx"] = Cf @ x
    mat_out[p + "Kmax"] = np.float64(abs(Ks.data).max())
    mat_out[p + "Mmax"] = np.float64(abs(Ms.data).max())
    mat_out[p + "Cmax"] = np.float64(abs(Cf.data).max())
    mat_out[p + "K_sha"] = np.array(sha(Ks.data))
    if Ks.nnz <= 20000:
        mat_out[p + "indptr"] = Ks.indptr.astype(np.int64)
        mat_out[p + "indices"] = Ks.indices.astype(np.int32)
        mat_out[p + "Kdata"] = Ks.data
        mat_out[p + "Mdata"] = Ms.data
    else:  # sampled entries
        rs = np.random.default_rng(7).choice(Ks.nnz, 4000, replace=False)
        rs.sort()
        mat_out[p + "sample_idx"] = rs
        mat_out[p + "Ksample"] = Ks.data[rs]
        mat_out[p + "Msample"] = Ms.data[rs]
    print(key, m.element_type, m.number_eq, Ks.nnz)
np.savez_compressed(os.path.join(OUT, "matrices.npz"), **mat_out)

# ---- element matrices (known answers from the reference's discretisation classes) -------------------------------
el_out = {}
rng = np.random.default_rng(20240607)
sys.path.insert(0, HERE)
import fem_np as orc  # noqa: E402  (only for the unit element coordinates table below)
c = orc._HEX_CORNER * 0.5
UNIT = {"hexa8": c, "quad4": np.c_[orc._QUAD_CORNER * 0.5, np.zeros(4)],
        "tri3": np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0.]]), "tetra4": np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.]])}
UNIT["hexa20"] = np.vstack([c] + [0.5 * (c[a] + c[b]) for a, b in orc._HEX20_EDGES])
q = UNIT["quad4"]
UNIT["quad8"] = np.vstack([q, 0.5 * (q[0] + q[1]), 0.5 * (q[1] + q[2]), 0.5 * (q[2] + q[3]), 0.5 * (q[3] + q[0])])
t = UNIT["tri3"]
UNIT["tri6"] = np.vstack([t, 0.5 * (t[0] + t[1]), 0.5 * (t[1] + t[2]), 0.5 * (t[0] + t[2])])
t = UNIT["tetra4"]
UNIT["tetra10"] = np.vstack([t, 0.5 * (t[0] + t[1]), 0.5 * (t[1] + t[2]), 0.5 * (t[0] + t[2]), 0.5 * (t[0] + t[3]),
                             0.5 * (t[2] + t[3]), 0.5 * (t[1] + t[3])])
for et, unit in UNIT.items():
    dim = 3 if et in ("hexa8", "hexa20", "tetra4", "tetra10") else 2
    fam_orders = [1, 2] if et.startswith("tetra") else [1, 2, 3]
    for order in fam_orders:
        xyz = unit + rng.uniform(-0.08, 0.08, unit.shape)
        if dim == 2:
            xyz[:, 2] = 0.0
        E, nu, rho = 30e6 * rng.uniform(0.5, 2), rng.uniform(0.0, 0.4), 1500 * rng.uniform(0.5, 2)
        el = (ref.discretisation.VolumeElement if dim == 3 else ref.discretisation.SurfaceElement)(et, order)
        el.generate(xyz)
        D = ref.material_models.stiffness_elasticity(E, nu, dim)
        p = f"{et}__o{order}__"
        el_out[p + "xyz"] = xyz
        el_out[p + "props"] = np.array([E, nu, rho])
        el_out[p + "Ke"] = el.compute_stiffness(D)
        el_out[p + "Me"] = el.compute_mass(rho)
        el_out[p + "N"] = np.array([np.asarray(n).ravel() for n in el.N])
        el_out[p + "dN"] = np.array(el.dN)
        el_out[p + "W"] = np.array(el.W)
np.savez_compressed(os.path.join(OUT, "elements.npz"), **el_out)

# ---- random-field bookkeeping (reference random_fields.RF, unmodified, with a stub in place of the gstools sampler) -----
import types  # noqa: E402


def stub_field(cen, mean, var):
    """Deterministic stand-in for one SRF realisation at the cell centres (shared with tests/test_host_logic.py)."""
    return mean + np.sqrt(var) * np.sin(cen @ np.array([1.3, 0.7, 0.4]) + 0.2)


class _StubModel:
    def __init__(self, dim, var, len_scale, angles):
        self.args = dict(dim=dim, var=float(var), len_scale=np.asarray(len_scale, dtype=float), angles=float(angles))


class _StubSRF:
    last = None

    def __init__(self, model, mean, seed):
        self.model, self.mean, self.seed = model, float(mean), int(seed)
        _StubSRF.last = self

    def mesh(self, mesh, points, name, seed):
        assert points == "centroids"
        (ctype, cells), = mesh.cells.items()
        cen = mesh.points[cells].mean(axis=1)
        self.cell_type = ctype
        mesh.cell_data = {name: [stub_field(cen, self.mean, self.model.args["var"])]}


class _StubMesh:
    def __init__(self, points, cells):
        self.points, self.cells, self.cell_data = np.asarray(points), cells, {}


gs = sys.modules["gstools"]
gs.SRF = _StubSRF
for _n in ("Exponential", "Gaussian", "Linear", "Matern"):
    setattr(gs, _n, type(_n, (_StubModel,), {}))
sys.modules["meshio"].Mesh = _StubMesh
import importlib  # noqa: E402
ref_rf = importlib.import_module("scatter.random_fields")
rf_out = {}
for key, model_name in (("rose_2D_side", "Gaussian"), ("cube", "Exponential")):
    fn, bc = cases.MATRIX_CASES[key]
    m = ref.mesher.ReadMesh(mesh_path(fn))
    m.read_gmsh(); m.read_bc(bc); m.mapping(); m.connectivities()
    mats = cases.case_materials(key)
    props = cases.rf_properties(key, model_name)
    out_dir = os.path.join("/tmp", "rf_" + key)
    rf = ref_rf.RF(props, mats, out_dir, m.element_type)
    idx = [mm[1] for mm in m.materials if mm[2] == props["material"]][0]
    rf.generate_gstools_rf(m.nodes, m.elem[m.materials_index == idx], m.dimension, angles=0.0)
    rf.dump()
    rf.update_material_list(mats, m, idx)
    mats.update(rf.new_material)
    srf = _StubSRF.last
    p = key + "__"
    rf_out[p + "model_args"] = np.concatenate([[srf.model.args["dim"], srf.model.args["var"], srf.model.args["angles"], srf.mean, srf.seed],
                                               srf.model.args["len_scale"]])
    rf_out[p + "model_class"] = np.array(type(srf.model).__name__)
    rf_out[p + "cell_type"] = np.array(srf.cell_type)
    rf_out[p + "field"] = np.asarray(rf.fields[0])
    rf_out[p + "materials_index"] = np.asarray(m.materials_index, dtype=np.int64)
    rf_out[p + "model_materials"] = np.array([f"{mm[0]}|{mm[1]}|{mm[2]}" for mm in m.materials])
    names = sorted(mats)
    rf_out[p + "material_names"] = np.array(names)
    rf_out[p + "material_values"] = np.array([[mats[n]["density"], mats[n]["Young"], mats[n]["poisson"]] for n in names], dtype=float)
    rf_out[p + "dump"] = np.array(open(os.path.join(out_dir, "rf_props.txt")).read())
    Er, nur, rhor = [], [], []
    dm = dict(np.array(m.materials)[:, 1:])                      # system_matrix.py:52 look-up
    for t in m.materials_index:
        nm = dm[str(t)]
        Er.append(mats[nm]["Young"]); nur.append(mats[nm]["poisson"]); rhor.append(mats[nm]["density"])
    rf_out[p + "E_elem"] = np.array(Er, dtype=float)
    rf_out[p + "rho_elem"] = np.array(rhor, dtype=float)
    print("rf", key, model_name, len(rf.fields[0]), rf_out[p + "model_args"])
np.savez_compressed(os.path.join(OUT, "random_field.npz"), **rf_out)

# ---- histories ------------------------------------------------------------------------------------------------
h = {}


def read_vtk_vectors(path, nn):
    with open(path) as f:
        L = f.read().splitlines()
    iu = L.index("VECTORS displacement double"); iv = L.index("VECTORS velocity double")
    u = np.array([[float(t) for t in l.split()] for l in L[iu + 1:iu + 1 + nn]])
    v = np.array([[float(t) for t in l.split()] for l in L[iv + 1:iv + 1 + nn]])
    return u, v


nn = 804
U = np.zeros((1001, nn, 3)); V = np.zeros_like(U)
for k in range(1001):
    U[k], V[k] = read_vtk_vectors(os.path.join(IT, f"results_mean/VTK/data_{k}.vtk"), nn)
assert abs(U[:, :, [0, 2]]).max() == 0 and abs(V[:, :, [0, 2]]).max() == 0   # 1-D problem: only y moves
# every 5th step for all nodes (y component) + the full-resolution history of 12 nodes along the column
h["hexa8_pulse__steps"] = np.arange(0, 1001, 5)
h["hexa8_pulse__uy"] = U[::5, :, 1]
h["hexa8_pulse__vy"] = V[::5, :, 1]
sel = np.array([2, 3, 6, 7, 100, 205, 300, 404, 500, 607, 700, 803])
h["hexa8_pulse__nodes_full"] = sel
h["hexa8_pulse__uy_full"] = U[:, sel, 1]
h["hexa8_pulse__vy_full"] = V[:, sel, 1]
with open(os.path.join(IT, "results_mean/VTK/data_7.vtk")) as f:
    h["hexa8_pulse__vtk_step7"] = np.array(f.read())          # one full golden VTK file: pins the output layout

with open(os.path.join(IT, "results_mean_2d/data.pickle"), "rb") as f:
    g = pickle.load(f)
ids = list(g["nodes"])
h["quad4_heaviside__nodes"] = np.array(ids)
h["quad4_heaviside__time"] = np.asarray(g["time"])
h["quad4_heaviside__position"] = np.asarray(g["position"])
for name in ("displacement", "velocity", "acceleration"):
    h["quad4_heaviside__" + name] = np.array([[g[name][str(n)][lab] for lab in "xy"] for n in ids])  # (nn, 2, nt)
with open(os.path.join(IT, "results_mean_2d/VTK/data_3.vtk")) as f:
    h["quad4_heaviside__vtk_step3"] = np.array(f.read())

for et, (nodes, nd) in cases.B2_NODES.items():
    with open(os.path.join(IT, f"test_data/column_{nd}D_{et}.pickle"), "rb") as f:
        g = pickle.load(f)
    h[et + "__time"] = np.asarray(g["time"])
    h[et + "__uy"] = np.asarray(g["displacement"]["3"]["y"])
    h[et + "__vy"] = np.asarray(g["velocity"]["3"]["y"])
    h[et + "__ay"] = np.asarray(g["acceleration"]["3"]["y"])
np.savez_compressed(os.path.join(OUT, "histories.npz"), **h)
for fn in sorted(os.listdir(OUT)):
    print(fn, os.path.getsize(os.path.join(OUT, fn)) // 1024, "KiB")
