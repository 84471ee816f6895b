"""Host-side logic of the product (no GPU): mesh reader, numbering, absorbing faces, loads, exporter, C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import cases
from conftest import ROOT, rel_l2


def test_library_exports_every_declared_symbol():
    from scatter_b200 import _lib
    header = open(os.path.join(ROOT, "include", "scatter_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(sc_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/scatter_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared
    assert _lib.load_library().sc_version() >= 100


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from scatter_b200 import _lib
    with pytest.raises(_lib.ScatterB200Error, match="no usable CUDA device"):
        _lib.Context(0)


def test_shape_tables_match_reference(golden_elements):
    from scatter_b200 import _lib
    G = golden_elements
    for key in sorted({k.rsplit("__", 1)[0] for k in G.files}):
        et, o = key.split("__")
        N, dN, w = _lib.shape_table(et, int(o[1:]))
        assert np.abs(N - G[key + "__N"]).max() <= 1e-14
        assert np.abs(dN - G[key + "__dN"]).max() <= 1e-14
        assert np.abs(w - G[key + "__W"]).max() <= 1e-15
    with pytest.raises(_lib.ScatterB200Error, match="integration order not supported"):
        _lib.shape_table("tetra4", 3)          # discretisation.py:491-492


@pytest.mark.parametrize("case", list(cases.MATRIX_CASES))
def test_mesher_numbering_is_bit_exact(case, golden_meshes, golden_matrices, oracle):
    from scatter_b200 import mesher
    fn, bc = cases.MATRIX_CASES[case]
    G = golden_matrices
    m = mesher.ReadMesh(golden_meshes[fn])
    m.read_gmsh(); m.read_bc(bc); m.mapping(); m.connectivities(); m.get_mesh_edges()
    assert m.number_eq == int(G[case + "__n_eq"])
    assert np.array_equal(m.equation_table_int(), G[case + "__eq_nb_dof"])
    assert np.array_equal(m.BC, G[case + "__BC"]) and np.array_equal(m.BC_dir, G[case + "__BC_dir"])
    om = oracle.build_model(golden_meshes[fn], bc)
    assert np.array_equal(np.nan_to_num(m.eq_nb_elem, nan=-1), np.nan_to_num(om.eq_nb_elem, nan=-1))
    assert np.array_equal(m.type_BC, om.type_BC)
    assert m.element_type == om.element_type and m.dimension == om.dimension
    assert m.lower_element_type == om.lower_element_type and m.nb_nodes_lower_elem == om.nb_nodes_lower_elem


def test_reader_skips_lower_dimensional_elements(golden_meshes, tmp_path):
    """The run-script meshes (mesh/rose_2D_side.msh, box2d.msh) list 1-D "rose" line elements before the domain elements;
    the reader keeps the elements of the highest dimension only (mesher.py:150-228)."""
    from scatter_b200 import mesher
    src = open(golden_meshes["box2d.msh"]).read().splitlines()
    i0 = src.index("$Elements")
    ne = int(src[i0 + 1])
    lines = [f"{k + 1} 1 2 2 7 {k + 1} {k + 2}" for k in range(5)]                 # 5 two-node lines, physical group 2 ("rose")
    body = []
    for k, l in enumerate(src[i0 + 2:i0 + 2 + ne]):
        t = l.split()
        body.append(" ".join([str(k + 6)] + t[1:]))
    out = src[:i0 + 1] + [str(ne + 5)] + lines + body + src[i0 + 2 + ne:]
    p = os.path.join(tmp_path, "mixed.msh")
    open(p, "w").write("\n".join(out) + "\n")
    a = mesher.ReadMesh(golden_meshes["box2d.msh"]); a.read_gmsh()
    b = mesher.ReadMesh(p); b.read_gmsh()
    assert b.element_type == "quad4" and np.array_equal(a.elem, b.elem) and np.array_equal(a.materials_index, b.materials_index)
    assert np.array_equal(a.nodes, b.nodes)


def test_reader_rejects_records_it_cannot_decode(golden_meshes, tmp_path):
    from scatter_b200 import gmsh_io
    src = open(golden_meshes["column_2D.msh"]).read().splitlines()
    i0 = src.index("$Elements")
    bad = list(src)
    t = bad[i0 + 2].split()
    bad[i0 + 2] = " ".join(t[:2] + ["3"] + t[3:5] + ["9"] + t[5:])           # three tags instead of two
    p = os.path.join(tmp_path, "tags.msh")
    open(p, "w").write("\n".join(bad) + "\n")
    with pytest.raises(SystemExit, match="exactly 2 tags"):
        gmsh_io.read_msh(p)
    bad = list(src)
    t = bad[i0 + 2].split()
    bad[i0 + 2] = " ".join([t[0], "93"] + t[2:])                              # unknown element type
    p = os.path.join(tmp_path, "type.msh")
    open(p, "w").write("\n".join(bad) + "\n")
    with pytest.raises(SystemExit, match="not supported"):
        gmsh_io.read_msh(p)


def test_mesher_errors(tmp_path):
    from scatter_b200 import mesher
    with pytest.raises(SystemExit, match="Mesh file does not exit"):
        mesher.ReadMesh(os.path.join(tmp_path, "missing.msh"))
    p = os.path.join(tmp_path, "a.txt")
    open(p, "w").write("x")
    with pytest.raises(SystemExit, match="not a valid file"):
        mesher.ReadMesh(p)


def test_cube_boundary_faces_and_top_surface(golden_meshes):
    from scatter_b200 import mesher
    m = mesher.ReadMesh(golden_meshes["cube.msh"])
    m.read_gmsh(); m.read_bc(cases.BC_CUBE); m.mapping(); m.connectivities(); m.get_mesh_edges()
    assert m.boundary_elem.shape == (600, 4)           # 6 faces x 10 x 10
    top = m.get_top_surface()
    ys = m.nodes[top - 1, 2]
    assert len(top) > 0 and np.allclose(ys, ys.max())


@pytest.mark.parametrize("case", ["column_abs", "column_high_order_abs", "cube_abs"])
def test_absorbing_entries_match_oracle(case, golden_meshes, oracle):
    from scatter_b200 import mesher, system_matrix
    fn, bc = cases.MATRIX_CASES[case]
    m = mesher.ReadMesh(golden_meshes[fn])
    m.read_gmsh(); m.read_bc(bc); m.mapping(); m.connectivities()
    E, nu, rho = system_matrix.resolve_element_properties(m, cases.materials())
    cd, kd = system_matrix.absorbing_entries(m, E, nu, rho, 2, [1, 1], 1e3)
    om = oracle.build_model(golden_meshes[fn], bc)
    Co, Ko = oracle.absorbing_matrices(om, E, nu, rho, 2, [1, 1], 1e3)
    k = np.array(list(cd.keys()))
    n = m.number_eq
    C = sp.csr_matrix((list(cd.values()), (k[:, 0], k[:, 1])), shape=(n, n))
    K = sp.csr_matrix((list(kd.values()), (k[:, 0], k[:, 1])), shape=(n, n))
    assert abs(C - Co).max() <= 1e-13 * abs(Co).max()
    assert abs(K - Ko).max() <= 1e-13 * abs(Ko).max()


def test_rayleigh_coefficients_closed_form():
    """unit_test/test_gen_matrix.py:13-60 of the reference: c0, c1 against the closed form."""
    from scatter_b200 import system_matrix
    f1, d1, f2, d2 = 1.0, 0.01, 30.0, 0.02
    c0, c1 = system_matrix.rayleigh_coefficients([f1, d1, f2, d2])
    w1, w2 = 2 * np.pi * f1, 2 * np.pi * f2
    a0 = 2 * w1 * w2 * (d1 * w2 - d2 * w1) / (w2 ** 2 - w1 ** 2)
    a1 = 2 * (d2 * w2 - d1 * w1) / (w2 ** 2 - w1 ** 2)
    np.testing.assert_array_almost_equal([c0, c1], [a0, a1], decimal=12)
    with pytest.raises(SystemExit, match="Frequencies for the Rayleigh damping are the same"):
        system_matrix.rayleigh_coefficients([1, 0.1, 1, 0.1])


@pytest.mark.parametrize("kind", ["pulse", "heaviside", "moving"])
def test_load_schedule_matches_oracle(kind, golden_meshes, oracle):
    from scatter_b200 import force_external, mesher
    m = mesher.ReadMesh(golden_meshes["cube.msh"])
    m.read_gmsh(); m.read_bc(cases.BC_CUBE); m.mapping(); m.connectivities()
    om = oracle.build_model(golden_meshes["cube.msh"], cases.BC_CUBE)
    time = np.linspace(0, 1, 201)
    load = {"force": [10.0, -1000.0, 3.0], "node": [8] if kind == "moving" else [3, 4, 8, 700], "time": 1, "type": kind,
            "speed": 10, "ini_steps": 50 if kind == "moving" else 7}
    if kind == "moving":
        load["node"] = 8
    F = force_external.Force()
    F.initialise_load(load, time, m, None)
    ref = oracle.LoadSchedule(om, load, time)
    ptr, dof, val = F.compile_schedule()
    assert len(ptr) == len(time) + 1
    for t in (0, 1, 3, 5, 6, 7, 49, 50, 51, 100, 150, 200):
        dense = np.zeros(m.number_eq)
        dense[dof[ptr[t]:ptr[t + 1]]] = val[ptr[t]:ptr[t + 1]]
        expect = ref(t)
        assert np.array_equal(dense, expect), (kind, t)
        assert np.array_equal(F.update_load_at_t(t), expect)


def test_moving_at_plane_load_matches_oracle(golden_meshes, oracle):
    """integration_test.py:551-602 (moving load on the top surface of cube.msh)."""
    from scatter_b200 import force_external, mesher
    m = mesher.ReadMesh(golden_meshes["cube.msh"])
    m.read_gmsh(); m.read_bc(cases.BC_CUBE); m.mapping(); m.connectivities(); m.get_mesh_edges()
    om = oracle.build_model(golden_meshes["cube.msh"], cases.BC_CUBE)
    faces = oracle.boundary_faces_hexa8(om)
    assert np.array_equal(faces, m.boundary_elem)
    top = oracle.top_surface_faces(om, faces)
    assert np.array_equal(top, m.get_top_surface())
    time = np.linspace(0, 1, 201)
    load = {"force": [0, -1000, 0], "start_coord": [0.5, -0.5], "time": 1, "type": "moving_at_plane", "direction": [0.5, -1],
            "speed": 10, "ini_steps": 50}
    F = force_external.Force()
    F.initialise_load(load, time, m, None, top_surface_elements=m.get_top_surface())
    ref = oracle.MovingAtPlaneLoad(om, load, time, top)
    ptr, dof, val = F.compile_schedule()
    for t in (0, 10, 49, 50, 51, 77, 120, 200):
        dense = np.zeros(m.number_eq)
        dense[dof[ptr[t]:ptr[t + 1]]] = val[ptr[t]:ptr[t + 1]]
        assert np.allclose(dense, ref(t), rtol=1e-13, atol=1e-13), t
        assert abs(dense.sum() - (-1000.0 * ref.sf[t])) < 1e-9          # the nodal forces add up to the point load


def _stub_field(cen, mean, var):
    # the stand-in sampler oracle/make_golden.py gave the unmodified reference RF class
    return mean + np.sqrt(var) * np.sin(cen @ np.array([1.3, 0.7, 0.4]) + 0.2)


@pytest.mark.parametrize("case,model_name", [("rose_2D_side", "Gaussian"), ("cube", "Exponential")])
def test_random_field_bookkeeping_matches_reference(case, model_name, golden_meshes, tmp_path, monkeypatch):
    """Everything around the sampler -- length scales, lognormal parameters, centroids, one material per element, the tag
    re-indexing, rf_props.txt, the per-element look-up of system_matrix.py:52-71 -- against the reference's RF class run
    with the same stand-in sampler (tests/golden/random_field.npz)."""
    from scatter_b200 import mesher, random_fields, system_matrix
    G = np.load(os.path.join(ROOT, "tests", "golden", "random_field.npz"))
    fn, bc = cases.MATRIX_CASES[case]
    m = mesher.ReadMesh(golden_meshes[fn])
    m.read_gmsh(); m.read_bc(bc); m.mapping(); m.connectivities()
    mats = cases.case_materials(case)
    props = cases.rf_properties(case, model_name)
    seen = {}

    def fake_call(self, pos, lognormal=False, device=0, ctx=None):
        seen.update(dim=self.dim, var=self.var, mean=self.mean, angles=self.angles, seed=self.seed, len_scale=self.len_scale,
                    model=self.model_name)
        f = _stub_field(np.asarray(pos), self.mean, self.var)
        return np.exp(f) if lognormal else f

    monkeypatch.setattr(random_fields.SpectralField, "__call__", fake_call)
    rf = random_fields.RF(props, mats, str(tmp_path), m.element_type)
    idx = [mm[1] for mm in m.materials if mm[2] == props["material"]][0]
    rf.generate_gstools_rf(m.nodes, m.elem[m.materials_index == idx], m.dimension, angles=0.0)
    E_arr, _, rho_arr = rf.element_properties(m, idx)            # array path, before the tags are re-indexed
    rf.dump()
    rf.update_material_list(mats, m, idx)
    mats.update(rf.new_material)
    p = case + "__"
    a = G[p + "model_args"]
    assert seen["dim"] == int(a[0]) and seen["angles"] == a[2] and seen["seed"] == int(a[4]) and seen["model"] == str(G[p + "model_class"])
    assert abs(seen["var"] - a[1]) <= 1e-15 * a[1] and abs(seen["mean"] - a[3]) <= 1e-15 * abs(a[3])
    assert np.array_equal(seen["len_scale"], a[5:5 + seen["dim"]])
    assert rf.element_type_to_meshio_element_type() == str(G[p + "cell_type"])
    assert np.abs(rf.fields[0] - G[p + "field"]).max() <= 1e-13 * np.abs(G[p + "field"]).max()
    assert np.array_equal(np.asarray(m.materials_index), G[p + "materials_index"])
    assert [f"{mm[0]}|{mm[1]}|{mm[2]}" for mm in m.materials] == list(G[p + "model_materials"])
    names = sorted(mats)
    assert names == list(G[p + "material_names"])
    vals = np.array([[mats[n]["density"], mats[n]["Young"], mats[n]["poisson"]] for n in names], dtype=float)
    assert np.abs(vals - G[p + "material_values"]).max() <= 1e-13 * np.abs(vals).max()
    assert open(os.path.join(tmp_path, "rf_props.txt")).read() == str(G[p + "dump"])
    E, nu, rho = system_matrix.resolve_element_properties(m, mats)
    assert np.abs(E - G[p + "E_elem"]).max() <= 1e-13 * E.max() and np.array_equal(rho, G[p + "rho_elem"])
    assert np.array_equal(E_arr, E) and np.array_equal(rho_arr, rho)


@pytest.mark.parametrize("model_name", ["Gaussian", "Exponential", "Matern"])
@pytest.mark.parametrize("dim", [2, 3])
def test_mode_sampler_has_the_model_spectrum(model_name, dim):
    """Bochner: E[cos(k . h)] over the sampled wave vectors is the correlation function of the covariance model."""
    from scatter_b200 import random_fields
    k, z1, z2 = random_fields.sample_modes(model_name, dim, seed=11, mode_no=400000)
    assert abs(z1.mean()) < 0.01 and abs(z1.std() - 1) < 0.01 and abs(np.mean(z1 * z2)) < 0.01
    if dim == 2:
        assert np.all(k[:, 2] == 0)
    for dist in (0.25, 0.7, 1.5, 3.0):
        for direction in ([1, 0, 0], [0.6, 0.8, 0], [0, 0, 1] if dim == 3 else [0, 1, 0]):
            h = dist * np.array(direction, dtype=float)
            assert abs(np.cos(k @ h).mean() - random_fields.correlation(model_name, dist)) < 6e-3
    with pytest.raises(NotImplementedError):
        random_fields.sample_modes("Linear", dim, 1)


def test_oracle_random_field_statistics(oracle):
    """The restated randomisation method reproduces mean, variance and the correlation at one lag (spatial averages over
    a domain of many correlation lengths)."""
    from scatter_b200 import random_fields
    sf = random_fields.SpectralField("Gaussian", 2, var=0.04, mean=1.5, len_scale=[2.0, 0.5, 1.0], angles=0.0, seed=5)
    gx, gy = np.meshgrid(np.arange(0, 400, 1.0), np.arange(0, 100, 0.25))
    pos = np.c_[gx.ravel(), gy.ravel(), np.zeros(gx.size)]
    f = oracle.srf_field(sf.isometrize(pos), sf.k, sf.z1, sf.z2, np.sqrt(sf.var / sf.mode_no), sf.mean, False).reshape(gx.shape)
    assert abs(f.mean() - 1.5) < 0.02 and abs(f.var() - 0.04) < 0.008
    c = ((f[:, 2:] - f.mean()) * (f[:, :-2] - f.mean())).mean() / f.var()           # lag 2 along x = one length scale
    assert abs(c - random_fields.correlation("Gaussian", 1.0)) < 0.08
    c = ((f[4:, :] - f.mean()) * (f[:-4, :] - f.mean())).mean() / f.var()           # lag 1 along y = two length scales
    assert abs(c - random_fields.correlation("Gaussian", 2.0)) < 0.08


def test_validator():
    from scatter_b200 import validator
    load = {"type": "pulse"}
    validator.ValidateLoad.validate(load)
    assert load["ini_steps"] == 5
    with pytest.raises(Exception, match="not supported"):
        validator.ValidateLoad.validate({"type": "earthquake"})


def test_exporter_reproduces_reference_vtk_and_pickle_layout(golden_meshes, golden_histories, oracle, tmp_path):
    """Feed oracle results through the product's exporter and compare with the reference's golden VTK file."""
    from scatter_b200 import export_results, mesher
    H = golden_histories
    c = cases.history_case("hexa8_pulse")
    load = dict(c["loading"], time=0.005)       # 11 steps are enough for file data_7
    model, _, (U, V, A, tt) = oracle.run_case(golden_meshes[c["mesh"]], c["materials"], c["bc"], c["settings"], load, c["time_step"])
    m = mesher.ReadMesh(golden_meshes[c["mesh"]])
    m.read_gmsh(); m.read_bc(c["bc"]); m.mapping(); m.connectivities()

    class Num:
        pass
    num = Num(); num.u, num.v, num.a, num.output_time = U, V, A, tt
    w = export_results.Write(os.path.join(tmp_path, "out"), m, c["materials"], num)
    w.pickle(write=True, nodes="all")
    w.vtk(write=True, binary=False)
    mine = open(os.path.join(tmp_path, "out", "VTK", "data_7.vtk")).read().splitlines()
    gold = str(H["hexa8_pulse__vtk_step7"]).splitlines()
    assert len(mine) == len(gold)
    for a, b in zip(mine, gold):
        tb = b.split()
        try:
            fb = [float(t) for t in tb]
        except ValueError:
            assert a == b, (a, b)
            continue
        np.testing.assert_almost_equal([float(t) for t in a.split()], fb, decimal=9)
    import pickle
    data = pickle.load(open(os.path.join(tmp_path, "out", "data.pickle"), "rb"))
    assert set(data) == {"time", "nodes", "position", "displacement", "velocity", "acceleration"}
    assert set(data["displacement"]["3"]) == {"x", "y", "z"}
    assert data["displacement"]["1"]["y"].shape == (len(tt),)
    # binary VTK: header + big-endian payload sizes
    w.vtk(name="bin", write=True, binary=True)
    raw = open(os.path.join(tmp_path, "out", "VTK", "bin_0.vtk"), "rb").read()
    assert raw.startswith(b"# vtk DataFile Version 2.0\nbin_0\nBINARY\nDATASET UNSTRUCTURED_GRID\nPOINTS 804 float\n")
    # subset pickle (test_benchmark_set_2.py: pickle_nodes=[3])
    w.pickle(name="sub", write=True, nodes=[3])
    sub = pickle.load(open(os.path.join(tmp_path, "out", "sub.pickle"), "rb"))
    assert sub["nodes"] == [3] and list(sub["displacement"].keys()) == ["3"]


def test_box_generator_matches_file_reader(tmp_path, oracle):
    from scatter_b200 import boxmesh, mesher
    for et in ("hexa8", "hexa20"):
        path = os.path.join(tmp_path, f"b_{et}.msh")
        boxmesh.write_box_msh(path, 3, 4, 2, 0.5, et)
        bc = boxmesh.box_boundaries(3, 4, 2, 0.5)
        a = boxmesh.box_model(3, 4, 2, 0.5, et); a.connectivities()
        b = mesher.ReadMesh(path); b.read_gmsh(); b.read_bc(bc); b.mapping(); b.connectivities()
        assert np.array_equal(a.nodes, b.nodes) and np.array_equal(a.elem, b.elem)
        assert np.array_equal(np.nan_to_num(a.eq_nb_elem, nan=-1), np.nan_to_num(b.eq_nb_elem, nan=-1))
        # every element is a positively oriented cube of volume h^3
        Ke, Me = oracle.element_matrices(et, 2, a.nodes[:, 1:][a.node_rows()], 1.0, 0.2, 1.0)
        assert np.allclose(Me[:, 0::3, 0::3].sum(axis=(1, 2)), 0.125)
    # slabs of a box are sub-boxes with shifted z
    n1, e1 = boxmesh.box_arrays(3, 4, 6, 0.5, "hexa8", z_range=(2, 5))
    n2, e2 = boxmesh.box_arrays(3, 4, 3, 0.5, "hexa8")
    assert np.array_equal(e1, e2) and np.allclose(n1[:, 3], n2[:, 3] + 1.0) and np.array_equal(n1[:, 1:3], n2[:, 1:3])


def test_interleaved_hexa20_numbering_is_the_same_mesh(oracle):
    """`hexa20_order="interleaved"` only renumbers the nodes: every element keeps its 20 coordinates in gmsh order, the
    equation count is unchanged, and the interior of the numbering is translation invariant (few distinct relative column
    lists, small bandwidth) -- what the column dictionary of the node-blocked SpMV needs."""
    from scatter_b200 import boxmesh
    s = 8
    a = boxmesh.box_model(s, s, s, 0.5, "hexa20"); a.connectivities()
    b = boxmesh.box_model(s, s, s, 0.5, "hexa20", hexa20_order="interleaved"); b.connectivities()
    assert a.number_eq == b.number_eq and len(a.nodes) == len(b.nodes)
    assert np.array_equal(b.nodes[:, 0], np.arange(1, len(b.nodes) + 1))
    assert np.allclose(a.nodes[:, 1:][a.node_rows()], b.nodes[:, 1:][b.node_rows()])
    ta, tb = boxmesh.top_centre_node(s, s, s), boxmesh.top_centre_node(s, s, s, model=b)
    assert np.allclose(a.nodes[ta - 1, 1:], b.nodes[tb - 1, 1:])
    assert boxmesh.top_centre_node(s, s, s, model=a) == ta

    def stats(m):
        eq_elem = np.nan_to_num(m.eq_nb_dof, nan=-1).astype(int)[np.asarray(m.node_rows())].reshape(len(m.elem), -1)
        P = oracle.structural_pattern(eq_elem, m.number_eq)
        bw = int(np.abs(P.indices - np.repeat(np.arange(P.shape[0]), np.diff(P.indptr))).max())
        pats = {}
        for row in m.eq_nb_dof:
            rows = row[~np.isnan(row)].astype(int)
            if len(rows):
                key = tuple(P.indices[P.indptr[rows[0]]:P.indptr[rows[0] + 1]] - rows[0])
                pats[key] = pats.get(key, 0) + 1
        return bw, sorted(pats.values(), reverse=True)
    bw_a, top_a = stats(a)
    bw_b, top_b = stats(b)
    assert bw_b * 4 < bw_a
    assert sum(top_b[:4]) > 10 * sum(top_a[:4])          # the four interior node types dominate
    with pytest.raises(ValueError):
        boxmesh.box_arrays(2, 2, 2, 0.5, "hexa20", hexa20_order="random")


def test_exporter_at_a_million_nodes(tmp_path):
    """the result dictionary, a node-subset pickle and one binary VTK frame of a 10^6-node mesh within a time
    budget (the reference's per-node Python loops, export_results.py:68-103,143-213, take minutes there), with the
    reference's key layout and values."""
    import pickle
    import time
    import types
    from scatter_b200 import boxmesh, export_results
    s = 100
    m = boxmesh.box_model(s, s, s, 0.5, "hexa8")
    n, nn = m.number_eq, len(m.nodes)
    assert nn == 101 ** 3
    rng = np.random.default_rng(0)
    num = types.SimpleNamespace(u=rng.standard_normal((3, n)), v=rng.standard_normal((3, n)), a=rng.standard_normal((3, n)),
                                output_time=np.array([0.0, 0.1, 0.2]))
    t0 = time.perf_counter()
    w = export_results.Write(str(tmp_path), m, {"solid": {"density": 1500.0, "Young": 30e6, "poisson": 0.2}}, num)
    top = boxmesh.top_centre_node(s, s, s)
    w.pickle(nodes=[top, 5, nn])
    d = w.data
    eq = m.eq_nb_dof
    got = d["velocity"][str(top)]["y"]
    assert np.array_equal(got, num.v[:, int(eq[top - 1, 1])]) and len(d["displacement"]) == nn and str(nn) in d["acceleration"]
    assert not d["displacement"]["1"]["y"].any()                     # bottom node: fixed dof -> zeros
    t_subset = time.perf_counter() - t0
    with open(os.path.join(tmp_path, "data.pickle"), "rb") as f:
        p = pickle.load(f)
    assert p["nodes"] == [top, 5, nn] and set(p["displacement"]) == {str(top), "5", str(nn)} and len(p["position"]) == 3
    assert np.array_equal(p["acceleration"][str(nn)]["y"], num.a[:, int(eq[nn - 1, 1])])      # x is a roller dof there
    assert not p["acceleration"][str(nn)]["x"].any()
    t0 = time.perf_counter()
    w.vtk(binary=True, output_interval=3)                            # one frame
    t_vtk = time.perf_counter() - t0
    path = os.path.join(tmp_path, "VTK", "data_0.vtk")
    head = open(path, "rb").read(120).decode("latin1")
    assert f"POINTS {nn} float" in head and os.path.getsize(path) > nn * (12 + 3 * 24)
    assert t_subset < 5.0 and t_vtk < 20.0, (t_subset, t_vtk)


def test_bench_reference_arm_runs_without_gpu():
    """`bench.py --impl reference` (the CPU path on the host cores) needs no GPU and prints the contract's JSON line."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-size", "6", "--cpu-steps", "3"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "dof_timesteps_per_s" and line["unit"] == "DOF*steps/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["dtype"] == "f64"
    # "reference": the reference's own assembler (baseline/_ref or /root/reference) was timed; "port": only the numpy
    # restatement was available
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1 and "sample" in line["cpu_baseline"]
    if line["cpu_baseline"]["kind"] == "reference":
        assert line["cpu_baseline"]["assembly_elem_per_s"] > 0 and "6x6x6" in line["config"]["workload"]
    assert line["e2e"] == {"value": line["value"], "unit": "DOF*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


# ---- property tests (hypothesis) ----------------------------------------------------------------------------------
from hypothesis import given, settings as hsettings, strategies as st  # noqa: E402


@hsettings(max_examples=25, deadline=None, derandomize=True, database=None)
@given(et=st.sampled_from(["tri3", "tri6", "quad4", "tetra4", "tetra10", "hexa8", "hexa20"]), nn=st.integers(30, 80),
       ne=st.integers(1, 40), nline=st.integers(0, 6), seed=st.integers(0, 2 ** 31 - 1))
def test_gmsh_write_read_round_trip(et, nn, ne, nline, seed, tmp_path_factory):
    """Any (nodes, connectivity, tags) written by the repo's writer comes back unchanged, with or without leading
    lower-dimensional records, and node coordinates survive to the last bit."""
    from scatter_b200 import _lib, gmsh_io
    rng = np.random.default_rng(seed)
    nne = _lib.ELEM_NNE[et]
    nodes = np.c_[np.arange(1, nn + 1), rng.normal(size=(nn, 3)) * 10.0 ** rng.integers(-3, 4)]
    elem = np.array([rng.choice(nn, nne, replace=False) + 1 for _ in range(ne)])
    tags = rng.integers(1, 4, ne)
    phys = [[float(_lib.ELEM_DIM[et]), 1, "a b"], [float(_lib.ELEM_DIM[et]), 2, "soil"], [float(_lib.ELEM_DIM[et]), 3, "c"]]
    path = os.path.join(tmp_path_factory.mktemp("rt"), "m.msh")
    gmsh_io.write_msh(path, nodes, elem, tags, phys, et)
    if nline:
        src = open(path).read().splitlines()
        i0 = src.index("$Elements")
        lines = [f"{k + 1} 1 2 9 9 {k + 1} {k + 2}" for k in range(nline)]
        body = [" ".join([str(k + 1 + nline)] + l.split()[1:]) for k, l in enumerate(src[i0 + 2:i0 + 2 + ne])]
        open(path, "w").write("\n".join(src[:i0 + 1] + [str(ne + nline)] + lines + body + src[i0 + 2 + ne:]) + "\n")
    r = gmsh_io.read_msh(path)
    assert np.array_equal(r["nodes"], nodes)
    code = gmsh_io.TYPE_TO_GMSH[et]
    blocks = {b[0]: b for b in r["elements"]}
    assert np.array_equal(blocks[code][2], elem) and np.array_equal(blocks[code][1], tags)
    assert (1 in blocks) == (nline > 0) and [p[2] for p in r["physical_names"]] == ["a b", "soil", "c"]


@hsettings(max_examples=30, deadline=None, derandomize=True, database=None)
@given(nn=st.integers(1, 400), world=st.integers(1, 9), seed=st.integers(0, 2 ** 31 - 1), dup=st.booleans())
def test_rcb_properties(nn, world, seed, dup):
    """Recursive coordinate bisection: every node gets a rank, the parts differ by at most one node per bisection level,
    and coincident points (ties) do not break the balance."""
    import types
    from scatter_b200 import partition
    rng = np.random.default_rng(seed)
    xyz = rng.normal(size=(nn, 3))
    if dup:
        xyz[:, 0] = np.round(xyz[:, 0])                         # many ties along x
        xyz[:, 1:] = 0.0
    model = types.SimpleNamespace(nodes=np.c_[np.arange(1, nn + 1), xyz])
    owner = partition.owner_by_rcb(model, world)
    counts = np.bincount(owner, minlength=world)
    assert owner.min() >= 0 and owner.max() < world and counts.sum() == nn
    assert counts.max() - counts.min() <= int(np.ceil(np.log2(max(world, 2)))) + 1
    assert np.array_equal(owner, partition.owner_by_rcb(model, world))


@hsettings(max_examples=30, deadline=None, derandomize=True, database=None)
@given(n_eq=st.integers(1, 60), steps=st.integers(1, 12), world=st.integers(1, 4), seed=st.integers(0, 2 ** 31 - 1))
def test_localised_schedules_partition_the_global_one(n_eq, steps, world, seed):
    import types
    from scatter_b200 import partition
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, 5, steps)
    ptr = np.concatenate([[0], np.cumsum(counts)])
    dof = rng.integers(0, n_eq, ptr[-1]); val = rng.normal(size=ptr[-1])
    owner_of_eq = rng.integers(0, world, n_eq)
    seen = []
    for r in range(world):
        geq = np.where(owner_of_eq == r)[0]
        loc = rng.permutation(len(geq) + 3)[:len(geq)]          # arbitrary local numbering (ghost dofs in between)
        dom = types.SimpleNamespace(n_global_eq=n_eq, global_eq_of_owned=geq, owned_eq=loc)
        lp, ld, lv = partition.localise_schedule(dom, ptr, dof, val)
        assert lp[0] == 0 and len(lp) == steps + 1 and lp[-1] == len(ld) == len(lv)
        back = {int(l): int(g) for l, g in zip(loc, geq)}
        for t in range(steps):
            seen += [(t, back[int(d)], float(v)) for d, v in zip(ld[lp[t]:lp[t + 1]], lv[lp[t]:lp[t + 1]])]
    want = [(t, int(d), float(v)) for t in range(steps) for d, v in zip(dof[ptr[t]:ptr[t + 1]], val[ptr[t]:ptr[t + 1]])]
    assert sorted(seen) == sorted(want)


def test_committed_bench_lines_carry_the_contract_keys():
    """The JSON lines kept under profiles/ are what `bench.py` printed on the B200: every key of the bench contract is there."""
    import json
    for fn, n in (("r1_bench_1gpu_final.json", 1), ("r1_bench_2gpu_final.json", 2)):
        line = json.loads(open(os.path.join(ROOT, "profiles", fn)).read().strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                    "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
            assert key in line, (fn, key)
        assert line["n_gpus"] == n and line["dtype"] == "f64" and line["scaling"] == "weak" and "workload" in line["config"]
        assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
        assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-12
        assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
        assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["value"] < line["value"] and line["gpu_launches"] > 0
        if n == 1:
            assert set(("value", "unit", "cores", "kind", "sample")) <= set(line["cpu_baseline"])


def test_absorbing_plan_groups_and_restriction(golden_meshes):
    """Index plan handed to sc_add_absorbing_faces: sorted unique keys, entries grouped in face order, restriction to owned rows."""
    from scatter_b200 import mesher, system_matrix
    m = mesher.ReadMesh(golden_meshes["cube.msh"])
    m.read_gmsh(); m.read_bc(cases.BC_CUBE_ABS); m.mapping(); m.connectivities()
    plan = system_matrix.absorbing_plan(m)
    nf, nl = plan.i1.shape
    assert plan.face_type == "quad4" and nl == 4 and plan.nodes.shape == (nf, 4) and plan.perp.shape == (nf, 4)
    key = plan.rows * m.number_eq + plan.cols
    assert (np.diff(key) > 0).all()                                                   # sorted, unique
    assert plan.grp_ptr[0] == 0 and plan.grp_ptr[-1] == nf * nl * nl == len(plan.grp_entry)
    assert sorted(plan.grp_entry.tolist()) == list(range(nf * nl * nl))              # every per-face entry exactly once
    for u in (0, len(plan.rows) // 2, len(plan.rows) - 1):
        ent = plan.grp_entry[plan.grp_ptr[u]:plan.grp_ptr[u + 1]]
        assert (np.diff(ent) > 0).all()                                               # face order inside a key
        f, a, b = ent // (nl * nl), (ent // nl) % nl, ent % nl
        assert (plan.i1[f, a] == plan.rows[u]).all() and (plan.i1[f, b] == plan.cols[u]).all()
    n_keys = len(plan.rows)
    half = np.unique(plan.rows)[::2]
    plan.restrict_rows(half)
    assert 0 < len(plan.rows) < n_keys and np.isin(plan.rows, half).all() and plan.grp_ptr[-1] == len(plan.grp_entry)
    plan.restrict_rows(np.zeros(0, dtype=np.int64))
    assert len(plan.rows) == 0 and len(plan.grp_entry) == 0 and plan.grp_ptr.tolist() == [0]
    # no absorbing dofs -> no plan
    m2 = mesher.ReadMesh(golden_meshes["cube.msh"])
    m2.read_gmsh(); m2.read_bc(cases.BC_CUBE); m2.mapping(); m2.connectivities()
    assert system_matrix.absorbing_plan(m2) is None


def test_absorbing_codes_on_2d_meshes_are_ignored_like_in_the_reference(golden_meshes, oracle):
    """2-D element types have no face element (`nb_nodes_lower_elem == []`), so the reference never finds an absorbing
    face on a 2-D mesh -- its "not implemented for 2D" exit is unreachable -- and the "2" dofs simply stay free."""
    from scatter_b200 import mesher, system_matrix
    bc = dict(cases.BC_2D); bc["bottom"] = ["02", bc["bottom"][1]]
    m = mesher.ReadMesh(golden_meshes["column_2D.msh"])
    m.read_gmsh(); m.read_bc(bc); m.mapping(); m.connectivities()
    assert (np.asarray(m.type_BC) == "Absorb").any() and m.nb_nodes_lower_elem == []
    assert system_matrix.absorbing_plan(m) is None
    ne = len(m.elem)
    assert system_matrix.absorbing_entries(m, np.full(ne, 30e6), np.full(ne, 0.2), np.full(ne, 1500.0), 2, [1, 1], 1e3) == ({}, {})
    om = oracle.build_model(golden_meshes["column_2D.msh"], bc)
    assert om.number_eq == m.number_eq
    K, M, C, _ = oracle.system_matrices(om, cases.materials(), cases.settings())     # runs, no absorbing contribution
    E, nu, rho = oracle.element_properties(om, cases.materials())
    K0, _ = oracle.assemble_global(om, E, nu, rho, 2)
    assert abs(K - K0).max() == 0


def test_header_is_plain_c():
    """include/scatter_b200.h is the drop-in boundary: it must compile as C (no C++ types, no torch types in the signatures)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    hdr = os.path.join(ROOT, "include", "scatter_b200.h")
    r = subprocess.run([gcc, "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror", hdr], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(hdr).read()
    assert "torch" not in text and "std::" not in text and 'extern "C"' in text
