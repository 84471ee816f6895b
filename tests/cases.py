"""Shared definitions of the reference's own test cases (settings copied as *data* from the reference tests).

Sources: integration_tests/integration_test.py:67-102 (hexa8 column pulse), :171-206 (absorbing bottom),
:216-251 (hexa20 column), :336-369 (quad4 column), :493-545 (cube, moving load);
integration_tests/test_benchmark_set_2.py:16-80 (tri3/tri6/tetra4/tetra10 columns).
"""


def _box_bc(x, y, z, bottom="010", front_quirk=False):
    # `front_quirk`: integration_test.py lists the second point of "front" as [z, 0, 0] (sic)
    return {"bottom": [bottom, [[0, 0, 0], [x, 0, 0], [0, 0, z], [x, 0, z]]],
            "left": ["100", [[0, 0, 0], [0, 0, z], [0, y, 0], [0, y, z]]],
            "right": ["100", [[x, 0, 0], [x, 0, z], [x, y, 0], [x, y, z]]],
            "front": ["001", [[0, 0, 0], [z if front_quirk else x, 0, 0], [0, y, 0], [x, y, 0]]],
            "back": ["001", [[0, 0, z], [x, 0, z], [0, y, z], [x, y, z]]]}


BC_COLUMN = _box_bc(0.1, 20, -0.1, front_quirk=True)
BC_COLUMN_ABS = _box_bc(0.1, 20, -0.1, bottom="020", front_quirk=True)
BC_CUBE = _box_bc(10, 10, -10, front_quirk=True)
BC_CUBE_ABS = dict(_box_bc(10, 10, -10, bottom="020", front_quirk=True))
BC_CUBE_ABS["left"] = ["200", BC_CUBE["left"][1]]
BC_B2_3D = _box_bc(1, 10, -1)
BC_2D = {"bottom": ["01", [[0, 0, 0], [1, 0, 0]]], "left": ["10", [[0, 0, 0], [0, 10, 0]]],
         "right": ["10", [[1, 0, 0], [1, 10, 0]]]}


def materials():
    return {"solid": {"density": 1500, "Young": 30e6, "poisson": 0.2},
            "bottom": {"density": 1200, "Young": 300e6, "poisson": 0.25}}


def materials_nu0():
    return {"solid": {"density": 1500, "Young": 30e6, "poisson": 0.0}}


def settings(**over):
    s = {"gamma": 0.5, "beta": 0.25, "int_order": 2, "damping": [1, 0.001, 30, 0.001], "absorbing_BC": [1, 1],
         "absorbing_BC_stiff": 1e3, "pickle": True, "pickle_nodes": "all", "VTK": False, "VTK_binary": True}
    s.update(over)
    return s


# (mesh file, BC) pairs whose assembled matrices are pinned in tests/golden/matrices.npz
MATRIX_CASES = {
    "column": ("column.msh", BC_COLUMN),
    "column_abs": ("column.msh", BC_COLUMN_ABS),
    "column_high_order": ("column_high_order.msh", BC_COLUMN),
    "column_high_order_abs": ("column_high_order.msh", BC_COLUMN_ABS),
    "cube": ("cube.msh", BC_CUBE),
    "cube_abs": ("cube.msh", BC_CUBE_ABS),
    "column_2D": ("column_2D.msh", BC_2D),
    "column_2D_tri3": ("column_2D_tri3.msh", BC_2D),
    "column_2D_tri6": ("column_2D_tri6.msh", BC_2D),
    "column_3D_tetra4": ("column_3D_tetra4.msh", BC_B2_3D),
    "column_3D_tetra10": ("column_3D_tetra10.msh", BC_B2_3D),
}

# meshes of the reference's run scripts (mesh/*.msh; settings from run_scatter_rose_2D.py:22-41): unstructured,
# multi-material plane-strain triangle / quad meshes that also carry 1-D "rose" line elements the reader must skip
def _bc_2d(x, y0, y1):
    return {"bottom": ["11", [[0, y0, 0], [x, y0, 0]]], "left": ["10", [[0, y0, 0], [0, y1, 0]]],
            "right": ["10", [[x, y0, 0], [x, y1, 0]]]}


MATRIX_CASES.update({
    "rose_2D_side": ("rose_2D_side.msh", _bc_2d(90, -3, 0.5)),
    "embankment_rose2D": ("embankment_rose2D.msh", _bc_2d(10, -5, 0.5)),
    "box2d": ("box2d.msh", _bc_2d(120, 0, 1.8)),
})


def materials_embankment():
    return {"embankment": {"density": 2000, "Young": 100e6, "poisson": 0.2},
            "soil1": {"density": 1700, "Young": 500e5, "poisson": 0.2},
            "soil2": {"density": 2000, "Young": 200e5, "poisson": 0.2}}


def case_materials(case):
    """Material dictionary of a MATRIX_CASE."""
    return materials_embankment() if case in ("rose_2D_side", "embankment_rose2D") else materials()


def rf_properties(case, model_name="Gaussian"):
    """Random-field settings in the shape of run_scatter_rose_2D.py:62-73 / integration_test.py:148-158."""
    return {"number_realisations": 1, "element_size": 1, "theta": 1.5, "seed_number": -26021981,
            "material": "soil1" if case in ("rose_2D_side", "embankment_rose2D") else "solid", "key_material": "Young",
            "std_value": 3e6, "aniso_x": 10, "aniso_z": 2, "model_name": model_name}


B2_NODES = {
    "tri3": ([3, 4, 25], 2),
    "tri6": ([3, 4, 47, 48, 49], 2),
    "tetra4": ([3, 4, 7, 8, 29, 69, 91, 92, 132], 3),
    "tetra10": ([3, 4, 7, 8, 51, 52, 53, 135, 136, 137, 183, 184, 185, 186, 187, 188, 432, 433, 434, 435, 436, 437,
                 438, 439, 440], 3),
}


def history_case(name):
    """-> dict(mesh, bc, materials, settings, loading, time_step) for the golden time histories."""
    if name == "hexa8_pulse":
        return dict(mesh="column.msh", bc=BC_COLUMN, materials=materials(), settings=settings(),
                    loading={"force": [0, -1000, 0], "node": [3, 4, 7, 8], "time": 0.5, "type": "pulse"},
                    time_step=0.5e-3)
    if name == "quad4_heaviside":
        return dict(mesh="column_2D.msh", bc=BC_2D, materials=materials(),
                    settings=settings(damping=[1, 0.005, 20, 0.005]),
                    loading={"force": [0, -1e6, 0], "node": [3, 4, 25], "time": 1.0, "type": "heaviside"},
                    time_step=5e-3)
    if name in B2_NODES:
        nodes, nd = B2_NODES[name]
        return dict(mesh=f"column_{nd}D_{name}.msh", bc=BC_2D if nd == 2 else BC_B2_3D, materials=materials_nu0(),
                    settings=settings(pickle_nodes=[3], output_interval=10),
                    loading={"force": [0, 1000 / len(nodes), 0], "node": list(nodes), "time": 1, "type": "heaviside"},
                    time_step=5e-4)
    raise KeyError(name)


HISTORY_CASES = ["hexa8_pulse", "quad4_heaviside", "tri3", "tri6", "tetra4", "tetra10"]
