"""`scatter.scatter(...)` -- the reference's only entry point (`scatter/scatter.py:36-171`), kept verbatim as the drop-in
boundary: same signature, same dict schemas, same return object (`export_results.Write`), same files
written (`<outfile_folder>/data.pickle`, `<outfile_folder>/VTK/data_<k>.vtk`).  The hot path underneath -- element
integration, assembly, damping, time integration -- runs on the B200 through libscatter_b200.so.
"""
import sys
from enum import Enum

import numpy as np

from . import export_results, force_external, mesher, random_fields, solvers, system_matrix, validator


class Solver(Enum):
    """Solver types (`scatter/scatter.py:16-33`)."""
    STATIC = "StaticSolver"
    NEWMARK_EXPLICIT = "NewmarkExplicit"
    NEWMARK_IMPLICIT = "NewmarkImplicitForce"
    CENTRAL_DIFFERENCE = "CentralDifferenceSolver"
    BATHE = "BatheSolver"


BANNER = r"""
   _____  _____       _______ _______ ______ _____             ____ ___   ___   ___
  / ____|/ ____|   /\|__   __|__   __|  ____|  __ \           |  _ \__ \ / _ \ / _ \
 | (___ | |       /  \  | |     | |  | |__  | |__) |  ______  | |_) | ) | | | | | | |
  \___ \| |      / /\ \ | |     | |  |  __| |  _  /  |______| |  _ < / /| | | | | | |
  ____) | |____ / ____ \| |     | |  | |____| | \ \           | |_) / /_| |_| | |_| |
 |_____/ \_____/_/    \_\_|     |_|  |______|_|  \_\          |____/____|\___/ \___/
"""


def scatter(mesh_file: str, outfile_folder: str, materials: dict, boundaries: dict,
            inp_settings: dict, loading: dict, time_step: float = 0.1, solver: Solver = Solver.NEWMARK_EXPLICIT,
            random_props: bool = False, device: int = 0) -> export_results.Write:
    r"""
    3D finite element code (B200 hot path).
                                                            ^  _
                                                          y |  /| z
    Mesh is generated with gmsh https://gmsh.info/          | /
    The coordinate system is the same as defined in gmsh    -----> x

    Consistent mass matrix (Newmark) / row-sum lumped mass (central difference).

    :param mesh_file: gmsh 2.2 mesh file, or a `ReadMesh`-shaped object built with `ReadMesh.from_arrays`
    :param outfile_folder: location of the output folder
    :param materials: dictionary with material properties
    :param boundaries: dictionary with boundary conditions
    :param inp_settings: dictionary with numerical settings
    :param loading: dictionary with loading conditions
    :param time_step: time step for the analysis (optional: default 0.1 s)
    :param solver: solver to use for the analysis, see `Solver` enum (optional: default Newmark explicit)
    :param random_props: random-field settings (optional: default False)
    :param device: CUDA device ordinal (extension; default 0)
    """
    print(BANNER)

    validator.ValidateLoad.validate(loading)

    if isinstance(mesh_file, mesher.ReadMesh):
        model = mesh_file
    else:
        model = mesher.ReadMesh(mesh_file)
        model.read_gmsh()
    model.read_bc(boundaries)
    model.mapping()
    model.connectivities()
    if loading["type"] == "rose":
        raise NotImplementedError("ROSE train-track coupling is outside the scope of the B200 hot path")
    model.get_mesh_edges()

    elem_props = None
    if random_props:
        print("Generating random field")
        rf = random_fields.RF(random_props, materials, outfile_folder, model.element_type)
        material_idx = [material[1] for material in model.materials if material[2] == random_props["material"]][0]
        elements = model.elem[model.materials_index == material_idx]
        rf.generate_gstools_rf(model.nodes, elements, model.dimension, angles=0.0)
        rf.dump()
        rf.update_material_list(materials, model, material_idx)
        materials.update(rf.new_material)

    print("Generating global matrices scatter")
    matrix = system_matrix.GenerateMatrix(model.number_eq, inp_settings['int_order'], device=device)
    explicit = solver == Solver.CENTRAL_DIFFERENCE
    matrix.want_full_mass = not explicit
    matrix.want_lumped_mass = explicit
    matrix.generate_stiffness_and_mass(model, materials, elem_props=elem_props)
    matrix.absorbing_boundaries(model, materials, inp_settings["absorbing_BC"], inp_settings["absorbing_BC_stiff"])
    matrix.damping_Rayleigh(inp_settings["damping"])

    time = np.linspace(0, loading["time"], int(np.ceil(loading["time"] / time_step) + 1))

    if solver == Solver.NEWMARK_EXPLICIT:
        numerical = solvers.NewmarkExplicit()
    elif solver == Solver.NEWMARK_IMPLICIT:
        numerical = solvers.NewmarkImplicitForce()
    elif solver == Solver.CENTRAL_DIFFERENCE:
        numerical = solvers.CentralDifferenceSolver()
    elif solver == Solver.BATHE:
        numerical = solvers.BatheSolver()
    elif solver == Solver.STATIC:
        numerical = solvers.StaticSolver()
    else:
        sys.exit(f"Error: {solver} not supported")

    numerical.output_interval = inp_settings["output_interval"] if "output_interval" in inp_settings.keys() else 1
    numerical.initialise(model.number_eq, time)
    numerical.bind(matrix)

    print("Setting load")
    F = force_external.Force()
    top_surface_elements = model.get_top_surface() if loading["type"] == "moving_at_plane" else []
    F.initialise_load(loading, time, model, numerical, top_surface_elements=top_surface_elements)
    numerical.update_rhs_at_time_step_func = F.update_load_at_t

    print("solver started")
    if solver == Solver.STATIC:
        numerical.calculate(None, F.force_vector, 0, len(F.time) - 1)
    else:
        numerical.update(0)
        numerical.calculate(None, None, None, F.force_vector, 0, len(F.time) - 1)

    results = export_results.Write(outfile_folder, model, materials, numerical)
    results.matrix = matrix
    results.pickle(write=inp_settings["pickle"], nodes=inp_settings["pickle_nodes"])
    results.vtk(write=inp_settings["VTK"], binary=inp_settings["VTK_binary"], output_interval=1)

    print("\n\n\n\x1B[3m" + "  Never tell me the odds. " + "\x1B[0m")
    print("\x1B[3m" + "--- Han Solo" + "\x1B[0m")
    return results
