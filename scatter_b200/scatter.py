"""Drop-in entry point: `scatter(...)` with the signature, dict schemas, printed progress lines, returned object and output
files of the reference's `scatter/scatter.py:36-171`; the hot path underneath (element integration, assembly, damping,
time integration) runs on the B200 through libscatter_b200.so.

The reference body is one long function; here it is a small pipeline object so that the stages can also be driven one by
one (benchmarks, multi-GPU drivers):  Pipeline.mesh -> .matrices -> .solver -> .loads -> .integrate -> .export
"""
import sys
from enum import Enum

import numpy as np

from . import export_results, force_external, mesher, random_fields, solvers, system_matrix, validator


class Solver(Enum):
    """Solver selection; member names and values as in `scatter/scatter.py:16-33`."""
    STATIC = "StaticSolver"
    NEWMARK_EXPLICIT = "NewmarkExplicit"
    NEWMARK_IMPLICIT = "NewmarkImplicitForce"
    CENTRAL_DIFFERENCE = "CentralDifferenceSolver"
    BATHE = "BatheSolver"


_SOLVER_CLASSES = {
    Solver.NEWMARK_EXPLICIT: solvers.NewmarkExplicit,
    Solver.NEWMARK_IMPLICIT: solvers.NewmarkImplicitForce,
    Solver.CENTRAL_DIFFERENCE: solvers.CentralDifferenceSolver,
    Solver.BATHE: solvers.BatheSolver,
    Solver.STATIC: solvers.StaticSolver,
}

BANNER = r"""
   ___  ___   _  _____ _____ ___ ___     _    ___ __   __
  / __|/ __| /_\|_   _|_   _| __| _ \   | |__|_  )  \ /  \
  \__ \ (__ / _ \ | |   | | | _||   /   | '_ \/ / () | () |
  |___/\___/_/ \_\|_|   |_| |___|_|_\   |_.__/___\__/ \__/
"""


class Pipeline:
    """One SCATTER analysis, stage by stage.  All heavy state lives on the GPU behind `self.matrix.ctx`."""

    def __init__(self, materials, boundaries, settings, loading, time_step, solver, random_props, device):
        self.materials, self.boundaries, self.settings, self.loading = materials, boundaries, settings, loading
        self.time_step, self.solver_kind, self.random_props, self.device = time_step, solver, random_props, device
        self.model = self.matrix = self.numerical = self.force = self.time = None

    def mesh(self, mesh_file):
        """gmsh file (or a prepared `ReadMesh`) -> nodes, elements, BC codes, equation numbers (`scatter.py:68-85`)."""
        if isinstance(mesh_file, mesher.ReadMesh):
            model = mesh_file
        else:
            model = mesher.ReadMesh(mesh_file)
            model.read_gmsh()
        if getattr(model, "prepared_bc", None) != self.boundaries:      # a prepared model already carries these BC / numbers
            model.read_bc(self.boundaries)
            model.mapping()
        model.connectivities()
        if self.loading["type"] == "rose":
            raise NotImplementedError("ROSE train-track coupling is outside the scope of the B200 hot path")
        model.get_mesh_edges()
        self.model = model
        return self

    def random_field(self, out_folder):
        """Per-element material from a random field (`scatter.py:87-99`); mutates `materials` like the reference."""
        rp = self.random_props
        if not rp:
            return self
        print("Generating random field")
        rf = random_fields.RF(rp, self.materials, out_folder, self.model.element_type, device=self.device)
        # physical tag of the material that gets the field, its elements, the field, the re-indexed material list
        tag = [m[1] for m in self.model.materials if m[2] == rp["material"]][0]
        rf.generate_gstools_rf(self.model.nodes, self.model.elem[self.model.materials_index == tag], self.model.dimension, angles=0.0)
        rf.dump()
        rf.update_material_list(self.materials, self.model, tag)
        self.rf = rf
        self.materials.update(rf.new_material)
        return self

    def matrices(self):
        """K, M (consistent or lumped), absorbing boundaries, Rayleigh damping on the device (`scatter.py:101-114`)."""
        print("Generating global matrices scatter")
        mx = system_matrix.GenerateMatrix(self.model.number_eq, self.settings['int_order'], device=self.device)
        explicit = self.solver_kind == Solver.CENTRAL_DIFFERENCE
        mx.want_full_mass, mx.want_lumped_mass = not explicit, explicit
        mx.generate_stiffness_and_mass(self.model, self.materials)
        mx.absorbing_boundaries(self.model, self.materials, self.settings["absorbing_BC"], self.settings["absorbing_BC_stiff"])
        mx.damping_Rayleigh(self.settings["damping"])
        self.matrix = mx
        return self

    def solver(self):
        """Time axis and solver object (`scatter.py:116-138`)."""
        total = self.loading["time"]
        self.time = np.linspace(0, total, int(np.ceil(total / self.time_step) + 1))
        if self.solver_kind not in _SOLVER_CLASSES:
            sys.exit(f"Error: {self.solver_kind} not supported")
        num = _SOLVER_CLASSES[self.solver_kind]()
        num.output_interval = self.settings.get("output_interval", 1)
        if "pcg_rtol" in self.settings and hasattr(num, "pcg_rtol"):     # optional: tolerance of the iterative solve
            num.pcg_rtol = float(self.settings["pcg_rtol"])
        num.initialise(self.model.number_eq, self.time)
        num.bind(self.matrix)
        self.numerical = num
        return self

    def loads(self):
        """External force object and the solver's per-step callback (`scatter.py:140-151`)."""
        print("Setting load")
        force = force_external.Force()
        top = self.model.get_top_surface() if self.loading["type"] == "moving_at_plane" else []
        force.initialise_load(self.loading, self.time, self.model, self.numerical, top_surface_elements=top)
        self.numerical.update_rhs_at_time_step_func = force.update_load_at_t
        self.force = force
        return self

    def integrate(self):
        """`scatter.py:153-159`"""
        print("solver started")
        last = len(self.force.time) - 1
        if self.solver_kind == Solver.STATIC:
            self.numerical.calculate(None, self.force.force_vector, 0, last)
        else:
            self.numerical.update(0)
            self.numerical.calculate(None, None, None, self.force.force_vector, 0, last)
        return self

    def export(self, out_folder):
        """`scatter.py:161-166`"""
        res = export_results.Write(out_folder, self.model, self.materials, self.numerical)
        res.matrix = self.matrix
        res.pickle(write=self.settings["pickle"], nodes=self.settings["pickle_nodes"])
        res.vtk(write=self.settings["VTK"], binary=self.settings["VTK_binary"], output_interval=1)
        return res


def scatter(mesh_file: str, outfile_folder: str, materials: dict, boundaries: dict,
            inp_settings: dict, loading: dict, time_step: float = 0.1, solver: Solver = Solver.NEWMARK_EXPLICIT,
            random_props: bool = False, device: int = 0) -> export_results.Write:
    r"""
    2D / 3D linear-elastic dynamic finite element analysis (B200 hot path).  Coordinate system as in gmsh: y is up.

    :param mesh_file: gmsh 2.2 mesh file, or a `ReadMesh` object built with `ReadMesh.from_arrays`
    :param outfile_folder: location of the output folder
    :param materials: {name: {"density", "Young", "poisson"}}
    :param boundaries: {name: [dof code string, [points]]}  (0 free, 1 fixed, 2 absorbing)
    :param inp_settings: numerical settings (int_order, damping, absorbing_BC, absorbing_BC_stiff, pickle, pickle_nodes,
                         VTK, VTK_binary, optional output_interval)
    :param loading: {"type", "force", "node", "time", ...}
    :param time_step: time step of the analysis (default 0.1 s)
    :param solver: member of `Solver` (default Newmark)
    :param random_props: random-field settings or False
    :param device: CUDA device ordinal (extension of the reference signature)
    """
    print(BANNER)
    validator.ValidateLoad.validate(loading)
    run = Pipeline(materials, boundaries, inp_settings, loading, time_step, solver, random_props, device)
    run.mesh(mesh_file).random_field(outfile_folder).matrices().solver().loads().integrate()
    results = run.export(outfile_folder)
    print("\n\n\n\x1B[3m" + "  Never tell me the odds. " + "\x1B[0m")
    print("\x1B[3m" + "--- Han Solo" + "\x1B[0m")
    return results
