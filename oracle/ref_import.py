"""TEST INFRASTRUCTURE (container-only helper) -- import the *unmodified* reference from /root/reference.

In the build container it imports /root/reference: `oracle/make_golden.py` generates the fixtures under tests/golden/
with it and `oracle/validate_against_reference.py` pins the numpy restatement in `oracle/fem_np.py`.  On the GPU box the
same code is available as `baseline/_ref` (an offline `pip install --target` of the reference tree made by
`__graft_entry__.build()`, git-ignored): only the CPU-baseline legs of `bench.py` (`cpu_baseline`, `--impl reference`)
import it from there -- to TIME the reference's own assembler.  No test marked gpu, smoke() or product code imports this file.

The reference imports third-party packages that are not installed here
(SURVEY.md Appendix B): `rose`, `solvers` (PuggleSolvers 1.0.1), `shapely`, `meshio`,
`gstools`, `vtk_tools`.  They are replaced by empty stub modules so that the pure
numpy/scipy part of the reference (mesher, system_matrix, discretisation, element_types,
material_models, utils) can run as-is.
"""
import sys
import types
import warnings

import os

# the reference tree in the build container; on the GPU box only `baseline/_ref` exists (pip install --target of that same
# tree, git-ignored, made by __graft_entry__.build()) -- bench.py's CPU-baseline legs use whichever is present
REF_ROOT = "/root/reference"
_REF_INSTALL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def install_stubs():
    class _Dummy:  # placeholder for classes that are only referenced, never used
        def __init__(self, *a, **k):
            pass

    _stub("rose")
    _stub("rose.model")
    _stub("rose.model.model_part", Material=_Dummy, Section=_Dummy)
    tti = _stub("rose.model.train_track_interaction", CoupledTrainTrack=_Dummy, List=list)
    tti.__all__ = ["CoupledTrainTrack", "List"]
    _stub("solvers")
    _stub("solvers.newmark_solver", NewmarkSolver=_Dummy, NewmarkExplicit=_Dummy, NewmarkImplicitForce=_Dummy)
    _stub("solvers.static_solver", StaticSolver=_Dummy)
    _stub("solvers.bathe_solver", BatheSolver=_Dummy)
    _stub("solvers.central_difference_solver", CentralDifferenceSolver=_Dummy)
    _stub("meshio")
    _stub("gstools", SRF=_Dummy, Exponential=_Dummy, Gaussian=_Dummy, Linear=_Dummy, Matern=_Dummy)
    _stub("vtk_tools", VTK_writer=_Dummy)
    _stub("shapely")
    _stub("shapely.geometry", Point=_Dummy, Polygon=_Dummy)
    _stub("shapely.geometry.polygon", Polygon=_Dummy)


def load_reference():
    """Return the reference modules (mesher, system_matrix, discretisation, element_types, material_models)."""
    root = REF_ROOT if os.path.isdir(os.path.join(REF_ROOT, "scatter")) else _REF_INSTALL
    if not os.path.isdir(os.path.join(root, "scatter")):
        raise RuntimeError("reference not present (neither /root/reference nor baseline/_ref)")
    install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    from scatter import mesher, system_matrix, discretisation, element_types, material_models, utils
    return types.SimpleNamespace(mesher=mesher, system_matrix=system_matrix, discretisation=discretisation,
                                 element_types=element_types, material_models=material_models, utils=utils)
