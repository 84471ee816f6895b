"""TEST INFRASTRUCTURE (container-only helper) -- import the *unmodified* reference from /root/reference.

Only usable where /root/reference exists (the build container).  It is used by
`oracle/make_golden.py` to generate the fixtures under tests/golden/ and by
`oracle/validate_against_reference.py` to pin the numpy restatement in `oracle/fem_np.py`.
Nothing in tests/ -m gpu, bench.py or the product imports this file.

The reference imports third-party packages that are not installed here
(SURVEY.md Appendix B): `rose`, `solvers` (PuggleSolvers 1.0.1), `shapely`, `meshio`,
`gstools`, `vtk_tools`.  They are replaced by empty stub modules so that the pure
numpy/scipy part of the reference (mesher, system_matrix, discretisation, element_types,
material_models, utils) can run as-is.
"""
import sys
import types
import warnings

REF_ROOT = "/root/reference"


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
    import os
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("reference tree not present: this helper only works in the build container")
    install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    from scatter import mesher, system_matrix, discretisation, element_types, material_models, utils
    return types.SimpleNamespace(mesher=mesher, system_matrix=system_matrix, discretisation=discretisation,
                                 element_types=element_types, material_models=material_models, utils=utils)
