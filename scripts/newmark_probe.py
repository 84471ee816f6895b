"""Newmark / PCG probe on structured boxes: iterations per step and time per iteration (run on the GPU box)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scatter_b200 import _lib, boxmesh, system_matrix

# usage: newmark_probe.py [element_type size dt]   (default: three small boxes)
CASES = [(sys.argv[1], int(sys.argv[2]), float(sys.argv[3]))] if len(sys.argv) > 3 else \
    [("hexa8", 128, 5e-4), ("hexa8", 128, 2e-3), ("hexa20", 48, 5e-4)]
for et, s, dt in CASES:
    model = boxmesh.box_model(s, s, s, 0.5, et)
    ne = len(model.elem)
    mx = system_matrix.GenerateMatrix(model.number_eq, 2)
    ctx = mx.ctx
    ctx.set_mesh(et, model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), model.number_eq, None)
    ctx.set_materials(boxmesh.lognormal_young(ne, 30e6, 1e6), np.full(ne, 0.2), np.full(ne, 1500.0))
    nnz = ctx.build_pattern()
    ctx.assemble(2, 3)
    mx.damping_Rayleigh([1, 0.01, 30, 0.01])
    d = int(model.eq_nb_dof[boxmesh.top_centre_node(s, s, s) - 1, 1])
    nt = 16
    ctx.set_load_schedule(np.arange(nt + 1, dtype=np.int64), np.full(nt, d, dtype=np.int64), -1000.0 * np.minimum(np.arange(nt) / 4.0, 1.0))
    ctx.set_state(None, None)
    print(f"device memory after assembly: {ctx.device_info()['free_mem'] / 1e9:.1f} GB free of {ctx.device_info()['total_mem'] / 1e9:.1f} GB", flush=True)
    for rtol in (1e-10, 1e-14):
        ctx.set_state(None, None)
        ctx.run_newmark(dt, 0, 2, 1, rtol=rtol, store=False)
        _, _, _, st = ctx.run_newmark(dt, 2, 5, 1, rtol=rtol, store=False)
        its = st["pcg_iterations"] / 5
        print(f"{et} {s}^3 n_eq {model.number_eq} nnz {nnz} dt {dt} rtol {rtol:g}: {its:.1f} its/step, "
              f"{1e3 * st['seconds_device'] / 5:.1f} ms/step, {1e3 * st['seconds_device'] / max(st['pcg_iterations'], 1):.3f} ms/iteration, "
              f"{model.number_eq * 5 / st['seconds_device']:.3e} DOF*steps/s", flush=True)
    ctx.close()
