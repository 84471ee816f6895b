"""Where a small Newmark run spends its time (reference-sized meshes): device vs wall, iterations and launches per step."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scatter_b200 import _lib, boxmesh, system_matrix

for et, s in (("hexa8", 8), ("hexa8", 24), ("hexa8", 32), ("hexa8", 48)):
    model = boxmesh.box_model(s, s, s, 0.5, et)
    ne = len(model.elem)
    mx = system_matrix.GenerateMatrix(model.number_eq, 2)
    ctx = mx.ctx
    ctx.set_mesh(et, model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), model.number_eq, None)
    ctx.set_materials(np.full(ne, 30e6), np.full(ne, 0.2), np.full(ne, 1500.0))
    ctx.build_pattern(); ctx.assemble(2, 3)
    mx.damping_Rayleigh([1, 0.01, 30, 0.01])
    d = int(model.eq_nb_dof[boxmesh.top_centre_node(s, s, s) - 1, 1])
    nt = 401
    ctx.set_load_schedule(np.arange(nt + 1, dtype=np.int64), np.full(nt, d, dtype=np.int64), -1000.0 * np.minimum(np.arange(nt) / 4.0, 1.0))
    for store in (False,):
        ctx.set_state(None, None)
        ctx.run_newmark(5e-4, 0, 2, 1, store=False)
        t0 = time.perf_counter()
        _, _, _, st = ctx.run_newmark(5e-4, 2, nt - 3, 1, store=store)
        wall = time.perf_counter() - t0
        n = nt - 3
        print(f"{et} {s}^3 n_eq {model.number_eq} store={store}: wall {1e3*wall/n:.3f} ms/step, device {1e3*st['seconds_device']/n:.3f} ms/step, "
              f"{st['pcg_iterations']/n:.1f} its/step, {st['kernel_launches']/n:.1f} launches/step", flush=True)
    ctx.close()
