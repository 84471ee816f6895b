"""GPU timing of the hexa20 Newmark/PCG workload (BASELINE config 4) for kernel options: ms per PCG iteration."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from scatter_b200 import _lib, boxmesh, system_matrix
s = int(sys.argv[1]) if len(sys.argv) > 1 else 94
variants = sys.argv[2:] or [""]
model = boxmesh.box_model(s, s, s, bench.H, "hexa20")
ne = len(model.elem)
E = boxmesh.lognormal_young(ne, bench.E_MEAN, bench.E_STD)
for v in variants:
    _lib.DEFAULT_OPTIONS.clear()
    for kv in filter(None, v.split(",")):
        k, val = kv.split("="); _lib.DEFAULT_OPTIONS[k] = int(val)
    mx = system_matrix.GenerateMatrix(model.number_eq, 2)
    ctx = mx.ctx
    ctx.set_mesh("hexa20", model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), model.number_eq, None)
    ctx.set_materials(E, np.full(ne, bench.NU), np.full(ne, bench.RHO))
    ctx.build_pattern(); ctx.assemble(2, _lib.ASM_K | _lib.ASM_M_FULL)
    mx.damping_Rayleigh(bench.DAMPING)
    d = int(model.eq_nb_dof[boxmesh.top_centre_node(s, s, s) - 1, 1])
    nt = 40
    ctx.set_load_schedule(np.arange(nt + 1, dtype=np.int64), np.full(nt, d, dtype=np.int64), np.full(nt, -1000.0))
    ctx.set_state(None, None)
    ctx.run_newmark(5e-4, 0, 1, 1, rtol=1e-10, store=False)
    _, _, _, st = ctx.run_newmark(5e-4, 1, 3, 1, rtol=1e-10, store=False)
    print(f"[{v or 'default'}] {st['pcg_iterations'] / 3:.1f} it/step, {1e3 * st['seconds_device'] / st['pcg_iterations']:.3f} ms/iteration, "
          f"{1e3 * st['seconds_device'] / 3:.1f} ms/step", flush=True)
    ctx.close()
