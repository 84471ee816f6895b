"""GPU A/B of the consumer-group layouts of k_spmv_node (options fixed before the pattern build): ms per explicit step."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from scatter_b200 import _lib, boxmesh, partition, system_matrix
s = int(sys.argv[1]) if len(sys.argv) > 1 else 255
et = sys.argv[2] if len(sys.argv) > 2 else "hexa8"
for v in sys.argv[3:] or [""]:
    _lib.DEFAULT_OPTIONS.clear()
    for kv in filter(None, v.split(",")):
        k, val = kv.split("="); _lib.DEFAULT_OPTIONS[k] = int(val)
    dom = partition.slab_partition(s, s, s, 0, 1, bench.H, et)
    model = dom.model
    ne = len(model.elem)
    E = boxmesh.lognormal_young(ne, bench.E_MEAN, bench.E_STD)
    mx = system_matrix.GenerateMatrix(model.number_eq, 2)
    ctx = mx.ctx
    ctx.set_mesh(et, model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), model.number_eq, None)
    ctx.set_materials(E, np.full(ne, bench.NU), np.full(ne, bench.RHO))
    ctx.build_pattern(); ctx.assemble(2, _lib.ASM_K | _lib.ASM_M_LUMPED)
    mx.damping_Rayleigh(bench.DAMPING)
    nt = 100000
    d = int(model.eq_nb_dof[boxmesh.top_centre_node(s, s, s) - 1, 1])
    ctx.set_load_schedule(np.arange(nt + 1, dtype=np.int64), np.full(nt, d, dtype=np.int64), np.full(nt, -1000.0))
    ctx.set_state(None, None)
    dt = bench.stable_dt() * (0.3 if et == "hexa20" else 1.0)
    t = 0
    ctx.run_central_difference(dt, t, 100, 100, store=False); t += 100
    for rep in range(2):
        _, _, _, st = ctx.run_central_difference(dt, t, 200, 200, store=False); t += 200
        print(f"[{et} {s} {v or 'default'}] {1e3 * st['seconds_device'] / 200:.3f} ms/step", flush=True)
    u = ctx.get_state()[0]
    print("checksum", float(np.abs(u).sum()), flush=True)
    ctx.close()
