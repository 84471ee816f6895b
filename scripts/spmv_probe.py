"""GPU A/B timing of the fused explicit step at the benchmark size for kernel options (sc_set_option).
    python scripts/spmv_probe.py [size] "opt=value,opt=value" "..."         one context, options changed between timings"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from scatter_b200 import _lib, boxmesh, partition, system_matrix
s = int(sys.argv[1]) if len(sys.argv) > 1 else 255
variants = sys.argv[2:] or [""]
dom = partition.slab_partition(s, s, s, 0, 1, bench.H, "hexa8")
model = dom.model
ne = len(model.elem)
E = boxmesh.lognormal_young(ne, bench.E_MEAN, bench.E_STD)
mx = system_matrix.GenerateMatrix(model.number_eq, 2)
ctx = mx.ctx
ctx.set_mesh("hexa8", model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), model.number_eq, None)
ctx.set_materials(E, np.full(ne, bench.NU), np.full(ne, bench.RHO))
ctx.build_pattern(); ctx.assemble(2, _lib.ASM_K | _lib.ASM_M_LUMPED)
mx.damping_Rayleigh(bench.DAMPING)
nt = 100000
d = int(model.eq_nb_dof[boxmesh.top_centre_node(s, s, s) - 1, 1])
ctx.set_load_schedule(np.arange(nt + 1, dtype=np.int64), np.full(nt, d, dtype=np.int64), np.full(nt, -1000.0))
ctx.set_state(None, None)
dt = bench.stable_dt()
t = 0
ctx.run_central_difference(dt, t, 100, 100, store=False); t += 100
for rep in range(2):
    for v in variants:
        for kv in filter(None, v.split(",")):
            k, val = kv.split("=")
            ctx.set_option(k, int(val)); ctx.set_state(None, None)
        ctx.run_central_difference(dt, t, 20, 20, store=False); t += 20
        _, _, _, st = ctx.run_central_difference(dt, t, 200, 200, store=False); t += 200
        print(f"[{v or 'default'}] {1e3 * st['seconds_device'] / 200:.3f} ms/step", flush=True)
u = ctx.get_state()[0]
print("checksum", float(np.abs(u).sum()))
