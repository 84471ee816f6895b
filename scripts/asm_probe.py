"""Assemble a box a few times (ncu target / quick timing).
usage: asm_probe.py [n=128] [reps=3] [element_type=hexa8] ["opt=v,opt=v" ...]   (one timing block per option set)"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from scatter_b200 import _lib, boxmesh

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
et = sys.argv[3] if len(sys.argv) > 3 else "hexa8"
variants = sys.argv[4:] or [""]
model = boxmesh.box_model(n, n, n, 0.5, et)
ne = model.elem.shape[0]
E = boxmesh.lognormal_young(ne)
ref = None
for v in variants:
    ctx = _lib.Context(0)
    for kv in filter(None, v.split(",")):
        k, val = kv.split("="); ctx.set_option(k, int(val))
    ctx.set_mesh(et, model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), model.number_eq, None)
    ctx.set_materials(E, np.full(ne, 0.2), np.full(ne, 1500.0))
    t0 = time.perf_counter()
    nnz = ctx.build_pattern()
    print(f"[{v or 'default'}] {et} {n}^3: pattern {time.perf_counter() - t0:.3f} s, nnz {nnz}", flush=True)
    for r in range(reps):
        s = ctx.assemble(2, _lib.ASM_K | _lib.ASM_M_LUMPED)
        print(f"   assemble {1e3 * s:.2f} ms  {ne / s / 1e6:.1f} Melem/s", flush=True)
    if n <= 128:
        k = ctx.get_values(_lib.MAT_K)
        if ref is None:
            ref = k
        else:
            print("   identical to the first variant:", bool(np.array_equal(ref, k)), flush=True)
    ctx.close()
