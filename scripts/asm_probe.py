"""Assemble a hexa8 box a few times (ncu target / quick timing).  usage: asm_probe.py [n=128] [reps=3]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from scatter_b200 import _lib, boxmesh

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
model = boxmesh.box_model(n, n, n, 0.5, "hexa8")
ne = model.elem.shape[0]
ctx = _lib.Context(0)
ctx.set_mesh("hexa8", model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), model.number_eq, None)
ctx.set_materials(boxmesh.lognormal_young(ne), np.full(ne, 0.2), np.full(ne, 1500.0))
t0 = time.perf_counter()
nnz = ctx.build_pattern()
print("pattern", time.perf_counter() - t0, "nnz", nnz)
for r in range(reps):
    s = ctx.assemble(2, _lib.ASM_K | _lib.ASM_M_LUMPED)
    print("assemble", s, "s", ne / s / 1e6, "Melem/s")
