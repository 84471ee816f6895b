"""hexa20 / tetra-free assembly timing: asm_probe20.py [n=48] (K + full M, order 2)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from scatter_b200 import _lib, boxmesh
n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
model = boxmesh.box_model(n, n, n, 0.5, "hexa20")
ne = model.elem.shape[0]
ctx = _lib.Context(0)
ctx.set_mesh("hexa20", model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), model.number_eq, None)
ctx.set_materials(boxmesh.lognormal_young(ne), np.full(ne, 0.2), np.full(ne, 1500.0))
nnz = ctx.build_pattern()
for order in (2, 3):
    t = min(ctx.assemble(order, _lib.ASM_K | _lib.ASM_M_FULL) for _ in range(3))
    print(f"hexa20 {n}^3 order {order}: {1e3*t:.2f} ms, {ne/t/1e6:.2f} Melem/s, nnz {nnz}")
