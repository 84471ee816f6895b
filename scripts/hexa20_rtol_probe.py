"""GPU experiment: which PCG tolerance keeps a 12^3 hexa20 Newmark history within 1e-8 of the
direct-solve oracle over 1000 steps?  Uses the stream-ordered PCG driver (the one the 94^3 benchmark box runs), not the
cooperative small-system kernel.  Prints one line per tolerance."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fem_np as oracle                                   # checker
from scatter_b200 import _lib, boxmesh, system_matrix

s = int(sys.argv[1]) if len(sys.argv) > 1 else 12
nst = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
H, dt = 0.5, 5e-4
pm = boxmesh.box_model(s, s, s, H, "hexa20"); pm.connectivities()
ne, n = len(pm.elem), pm.number_eq
E = boxmesh.lognormal_young(ne, 30e6, 1e6, seed=20)
Ko, Mo = oracle.assemble_global(oracle.model_from_readmesh(pm), E, np.full(ne, 0.2), np.full(ne, 1500.0), 2)
c0, c1 = oracle.rayleigh_coefficients([1, 0.01, 30, 0.01])
d = int(pm.eq_nb_dof[boxmesh.top_centre_node(s, s, s) - 1, 1])


def force(t):
    f = np.zeros(n); f[d] = -1000.0 * min(1.0, t / 4.0)
    return f


t0 = time.perf_counter()
Uo, Vo, _, _ = oracle.newmark(Mo, Mo * c0 + Ko * c1, Ko, force, np.arange(nst + 1) * dt, 50)
print(f"oracle: {n} dof, {nst} steps in {time.perf_counter() - t0:.1f} s", flush=True)
for rtol in (1e-10, 1e-11, 1e-12, 1e-13, 1e-14):
    mx = system_matrix.GenerateMatrix(n, 2)
    ctx = mx.ctx
    ctx.set_option("small_pcg", 0)
    ctx.set_mesh("hexa20", pm.nodes[:, 1:], pm.node_rows(), pm.equation_table_int(), n, None)
    ctx.set_materials(E, np.full(ne, 0.2), np.full(ne, 1500.0))
    ctx.build_pattern()
    ctx.assemble(2, _lib.ASM_K | _lib.ASM_M_FULL)
    mx.damping_Rayleigh([1, 0.01, 30, 0.01])
    ctx.set_load_schedule(np.arange(nst + 2, dtype=np.int64), np.full(nst + 1, d, dtype=np.int64), -1000.0 * np.minimum(1.0, np.arange(nst + 1) / 4.0))
    ctx.set_state(None, None)
    u, v, _, st = ctx.run_newmark(dt, 0, nst, 50, rtol=rtol)
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    rows = [rel(u[k], Uo[k]) for k in (1, len(u) // 2, len(u) - 1)]
    print(f"rtol {rtol:g}: u rel L2 {rel(u, Uo):.2e}  v rel L2 {rel(v, Vo):.2e}  rows(first/mid/last) {rows[0]:.1e} {rows[1]:.1e} {rows[2]:.1e}  "
          f"iterations/step {st['pcg_iterations'] / nst:.1f}  stagnations {st['pcg_stagnations']}", flush=True)
    ctx.close()
