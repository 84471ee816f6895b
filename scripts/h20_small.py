import os, sys
import numpy as np, scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scatter_b200 import _lib, boxmesh, system_matrix
s = int(sys.argv[1]) if len(sys.argv) > 1 else 8
et = sys.argv[2] if len(sys.argv) > 2 else "hexa20"
model = boxmesh.box_model(s, s, s, 0.5, et)
ne = len(model.elem)
for groups in (1, 2):
    mx = system_matrix.GenerateMatrix(model.number_eq, 2)
    ctx = mx.ctx
    ctx.set_option("small_pcg", 0); ctx.set_option("spmv_groups", groups)
    ctx.set_mesh(et, model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), model.number_eq, None)
    ctx.set_materials(np.full(ne, 30e6), np.full(ne, 0.2), np.full(ne, 1500.0))
    ctx.build_pattern(); ctx.assemble(2, _lib.ASM_K | _lib.ASM_M_FULL)
    n = model.number_eq
    rp, col = ctx.get_pattern()
    K = sp.csr_matrix((ctx.get_values(_lib.MAT_K), col, rp), shape=(n, n))
    x = np.sin(0.37 * np.arange(n) + 0.11)
    y = ctx.spmv(_lib.MAT_K, x)
    ref = K @ x
    bad = np.where(np.abs(y - ref) > 1e-9 * np.abs(ref).max())[0]
    print(f"groups {groups}: stats {ctx.pattern_stats()['dict_patterns']} dict, spmv max err {np.abs(y - ref).max() / np.abs(ref).max():.2e}, bad rows {len(bad)} first {bad[:8]}")
    mx.damping_Rayleigh([1, 0.01, 30, 0.01])
    d = int(model.eq_nb_dof[boxmesh.top_centre_node(s, s, s) - 1, 1])
    ctx.set_load_schedule(np.arange(6, dtype=np.int64), np.full(5, d, dtype=np.int64), np.full(5, -1000.0))
    ctx.set_state(None, None)
    try:
        u, v, a, st = ctx.run_newmark(5e-4, 0, 2, 1, rtol=1e-10, maxit=2000)
        print("   newmark ok", st["pcg_iterations"], float(np.abs(u).max()))
    except Exception as e:
        print("   newmark FAILED", str(e)[:120])
    ctx.close()
