import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fem_np as oracle
from scatter_b200 import _lib, boxmesh, system_matrix

for s in [int(a) for a in sys.argv[1:]] or [4, 8, 16]:
    model = boxmesh.box_model(s, s, s, 0.5, "hexa20")
    model.connectivities()
    ne = len(model.elem)
    E = boxmesh.lognormal_young(ne, 30e6, 1e6)
    mx = system_matrix.GenerateMatrix(model.number_eq, 2)
    ctx = mx.ctx
    ctx.set_mesh("hexa20", model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), model.number_eq, None)
    ctx.set_materials(E, np.full(ne, 0.2), np.full(ne, 1500.0))
    nnz = ctx.build_pattern()
    ctx.assemble(2, 3)
    mx.damping_Rayleigh([1, 0.01, 30, 0.01])
    n = model.number_eq
    x = np.random.default_rng(0).standard_normal(n)
    msg = f"size {s}: n_eq {n} nnz {nnz}"
    if s <= 10:
        om = oracle.model_from_readmesh(model)
        K, M = oracle.assemble_global(om, E, np.full(ne, 0.2), np.full(ne, 1500.0), 2)
        kx = ctx.spmv(0, x); mxv = ctx.spmv(1, x)
        msg += f" | K.x err {np.abs(kx - K @ x).max() / np.abs(K @ x).max():.2e} M.x err {np.abs(mxv - M @ x).max() / np.abs(M @ x).max():.2e}"
    kx = ctx.spmv(0, x); y = np.random.default_rng(1).standard_normal(n); ky = ctx.spmv(0, y)
    mxx = ctx.spmv(1, x)
    msg += f" | sym {abs(y @ kx - x @ ky) / abs(y @ kx):.2e} xKx {x @ kx:.3e} xMx {x @ mxx:.3e}"
    d = int(model.eq_nb_dof[boxmesh.top_centre_node(s, s, s) - 1, 1])
    ctx.set_load_schedule(np.arange(11, dtype=np.int64), np.full(10, d, dtype=np.int64), -1000.0 * np.minimum(np.arange(10) / 4.0, 1.0))
    ctx.set_state(None, None)
    try:
        t0 = time.time()
        _, _, _, st = ctx.run_newmark(5e-4, 0, 3, 1, rtol=1e-10, maxit=2000, store=False)
        _, _, _, st = ctx.run_newmark(5e-4, 3, 3, 1, rtol=1e-10, maxit=2000, store=False)
        msg += f" | newmark ok: its/step {st['pcg_iterations'] / 3:.1f} res {st['last_residual']:.2e} {time.time() - t0:.2f}s"
    except Exception as e:
        msg += f" | newmark FAILED: {e}"
    print(msg, flush=True)
    ctx.close()
