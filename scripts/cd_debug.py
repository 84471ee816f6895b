"""GPU debugging aid: central difference on cube.msh step by step against the oracle, for each SpMV kernel family."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fem_np as oracle
import scipy.sparse as sp
from scatter_b200 import _lib, boxmesh, system_matrix
s = 6
model = boxmesh.box_model(s, s, s, 0.5, "hexa8"); model.connectivities()
ne, n = len(model.elem), model.number_eq
E = boxmesh.lognormal_young(ne)
om = oracle.model_from_readmesh(model)
Ko, Mo = oracle.assemble_global(om, E, np.full(ne, 0.2), np.full(ne, 1500.0), 2)
Ko = sp.csr_matrix(Ko); ml = oracle.lump_rows(Mo)
d = int(model.eq_nb_dof[boxmesh.top_centre_node(s, s, s) - 1, 1])
nt = 12
def force(t):
    f = np.zeros(n); f[d] = -1000.0 * min(1.0, t / 4.0); return f
sched = (np.arange(nt + 1, dtype=np.int64), np.full(nt, d, dtype=np.int64), np.array([-1000.0 * min(1.0, t / 4.0) for t in range(nt)]))
dt = 5e-4
for damping in ([1, 0.0, 30, 0.0], [1, 0.01, 30, 0.01]):
    c0, c1 = oracle.rayleigh_coefficients(damping)
    U, V, A, _ = oracle.central_difference(sp.diags(ml), sp.diags(ml) * c0 + Ko * c1, Ko, force, np.arange(nt) * dt, 1, c1=c1)
    for opts in ({}, {"node_spmv": 0}, {"node_spmv": 0, "tma_spmv": 0}):
        _lib.DEFAULT_OPTIONS.clear(); _lib.DEFAULT_OPTIONS.update(opts)
        mx = system_matrix.GenerateMatrix(n, 2)
        mx.want_full_mass, mx.want_lumped_mass = False, True
        mx.generate_stiffness_and_mass(model, None, elem_props=(E, np.full(ne, 0.2), np.full(ne, 1500.0)))
        mx.damping_Rayleigh(damping)
        ctx = mx.ctx
        ctx.set_load_schedule(*sched); ctx.set_state(None, None)
        try:
            u, v, a, st = ctx.run_central_difference(dt, 0, nt - 1, 1)
            err = [float(np.linalg.norm(u[k] - U[k]) / max(np.linalg.norm(U[k]), 1e-300)) for k in range(nt)]
            print(damping, opts, "row errors", ["%.1e" % e for e in err])
        except Exception as exc:
            print(damping, opts, "FAILED", exc)
            uu = ctx.get_state()[0]
            print("   nonfinite entries:", int((~np.isfinite(uu)).sum()), "of", n)
        ctx.close()
