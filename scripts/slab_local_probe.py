"""One GPU, no communication: rank r's sub-domain of an N-slab partition of the 24^3 parity box run alone (ghost dofs stay
zero) through the stream-ordered PCG -- isolates the local kernels (FSAI on rows with ghost columns, ring kernels on thin
slabs) from the halo exchange."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from scatter_b200 import _lib, boxmesh, partition, system_matrix
S = 24
for world in (int(a) for a in (sys.argv[1:] or ["8"])):
    for rank in range(world):
        for opts in ({}, {"fsai": 0, "pcg_projection": 0}, {"fsai": 0, "spmv_groups": 1}, {"fsai": 0, "node_spmv": 0}, {"fsai": 0, "node_spmv": 0, "tma_spmv": 0}, {"fsai": 0, "pcg_graph": 0}):
            if rank not in (0, 1, world - 1): continue
            _lib.DEFAULT_OPTIONS.clear(); _lib.DEFAULT_OPTIONS.update(opts); _lib.DEFAULT_OPTIONS["small_pcg"] = 0
            per = S // world
            dom = partition.slab_partition(S, S, per, rank, world, bench.H, "hexa8")
            model = dom.model
            ne = len(model.elem)
            mx = system_matrix.GenerateMatrix(model.number_eq, 2)
            ctx = mx.ctx
            ctx.set_mesh("hexa8", model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), model.number_eq, dom.active)
            ctx.set_materials(np.full(ne, 30e6), np.full(ne, 0.2), np.full(ne, 1500.0))
            ctx.build_pattern(); ctx.assemble(2, _lib.ASM_K | _lib.ASM_M_FULL)
            mx.damping_Rayleigh(bench.DAMPING)
            nt = 12
            d = int(dom.owned_eq[len(dom.owned_eq) // 2])
            ctx.set_load_schedule(np.arange(nt + 1, dtype=np.int64), np.full(nt, d, dtype=np.int64), -1000.0 * np.minimum(1.0, np.arange(nt) / 4.0))
            ctx.set_state(None, None)
            try:
                u, v, a, st = ctx.run_newmark(5e-4, 0, 10, 5, rtol=1e-12)
                msg = f"ok its/step {st['pcg_iterations'] / 10:.1f} |u| {np.abs(u).max():.3e} precond {ctx.precond_info()}"
            except Exception as e:
                msg = "FAILED " + str(e)[:100]
            print(f"world {world} rank {rank} n_eq {model.number_eq} opts {opts}: {msg}", flush=True)
            ctx.close()
