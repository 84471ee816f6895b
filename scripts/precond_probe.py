"""GPU probe of the implicit solve's preconditioner / initial-guess options.
  part 1: 12^3 hexa20 Newmark history vs the oracle's direct solve for every option set (stream-ordered PCG driver)
  part 2: 94^3 hexa20 (BASELINE config 4) iterations and ms per step for every option set
    python scripts/precond_probe.py [size] [steps] "opt=v,opt=v" ..."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench
from scatter_b200 import _lib, boxmesh, system_matrix
s = int(sys.argv[1]) if len(sys.argv) > 1 else 94
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
et = sys.argv[3] if len(sys.argv) > 3 else "hexa20"
variants = sys.argv[4:] or ["fsai=0,pcg_projection=0", "pcg_projection=0", ""]
dt, rtol = 5e-4, 1e-12


def setup(model, E, opts):
    _lib.DEFAULT_OPTIONS.clear()
    for kv in filter(None, opts.split(",")):
        k, val = kv.split("="); _lib.DEFAULT_OPTIONS[k] = int(val)
    mx = system_matrix.GenerateMatrix(model.number_eq, 2)
    ctx = mx.ctx
    ctx.set_option("small_pcg", 0)
    ne = len(model.elem)
    ctx.set_mesh(et, model.nodes[:, 1:], model.node_rows(), model.equation_table_int(), model.number_eq, None)
    ctx.set_materials(E, np.full(ne, bench.NU), np.full(ne, bench.RHO))
    ctx.build_pattern(); ctx.assemble(2, _lib.ASM_K | _lib.ASM_M_FULL)
    mx.damping_Rayleigh(bench.DAMPING)
    return mx, ctx


if os.environ.get("PARITY", "1") == "1":
    import fem_np as oracle
    sp_, nst = int(os.environ.get("PSIZE", "12")), int(os.environ.get("PSTEPS", "60"))
    pm = boxmesh.box_model(sp_, sp_, sp_, bench.H, et, hexa20_order=os.environ.get("ORDER20", "grouped")); pm.connectivities()
    pne, pn = len(pm.elem), pm.number_eq
    pE = boxmesh.lognormal_young(pne, bench.E_MEAN, bench.E_STD, seed=20)
    Ko, Mo = oracle.assemble_global(oracle.model_from_readmesh(pm), pE, np.full(pne, bench.NU), np.full(pne, bench.RHO), 2)
    c0, c1 = oracle.rayleigh_coefficients(bench.DAMPING)
    pd = int(pm.eq_nb_dof[boxmesh.top_centre_node(sp_, sp_, sp_, model=pm, h=bench.H) - 1, 1])

    def force(t):
        f = np.zeros(pn); f[pd] = -1000.0 * min(1.0, t / 4.0)
        return f
    Uo, Vo, _, _ = oracle.newmark(Mo, Mo * c0 + Ko * c1, Ko, force, np.arange(nst + 1) * dt, 5)
    for v in variants:
        mx, ctx = setup(pm, pE, v)
        ctx.set_load_schedule(np.arange(nst + 2, dtype=np.int64), np.full(nst + 1, pd, dtype=np.int64), -1000.0 * np.minimum(1.0, np.arange(nst + 1) / 4.0))
        ctx.set_state(None, None)
        u, vv, _, st = ctx.run_newmark(dt, 0, nst, 5, rtol=rtol)
        rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
        print(f"parity 12^3 [{v or 'default'}]: u {rel(u, Uo):.2e} v {rel(vv, Vo):.2e}, {st['pcg_iterations'] / nst:.1f} it/step, "
              f"precond {ctx.precond_info()}", flush=True)
        ctx.close()

order20 = os.environ.get("ORDER20", "grouped")
model = boxmesh.box_model(s, s, s, bench.H, et, hexa20_order=order20)
ne = len(model.elem)
E = boxmesh.lognormal_young(ne, bench.E_MEAN, bench.E_STD)
for v in variants:
    mx, ctx = setup(model, E, v)
    total = nsteps + 12
    d = int(model.eq_nb_dof[boxmesh.top_centre_node(s, s, s, model=model, h=bench.H) - 1, 1])
    ramp = np.ones(total); ramp[:5] = np.linspace(0, 1, 5)
    ctx.set_load_schedule(np.arange(total + 1, dtype=np.int64), np.full(total, d, dtype=np.int64), -1000.0 * ramp)
    ctx.set_state(None, None)
    t0 = time.perf_counter()
    _, _, _, st0 = ctx.run_newmark(dt, 0, 2, 1, rtol=rtol, store=False)
    t_first = time.perf_counter() - t0
    _, _, _, st = ctx.run_newmark(dt, 2, nsteps, 1, rtol=rtol, store=False)
    its = st["pcg_iterations"] / nsteps
    print(f"{et} {s}^3 [{v or 'default'}] {model.number_eq} dof: {its:.1f} it/step, {1e3 * st['seconds_device'] / nsteps:.1f} ms/step, "
          f"{1e3 * st['seconds_device'] / max(st['pcg_iterations'], 1):.3f} ms/iteration, first call {t_first:.2f} s "
          f"(fsai set-up {st0['fsai_setup_seconds']:.3f} s), precond {ctx.precond_info()}, last residual {st['last_residual']:.2e}, "
          f"numbering {order20}, dict patterns {ctx.pattern_stats().get('dict_patterns')}", flush=True)
    u = ctx.get_state()[0]
    print("   checksum |u|", float(np.abs(u).sum()), flush=True)
    ctx.close()
