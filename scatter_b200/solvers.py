"""Device time integrators with the solver protocol `scatter.scatter` relies on.

The reference instantiates a class from the external package PuggleSolvers==1.0.1 (`scatter/scatter.py:120-131`) and
uses exactly these members: zero-argument constructor, `.output_interval`, `.initialise(n_eq, time)`,
`.update_rhs_at_time_step_func`, `.update(t_start_idx)`, `.calculate(M, C, K, F, t_start_idx, t_end_idx)` and the
results `.u .v .a .time .output_time` (`scatter/export_results.py:52-55`).  The classes below keep those names; the
matrices stay on the GPU, so `calculate` takes them from the bound `GenerateMatrix` (`bind(matrix)`), and the M, C, K
arguments are accepted only for signature compatibility.  No CPU path exists: an unbound solver raises.

* `NewmarkExplicit` / `NewmarkImplicitForce`: incremental constant-average-acceleration Newmark (beta=1/4,
  gamma=1/2), effective-stiffness solve by Jacobi-PCG on the device (`sc_run_newmark`).  For a linear system the
  total-force form (`NewmarkImplicitForce`) is algebraically the same recurrence; both map to the same kernels.
* `CentralDifferenceSolver`: explicit central difference with row-sum lumped M and C (`sc_run_central_difference`).
"""
from __future__ import annotations

import numpy as np


class _DeviceSolver:
    def __init__(self):
        self.output_interval = 1
        self.u = self.v = self.a = None
        self.time = None
        self.output_time = None
        self.number_equations = None
        self.update_rhs_at_time_step_func = None
        self.update_rhs_at_non_linear_iteration_func = None
        self.matrix = None
        self.u0 = self.v0 = None
        self.stats = []
        self.load_schedule = None       # optional precompiled (step_ptr, dof, val)

    # ---- protocol -----------------------------------------------------------------------------------------------
    def initialise(self, number_equations, time):
        self.number_equations = int(number_equations)
        self.time = np.array(time)
        self.output_time = self.time[::self.output_interval]
        n_out = len(self.output_time)
        self.u = np.zeros((n_out, self.number_equations))
        self.v = np.zeros((n_out, self.number_equations))
        self.a = np.zeros((n_out, self.number_equations))
        self.u0 = np.zeros(self.number_equations)
        self.v0 = np.zeros(self.number_equations)

    def update(self, t_start_idx):
        """Restart hook (`scatter.py:158`): u0, v0 <- stored row.  The rows are taken as views, so the next `calculate`
        uploads them straight from the (possibly page-locked) history arrays without an extra host copy."""
        row = int(t_start_idx) // self.output_interval
        self.u0 = self.u[row]
        self.v0 = self.v[row]
        self._state_dirty = True

    def bind(self, matrix):
        """Attach the `GenerateMatrix` whose device matrices this solver integrates."""
        self.matrix = matrix
        return self

    # ---- helpers ------------------------------------------------------------------------------------------------
    def _ctx(self):
        if self.matrix is None:
            raise RuntimeError("solver is not bound to device matrices: call solver.bind(matrix) "
                               "(scatter_b200 has no CPU time-integration path)")
        return self.matrix.ctx

    def _dt(self, t0, t1):
        return float((self.time[t1] - self.time[t0]) / (t1 - t0))

    def _upload_loads(self, ctx):
        if self.load_schedule is not None:
            ptr, dof, val = self.load_schedule
        else:
            f = self.update_rhs_at_time_step_func
            owner = getattr(f, "__self__", None)
            if owner is not None and hasattr(owner, "compile_schedule"):
                ptr, dof, val = owner.compile_schedule()
            else:   # generic callback: evaluate once per step and keep the non-zeros
                ptr, dofs, vals = [0], [], []
                for t in range(len(self.time)):
                    vec = np.asarray(f(t))
                    nz = np.nonzero(vec)[0]
                    dofs.append(nz); vals.append(vec[nz]); ptr.append(ptr[-1] + len(nz))
                ptr, dof, val = np.array(ptr), np.concatenate(dofs), np.concatenate(vals)
            self.load_schedule = (ptr, dof, val)
        ctx.set_load_schedule(ptr, dof, val)

    def _out_views(self, t0):
        oi = self.output_interval
        first = -(-t0 // oi)
        return self.u[first:], self.v[first:], self.a[first:]


class NewmarkExplicit(_DeviceSolver):
    """Default solver of the reference (`Solver.NEWMARK_EXPLICIT`)."""
    beta = 0.25
    gamma = 0.5

    def __init__(self):
        super().__init__()
        self.pcg_rtol = 1e-14
        self.pcg_maxit = 20000

    def calculate(self, M, C, K, F, t_start_idx, t_end_idx):
        ctx = self._ctx()
        self._upload_loads(ctx)
        if getattr(self, "_state_dirty", True):      # a stage that continues the previous one keeps u, v, a on the device
            ctx.set_state(self.u0, self.v0)
            self._state_dirty = False
        n_steps = int(t_end_idx) - int(t_start_idx)
        uo, vo, ao = self._out_views(int(t_start_idx))
        _, _, _, st = ctx.run_newmark(self._dt(t_start_idx, t_end_idx), int(t_start_idx), n_steps, self.output_interval,
                                      self.beta, self.gamma, self.pcg_rtol, self.pcg_maxit, uo, vo, ao)
        self.stats.append(st)


class NewmarkImplicitForce(NewmarkExplicit):
    pass


class CentralDifferenceSolver(_DeviceSolver):
    def calculate(self, M, C, K, F, t_start_idx, t_end_idx):
        ctx = self._ctx()
        self._upload_loads(ctx)
        if getattr(self, "_state_dirty", True):
            ctx.set_state(self.u0, self.v0)
            self._state_dirty = False
        n_steps = int(t_end_idx) - int(t_start_idx)
        uo, vo, ao = self._out_views(int(t_start_idx))
        _, _, _, st = ctx.run_central_difference(self._dt(t_start_idx, t_end_idx), int(t_start_idx), n_steps,
                                                 self.output_interval, uo, vo, ao)
        self.stats.append(st)


class BatheSolver(_DeviceSolver):
    """Bathe composite scheme (trapezoidal rule over dt/2 + 3-point backward Euler); see csrc/timeloop.cu `tl_bathe`."""

    def __init__(self):
        super().__init__()
        self.pcg_rtol = 1e-14
        self.pcg_maxit = 20000

    def calculate(self, M, C, K, F, t_start_idx, t_end_idx):
        ctx = self._ctx()
        self._upload_loads(ctx)
        ctx.set_state(self.u0, self.v0)
        uo, vo, ao = self._out_views(int(t_start_idx))
        _, _, _, st = ctx.run_bathe(self._dt(t_start_idx, t_end_idx), int(t_start_idx), int(t_end_idx) - int(t_start_idx),
                                    self.output_interval, self.pcg_rtol, self.pcg_maxit, uo, vo, ao)
        self.stats.append(st)


class StaticSolver(_DeviceSolver):
    """K u(t) = F(t) at every step (`calculate(K, F, t_start_idx, t_end_idx)`, scatter.py:156)."""

    def __init__(self):
        super().__init__()
        self.pcg_rtol = 1e-12
        self.pcg_maxit = 200000

    def calculate(self, K, F, t_start_idx, t_end_idx):
        ctx = self._ctx()
        self._upload_loads(ctx)
        ctx.set_state(self.u0, None)
        uo, _, _ = self._out_views(int(t_start_idx))
        _, st = ctx.run_static(int(t_start_idx), int(t_end_idx) - int(t_start_idx), self.output_interval, self.pcg_rtol,
                               self.pcg_maxit, uo)
        self.stats.append(st)
