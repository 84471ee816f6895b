"""ctypes binding of libscatter_b200.so (C ABI declared in include/scatter_b200.h).

The library is the product: there is no CPU fallback anywhere in this package.  If the shared object is missing or no
CUDA device is usable, the calls below raise -- loudly -- instead of computing something else.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscatter_b200.so")

# enum sc_elem_type
ELEM_TYPE_ID = {"tri3": 0, "tri6": 1, "quad4": 2, "quad8": 3, "tetra4": 4, "tetra10": 5, "hexa8": 6, "hexa20": 7}
ELEM_NNE = {"tri3": 3, "tri6": 6, "quad4": 4, "quad8": 8, "tetra4": 4, "tetra10": 10, "hexa8": 8, "hexa20": 20}
ELEM_DIM = {"tri3": 2, "tri6": 2, "quad4": 2, "quad8": 2, "tetra4": 3, "tetra10": 3, "hexa8": 3, "hexa20": 3}
MAT_K, MAT_M, MAT_C, MAT_KHAT = 0, 1, 2, 3
ASM_K, ASM_M_FULL, ASM_M_LUMPED = 1, 2, 4

# every symbol include/scatter_b200.h declares (checked by tests/test_host_logic.py)
SYMBOLS = [
    "sc_create", "sc_destroy", "sc_last_error", "sc_version", "sc_device_info", "sc_kernel_launches", "sc_precond_info", "sc_set_option", "sc_shape_table",
    "sc_host_alloc", "sc_host_free",
    "sc_set_mesh", "sc_set_csr", "sc_set_output_dofs", "sc_set_materials", "sc_build_pattern", "sc_get_pattern", "sc_pattern_stats", "sc_assemble", "sc_add_entries",
    "sc_set_rayleigh", "sc_get_values", "sc_get_lumped_mass", "sc_spmv", "sc_set_load_schedule", "sc_set_state",
    "sc_get_state", "sc_set_final_output_step", "sc_run_newmark", "sc_run_central_difference", "sc_run_bathe", "sc_run_static", "sc_nccl_unique_id", "sc_dist_init", "sc_set_halo",
    "sc_halo_exchange", "sc_srf_sample", "sc_add_absorbing_faces",
]


class Stats(C.Structure):
    _fields_ = [("seconds_total", C.c_double), ("seconds_device", C.c_double), ("seconds_halo", C.c_double),
                ("steps", C.c_int64), ("pcg_iterations", C.c_int64), ("kernel_launches", C.c_int64),
                ("last_residual", C.c_double), ("reserved", C.c_double * 4)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}
        d["step_bytes"] = self.reserved[0]
        d["pcg_stagnations"] = int(self.reserved[2])
        d["fsai_setup_seconds"] = self.reserved[3]         # implicit loops
        d["halo_overlapped"] = bool(self.reserved[3])      # central difference on several GPUs: exchange ran beside the interior tiles
        d["step_kernel"] = {0: "k_spmv (register staged)", 1: "k_spmv_tma (row tiles)", 2: "k_spmv_node (node-blocked, TMA)"}.get(int(self.reserved[1]), "?")
        return d


class ScatterB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libscatter_b200 error {code}: {message}")
        self.code = code


_lib = None


def load_library():
    """dlopen the in-tree shared object (built by `__graft_entry__.build()` / scatter_b200/csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ScatterB200Error(-100, f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                     "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    P = C.POINTER
    lib.sc_create.argtypes = [i32, P(vp)]
    lib.sc_destroy.argtypes = [vp]
    lib.sc_destroy.restype = None
    lib.sc_last_error.argtypes = [vp]
    lib.sc_last_error.restype = C.c_char_p
    lib.sc_version.argtypes = []
    lib.sc_device_info.argtypes = [vp, P(i32), P(i64), P(i64), C.c_char_p, i32]
    lib.sc_kernel_launches.argtypes = [vp]
    lib.sc_kernel_launches.restype = i64
    lib.sc_precond_info.argtypes = [vp, i32, P(i64), P(dbl), P(i32)]
    lib.sc_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.sc_shape_table.argtypes = [i32, i32, P(i32), P(i32), P(i32), vp, vp, vp]
    lib.sc_host_alloc.argtypes = [P(vp), i64]
    lib.sc_host_free.argtypes = [vp]
    lib.sc_set_mesh.argtypes = [vp, i32, i64, vp, i64, vp, vp, i64, vp]
    lib.sc_set_csr.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    lib.sc_set_output_dofs.argtypes = [vp, i64, vp]
    lib.sc_set_materials.argtypes = [vp, vp, vp, vp]
    lib.sc_build_pattern.argtypes = [vp, P(i64)]
    lib.sc_get_pattern.argtypes = [vp, vp, vp]
    lib.sc_pattern_stats.argtypes = [vp, vp]
    lib.sc_assemble.argtypes = [vp, i32, i32, P(dbl)]
    lib.sc_add_entries.argtypes = [vp, i32, i64, vp, vp, vp]
    lib.sc_set_rayleigh.argtypes = [vp, dbl, dbl]
    lib.sc_get_values.argtypes = [vp, i32, vp]
    lib.sc_get_lumped_mass.argtypes = [vp, vp]
    lib.sc_spmv.argtypes = [vp, i32, vp, vp]
    lib.sc_set_load_schedule.argtypes = [vp, i64, vp, vp, vp]
    lib.sc_set_state.argtypes = [vp, vp, vp]
    lib.sc_get_state.argtypes = [vp, vp, vp, vp]
    lib.sc_set_final_output_step.argtypes = [vp, i64]
    lib.sc_run_newmark.argtypes = [vp, dbl, i64, i64, i64, dbl, dbl, dbl, i32, i64, vp, vp, vp, P(Stats)]
    lib.sc_run_central_difference.argtypes = [vp, dbl, i64, i64, i64, i64, vp, vp, vp, P(Stats)]
    lib.sc_run_bathe.argtypes = [vp, dbl, i64, i64, i64, dbl, i32, i64, vp, vp, vp, P(Stats)]
    lib.sc_run_static.argtypes = [vp, i64, i64, i64, dbl, i32, i64, vp, P(Stats)]
    lib.sc_nccl_unique_id.argtypes = [vp]
    lib.sc_dist_init.argtypes = [vp, i32, i32, vp]
    lib.sc_set_halo.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.sc_halo_exchange.argtypes = [vp, vp]
    lib.sc_add_absorbing_faces.argtypes = [vp, i32, i32, i64, vp, vp, vp, vp, i64, vp, vp, vp, vp, dbl, dbl, dbl]
    lib.sc_srf_sample.argtypes = [vp, i64, vp, i32, vp, vp, vp, dbl, dbl, i32, vp, P(dbl)]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("sc_destroy", "sc_last_error", "sc_kernel_launches"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def shape_table(elem_type: str, order: int):
    """(N[ngp,nne], dN[ngp,nne,dim], w[ngp]) from the library's host-side tables (no GPU needed)."""
    lib = load_library()
    nne, dim, ngp = C.c_int(), C.c_int(), C.c_int()
    rc = lib.sc_shape_table(ELEM_TYPE_ID[elem_type], order, C.byref(nne), C.byref(dim), C.byref(ngp), None, None, None)
    if rc != 0:
        raise ScatterB200Error(rc, lib.sc_last_error(None).decode())
    N = np.zeros((ngp.value, nne.value)); dN = np.zeros((ngp.value, nne.value, dim.value)); w = np.zeros(ngp.value)
    rc = lib.sc_shape_table(ELEM_TYPE_ID[elem_type], order, None, None, None, _ptr(N), _ptr(dN), _ptr(w))
    if rc != 0:
        raise ScatterB200Error(rc, lib.sc_last_error(None).decode())
    return N, dN, w


class _PinnedOwner:
    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        try:
            self.lib.sc_host_free(self.ptr)
        except Exception:
            pass


def pinned_zeros(shape, dtype=np.float64) -> np.ndarray:
    """numpy array backed by page-locked host memory (cudaMallocHost) -- needs a CUDA device."""
    lib = load_library()
    shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    rc = lib.sc_host_alloc(C.byref(p), nbytes)
    if rc != 0:
        raise ScatterB200Error(rc, lib.sc_last_error(None).decode())
    buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    arr[...] = 0
    _PINNED[id(buf)] = (_PinnedOwner(lib, p), buf)
    return arr


_PINNED = {}


def nccl_unique_id() -> bytes:
    lib = load_library()
    buf = C.create_string_buffer(128)
    rc = lib.sc_nccl_unique_id(buf)
    if rc != 0:
        raise ScatterB200Error(rc, lib.sc_last_error(None).decode())
    return buf.raw


# options applied to every new Context (`sc_set_option`); empty in production, filled by tests / A-B measurements
DEFAULT_OPTIONS: dict = {}


class Context:
    """One GPU context (`sc_ctx`).  Thin, 1:1 over the C ABI; all arrays are numpy, all heavy work is on the device."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.sc_create(int(device), C.byref(h))
        if rc != 0:
            raise ScatterB200Error(rc, self.lib.sc_last_error(None).decode())
        self.h = h
        self.device = device
        self.n_eq = 0
        self.nnz = 0
        self.elem_type = None
        for name, value in DEFAULT_OPTIONS.items():
            self.set_option(name, value)
        self.state_epoch = 0            # bumped by everything that changes u, v, a on the device (see solvers._prepare)
        self.final_output_step = -1
        self.output_dofs = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.sc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise ScatterB200Error(rc, self.lib.sc_last_error(self.h).decode())

    # ---- info
    def device_info(self):
        sm, tot, free = C.c_int(), C.c_int64(), C.c_int64()
        name = C.create_string_buffer(256)
        self._ck(self.lib.sc_device_info(self.h, C.byref(sm), C.byref(tot), C.byref(free), name, 256))
        return {"sm_count": sm.value, "total_mem": tot.value, "free_mem": free.value, "name": name.value.decode()}

    def kernel_launches(self) -> int:
        return int(self.lib.sc_kernel_launches(self.h))

    def precond_info(self, slot: int = 0):
        """FSAI factor / projection basis currently held for the effective matrix (`sc_precond_info`)."""
        nnz, sec, nv = C.c_int64(), C.c_double(), C.c_int()
        self._ck(self.lib.sc_precond_info(self.h, slot, C.byref(nnz), C.byref(sec), C.byref(nv)))
        return {"fsai_nnz": nnz.value, "fsai_seconds": sec.value, "projection_vectors": nv.value}

    def set_option(self, name: str, value: int):
        """Kernel-selection switch (tests / A-B measurements); see `sc_set_option` in include/scatter_b200.h."""
        self._ck(self.lib.sc_set_option(self.h, name.encode(), int(value)))

    # ---- mesh / matrices
    def set_mesh(self, elem_type: str, xyz, conn, eq, n_eq: int, active=None):
        xyz = _arr(xyz, np.float64); conn = _arr(conn, np.int32); eq = _arr(eq, np.int64)
        active = _arr(active, np.uint8)
        if xyz.ndim != 2 or xyz.shape[1] != 3:
            raise ValueError("xyz must be (n_nodes, 3)")
        self._ck(self.lib.sc_set_mesh(self.h, ELEM_TYPE_ID[elem_type], xyz.shape[0], _ptr(xyz), conn.shape[0], _ptr(conn),
                                      _ptr(eq), int(n_eq), _ptr(active)))
        self.n_eq = int(n_eq)
        self.elem_type = elem_type
        self.state_epoch += 1
        self.output_dofs = None
        self.n_elem = conn.shape[0]
        self.n_nodes = xyz.shape[0]

    def set_csr(self, M, C, K):
        """Upload caller-supplied scipy matrices (any sparse format) on the union of their patterns (`sc_set_csr`)."""
        import scipy.sparse as sp
        mats = [None if m is None else sp.csr_matrix(m) for m in (M, C, K)]
        if mats[2] is None:
            raise ValueError("K is required")
        n = mats[2].shape[0]
        keys = []
        for m in mats:
            if m is None:
                continue
            if m.shape != (n, n):
                raise ValueError("M, C, K must be square matrices of equal shape")
            m.sum_duplicates()
            rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(m.indptr))
            keys.append(rows * n + m.indices.astype(np.int64))
        keys.append(np.arange(n, dtype=np.int64) * (n + 1))             # the diagonal is always stored (Jacobi scaling)
        union = np.unique(np.concatenate(keys))
        rowptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.bincount(union // n, minlength=n), out=rowptr[1:])
        col = (union % n).astype(np.int32)

        def values(m):
            if m is None:
                return None
            rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(m.indptr))
            v = np.zeros(len(union))
            v[np.searchsorted(union, rows * n + m.indices.astype(np.int64))] = m.data
            return v
        vM, vC, vK = (values(m) for m in mats)
        self._ck(self.lib.sc_set_csr(self.h, n, _ptr(rowptr), _ptr(col), _ptr(vK), _ptr(vM), _ptr(vC)))
        self.n_eq, self.nnz, self.elem_type = n, int(len(union)), None
        self.state_epoch += 1

    def set_materials(self, young, poisson, density):
        E = _arr(np.broadcast_to(young, (self.n_elem,)), np.float64)
        nu = _arr(np.broadcast_to(poisson, (self.n_elem,)), np.float64)
        rho = _arr(np.broadcast_to(density, (self.n_elem,)), np.float64)
        self._ck(self.lib.sc_set_materials(self.h, _ptr(E), _ptr(nu), _ptr(rho)))

    def build_pattern(self) -> int:
        nnz = C.c_int64()
        self._ck(self.lib.sc_build_pattern(self.h, C.byref(nnz)))
        self.nnz = nnz.value
        return self.nnz

    def get_pattern(self):
        rowptr = np.empty(self.n_eq + 1, dtype=np.int64)
        col = np.empty(self.nnz, dtype=np.int32)
        self._ck(self.lib.sc_get_pattern(self.h, _ptr(rowptr), _ptr(col)))
        return rowptr, col

    def pattern_stats(self) -> dict:
        out = np.zeros(8, dtype=np.int64)
        self._ck(self.lib.sc_pattern_stats(self.h, _ptr(out)))
        keys = ("nnz", "node_col_entries", "n_nodes", "max_row_len", "max_node_neighbours", "max_elems_per_node", "node_blocked", "dict_patterns")
        return {k: int(v) for k, v in zip(keys, out)}

    def assemble(self, order: int, flags: int = ASM_K | ASM_M_FULL) -> float:
        sec = C.c_double()
        self._ck(self.lib.sc_assemble(self.h, int(order), int(flags), C.byref(sec)))
        return sec.value

    def add_entries(self, which: int, rows, cols, vals):
        rows = _arr(rows, np.int64); cols = _arr(cols, np.int64); vals = _arr(vals, np.float64)
        self._ck(self.lib.sc_add_entries(self.h, which, len(vals), _ptr(rows), _ptr(cols), _ptr(vals)))

    def add_absorbing_faces(self, plan, order: int, p0: float, p1: float, stiff: float):
        """Evaluate an `system_matrix.AbsorbingPlan` on the device: C_abs and the spring terms of K."""
        self._ck(self.lib.sc_add_absorbing_faces(self.h, ELEM_TYPE_ID[plan.face_type], int(order), len(plan.elem), _ptr(plan.nodes),
                                                 _ptr(plan.elem), _ptr(plan.direction), _ptr(plan.perp), len(plan.rows), _ptr(plan.rows),
                                                 _ptr(plan.cols), _ptr(plan.grp_ptr), _ptr(plan.grp_entry), float(p0), float(p1), float(stiff)))

    def set_rayleigh(self, c0: float, c1: float):
        self._ck(self.lib.sc_set_rayleigh(self.h, float(c0), float(c1)))

    def get_values(self, which: int):
        v = np.empty(self.nnz, dtype=np.float64)
        self._ck(self.lib.sc_get_values(self.h, which, _ptr(v)))
        return v

    def get_lumped_mass(self):
        v = np.empty(self.n_eq, dtype=np.float64)
        self._ck(self.lib.sc_get_lumped_mass(self.h, _ptr(v)))
        return v

    def spmv(self, which: int, x):
        x = _arr(x, np.float64)
        y = np.empty(self.n_eq, dtype=np.float64)
        self._ck(self.lib.sc_spmv(self.h, which, _ptr(x), _ptr(y)))
        return y

    def srf_sample(self, pos, k, z1, z2, scale: float, mean: float, lognormal: bool):
        """Random field at `pos` (n,3) from wave vectors k (m,3) and amplitudes z1, z2 (m,) -> (values (n,), kernel seconds)."""
        pos = _arr(pos, np.float64); k = _arr(k, np.float64); z1 = _arr(z1, np.float64); z2 = _arr(z2, np.float64)
        if pos.ndim != 2 or pos.shape[1] != 3 or k.shape != (len(z1), 3) or z2.shape != z1.shape:
            raise ValueError("pos must be (n,3), k (m,3), z1 and z2 (m,)")
        out = np.empty(pos.shape[0], dtype=np.float64)
        sec = C.c_double()
        self._ck(self.lib.sc_srf_sample(self.h, pos.shape[0], _ptr(pos), len(z1), _ptr(k), _ptr(z1), _ptr(z2), float(scale),
                                        float(mean), int(bool(lognormal)), _ptr(out), C.byref(sec)))
        return out, sec.value

    # ---- loads / state / time loop
    def set_load_schedule(self, step_ptr, dof, val):
        step_ptr = _arr(step_ptr, np.int64); dof = _arr(dof, np.int64); val = _arr(val, np.float64)
        self._ck(self.lib.sc_set_load_schedule(self.h, len(step_ptr) - 1, _ptr(step_ptr), _ptr(dof), _ptr(val)))

    def set_state(self, u=None, v=None):
        u = _arr(u, np.float64); v = _arr(v, np.float64)
        self._ck(self.lib.sc_set_state(self.h, _ptr(u), _ptr(v)))
        self.state_epoch += 1

    def set_output_dofs(self, dofs=None):
        """Only these equations are copied to the host per output row (None: full rows of n_eq values)."""
        if dofs is None:
            if self.output_dofs is not None:
                self._ck(self.lib.sc_set_output_dofs(self.h, 0, None))
            self.output_dofs = None
            return
        if dofs is self.output_dofs:
            return
        dofs = _arr(dofs, np.int64)
        self._ck(self.lib.sc_set_output_dofs(self.h, len(dofs), _ptr(dofs) if len(dofs) else _ptr(np.zeros(1, dtype=np.int64))))
        self.output_dofs = dofs

    def _row_len(self):
        return self.n_eq if self.output_dofs is None else len(self.output_dofs)

    def get_state(self):
        u = np.empty(self.n_eq); v = np.empty(self.n_eq); a = np.empty(self.n_eq)
        self._ck(self.lib.sc_get_state(self.h, _ptr(u), _ptr(v), _ptr(a)))
        return u, v, a

    def set_final_output_step(self, step: int | None):
        """Also store the state at this step (None / -1: only multiples of the output interval)."""
        self.final_output_step = -1 if step is None else int(step)
        self._ck(self.lib.sc_set_final_output_step(self.h, self.final_output_step))

    @staticmethod
    def n_output_rows(t_start: int, n_steps: int, out_interval: int, final_step: int = -1) -> int:
        first = -(-t_start // out_interval) * out_interval
        last = t_start + n_steps
        n = 0 if first > last else (last - first) // out_interval + 1
        if t_start <= final_step <= last and final_step % out_interval != 0:
            n += 1
        return n

    def run_newmark(self, dt, t_start, n_steps, out_interval=1, beta=0.25, gamma=0.5, rtol=1e-14, maxit=20000,
                    u_out=None, v_out=None, a_out=None, store=True):
        n_out = self.n_output_rows(t_start, n_steps, out_interval, self.final_output_step) if store else 0
        if store:
            u_out = np.zeros((n_out, self._row_len())) if u_out is None else u_out
            v_out = np.zeros((n_out, self._row_len())) if v_out is None else v_out
            a_out = np.zeros((n_out, self._row_len())) if a_out is None else a_out
        st = Stats()
        self.state_epoch += 1
        self._ck(self.lib.sc_run_newmark(self.h, float(dt), int(t_start), int(n_steps), int(out_interval), float(beta),
                                         float(gamma), float(rtol), int(maxit), n_out, _ptr(u_out), _ptr(v_out), _ptr(a_out),
                                         C.byref(st)))
        return u_out, v_out, a_out, st.as_dict()

    def run_central_difference(self, dt, t_start, n_steps, out_interval=1, u_out=None, v_out=None, a_out=None, store=True):
        n_out = self.n_output_rows(t_start, n_steps, out_interval, self.final_output_step) if store else 0
        if store:
            u_out = np.zeros((n_out, self._row_len())) if u_out is None else u_out
            v_out = np.zeros((n_out, self._row_len())) if v_out is None else v_out
            a_out = np.zeros((n_out, self._row_len())) if a_out is None else a_out
        st = Stats()
        self.state_epoch += 1
        self._ck(self.lib.sc_run_central_difference(self.h, float(dt), int(t_start), int(n_steps), int(out_interval), n_out,
                                                    _ptr(u_out), _ptr(v_out), _ptr(a_out), C.byref(st)))
        return u_out, v_out, a_out, st.as_dict()

    def run_bathe(self, dt, t_start, n_steps, out_interval=1, rtol=1e-14, maxit=20000, u_out=None, v_out=None, a_out=None):
        n_out = self.n_output_rows(t_start, n_steps, out_interval, self.final_output_step)
        u_out = np.zeros((n_out, self._row_len())) if u_out is None else u_out
        v_out = np.zeros((n_out, self._row_len())) if v_out is None else v_out
        a_out = np.zeros((n_out, self._row_len())) if a_out is None else a_out
        st = Stats()
        self.state_epoch += 1
        self._ck(self.lib.sc_run_bathe(self.h, float(dt), int(t_start), int(n_steps), int(out_interval), float(rtol), int(maxit),
                                       n_out, _ptr(u_out), _ptr(v_out), _ptr(a_out), C.byref(st)))
        return u_out, v_out, a_out, st.as_dict()

    def run_static(self, t_start, n_steps, out_interval=1, rtol=1e-12, maxit=100000, u_out=None):
        n_out = self.n_output_rows(t_start, n_steps, out_interval, self.final_output_step)
        u_out = np.zeros((n_out, self._row_len())) if u_out is None else u_out
        st = Stats()
        self.state_epoch += 1
        self._ck(self.lib.sc_run_static(self.h, int(t_start), int(n_steps), int(out_interval), float(rtol), int(maxit), n_out,
                                        _ptr(u_out), C.byref(st)))
        return u_out, st.as_dict()

    # ---- multi-GPU
    def dist_init(self, rank: int, world: int, unique_id: bytes | None):
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        self._ck(self.lib.sc_dist_init(self.h, rank, world, buf))

    def set_halo(self, neighbor_rank, send_ptr, send_idx, recv_ptr, recv_idx):
        nr = _arr(neighbor_rank, np.int32); sp = _arr(send_ptr, np.int64); si = _arr(send_idx, np.int64)
        rp = _arr(recv_ptr, np.int64); ri = _arr(recv_idx, np.int64)
        self._ck(self.lib.sc_set_halo(self.h, len(nr), _ptr(nr), _ptr(sp), _ptr(si), _ptr(rp), _ptr(ri)))

    def halo_exchange(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).copy()
        self._ck(self.lib.sc_halo_exchange(self.h, _ptr(x)))
        return x
