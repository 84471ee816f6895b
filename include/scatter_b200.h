/*
 * scatter_b200.h -- C ABI of libscatter_b200.so: the B200 (sm_100a) implementation of SCATTER's hot path.
 *
 * The reference (PlatypusBytes/scatter) is pure Python and has no FFI of its own; its seams for this path are
 * plain Python call sites.  Every entry point below names the reference call it replaces
 * (file:line relative to the reference tree).  The Python shims in scatter_b200/ (GenerateMatrix, the solver
 * classes, scatter.scatter) bind these symbols with ctypes -- see INTEGRATION.md for the binding.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every host buffer, the library owns all device memory
 *   - one opaque context per GPU / rank; a context is not thread-safe; calls are synchronous at return
 *   - every function returns 0 on success or a negative sc_status; sc_last_error() gives the message
 *   - there is no CPU fallback: sc_create fails if no CUDA device is usable
 */
#ifndef SCATTER_B200_H
#define SCATTER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sc_ctx sc_ctx;

enum sc_status {
    SC_OK = 0,
    SC_ERR_CUDA = -1,      /* CUDA runtime error (message has the call site) */
    SC_ERR_ARG = -2,       /* invalid argument / call order */
    SC_ERR_STATE = -3,     /* required earlier step missing (mesh, pattern, assembly ...) */
    SC_ERR_UNSUPPORTED = -4,
    SC_ERR_NOCONV = -5,    /* PCG did not reach the requested tolerance, or an explicit run diverged (non-finite state) */
    SC_ERR_NCCL = -6
};

/* element types, order as in SURVEY.md 8a; gmsh node ordering (scatter/element_types.py) */
enum sc_elem_type {
    SC_TRI3 = 0, SC_TRI6 = 1, SC_QUAD4 = 2, SC_QUAD8 = 3, SC_TETRA4 = 4, SC_TETRA10 = 5, SC_HEXA8 = 6, SC_HEXA20 = 7
};

enum sc_matrix { SC_MAT_K = 0, SC_MAT_M = 1, SC_MAT_C = 2, SC_MAT_KHAT = 3 };

/* flags for sc_assemble */
enum sc_assemble_flags {
    SC_ASM_K = 1,            /* stiffness values on the structural pattern */
    SC_ASM_M_FULL = 2,       /* consistent mass on the same pattern (explicit zeros kept, system_matrix.py:98-103) */
    SC_ASM_M_LUMPED = 4      /* row sums of the consistent mass (central-difference path) */
};

typedef struct sc_stats {
    double  seconds_total;       /* wall time inside the call (host clock) */
    double  seconds_device;      /* CUDA-event time of the stepping loop */
    double  seconds_halo;        /* CUDA-event time spent in halo exchange (multi-GPU) */
    int64_t steps;               /* time steps taken */
    int64_t pcg_iterations;      /* total PCG iterations (Newmark) */
    int64_t kernel_launches;     /* kernels of this library launched inside the call */
    double  last_residual;       /* last relative PCG residual */
    double  reserved[4];
} sc_stats;

/* ---- life cycle ------------------------------------------------------------------------------------------- */
int         sc_create(int device, sc_ctx** out);
void        sc_destroy(sc_ctx* ctx);
const char* sc_last_error(sc_ctx* ctx);          /* ctx may be NULL: error of a failed sc_create */
int         sc_version(void);
int         sc_device_info(sc_ctx* ctx, int* sm_count, int64_t* total_mem, int64_t* free_mem, char* name, int name_len);
int64_t     sc_kernel_launches(sc_ctx* ctx);     /* running count of kernels launched by this context */
/* preconditioner currently held for the effective matrix (slot 0: Newmark / static / Bathe sub-step 1, slot 1: Bathe
 * sub-step 2): entries of the FSAI factor G (0: none, Jacobi in use), its set-up time, vectors in the projection basis */
int         sc_precond_info(sc_ctx* ctx, int slot, int64_t* fsai_nnz, double* fsai_seconds, int* projection_vectors);
/* kernel-selection switches for tests and A/B measurements (the defaults are the product path; nothing reads the
 * environment): "node_spmv", "tma_spmv", "column_dictionary", "small_pcg", "pcg_graph", "assembly_records" (default 1:
 * element records + persistent TMA-fed assembly; 0: block kernel without scratch, same bits), "generic_assembly"
 * (default 0), "spmv_groups" (consumer groups per CTA of the node-blocked SpMV: 2, or 1 for the two-CTA layout);
 * solver options of the implicit integrators: "fsai" (default 1: factorised sparse approximate inverse preconditioner
 * for systems beyond the cooperative small-system kernel; 0: Jacobi), "fsai_tau_permille" (pattern filter, default 50), "fsai_vertex_first"
 * (default 1: vertex equations of quadratic meshes precede mid-side equations in the factor's elimination order),
 * "pcg_projection" (previous solutions the right-hand side is projected on before PCG, default 16, 0: off) */
int         sc_set_option(sc_ctx* ctx, const char* name, int64_t value);

/* page-locked host buffers for result rows / initial states (faster, asynchronous host<->device copies) */
int         sc_host_alloc(void** out, int64_t bytes);
int         sc_host_free(void* p);

/* integration tables of one (element type, Gauss order): N[ngp*nne], dN[ngp*nne*dim], w[ngp]; runs on the host
 * (no GPU needed).  element_types.py:52-802 + discretisation.py:436-497.  Buffers may be NULL to query sizes. */
int         sc_shape_table(int elem_type, int order, int* nne, int* dim, int* ngp, double* N, double* dN, double* w);

/* ---- mesh / numbering (replaces the ReadMesh attributes consumed at system_matrix.py:35-103:
 *      model.nodes, model.elem, model.eq_nb_dof, model.number_eq, model.element_type) ---------------------------
 *  xyz   [n_nodes*3] row-major (2-D elements ignore z, discretisation.py:329)
 *  conn  [n_elem*nne] 0-based node rows, gmsh node order
 *  eq    [n_nodes*dim] equation number or -1 for a fixed dof (mesher.py:276-310 uses NaN); must increase with
 *        (node row, dof) exactly as the reference numbers them
 *  active[n_nodes] or NULL: 1 = rows of this node are assembled/integrated on this rank, 0 = ghost node whose
 *        values arrive through the halo exchange (domain decomposition; NULL = all active)                      */
int sc_set_mesh(sc_ctx* ctx, int elem_type, int64_t n_nodes, const double* xyz, int64_t n_elem, const int32_t* conn,
                const int64_t* eq, int64_t n_eq, const uint8_t* active);

/* ---- random material field (random_fields.py:59-104: gstools SRF sampled at the element centroids; gstools==1.7.0 is
 *      un-vendored, its RandMeth generator is restated here).  Randomisation method:
 *        out[p] = mean + scale * sum_j ( z1[j] cos(k_j . x_p) + z2[j] sin(k_j . x_p) ),   scale = sqrt(var / n_modes)
 *      and out = exp(out) when `lognormal` (random_fields.py:103-104).  The sum runs over j in order (reproducible).
 *  pos [n_points*3] row-major point coordinates (already divided by the anisotropic length scales),
 *  k [n_modes*3] wave vectors of the unit-length-scale covariance model, z1/z2 [n_modes] standard-normal amplitudes,
 *  out [n_points] host buffer -> per-element values that go into sc_set_materials                                   */
int sc_srf_sample(sc_ctx* ctx, int64_t n_points, const double* pos, int n_modes, const double* k, const double* z1,
                  const double* z2, double scale, double mean, int lognormal, double* out, double* seconds_device /*may be NULL*/);

/* per-element Young's modulus, Poisson ratio, density (system_matrix.py:64-71; random_fields.py:46-57) */
int sc_set_materials(sc_ctx* ctx, const double* young, const double* poisson, const double* density);

/* ---- structural CSR pattern (the key set of k_dict, system_matrix.py:98-121): sorted rows, sorted columns ----- */
int sc_build_pattern(sc_ctx* ctx, int64_t* nnz_out);
int sc_get_pattern(sc_ctx* ctx, int64_t* rowptr /*[n_eq+1]*/, int32_t* col /*[nnz]*/);
/* sizes of the device structures: out[0] nnz, [1] entries of the node-blocked column lists, [2] n_nodes, [3] longest row,
 * [4] most neighbour nodes, [5] most elements per node, [6] 1 if the node-blocked SpMV is in use, [7] entries of the
 * column-pattern dictionary (nodes whose relative column list is in it have no explicit list: [1] counts the rest) */
int sc_pattern_stats(sc_ctx* ctx, int64_t* out8);

/* ---- element integration + deterministic assembly (GenerateMatrix.generate_stiffness_and_mass,
 *      system_matrix.py:35-121 over discretisation.py:83-222,290-417) -------------------------------------------- */
int sc_assemble(sc_ctx* ctx, int gauss_order, int flags, double* seconds_device /*may be NULL*/);

/* add pre-summed COO entries (must lie inside the pattern): absorbing dashpots -> which = SC_MAT_C (kept as the
 * separate term C_abs), absorbing springs -> which = SC_MAT_K (system_matrix.py:360-376) */
int sc_add_entries(sc_ctx* ctx, int which, int64_t n, const int64_t* rows, const int64_t* cols, const double* vals);

/* absorbing boundary faces (GenerateMatrix.absorbing_boundaries, system_matrix.py:256-376, with the face matrices of
 * compute_abs_bound, discretisation.py:419-433).  The caller plans, the device computes:
 *   face_nodes [n_faces*nl] node rows of every absorbing face in the reference's face-node order (utils.py:141-175),
 *   face_elem / face_dir [n_faces] element (material) and normal direction 0..2 of the face,
 *   perp [n_faces*nl] 1 where the dof paired with face position b is perpendicular to its boundary (p0 rho vp, Ec) else
 *        (p1 rho vs, G),
 *   rows/cols [n_unique] sorted, unique matrix positions that receive entries; grp_ptr [n_unique+1] / grp_entry: the
 *        per-face entries (id = (face*nl + a)*nl + b) of every position, ascending = the reference's accumulation order.
 * Result: C_abs := sum of S_ab * coefficient_b (replaces any earlier C_abs), K += sum |S_ab| * modulus_b / stiff.      */
int sc_add_absorbing_faces(sc_ctx* ctx, int face_type, int gauss_order, int64_t n_faces, const int32_t* face_nodes,
                           const int32_t* face_elem, const int32_t* face_dir, const uint8_t* perp, int64_t n_unique,
                           const int64_t* rows, const int64_t* cols, const int64_t* grp_ptr, const int64_t* grp_entry,
                           double p0, double p1, double stiff);

/* Rayleigh damping C = C_abs + c0 M + c1 K (system_matrix.py:198); never materialised in the time loop */
int sc_set_rayleigh(sc_ctx* ctx, double c0, double c1);

int sc_get_values(sc_ctx* ctx, int which, double* vals /*[nnz]*/);
int sc_get_lumped_mass(sc_ctx* ctx, double* diag /*[n_eq]*/);
/* y = A x with the device SpMV kernel (test hook; A = K, M, C or KHAT) */
int sc_spmv(sc_ctx* ctx, int which, const double* x, double* y);

/* ---- caller-supplied matrices: the time-loop seam with matrices this library did not assemble
 *      (solvers.NewmarkExplicit().calculate(matrix.M, matrix.C, matrix.K, F, t0, t1), scatter.py:159, with the scipy
 *      matrices of the reference's own GenerateMatrix).  One CSR pattern (sorted, unique columns; the union of the three
 *      patterns), values of K and optionally M and C on it.  Replaces any mesh / pattern / matrices of the context; the
 *      time loops then run the row-wise SpMV kernels, C enters as given (no Rayleigh split: the central-difference solver
 *      lumps all of it by row sums). */
int sc_set_csr(sc_ctx* ctx, int64_t n_eq, const int64_t* rowptr /*[n_eq+1]*/, const int32_t* col /*[nnz]*/,
               const double* K /*[nnz]*/, const double* M /*[nnz] or NULL*/, const double* C /*[nnz] or NULL*/);

/* ---- loads (replaces the per-step callback Force.update_load_at_t, force_external.py:55-74, scatter.py:151):
 *      total external force of step t = entries step_ptr[t]..step_ptr[t+1] of (dof, val); other dofs zero -------- */
int sc_set_load_schedule(sc_ctx* ctx, int64_t n_steps, const int64_t* step_ptr, const int64_t* dof, const double* val);

/* ---- state (solver.update(t_start) restart hook, scatter.py:158) -------------------------------------------- */
int sc_set_state(sc_ctx* ctx, const double* u, const double* v /*each [n_eq] or NULL = zeros*/);
int sc_get_state(sc_ctx* ctx, double* u, double* v, double* a /*each [n_eq] or NULL*/);

/* ---- time integration (replaces solvers.NewmarkExplicit.calculate(M, C, K, F, t0, t1), scatter.py:159) -------
 *  Integrates load-schedule steps t_start+1 .. t_start+n_steps.  Output row r (r = 0 .. n_out-1) receives the
 *  state at step t_start + r*out_interval ... only steps with (t % out_interval == 0) are stored, the first stored
 *  row is the state at t_start itself when t_start % out_interval == 0.  u_out/v_out/a_out are host buffers of
 *  n_out*n_eq doubles or NULL.  The implicit solve uses Jacobi-preconditioned CG to relative residual pcg_rtol.   */
/* Output selection (export_results.Write.pickle(nodes=[...]), export_results.py:117-135: only a few nodes are kept): the
 * output rows of the run functions then hold just these equations, in this order -- u_out/v_out/a_out are n_out*n doubles.
 * dofs = NULL restores full rows of n_eq doubles. */
int sc_set_output_dofs(sc_ctx* ctx, int64_t n, const int64_t* dofs);

/* One extra output step: the state at load-schedule step `step` is stored as an output row even when it is not a
 * multiple of out_interval (the solver objects pass the last index of the time axis, so that the final state of a run is
 * always kept); -1 switches it off.  n_out of the run functions counts it when it falls inside the stage. */
int sc_set_final_output_step(sc_ctx* ctx, int64_t step);

int sc_run_newmark(sc_ctx* ctx, double dt, int64_t t_start, int64_t n_steps, int64_t out_interval, double beta,
                   double gamma, double pcg_rtol, int pcg_maxit, int64_t n_out, double* u_out, double* v_out,
                   double* a_out, sc_stats* stats);

/* explicit central difference (solvers.CentralDifferenceSolver; the reference ships no fixture for it and the solver
 * source is not in its tree, so the scheme is this library's own -- DESIGN.md 3.3):  row-sum lumped mass m; damping
 * C = C_abs + c0 M + c1 K split into the diagonal c_d = c0 m + rowsum(C_abs), centred in time, and the
 * stiffness-proportional part c1 K applied to the lagged velocity (u(t) - u(t-dt))/dt through the step's SpMV:
 *   (m/dt^2 + c_d/2dt) u(t+dt) = F(t) - K [(1 + c1/dt) u(t) - (c1/dt) u(t-dt)] + 2 m/dt^2 u(t) - (m/dt^2 - c_d/2dt) u(t-dt) */
int sc_run_central_difference(sc_ctx* ctx, double dt, int64_t t_start, int64_t n_steps, int64_t out_interval,
                              int64_t n_out, double* u_out, double* v_out, double* a_out, sc_stats* stats);

/* Bathe composite scheme (solvers.BatheSolver, scatter.py:126-127) and static solver (solvers.StaticSolver.calculate(K, F,
 * t0, t1), scatter.py:128-129,156).  Neither is pinned by a reference fixture; schemes documented in DESIGN.md. */
int sc_run_bathe(sc_ctx* ctx, double dt, int64_t t_start, int64_t n_steps, int64_t out_interval, double pcg_rtol,
                 int pcg_maxit, int64_t n_out, double* u_out, double* v_out, double* a_out, sc_stats* stats);
int sc_run_static(sc_ctx* ctx, int64_t t_start, int64_t n_steps, int64_t out_interval, double pcg_rtol, int pcg_maxit,
                  int64_t n_out, double* u_out, sc_stats* stats);

/* ---- multi-GPU (domain decomposition; one context per rank) --------------------------------------------------- */
int sc_nccl_unique_id(void* out128 /*128 bytes*/);
int sc_dist_init(sc_ctx* ctx, int rank, int world, const void* nccl_unique_id128);
/* neighbour r sends dofs send_idx[send_ptr[r]..send_ptr[r+1]) of the local vector and receives into
 * recv_idx[recv_ptr[r]..recv_ptr[r+1]) (local equation numbers) */
int sc_set_halo(sc_ctx* ctx, int n_neighbors, const int32_t* neighbor_rank, const int64_t* send_ptr,
                const int64_t* send_idx, const int64_t* recv_ptr, const int64_t* recv_idx);
int sc_halo_exchange(sc_ctx* ctx, double* x_host /*[n_eq] in/out, test hook*/);

#ifdef __cplusplus
}
#endif
#endif /* SCATTER_B200_H */
