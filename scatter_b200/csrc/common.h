// Internal declarations shared by the translation units of libscatter_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/scatter_b200.h"

#define SC_MAX_NNE 20
#define SC_MAX_GP 27

// element-independent integration tables, evaluated on the host once per (element type, order)
struct ShapeTable {
    int nne = 0, dim = 0, ngp = 0;
    std::vector<double> N;    // [ngp][nne]
    std::vector<double> dN;   // [ngp][nne][dim]
    std::vector<double> w;    // [ngp]
};
// shape_tables.cpp
int sc_elem_nne(int elem_type);
int sc_elem_dim(int elem_type);
bool sc_make_shape_table(int elem_type, int order, ShapeTable& out, std::string& err);

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
};

struct NcclApi;   // dist.cu

// Factorised sparse approximate inverse of an SPD matrix A in the CSR pattern of the context:  A^-1 ~ G^T G  with G lower
// triangular on a filtered subset of A's lower pattern (fsai.cu).  Both factors are kept row-wise (gather form, FP32 values:
// a preconditioner only has to be symmetric positive definite, the residuals stay FP64).
struct sc_fsai {
    const double* for_vals = nullptr;   // values array the factor was computed from (null: empty slot)
    int64_t nnz = 0;
    int64_t* rowptr = nullptr;   int2* cv = nullptr;      // G: (column, FP32 value bits) pairs, rows ascending in column, diagonal last
    int64_t* t_rowptr = nullptr; int2* t_cv = nullptr;    // G^T, diagonal first
    int32_t* perm = nullptr;            // [n_eq] component-major number of every equation; rows and columns of both factors
                                        // are stored in that numbering (null: the context's own numbering)
    int lanes = 8;                      // lanes per row of the apply kernels
    int max_row = 0, t_max_row = 0;     // longest row of G / G^T
    int tile_max = 0, t_tile_max = 0;   // most entries in one tile of rows of G / G^T (ring stage size of the TMA-fed products)
    int jacobi_rows = 0;                // rows whose local system was not positive definite (diagonal scaling instead)
    double seconds = 0.0;               // set-up time (device)
};

struct sc_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    std::string err;
    int64_t launches = 0;

    // mesh
    int elem_type = -1, nne = 0, dim = 0;
    int64_t n_nodes = 0, n_elem = 0, n_eq = 0;
    double* d_xyz = nullptr;        // [n_nodes*3]
    int32_t* d_conn = nullptr;      // [n_elem*nne]
    int32_t* d_eq = nullptr;        // [n_nodes*dim], -1 fixed
    uint8_t* d_active = nullptr;    // [n_nodes] or null
    double *d_E = nullptr, *d_nu = nullptr, *d_rho = nullptr;   // [n_elem]
    bool have_mesh = false, have_mat = false;

    // node-level structure (pattern.cu)
    int64_t* d_n2e_ptr = nullptr;   // [n_nodes+1]
    int32_t* d_n2e = nullptr;       // element ids, ascending per node
    int64_t* d_nbr_ptr = nullptr;   // [n_nodes+1]
    int32_t* d_nbr = nullptr;       // neighbour node rows (incl. self), ascending
    uint16_t* d_nbr_off = nullptr;  // dof offset of neighbour inside a row of the node
    uint8_t* d_nbr_free = nullptr;  // bit j set: dof j of the neighbour node is free (has a column)
    int32_t* d_node_rl = nullptr;   // [n_nodes] row length of the node's rows (0 if inactive)
    int64_t* d_node_row0 = nullptr; // [n_nodes+1] number of free dofs before the node (= first row of the node)
    // node-blocked column structure for the time loop: all rows of a node share one column list
    struct NodeDesc { int64_t val_off; int64_t col_off; int32_t row0; int32_t len_nfree; };   // len | pattern id << 16 | nfree << 24
    NodeDesc* d_nd = nullptr;       // [n_nodes + 32] (padded with empty descriptors)
    int32_t* d_ncol = nullptr;      // [sum node_rl] column list per node
    uint8_t* d_pair_pos = nullptr;  // [n_pairs*nne] position of every element node in the pair's node neighbour list (or null)
    uint8_t* d_pair_al = nullptr;   // [n_pairs] local index of the pair's node in its element
    // block descriptors of the record-fed assembly kernel (assemble.cu: asm_build_block_desc, built with the pattern)
    int32_t* d_blk_elem = nullptr;  // [n_blocks][blk_ppb] distinct elements of every block of blk_npb consecutive nodes, ascending
    int32_t* d_blk_U = nullptr;     // [n_blocks] how many                      (these three only live until k_blk_pack has
    uint8_t* d_pair_ui = nullptr;   // [n_pairs] index of the pair's element in its block's list (255: none)    packed them)
    int blk_npb = 0, blk_imax = 0, blk_ppb = 0, blk_umax = 0;   // most nodes / (node, neighbour) items / pair lanes / distinct elements of a block
    int64_t blk_count = 0;          // node blocks (consecutive nodes packed greedily into the pair lanes of a CTA)
    unsigned char* d_blk_desc = nullptr;   // [n_blocks][blk_desc_stride] packed block descriptors of the persistent kernel (k_blk_pack)
    int blk_desc_stride = 0;
    double* d_asm_rec = nullptr;    // [n_elem][REC] element records (scratch of the assembly, see asm_release_scratch)
    size_t asm_rec_cap = 0;         // doubles allocated
    bool no_asm_records = false;    // sc_set_option("assembly_records", 0): k_assemble_blk (Jacobian set-up inside every block, no scratch)
    int64_t ncol_total = 0;
    int32_t* d_dict = nullptr;      // [n_dict*dict_stride] most frequent relative column lists (node_dict.cu); nodes with one of
    int n_dict = 0, dict_stride = 0;   // them carry its id in their descriptor and have no explicit list in d_ncol
    bool no_dict = false;           // sc_set_option("column_dictionary", 0): keep every explicit node column list
    int max_nbr = 0, max_rl = 0, max_valence = 0;   // max neighbours / row length / elements per node

    // dof-level CSR
    int64_t nnz = 0;
    int64_t* d_rowptr = nullptr;    // [n_eq+1]
    int32_t* d_col = nullptr;       // [nnz]
    bool have_pattern = false;

    // values
    double* d_K = nullptr;          // [nnz]
    double* d_M = nullptr;          // [nnz] (optional)
    double* d_Ml = nullptr;         // [n_eq] lumped mass (optional)
    double* d_Khat = nullptr;       // [nnz] effective matrix (Newmark / Bathe sub-step 1)
    double* d_Khat2 = nullptr;      // [nnz] effective matrix of Bathe sub-step 2
    double* d_C = nullptr;          // [nnz] explicit damping matrix of sc_set_csr (null: C = C_abs + c0 M + c1 K, never stored)
    bool csr_only = false;          // matrices came from sc_set_csr: no mesh, no node structure, row-wise kernels only
    bool have_K = false, have_M = false, have_Ml = false;
    double c0 = 0.0, c1 = 0.0;

    // C_abs (absorbing dashpots) kept as a small row-compressed list
    int64_t cabs_n = 0, cabs_rows = 0;
    int64_t* d_cabs_rowid = nullptr;   // [cabs_rows] row index
    int64_t* d_cabs_rptr = nullptr;    // [cabs_rows+1]
    int32_t* d_cabs_col = nullptr;     // [cabs_n]
    int64_t* d_cabs_slot = nullptr;    // [cabs_n] position in the main CSR
    double* d_cabs_val = nullptr;      // [cabs_n]

    // loads
    int64_t load_steps = 0;
    std::vector<int64_t> h_load_ptr;
    int32_t* d_load_dof = nullptr;
    double* d_load_val = nullptr;

    // PCG preconditioner and initial guess (timeloop.cu, fsai.cu)
    sc_fsai fsai[2];                // factors of d_Khat / d_Khat2 (or d_K for the static solver)
    bool no_fsai = false;           // sc_set_option("fsai", 0): Jacobi preconditioner
    bool fsai_no_perm = false;      // sc_set_option("fsai_component_major", 0): factors in the context's interleaved numbering
    bool fsai_no_vertex_first = false;   // sc_set_option("fsai_vertex_first", 0): eliminate in plain equation order (quadratic meshes)
    double fsai_tau = 0.05;         // pattern filter: |a_ij| >= tau sqrt(a_ii a_jj)   (sc_set_option("fsai_tau_permille", ..))
    int proj_k = 16;                // sc_set_option("pcg_projection", k): A-orthonormal basis of up to k previous solutions (0: off)
    int proj_n = 0;                 // vectors currently in the basis
    const double* proj_for = nullptr;   // matrix values the basis belongs to
    std::vector<double*> proj_x, proj_ax;   // [proj_k] basis vectors and their products with the matrix
    double** d_proj_ptr = nullptr;  // device copy of the 2 proj_k pointers (x first, then ax)
    double* d_proj_coef = nullptr;  // [proj_k + 2] projection coefficients / norms

    int64_t pcg_stagnations = 0;    // solves accepted at a stagnated residual below 1e-9 (see timeloop.cu: pcg)
    int64_t extra_out_step = -1;    // sc_set_final_output_step: this step is stored even if it is no multiple of the interval

    // state
    double *d_u = nullptr, *d_v = nullptr, *d_a = nullptr;   // [n_eq]
    std::vector<double*> work;      // work vectors [n_eq], allocated on demand
    double* d_scal = nullptr;       // small device scalar block
    double* d_partial = nullptr;    // reduction partials
    double* h_pinned = nullptr;     // pinned host scalars
    cudaEvent_t ev_rows_ready = nullptr, ev_rows_done = nullptr;   // output rows: snapshot ready / D2H finished
    bool rows_pending = false;
    double* d_snap[3] = {nullptr, nullptr, nullptr};               // snapshots of u, v, a while their D2H copy is in flight
    int32_t* d_sel = nullptr;                                      // sc_set_output_dofs: equations copied out per output row
    int64_t n_sel = -1;                                            // -1: full rows
    double* d_selbuf[3] = {nullptr, nullptr, nullptr};             // gathered u, v, a of the current output row
    bool force_one_group = false;          // sc_set_option("spmv_groups", 1): two CTAs per SM with their own short rings (round-1 layout)
    bool force_no_node = false;            // sc_set_option("node_spmv", 0): row-wise kernels instead of the node-blocked one
    bool force_no_tma = false;             // sc_set_option("tma_spmv", 0): register-staged SpMV instead of the TMA ring
    bool no_small_pcg = false;             // sc_set_option("small_pcg", 0): never use the cooperative single-kernel PCG
    int small_pcg_grid = 0;                // co-resident grid limit of k_pcg_small (0: not queried yet)
    bool no_graph = false;                 // sc_set_option("pcg_graph", 0): launch the PCG iteration kernel by kernel
    // captured PCG iterations (single-GPU), each valid for the pointers in its key; two slots, because Bathe alternates
    // between two effective matrices / preconditioners every step
    struct PcgGraph { cudaGraphExec_t exec = nullptr; const void* key[8] = {}; int64_t n = 0; int launches = 0; uint64_t used = 0; };
    PcgGraph pcg_graphs[2];
    uint64_t pcg_graph_clock = 0;
    bool force_generic_assembly = false;   // sc_set_option("generic_assembly", 1): warp-per-node kernel for every element type
    bool nm_resume_valid = false;   // d_a holds the Newmark acceleration of step nm_resume_t (stage continuation)
    int64_t nm_resume_t = 0;
    double khat_a1 = -1.0, khat_a4 = -1.0;   // parameters d_Khat was built with (-1: invalid)
    double cd_coef_dt = -1.0;       // work[2..4] hold inv_d, alpha, lumped c of the central difference for this dt (-1: invalid)
    bool cd_resume_valid = false;   // work[0] holds u(t - dt) of the central-difference state at step cd_resume_t
    int64_t cd_resume_t = 0;
    double cd_resume_dt = 0.0;

    // multi-GPU
    // halo overlap of the explicit step (spmv_node.cu: la_node_overlap_plan): tiles [ov_tile_lo, ov_tile_hi) = rows
    // [ov_row_lo, ov_row_hi) neither send nor read halo values; they run while the exchange of the other tiles is in flight
    bool ov_planned = false, ov_ok = false, no_overlap = true;    // sc_set_option("halo_overlap", 1) enables it (default: serial)
    int64_t ov_tile_lo = 0, ov_tile_hi = 0, ov_row_lo = 0, ov_row_hi = 0;
    int ov_spare_sms = 4;           // SMs the interior launch leaves to NCCL (sc_set_option("halo_spare_sms", k))
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_bnd = nullptr, ev_halo = nullptr;
    // peer-memory halo exchange (dist.cu): receive window + flags in this rank's HBM, mapped neighbours' windows / flags
    bool peer_tried = false, peer_ok = false, no_peer_halo = false;   // sc_set_option("peer_halo", 0): NCCL send / recv instead
    double* d_peer_win = nullptr;                   // [2][n_recv]
    unsigned long long* d_peer_flags = nullptr;     // [n_nbr] exchange counters raised by the neighbours, + counter, + time-out mark
    double** d_peer_win_ptr = nullptr;              // [n_nbr] neighbours' windows
    unsigned long long** d_peer_flag_ptr = nullptr; // [n_nbr] this rank's flag in every neighbour's flag array
    int64_t* d_peer_off = nullptr;                  // [2 n_nbr] offset of this rank's values in the neighbour's window; its n_recv
    int64_t* d_send_ptr = nullptr;                  // [n_nbr + 1]
    std::vector<void*> peer_maps;                   // IPC mappings to close
    unsigned long long peer_epoch = 0;
    int rank = 0, world = 1;
    NcclApi* nccl = nullptr;
    void* comm = nullptr;
    int n_nbr_ranks = 0;
    std::vector<int> nbr_rank;
    std::vector<int64_t> send_ptr, recv_ptr;
    int64_t* d_send_idx = nullptr;
    int64_t* d_recv_idx = nullptr;
    double* d_send_buf = nullptr;
    double* d_recv_buf = nullptr;
};

int sc_fail(sc_ctx* ctx, int code, const char* fmt, ...);
void sc_set_global_error(const char* msg);

#define SC_CUDA(ctx, call)                                                                           \
    do {                                                                                             \
        cudaError_t _e = (call);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return sc_fail((ctx), SC_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,  \
                           cudaGetErrorString(_e));                                                  \
    } while (0)

#define SC_TRY(expr)              \
    do {                          \
        int _r = (expr);          \
        if (_r != SC_OK) return _r; \
    } while (0)

#define SC_CHECK_LAUNCH(ctx)                                   \
    do {                                                       \
        (ctx)->launches++;                                     \
        SC_CUDA((ctx), cudaGetLastError());                    \
    } while (0)

// Pair of timing events on one stream; destroyed on every exit path of the caller (the SC_TRY / SC_CUDA macros return early).
struct sc_gpu_timer {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t st;
    explicit sc_gpu_timer(cudaStream_t s) : st(s) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
    ~sc_gpu_timer() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
    sc_gpu_timer(const sc_gpu_timer&) = delete;
    sc_gpu_timer& operator=(const sc_gpu_timer&) = delete;
    void start() { cudaEventRecord(e0, st); }
    void stop() { cudaEventRecord(e1, st); }
    float ms() const { float v = 0.f; cudaEventElapsedTime(&v, e0, e1); return v; }   // after the stream was synchronised
};

template <typename T>
int sc_alloc(sc_ctx* ctx, T** p, size_t n) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (n == 0) n = 1;
    // 64 bytes of slack: the TMA bulk copies round slice ends up to 16 bytes
    cudaError_t e = cudaMalloc((void**)p, n * sizeof(T) + 64);
    if (e != cudaSuccess) {
        *p = nullptr;
        return sc_fail(ctx, SC_ERR_CUDA, "cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
    }
    return SC_OK;
}
template <typename T>
void sc_free(T** p) {
    if (*p) { cudaFree(*p); *p = nullptr; }
}

// pattern.cu
int sc_pattern_build(sc_ctx* ctx);
// assemble.cu
int sc_assemble_run(sc_ctx* ctx, int order, int flags, double* seconds);
void asm_release_scratch(sc_ctx* ctx);                                  // drop the element-record scratch (time loops, pattern change)
int asm_build_block_desc(sc_ctx* ctx);                                  // called at the end of sc_pattern_build
// threads and lanes per (node, element) pair of the block assembly kernels for an element type
inline void asm_blk_shape(int nne, int dim, int* tpb, int* lpp) {
    if (dim * nne * dim > 72 && nne % 5 == 0) { *tpb = 160; *lpp = 5; }         // tetra10 / hexa20
    else if (dim * nne * dim > 36 && nne % 2 == 0) { *tpb = 128; *lpp = 2; }    // hexa8, quad8
    else { *tpb = 128; *lpp = 1; }
}
// linalg.cu
int sc_work(sc_ctx* ctx, int idx, double** out);                       // work vector idx
int la_spmv(sc_ctx* ctx, const double* vals, const double* x, double* y);                 // y = A x
int la_spmv2(sc_ctx* ctx, const double* va, const double* xa, const double* vb, const double* xb, double* y); // y = A xa + B xb
int la_cabs_spmv_add(sc_ctx* ctx, const double* x, double* y, double scale);   // y += scale * C_abs x
int la_cabs_add_values(sc_ctx* ctx, double* vals, double scale);               // vals[slot] += scale * C_abs
int la_axpby_vals(sc_ctx* ctx, double* out, double a, const double* x, double b, const double* y, int64_t n);
int la_extract_diag(sc_ctx* ctx, const double* vals, double* diag, bool invert);
int la_fill(sc_ctx* ctx, double* x, double v, int64_t n);
int la_dot(sc_ctx* ctx, const double* x, const double* y, double* d_out);     // deterministic 2-stage, result on device
int la_scratch(sc_ctx* ctx);
constexpr size_t SC_PARTIAL_DOUBLES = 4096;                              // lower bound of the size of sc_ctx::d_partial
constexpr int64_t SMALL_PCG_MAX_N = 250000;                              // largest system solved by the cooperative PCG
// pcg_small.cu
bool pcg_small_usable(sc_ctx* ctx);
int pcg_small(sc_ctx* ctx, const double* vals, const double* dinv, const double* b, double* x, double* r, double* p, double* q,
              double rtol, int maxit, int* iters, double* relres, double ref_norm2);
void pcg_graph_drop(sc_ctx* ctx);                                        // timeloop.cu: forget the captured PCG iteration
void precond_drop(sc_ctx* ctx);                                          // timeloop.cu: forget FSAI factors / projection basis (matrix values changed)
void precond_destroy(sc_ctx* ctx);                                       // ... and release their buffers (equation count changes)
int la_cd_step(sc_ctx* ctx, const double* K, const double* w, const double* u, double* uprev_next, const double* inv_d,
               const double* alpha, double g, double* w_next);
int la_spmv_dot(sc_ctx* ctx, const double* vals, const double* p, double* q, double* d_out);
// fsai.cu
int fsai_build(sc_ctx* ctx, sc_fsai* f, const double* vals);            // factor for the matrix `vals` (replaces the slot's content)
void fsai_free(sc_fsai* f);
int fsai_apply(sc_ctx* ctx, const sc_fsai* f, const double* r, double* t, double* z, double* partial, int nb, double* rp);
                                                                        // z = G^T G r in the factor's numbering (f->perm), partial[0..nb) =
                                                                        // per-block sums of r.z; rp: scratch vector (renumbered r)
int fsai_unpermute(sc_ctx* ctx, const sc_fsai* f, const double* zp, double* z);   // z[i] = zp[f->perm[i]]
// spmv_tma.cu
bool la_tma_usable(sc_ctx* ctx);
// spmv_node.cu
bool la_node_usable(sc_ctx* ctx);
int la_node_spmv(sc_ctx* ctx, const double* vals, const double* x, double* y);
int la_node_cd_step(sc_ctx* ctx, const double* K, const double* w, const double* u, double* uprev_next, const double* inv_d,
                    const double* alpha, double g, double* w_next, int part = 0);
int la_node_overlap_plan(sc_ctx* ctx);
int la_node_spmv_dot(sc_ctx* ctx, const double* vals, const double* p, double* q, double* partial, unsigned* nblocks);
int64_t la_node_step_bytes(sc_ctx* ctx, bool lagged);
int la_tma_spmv(sc_ctx* ctx, const double* vals, const double* x, double* y);
int la_tma_cd_step(sc_ctx* ctx, const double* K, const double* w, const double* u, double* uprev_next, const double* inv_d,
                   const double* alpha, double g, double* w_next);
int la_tma_spmv_dot(sc_ctx* ctx, const double* vals, const double* p, double* q, double* partial, unsigned* nblocks);
// timeloop.cu
int tl_newmark(sc_ctx* ctx, double dt, int64_t t0, int64_t n_steps, int64_t oi, double beta, double gamma, double rtol,
               int maxit, int64_t n_out, double* u_out, double* v_out, double* a_out, sc_stats* st);
int tl_central_difference(sc_ctx* ctx, double dt, int64_t t0, int64_t n_steps, int64_t oi, int64_t n_out, double* u_out,
                          double* v_out, double* a_out, sc_stats* st);
int tl_bathe(sc_ctx* ctx, double dt, int64_t t0, int64_t n_steps, int64_t oi, double rtol, int maxit, int64_t n_out, double* u_out,
             double* v_out, double* a_out, sc_stats* st);
int tl_static(sc_ctx* ctx, int64_t t0, int64_t n_steps, int64_t oi, double rtol, int maxit, int64_t n_out, double* u_out, sc_stats* st);
// absorb.cu: face matrices, dashpot / spring coefficients and the ordered per-key sums -> device arrays [n_unique]
int abs_faces_eval(sc_ctx* ctx, int face_type, int order, int64_t n_faces, const int32_t* face_nodes, const int32_t* face_elem,
                   const int32_t* face_dir, const uint8_t* perp, int64_t n_unique, const int64_t* grp_ptr, const int64_t* grp_entry,
                   double p0, double p1, double stiff, double** d_csum, double** d_ksum);
// node_dict.cu / spmv_node.cu
constexpr int SC_DICT_MAX = 32;
int node_dict_build(sc_ctx* ctx);
int64_t node_dict_room(sc_ctx* ctx);
// srf.cu
int srf_sample(sc_ctx* ctx, int64_t n_points, const double* pos, int n_modes, const double* k, const double* z1, const double* z2,
               double scale, double mean, int lognormal, double* out, double* seconds);
// dist.cu
int dist_init(sc_ctx* ctx, int rank, int world, const void* id);
int dist_unique_id(void* out);
int dist_halo(sc_ctx* ctx, double* d_x, cudaStream_t s);      // exchange ghost values of device vector x (in place)
int dist_allreduce_sum(sc_ctx* ctx, double* d_vals, int n, cudaStream_t s);
void dist_destroy(sc_ctx* ctx);
void dist_peer_release(sc_ctx* ctx);                          // drop the peer-memory windows (halo plan changed)
int dist_peer_check(sc_ctx* ctx);                             // after a stream synchronisation: did a wait on a neighbour time out?
