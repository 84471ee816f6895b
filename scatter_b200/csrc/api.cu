// C ABI glue of libscatter_b200.so (see include/scatter_b200.h for the contract of every entry point).
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include "common.h"

namespace {
std::string g_error;

__global__ void k_find_slots(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int64_t* __restrict__ rows,
                             const int32_t* __restrict__ cols, int64_t n, int64_t* __restrict__ slot, int* __restrict__ missing) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int64_t r = rows[t];
    const int c = cols[t];
    int64_t lo = rowptr[r], hi = rowptr[r + 1], found = -1;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        int cm = col[mid];
        if (cm == c) { found = mid; break; }
        if (cm < c) lo = mid + 1; else hi = mid;
    }
    slot[t] = found;
    if (found < 0) atomicExch(missing, 1);
}
__global__ void k_add_at(double* __restrict__ vals, const int64_t* __restrict__ slot, const double* __restrict__ v, int64_t n) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n && slot[t] >= 0) vals[slot[t]] += v[t];
}

template <typename T>
int upload(sc_ctx* ctx, T** dst, const T* src, size_t n) {
    SC_TRY(sc_alloc(ctx, dst, n));
    if (n) SC_CUDA(ctx, cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return SC_OK;
}

void free_pattern(sc_ctx* c) {
    pcg_graph_drop(c);
    precond_destroy(c);
    c->ov_planned = c->ov_ok = false;
    sc_free(&c->d_n2e_ptr); sc_free(&c->d_n2e); sc_free(&c->d_nbr_ptr); sc_free(&c->d_nbr); sc_free(&c->d_nbr_off); sc_free(&c->d_nbr_free);
    sc_free(&c->d_node_rl); sc_free(&c->d_node_row0); sc_free(&c->d_rowptr); sc_free(&c->d_col); sc_free(&c->d_nd); sc_free(&c->d_ncol); sc_free(&c->d_dict); c->n_dict = 0; sc_free(&c->d_pair_pos); sc_free(&c->d_pair_al);
    asm_release_scratch(c);
    sc_free(&c->d_blk_elem); sc_free(&c->d_blk_U); sc_free(&c->d_pair_ui); sc_free(&c->d_blk_desc);
    c->blk_npb = c->blk_imax = c->blk_ppb = c->blk_umax = c->blk_desc_stride = 0; c->blk_count = 0;
    sc_free(&c->d_K); sc_free(&c->d_M); sc_free(&c->d_Ml); sc_free(&c->d_Khat); sc_free(&c->d_Khat2); sc_free(&c->d_C); c->csr_only = false;
    sc_free(&c->d_cabs_rowid); sc_free(&c->d_cabs_rptr); sc_free(&c->d_cabs_col); sc_free(&c->d_cabs_slot); sc_free(&c->d_cabs_val);
    c->cabs_n = c->cabs_rows = 0;
    c->have_pattern = c->have_K = c->have_M = c->have_Ml = false;
    c->khat_a1 = c->khat_a4 = -1.0; c->cd_coef_dt = -1.0;
    c->nnz = 0;
}
void free_vectors(sc_ctx* c) {
    pcg_graph_drop(c);
    precond_destroy(c);
    sc_free(&c->d_u); sc_free(&c->d_v); sc_free(&c->d_a);
    for (auto& w : c->work) sc_free(&w);
    c->work.clear();
    sc_free(&c->d_partial);
    for (int k = 0; k < 3; ++k) { sc_free(&c->d_snap[k]); sc_free(&c->d_selbuf[k]); }
    sc_free(&c->d_sel); c->n_sel = -1;
    c->rows_pending = false;
    c->cd_resume_valid = false; c->nm_resume_valid = false; c->cd_coef_dt = -1.0;
}

// values of C = C_abs + c0 M + c1 K into a fresh device buffer
int build_C(sc_ctx* ctx, double** out) {
    if (!ctx->d_C && (!ctx->have_K || !ctx->have_M)) return sc_fail(ctx, SC_ERR_STATE, "C needs assembled K and full M");
    double* tmp = nullptr;
    SC_TRY(sc_alloc(ctx, &tmp, (size_t)ctx->nnz));
    int rc = ctx->d_C ? la_axpby_vals(ctx, tmp, 1.0, ctx->d_C, 0.0, nullptr, ctx->nnz)
                      : la_axpby_vals(ctx, tmp, ctx->c0, ctx->d_M, ctx->c1, ctx->d_K, ctx->nnz);
    if (rc == SC_OK) rc = la_cabs_add_values(ctx, tmp, 1.0);
    if (rc != SC_OK) { sc_free(&tmp); return rc; }
    *out = tmp;
    return SC_OK;
}
}  // namespace

void sc_set_global_error(const char* msg) { g_error = msg ? msg : ""; }

int sc_fail(sc_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    g_error = buf;
    return code;
}

extern "C" {

int sc_version(void) { return 100; }

const char* sc_last_error(sc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_error.c_str(); }

int sc_create(int device, sc_ctx** out) {
    if (!out) return SC_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return sc_fail(nullptr, SC_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return sc_fail(nullptr, SC_ERR_ARG, "device %d out of range (%d devices)", device, count);
    sc_ctx* ctx = new sc_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        sc_fail(nullptr, SC_ERR_CUDA, "cannot initialise device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
        delete ctx;
        return SC_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = ctx;
    return SC_OK;
}

void sc_destroy(sc_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    dist_destroy(ctx);
    free_pattern(ctx);
    free_vectors(ctx);
    sc_free(&ctx->d_xyz); sc_free(&ctx->d_conn); sc_free(&ctx->d_eq); sc_free(&ctx->d_active);
    sc_free(&ctx->d_E); sc_free(&ctx->d_nu); sc_free(&ctx->d_rho);
    sc_free(&ctx->d_load_dof); sc_free(&ctx->d_load_val); sc_free(&ctx->d_scal);
    sc_free(&ctx->d_send_idx); sc_free(&ctx->d_recv_idx); sc_free(&ctx->d_send_buf); sc_free(&ctx->d_recv_buf);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->ev_rows_ready) cudaEventDestroy(ctx->ev_rows_ready);
    if (ctx->ev_rows_done) cudaEventDestroy(ctx->ev_rows_done);
    if (ctx->ev_bnd) cudaEventDestroy(ctx->ev_bnd);
    if (ctx->ev_halo) cudaEventDestroy(ctx->ev_halo);
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

int sc_device_info(sc_ctx* ctx, int* sm_count, int64_t* total_mem, int64_t* free_mem, char* name, int name_len) {
    if (!ctx) return SC_ERR_ARG;
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    SC_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    size_t f = 0, t = 0;
    SC_CUDA(ctx, cudaMemGetInfo(&f, &t));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (total_mem) *total_mem = (int64_t)t;
    if (free_mem) *free_mem = (int64_t)f;
    if (name && name_len > 0) { std::strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
    return SC_OK;
}

int64_t sc_kernel_launches(sc_ctx* ctx) { return ctx ? ctx->launches : 0; }

int sc_precond_info(sc_ctx* ctx, int slot, int64_t* fsai_nnz, double* fsai_seconds, int* projection_vectors) {
    if (!ctx || slot < 0 || slot > 1) return SC_ERR_ARG;
    const sc_fsai& f = ctx->fsai[slot];
    if (fsai_nnz) *fsai_nnz = f.for_vals ? f.nnz : 0;
    if (fsai_seconds) *fsai_seconds = f.for_vals ? f.seconds : 0.0;
    if (projection_vectors) *projection_vectors = ctx->proj_n;
    return SC_OK;
}

// Kernel-selection switches for tests and A/B measurements; the defaults are the product path.
int sc_set_option(sc_ctx* ctx, const char* name, int64_t value) {
    if (!ctx || !name) return SC_ERR_ARG;
    const std::string k(name);
    const bool on = value != 0;
    if (k == "node_spmv") ctx->force_no_node = !on;                 // node-blocked TMA SpMV (default on)
    else if (k == "tma_spmv") ctx->force_no_tma = !on;              // row-tile TMA SpMV as the second choice (default on)
    else if (k == "column_dictionary") ctx->no_dict = !on;          // relative column lists in shared memory (default on); takes
                                                                    // effect at the next sc_build_pattern
    else if (k == "small_pcg") ctx->no_small_pcg = !on;             // cooperative single-kernel PCG below 250 k equations (default on)
    else if (k == "pcg_graph") ctx->no_graph = !on;                 // CUDA-graph replay of the PCG iteration (default on)
    else if (k == "spmv_groups") ctx->force_one_group = value == 1;      // 1: one consumer group per CTA, two CTAs per SM; default 2
    else if (k == "fsai") { ctx->no_fsai = !on; precond_drop(ctx); }   // FSAI preconditioner of the stream-ordered PCG (default on; 0: Jacobi)
    else if (k == "fsai_component_major") { ctx->fsai_no_perm = !on; precond_drop(ctx); }   // factors stored x-, y-, z-equations first (default on)
    else if (k == "fsai_vertex_first") { ctx->fsai_no_vertex_first = !on; precond_drop(ctx); }   // quadratic meshes: vertex equations precede mid-side ones in the FSAI triangle (default on)
    else if (k == "fsai_tau_permille") {                              // FSAI pattern filter tau in 1/1000 (default 50)
        if (value < 0 || value > 1000) return sc_fail(ctx, SC_ERR_ARG, "fsai_tau_permille must lie in [0, 1000]");
        ctx->fsai_tau = (double)value / 1000.0; precond_drop(ctx);
    }
    else if (k == "pcg_projection") {                                 // previous solutions the right-hand side is projected on (default 16; 0: off)
        if (value < 0 || value > 32) return sc_fail(ctx, SC_ERR_ARG, "pcg_projection must lie in [0, 32]");
        ctx->proj_k = (int)value; precond_drop(ctx);
    }
    else if (k == "peer_halo") { ctx->no_peer_halo = !on; dist_peer_release(ctx); }   // halo values stored straight into the neighbours' HBM (default on; 0: NCCL send / recv)
    else if (k == "halo_spare_sms") {
        if (value < 0 || value > 64) return sc_fail(ctx, SC_ERR_ARG, "halo_spare_sms must lie in [0, 64]");
        ctx->ov_spare_sms = (int)value;
    }
    else if (k == "halo_overlap") { ctx->no_overlap = !on; ctx->ov_planned = false; }   // interior tiles step beside the halo exchange (default off: measured, no gain -- DESIGN.md 3.5)
    else if (k == "assembly_records") ctx->no_asm_records = !on;          // element records + persistent TMA-fed row-gather kernel (default on)
    else if (k == "generic_assembly") ctx->force_generic_assembly = on;   // warp-per-node assembly for every element type (default off)
    else return sc_fail(ctx, SC_ERR_ARG, "unknown option '%s'", name);
    pcg_graph_drop(ctx);
    return SC_OK;
}

int sc_host_alloc(void** out, int64_t bytes) {
    if (!out || bytes < 0) return SC_ERR_ARG;
    cudaError_t e = cudaMallocHost(out, (size_t)(bytes > 0 ? bytes : 1));
    if (e != cudaSuccess) return sc_fail(nullptr, SC_ERR_CUDA, "cudaMallocHost(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
    return SC_OK;
}

int sc_host_free(void* p) {
    if (p && cudaFreeHost(p) != cudaSuccess) return sc_fail(nullptr, SC_ERR_CUDA, "cudaFreeHost failed");
    return SC_OK;
}

int sc_shape_table(int elem_type, int order, int* nne, int* dim, int* ngp, double* N, double* dN, double* w) {
    ShapeTable t;
    std::string err;
    if (!sc_make_shape_table(elem_type, order, t, err)) return sc_fail(nullptr, SC_ERR_ARG, "%s", err.c_str());
    if (nne) *nne = t.nne;
    if (dim) *dim = t.dim;
    if (ngp) *ngp = t.ngp;
    if (N) std::memcpy(N, t.N.data(), t.N.size() * sizeof(double));
    if (dN) std::memcpy(dN, t.dN.data(), t.dN.size() * sizeof(double));
    if (w) std::memcpy(w, t.w.data(), t.w.size() * sizeof(double));
    return SC_OK;
}

int sc_set_mesh(sc_ctx* ctx, int elem_type, int64_t n_nodes, const double* xyz, int64_t n_elem, const int32_t* conn,
                const int64_t* eq, int64_t n_eq, const uint8_t* active) {
    if (!ctx) return SC_ERR_ARG;
    const int nne = sc_elem_nne(elem_type), dim = sc_elem_dim(elem_type);
    if (nne == 0) return sc_fail(ctx, SC_ERR_ARG, "ERROR: Element type not supported");
    if (n_nodes <= 0 || n_elem <= 0 || !xyz || !conn || !eq) return sc_fail(ctx, SC_ERR_ARG, "empty mesh");
    if (n_nodes >= (int64_t)1 << 31 || n_eq >= (int64_t)1 << 31) return sc_fail(ctx, SC_ERR_UNSUPPORTED, "more than 2^31 nodes/equations per rank");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    // equation numbers must increase with (node, dof): the pattern builder relies on it for sorted columns
    std::vector<int32_t> eq32((size_t)n_nodes * dim);
    int64_t next = 0;
    for (int64_t i = 0; i < n_nodes * dim; ++i) {
        const int64_t v = eq[i];
        if (v < 0) { eq32[i] = -1; continue; }
        if (v != next) return sc_fail(ctx, SC_ERR_ARG, "equation numbers must be consecutive in (node, dof) order (entry %lld is %lld, expected %lld)",
                                      (long long)i, (long long)v, (long long)next);
        eq32[i] = (int32_t)v;
        ++next;
    }
    if (next != n_eq) return sc_fail(ctx, SC_ERR_ARG, "eq table holds %lld equations, n_eq = %lld", (long long)next, (long long)n_eq);
    for (int64_t i = 0; i < n_elem * nne; ++i)
        if (conn[i] < 0 || conn[i] >= n_nodes) return sc_fail(ctx, SC_ERR_ARG, "connectivity entry %lld out of range", (long long)i);
    free_pattern(ctx);
    free_vectors(ctx);
    ctx->elem_type = elem_type; ctx->nne = nne; ctx->dim = dim;
    ctx->n_nodes = n_nodes; ctx->n_elem = n_elem; ctx->n_eq = n_eq;
    SC_TRY(upload(ctx, &ctx->d_xyz, xyz, (size_t)n_nodes * 3));
    SC_TRY(upload(ctx, &ctx->d_conn, conn, (size_t)n_elem * nne));
    SC_TRY(upload(ctx, &ctx->d_eq, eq32.data(), eq32.size()));
    if (active) SC_TRY(upload(ctx, &ctx->d_active, active, (size_t)n_nodes));
    else sc_free(&ctx->d_active);
    ctx->have_mesh = true;
    ctx->have_mat = false;
    return SC_OK;
}

int sc_set_materials(sc_ctx* ctx, const double* young, const double* poisson, const double* density) {
    if (!ctx || !ctx->have_mesh) return sc_fail(ctx, SC_ERR_STATE, "sc_set_mesh must be called first");
    if (!young || !poisson || !density) return sc_fail(ctx, SC_ERR_ARG, "null material array");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    SC_TRY(upload(ctx, &ctx->d_E, young, (size_t)ctx->n_elem));
    SC_TRY(upload(ctx, &ctx->d_nu, poisson, (size_t)ctx->n_elem));
    SC_TRY(upload(ctx, &ctx->d_rho, density, (size_t)ctx->n_elem));
    ctx->have_mat = true;
    return SC_OK;
}

int sc_build_pattern(sc_ctx* ctx, int64_t* nnz_out) {
    if (!ctx || !ctx->have_mesh) return sc_fail(ctx, SC_ERR_STATE, "sc_set_mesh must be called first");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    free_pattern(ctx);
    SC_TRY(sc_pattern_build(ctx));
    if (nnz_out) *nnz_out = ctx->nnz;
    return SC_OK;
}

int sc_get_pattern(sc_ctx* ctx, int64_t* rowptr, int32_t* col) {
    if (!ctx || !ctx->have_pattern) return sc_fail(ctx, SC_ERR_STATE, "sc_build_pattern must be called first");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    if (rowptr) SC_CUDA(ctx, cudaMemcpy(rowptr, ctx->d_rowptr, sizeof(int64_t) * (ctx->n_eq + 1), cudaMemcpyDeviceToHost));
    if (col) SC_CUDA(ctx, cudaMemcpy(col, ctx->d_col, sizeof(int32_t) * ctx->nnz, cudaMemcpyDeviceToHost));
    return SC_OK;
}

int sc_pattern_stats(sc_ctx* ctx, int64_t* out8) {
    if (!ctx || !ctx->have_pattern || !out8) return sc_fail(ctx, SC_ERR_STATE, "sc_build_pattern must be called first");
    out8[0] = ctx->nnz; out8[1] = ctx->ncol_total; out8[2] = ctx->n_nodes; out8[3] = ctx->max_rl; out8[4] = ctx->max_nbr;
    out8[5] = ctx->max_valence; out8[6] = la_node_usable(ctx) ? 1 : 0; out8[7] = ctx->n_dict;
    return SC_OK;
}

int sc_assemble(sc_ctx* ctx, int gauss_order, int flags, double* seconds_device) {
    if (!ctx || !ctx->have_pattern) return sc_fail(ctx, SC_ERR_STATE, "sc_build_pattern must be called first");
    if (!ctx->have_mat) return sc_fail(ctx, SC_ERR_STATE, "sc_set_materials must be called first");
    if (!(flags & (SC_ASM_K | SC_ASM_M_FULL | SC_ASM_M_LUMPED))) return sc_fail(ctx, SC_ERR_ARG, "nothing to assemble");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->cd_resume_valid = false; ctx->nm_resume_valid = false; ctx->khat_a1 = ctx->khat_a4 = -1.0; ctx->cd_coef_dt = -1.0; precond_drop(ctx);
    return sc_assemble_run(ctx, gauss_order, flags, seconds_device);
}

// (row, col) keys sorted and free of duplicates on the host, their values already on the device: slots by binary search in
// the structural pattern, then K[slot] += v, or C_abs := the row-compressed list (takes ownership of *d_v)
static int add_sorted_entries(sc_ctx* ctx, int which, int64_t n, const std::vector<int64_t>& r, const std::vector<int32_t>& c, double** d_v) {
    int64_t *d_r = nullptr, *d_slot = nullptr;
    int32_t* d_c = nullptr;
    int* d_flag = nullptr;
    int rc = SC_OK;
    auto body = [&]() -> int {
        SC_TRY(upload(ctx, &d_r, r.data(), (size_t)n));
        SC_TRY(upload(ctx, &d_c, c.data(), (size_t)n));
        SC_TRY(sc_alloc(ctx, &d_slot, (size_t)n));
        SC_TRY(sc_alloc(ctx, &d_flag, 1));
        SC_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream));
        k_find_slots<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_rowptr, ctx->d_col, d_r, d_c, n, d_slot, d_flag);
        SC_CHECK_LAUNCH(ctx);
        int flag = 0;
        SC_CUDA(ctx, cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (flag) return sc_fail(ctx, SC_ERR_ARG, "an added entry lies outside the structural pattern");
        if (which == SC_MAT_K) {
            if (!ctx->have_K) return sc_fail(ctx, SC_ERR_STATE, "K is not assembled");
            k_add_at<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_K, d_slot, *d_v, n);
            SC_CHECK_LAUNCH(ctx);
            SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            return SC_OK;
        }
        // C_abs replaces any previous list: row-compressed form
        std::vector<int64_t> rowid, rptr;
        for (int64_t i = 0; i < n; ++i) {
            if (i == 0 || r[i] != r[i - 1]) { rowid.push_back(r[i]); rptr.push_back(i); }
        }
        rptr.push_back(n);
        sc_free(&ctx->d_cabs_rowid); sc_free(&ctx->d_cabs_rptr); sc_free(&ctx->d_cabs_col); sc_free(&ctx->d_cabs_slot); sc_free(&ctx->d_cabs_val);
        ctx->cabs_n = ctx->cabs_rows = 0;
        SC_TRY(upload(ctx, &ctx->d_cabs_rowid, rowid.data(), rowid.size()));
        SC_TRY(upload(ctx, &ctx->d_cabs_rptr, rptr.data(), rptr.size()));
        ctx->d_cabs_col = d_c; d_c = nullptr;
        ctx->d_cabs_slot = d_slot; d_slot = nullptr;
        ctx->d_cabs_val = *d_v; *d_v = nullptr;
        ctx->cabs_n = n;
        ctx->cabs_rows = (int64_t)rowid.size();
        return SC_OK;
    };
    rc = body();
    sc_free(&d_r); sc_free(&d_c); sc_free(&d_slot); sc_free(&d_flag);
    ctx->cd_resume_valid = false; ctx->nm_resume_valid = false; ctx->khat_a1 = ctx->khat_a4 = -1.0; ctx->cd_coef_dt = -1.0; precond_drop(ctx);
    return rc;
}

int sc_add_entries(sc_ctx* ctx, int which, int64_t n, const int64_t* rows, const int64_t* cols, const double* vals) {
    if (!ctx || !ctx->have_pattern) return sc_fail(ctx, SC_ERR_STATE, "sc_build_pattern must be called first");
    if (which != SC_MAT_K && which != SC_MAT_C) return sc_fail(ctx, SC_ERR_ARG, "entries can be added to K or C only");
    if (n == 0) return SC_OK;
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    // sort by (row, col); duplicates are a contract violation (the caller pre-sums them in a fixed order)
    std::vector<int64_t> order((size_t)n);
    std::iota(order.begin(), order.end(), (int64_t)0);
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return rows[a] != rows[b] ? rows[a] < rows[b] : cols[a] < cols[b]; });
    std::vector<int64_t> r((size_t)n);
    std::vector<int32_t> c((size_t)n);
    std::vector<double> v((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t o = order[i];
        if (rows[o] < 0 || rows[o] >= ctx->n_eq || cols[o] < 0 || cols[o] >= ctx->n_eq) return sc_fail(ctx, SC_ERR_ARG, "entry %lld out of range", (long long)o);
        r[i] = rows[o]; c[i] = (int32_t)cols[o]; v[i] = vals[o];
        if (i > 0 && r[i] == r[i - 1] && c[i] == c[i - 1]) return sc_fail(ctx, SC_ERR_ARG, "duplicate entry (%lld, %lld)", (long long)r[i], (long long)c[i]);
    }
    double* d_v = nullptr;
    SC_TRY(upload(ctx, &d_v, v.data(), (size_t)n));
    const int rc = add_sorted_entries(ctx, which, n, r, c, &d_v);
    sc_free(&d_v);
    return rc;
}

int sc_add_absorbing_faces(sc_ctx* ctx, int face_type, int gauss_order, int64_t n_faces, const int32_t* face_nodes, const int32_t* face_elem,
                           const int32_t* face_dir, const uint8_t* perp, int64_t n_unique, const int64_t* rows, const int64_t* cols,
                           const int64_t* grp_ptr, const int64_t* grp_entry, double p0, double p1, double stiff) {
    if (!ctx || !ctx->have_pattern || !ctx->have_mat) return sc_fail(ctx, SC_ERR_STATE, "mesh, materials and pattern must be set first");
    if (!ctx->have_K) return sc_fail(ctx, SC_ERR_STATE, "K is not assembled");
    if (ctx->dim != 3) return sc_fail(ctx, SC_ERR_UNSUPPORTED, "absorbing faces exist for 3-D meshes only (system_matrix.py:324-326)");
    if (n_faces <= 0 || n_unique <= 0) return SC_OK;
    if (!face_nodes || !face_elem || !face_dir || !perp || !rows || !cols || !grp_ptr || !grp_entry || stiff == 0.0)
        return sc_fail(ctx, SC_ERR_ARG, "null argument");
    const int nl = sc_elem_nne(face_type);
    if (nl <= 0 || sc_elem_dim(face_type) != 2) return sc_fail(ctx, SC_ERR_ARG, "the face element must be a 2-D element type");
    const int64_t n_entries = n_faces * nl * nl;
    if (grp_ptr[0] != 0 || grp_ptr[n_unique] > n_entries) return sc_fail(ctx, SC_ERR_ARG, "bad entry grouping");
    std::vector<int64_t> r((size_t)n_unique);
    std::vector<int32_t> c((size_t)n_unique);
    for (int64_t i = 0; i < n_unique; ++i) {
        if (rows[i] < 0 || rows[i] >= ctx->n_eq || cols[i] < 0 || cols[i] >= ctx->n_eq) return sc_fail(ctx, SC_ERR_ARG, "key %lld out of range", (long long)i);
        if (grp_ptr[i + 1] < grp_ptr[i]) return sc_fail(ctx, SC_ERR_ARG, "bad entry grouping");
        r[i] = rows[i]; c[i] = (int32_t)cols[i];
        if (i > 0 && !(r[i] > r[i - 1] || (r[i] == r[i - 1] && c[i] > c[i - 1]))) return sc_fail(ctx, SC_ERR_ARG, "keys must be sorted and unique");
    }
    for (int64_t i = 0; i < n_faces; ++i)
        if (face_elem[i] < 0 || face_elem[i] >= ctx->n_elem || face_dir[i] < 0 || face_dir[i] > 2) return sc_fail(ctx, SC_ERR_ARG, "face %lld out of range", (long long)i);
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    double *d_c = nullptr, *d_k = nullptr;
    int rc = abs_faces_eval(ctx, face_type, gauss_order, n_faces, face_nodes, face_elem, face_dir, perp, n_unique, grp_ptr, grp_entry,
                            p0, p1, stiff, &d_c, &d_k);
    if (rc == SC_OK) rc = add_sorted_entries(ctx, SC_MAT_C, n_unique, r, c, &d_c);
    if (rc == SC_OK) rc = add_sorted_entries(ctx, SC_MAT_K, n_unique, r, c, &d_k);
    sc_free(&d_c); sc_free(&d_k);
    return rc;
}

int sc_set_rayleigh(sc_ctx* ctx, double c0, double c1) {
    if (!ctx) return SC_ERR_ARG;
    if (ctx->d_C && (c0 != 0.0 || c1 != 0.0)) return sc_fail(ctx, SC_ERR_UNSUPPORTED, "sc_set_csr supplied C explicitly; add Rayleigh terms to it before the upload");
    ctx->c0 = c0; ctx->c1 = c1;
    ctx->cd_resume_valid = false; ctx->nm_resume_valid = false; ctx->khat_a1 = ctx->khat_a4 = -1.0; ctx->cd_coef_dt = -1.0; precond_drop(ctx);
    return SC_OK;
}

int sc_get_values(sc_ctx* ctx, int which, double* vals) {
    if (!ctx || !vals) return sc_fail(ctx, SC_ERR_ARG, "null argument");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    const double* src = nullptr;
    double* tmp = nullptr;
    switch (which) {
        case SC_MAT_K: if (!ctx->have_K) return sc_fail(ctx, SC_ERR_STATE, "K is not assembled"); src = ctx->d_K; break;
        case SC_MAT_M: if (!ctx->have_M) return sc_fail(ctx, SC_ERR_STATE, "full M is not assembled"); src = ctx->d_M; break;
        case SC_MAT_C: SC_TRY(build_C(ctx, &tmp)); src = tmp; break;
        case SC_MAT_KHAT: if (!ctx->d_Khat) return sc_fail(ctx, SC_ERR_STATE, "no effective matrix yet"); src = ctx->d_Khat; break;
        default: return sc_fail(ctx, SC_ERR_ARG, "unknown matrix id %d", which);
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpy(vals, src, sizeof(double) * ctx->nnz, cudaMemcpyDeviceToHost);
    sc_free(&tmp);
    if (e != cudaSuccess) return sc_fail(ctx, SC_ERR_CUDA, "copy of matrix values failed: %s", cudaGetErrorString(e));
    return SC_OK;
}

int sc_get_lumped_mass(sc_ctx* ctx, double* diag) {
    if (!ctx || !ctx->have_Ml) return sc_fail(ctx, SC_ERR_STATE, "lumped mass is not assembled");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    SC_CUDA(ctx, cudaMemcpy(diag, ctx->d_Ml, sizeof(double) * ctx->n_eq, cudaMemcpyDeviceToHost));
    return SC_OK;
}

int sc_spmv(sc_ctx* ctx, int which, const double* x, double* y) {
    if (!ctx || !ctx->have_pattern) return sc_fail(ctx, SC_ERR_STATE, "no pattern");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    double *dx = nullptr, *dy = nullptr, *tmp = nullptr;
    const double* vals = nullptr;
    switch (which) {
        case SC_MAT_K: if (!ctx->have_K) return sc_fail(ctx, SC_ERR_STATE, "K is not assembled"); vals = ctx->d_K; break;
        case SC_MAT_M: if (!ctx->have_M) return sc_fail(ctx, SC_ERR_STATE, "full M is not assembled"); vals = ctx->d_M; break;
        case SC_MAT_C: SC_TRY(build_C(ctx, &tmp)); vals = tmp; break;
        case SC_MAT_KHAT: if (!ctx->d_Khat) return sc_fail(ctx, SC_ERR_STATE, "no effective matrix yet"); vals = ctx->d_Khat; break;
        default: return sc_fail(ctx, SC_ERR_ARG, "unknown matrix id %d", which);
    }
    int rc = sc_alloc(ctx, &dx, (size_t)ctx->n_eq);
    if (rc == SC_OK) rc = sc_alloc(ctx, &dy, (size_t)ctx->n_eq);
    if (rc == SC_OK && cudaMemcpy(dx, x, sizeof(double) * ctx->n_eq, cudaMemcpyHostToDevice) != cudaSuccess) rc = sc_fail(ctx, SC_ERR_CUDA, "H2D failed");
    if (rc == SC_OK) rc = la_spmv(ctx, vals, dx, dy);
    if (rc == SC_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = sc_fail(ctx, SC_ERR_CUDA, "spmv failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc == SC_OK && cudaMemcpy(y, dy, sizeof(double) * ctx->n_eq, cudaMemcpyDeviceToHost) != cudaSuccess) rc = sc_fail(ctx, SC_ERR_CUDA, "D2H failed");
    sc_free(&dx); sc_free(&dy); sc_free(&tmp);
    return rc;
}

int sc_set_load_schedule(sc_ctx* ctx, int64_t n_steps, const int64_t* step_ptr, const int64_t* dof, const double* val) {
    if (!ctx || !(ctx->have_mesh || ctx->csr_only)) return sc_fail(ctx, SC_ERR_STATE, "sc_set_mesh (or sc_set_csr) must be called first");
    if (n_steps < 0 || (n_steps > 0 && !step_ptr)) return sc_fail(ctx, SC_ERR_ARG, "bad load schedule");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->load_steps = n_steps;
    ctx->h_load_ptr.assign(step_ptr, step_ptr + n_steps + 1);
    const int64_t n = n_steps > 0 ? step_ptr[n_steps] : 0;
    std::vector<int32_t> d32((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        if (dof[i] < 0 || dof[i] >= ctx->n_eq) return sc_fail(ctx, SC_ERR_ARG, "load dof %lld out of range", (long long)dof[i]);
        d32[i] = (int32_t)dof[i];
    }
    SC_TRY(upload(ctx, &ctx->d_load_dof, d32.data(), (size_t)n));
    SC_TRY(upload(ctx, &ctx->d_load_val, val, (size_t)n));
    return SC_OK;
}

int sc_set_state(sc_ctx* ctx, const double* u, const double* v) {
    if (!ctx || !(ctx->have_mesh || ctx->csr_only)) return sc_fail(ctx, SC_ERR_STATE, "sc_set_mesh (or sc_set_csr) must be called first");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->n_eq;
    for (int k = 0; k < 3; ++k) {
        double** d = k == 0 ? &ctx->d_u : k == 1 ? &ctx->d_v : &ctx->d_a;
        const double* h = k == 0 ? u : k == 1 ? v : nullptr;
        if (!*d) SC_TRY(sc_alloc(ctx, d, n));
        if (h) SC_CUDA(ctx, cudaMemcpy(*d, h, n * sizeof(double), cudaMemcpyHostToDevice));
        else SC_CUDA(ctx, cudaMemset(*d, 0, n * sizeof(double)));
    }
    ctx->cd_resume_valid = false; ctx->nm_resume_valid = false;
    return SC_OK;
}

// Caller-supplied matrices (seam B2 of SURVEY.md 8b: solvers.*.calculate(M, C, K, F, t0, t1) with scipy matrices that were
// not assembled by this library, scatter/scatter.py:159).  One CSR pattern, three value arrays.
int sc_set_csr(sc_ctx* ctx, int64_t n_eq, const int64_t* rowptr, const int32_t* col, const double* K, const double* M, const double* C) {
    if (!ctx || n_eq <= 0 || !rowptr || !col || !K) return sc_fail(ctx, SC_ERR_ARG, "sc_set_csr needs n_eq > 0, a pattern and K");
    if (n_eq >= (int64_t)1 << 31) return sc_fail(ctx, SC_ERR_UNSUPPORTED, "more than 2^31 equations per rank");
    if (rowptr[0] != 0) return sc_fail(ctx, SC_ERR_ARG, "rowptr[0] must be 0");
    int max_rl = 0;
    for (int64_t i = 0; i < n_eq; ++i) {
        const int64_t len = rowptr[i + 1] - rowptr[i];
        if (len < 0 || len > 65535) return sc_fail(ctx, SC_ERR_ARG, "row %lld has a bad length %lld", (long long)i, (long long)len);
        max_rl = std::max(max_rl, (int)len);
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            if (col[k] < 0 || col[k] >= n_eq) return sc_fail(ctx, SC_ERR_ARG, "column index out of range in row %lld", (long long)i);
            if (k > rowptr[i] && col[k] <= col[k - 1]) return sc_fail(ctx, SC_ERR_ARG, "columns of row %lld are not sorted / unique", (long long)i);
        }
    }
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    free_pattern(ctx);
    free_vectors(ctx);
    sc_free(&ctx->d_xyz); sc_free(&ctx->d_conn); sc_free(&ctx->d_eq); sc_free(&ctx->d_active);
    ctx->have_mesh = false; ctx->have_mat = false;
    ctx->elem_type = -1; ctx->nne = 0; ctx->dim = 0; ctx->n_nodes = 0; ctx->n_elem = 0;
    ctx->n_eq = n_eq; ctx->nnz = rowptr[n_eq]; ctx->max_rl = max_rl;
    ctx->c0 = ctx->c1 = 0.0;
    SC_TRY(upload(ctx, &ctx->d_rowptr, rowptr, (size_t)n_eq + 1));
    SC_TRY(upload(ctx, &ctx->d_col, col, (size_t)ctx->nnz));
    SC_TRY(upload(ctx, &ctx->d_K, K, (size_t)ctx->nnz));
    ctx->have_pattern = true; ctx->have_K = true; ctx->csr_only = true;
    if (M) {
        SC_TRY(upload(ctx, &ctx->d_M, M, (size_t)ctx->nnz));
        ctx->have_M = true;
        // row sums of M = lumped mass of the explicit scheme
        double* ones = nullptr;
        SC_TRY(sc_alloc(ctx, &ones, (size_t)n_eq));
        SC_TRY(sc_alloc(ctx, &ctx->d_Ml, (size_t)n_eq));
        int rc = la_fill(ctx, ones, 1.0, n_eq);
        if (rc == SC_OK) rc = la_spmv(ctx, ctx->d_M, ones, ctx->d_Ml);
        if (rc == SC_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = sc_fail(ctx, SC_ERR_CUDA, "lumping M failed");
        sc_free(&ones);
        SC_TRY(rc);
        ctx->have_Ml = true;
    }
    if (C) SC_TRY(upload(ctx, &ctx->d_C, C, (size_t)ctx->nnz));
    return SC_OK;
}

// Only these equations are copied to the host per output row (n = 0 .. n_eq, sorted or not; NULL restores full rows).
int sc_set_output_dofs(sc_ctx* ctx, int64_t n, const int64_t* dofs) {
    if (!ctx) return SC_ERR_ARG;
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->rows_pending) { cudaStreamSynchronize(ctx->copy_stream); ctx->rows_pending = false; }
    sc_free(&ctx->d_sel);
    for (int k = 0; k < 3; ++k) sc_free(&ctx->d_selbuf[k]);
    ctx->n_sel = -1;
    if (!dofs) return SC_OK;
    if (n < 0) return sc_fail(ctx, SC_ERR_ARG, "negative count");
    std::vector<int32_t> d32((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        if (dofs[i] < 0 || dofs[i] >= ctx->n_eq) return sc_fail(ctx, SC_ERR_ARG, "output dof %lld out of range", (long long)dofs[i]);
        d32[i] = (int32_t)dofs[i];
    }
    SC_TRY(upload(ctx, &ctx->d_sel, d32.data(), (size_t)n));
    ctx->n_sel = n;
    return SC_OK;
}

int sc_set_final_output_step(sc_ctx* ctx, int64_t step) {
    if (!ctx) return SC_ERR_ARG;
    ctx->extra_out_step = step < 0 ? -1 : step;
    return SC_OK;
}

int sc_get_state(sc_ctx* ctx, double* u, double* v, double* a) {
    if (!ctx || !ctx->d_u) return sc_fail(ctx, SC_ERR_STATE, "no state yet");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t b = sizeof(double) * ctx->n_eq;
    if (u) SC_CUDA(ctx, cudaMemcpy(u, ctx->d_u, b, cudaMemcpyDeviceToHost));
    if (v) SC_CUDA(ctx, cudaMemcpy(v, ctx->d_v, b, cudaMemcpyDeviceToHost));
    if (a) SC_CUDA(ctx, cudaMemcpy(a, ctx->d_a, b, cudaMemcpyDeviceToHost));
    return SC_OK;
}

int sc_run_newmark(sc_ctx* ctx, double dt, int64_t t_start, int64_t n_steps, int64_t out_interval, double beta, double gamma,
                   double pcg_rtol, int pcg_maxit, int64_t n_out, double* u_out, double* v_out, double* a_out, sc_stats* stats) {
    if (!ctx || !ctx->have_K || !ctx->have_M) return sc_fail(ctx, SC_ERR_STATE, "Newmark needs assembled K and full M");
    if (dt <= 0 || n_steps < 0 || out_interval < 1 || beta <= 0) return sc_fail(ctx, SC_ERR_ARG, "bad time-integration arguments");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    if (stats) std::memset(stats, 0, sizeof(*stats));
    ctx->cd_resume_valid = false; ctx->cd_coef_dt = -1.0;   // Newmark reuses the work vectors of the cached coefficients
    if (beta != 0.25 || gamma != 0.5) { ctx->khat_a1 = -1.0; precond_drop(ctx); }   // (a1, a4) identify the matrix only together with beta, gamma
    return tl_newmark(ctx, dt, t_start, n_steps, out_interval, beta, gamma, pcg_rtol, pcg_maxit, n_out, u_out, v_out, a_out, stats);
}

int sc_run_central_difference(sc_ctx* ctx, double dt, int64_t t_start, int64_t n_steps, int64_t out_interval, int64_t n_out,
                              double* u_out, double* v_out, double* a_out, sc_stats* stats) {
    if (!ctx || !ctx->have_K || !ctx->have_Ml) return sc_fail(ctx, SC_ERR_STATE, "central difference needs assembled K and lumped M");
    if (dt <= 0 || n_steps < 0 || out_interval < 1) return sc_fail(ctx, SC_ERR_ARG, "bad time-integration arguments");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    if (stats) std::memset(stats, 0, sizeof(*stats));
    ctx->nm_resume_valid = false;
    return tl_central_difference(ctx, dt, t_start, n_steps, out_interval, n_out, u_out, v_out, a_out, stats);
}

int sc_run_bathe(sc_ctx* ctx, double dt, int64_t t_start, int64_t n_steps, int64_t out_interval, double pcg_rtol, int pcg_maxit,
                 int64_t n_out, double* u_out, double* v_out, double* a_out, sc_stats* stats) {
    if (!ctx || !ctx->have_K || !ctx->have_M) return sc_fail(ctx, SC_ERR_STATE, "Bathe needs assembled K and full M");
    if (dt <= 0 || n_steps < 0 || out_interval < 1) return sc_fail(ctx, SC_ERR_ARG, "bad time-integration arguments");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    if (stats) std::memset(stats, 0, sizeof(*stats));
    ctx->cd_resume_valid = false; ctx->nm_resume_valid = false; ctx->khat_a1 = ctx->khat_a4 = -1.0; ctx->cd_coef_dt = -1.0; precond_drop(ctx);
    return tl_bathe(ctx, dt, t_start, n_steps, out_interval, pcg_rtol, pcg_maxit, n_out, u_out, v_out, a_out, stats);
}

int sc_run_static(sc_ctx* ctx, int64_t t_start, int64_t n_steps, int64_t out_interval, double pcg_rtol, int pcg_maxit, int64_t n_out,
                  double* u_out, sc_stats* stats) {
    if (!ctx || !ctx->have_K) return sc_fail(ctx, SC_ERR_STATE, "the static solver needs assembled K");
    if (n_steps < 0 || out_interval < 1) return sc_fail(ctx, SC_ERR_ARG, "bad arguments");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    if (stats) std::memset(stats, 0, sizeof(*stats));
    ctx->cd_resume_valid = false; ctx->nm_resume_valid = false; ctx->cd_coef_dt = -1.0;
    return tl_static(ctx, t_start, n_steps, out_interval, pcg_rtol, pcg_maxit, n_out, u_out, stats);
}

int sc_srf_sample(sc_ctx* ctx, int64_t n_points, const double* pos, int n_modes, const double* k, const double* z1, const double* z2,
                  double scale, double mean, int lognormal, double* out, double* seconds_device) {
    if (!ctx) return SC_ERR_ARG;
    if (n_points < 0 || n_modes <= 0 || n_modes > (1 << 20)) return sc_fail(ctx, SC_ERR_ARG, "bad random-field sizes (%lld points, %d modes)", (long long)n_points, n_modes);
    if (seconds_device) *seconds_device = 0.0;
    if (n_points == 0) return SC_OK;
    if (!pos || !k || !z1 || !z2 || !out) return sc_fail(ctx, SC_ERR_ARG, "null argument");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    return srf_sample(ctx, n_points, pos, n_modes, k, z1, z2, scale, mean, lognormal, out, seconds_device);
}

int sc_nccl_unique_id(void* out128) { return out128 ? dist_unique_id(out128) : SC_ERR_ARG; }

int sc_dist_init(sc_ctx* ctx, int rank, int world, const void* id) {
    if (!ctx || (world > 1 && !id) || rank < 0 || rank >= world) return sc_fail(ctx, SC_ERR_ARG, "bad rank/world");
    return dist_init(ctx, rank, world, id);
}

int sc_set_halo(sc_ctx* ctx, int n_neighbors, const int32_t* neighbor_rank, const int64_t* send_ptr, const int64_t* send_idx,
                const int64_t* recv_ptr, const int64_t* recv_idx) {
    if (!ctx || !ctx->have_mesh) return sc_fail(ctx, SC_ERR_STATE, "sc_set_mesh must be called first");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    dist_peer_release(ctx);
    ctx->n_nbr_ranks = n_neighbors;
    ctx->ov_planned = ctx->ov_ok = false;
    ctx->nbr_rank.assign(neighbor_rank, neighbor_rank + n_neighbors);
    ctx->send_ptr.assign(send_ptr, send_ptr + n_neighbors + 1);
    ctx->recv_ptr.assign(recv_ptr, recv_ptr + n_neighbors + 1);
    const int64_t ns = ctx->send_ptr.back(), nr = ctx->recv_ptr.back();
    for (int64_t i = 0; i < ns; ++i) if (send_idx[i] < 0 || send_idx[i] >= ctx->n_eq) return sc_fail(ctx, SC_ERR_ARG, "send index out of range");
    for (int64_t i = 0; i < nr; ++i) if (recv_idx[i] < 0 || recv_idx[i] >= ctx->n_eq) return sc_fail(ctx, SC_ERR_ARG, "recv index out of range");
    SC_TRY(upload(ctx, &ctx->d_send_idx, send_idx, (size_t)ns));
    SC_TRY(upload(ctx, &ctx->d_recv_idx, recv_idx, (size_t)nr));
    SC_TRY(sc_alloc(ctx, &ctx->d_send_buf, (size_t)ns));
    SC_TRY(sc_alloc(ctx, &ctx->d_recv_buf, (size_t)nr));
    return SC_OK;
}

int sc_halo_exchange(sc_ctx* ctx, double* x_host) {
    if (!ctx || !ctx->have_mesh || !x_host) return sc_fail(ctx, SC_ERR_ARG, "bad argument");
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    double* d = nullptr;
    SC_TRY(sc_alloc(ctx, &d, (size_t)ctx->n_eq));
    int rc = SC_OK;
    if (cudaMemcpy(d, x_host, sizeof(double) * ctx->n_eq, cudaMemcpyHostToDevice) != cudaSuccess) rc = sc_fail(ctx, SC_ERR_CUDA, "H2D failed");
    if (rc == SC_OK) rc = dist_halo(ctx, d, ctx->stream);
    if (rc == SC_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = sc_fail(ctx, SC_ERR_CUDA, "halo exchange failed");
    if (rc == SC_OK && cudaMemcpy(x_host, d, sizeof(double) * ctx->n_eq, cudaMemcpyDeviceToHost) != cudaSuccess) rc = sc_fail(ctx, SC_ERR_CUDA, "D2H failed");
    sc_free(&d);
    return rc;
}

}  // extern "C"
