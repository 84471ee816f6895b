// Structural sparsity pattern on the GPU.
//
// The reference builds the pattern implicitly as the key set of a Python dict filled with every free (i,k) pair of
// every element (scatter/system_matrix.py:90-103) and converts it with coo_matrix(...).tolil() (:120-121).  Here the
// same set is built at node level: two dofs couple iff their nodes share an element, so
//     row(a,i) = concat over neighbours b of a (ascending node row) of the free dofs of b,
// which is bit-identical to the sorted CSR of that key set because the reference's equation numbers increase with
// (node row, dof) (scatter/mesher.py:289-309).
//
// Stages: node->element lists (ascending element id) -> node->node lists (ascending, incl. self) -> per-node row
// length -> int64 row pointer (CUB scan) -> int32 column fill.  All integer work; integer atomics are used only for
// counting / slot reservation and every list is sorted afterwards, so the result is deterministic.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_reduce.cuh>
#include "common.h"

namespace {

__global__ void k_count_node_elems(const int32_t* __restrict__ conn, int64_t n_elem, int nne, int* __restrict__ cnt) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_elem * nne) return;
    atomicAdd(&cnt[conn[i]], 1);
}

__global__ void k_fill_node_elems(const int32_t* __restrict__ conn, int64_t n_elem, int nne, const int64_t* __restrict__ ptr,
                                  int* __restrict__ cursor, int32_t* __restrict__ n2e) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_elem * nne) return;
    int node = conn[i];
    int pos = atomicAdd(&cursor[node], 1);
    n2e[ptr[node] + pos] = (int32_t)(i / nne);
}

// ascending insertion sort of every node's (short) element list; also removes nothing: an element lists a node once
__global__ void k_sort_node_elems(const int64_t* __restrict__ ptr, int32_t* __restrict__ n2e, int64_t n_nodes) {
    int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (a >= n_nodes) return;
    int64_t s = ptr[a], e = ptr[a + 1];
    for (int64_t i = s + 1; i < e; ++i) {
        int32_t v = n2e[i];
        int64_t j = i - 1;
        while (j >= s && n2e[j] > v) { n2e[j + 1] = n2e[j]; --j; }
        n2e[j + 1] = v;
    }
}

// Pass 1 (FILL=false): count unique neighbours.  Pass 2 (FILL=true): write them (ascending) to nbr[nbr_ptr[a]..].
// The sorted unique set is kept in a per-thread local array (interleaved local memory => coalesced across threads).
template <int CAP, bool FILL>
__global__ void k_node_neighbours(const int32_t* __restrict__ conn, int nne, const int64_t* __restrict__ n2e_ptr,
                                  const int32_t* __restrict__ n2e, int64_t n_nodes, const uint8_t* __restrict__ active,
                                  int* __restrict__ count, const int64_t* __restrict__ nbr_ptr, int32_t* __restrict__ nbr,
                                  int* __restrict__ overflow) {
    int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (a >= n_nodes) return;
    if (active && !active[a]) {
        if (!FILL) count[a] = 0;
        return;
    }
    int32_t set[CAP];
    int n = 0;
    for (int64_t k = n2e_ptr[a]; k < n2e_ptr[a + 1]; ++k) {
        const int32_t* c = conn + (int64_t)n2e[k] * nne;
        for (int b = 0; b < nne; ++b) {
            int32_t v = c[b];
            // binary search for the insertion point
            int lo = 0, hi = n;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (set[mid] < v) lo = mid + 1; else hi = mid;
            }
            if (lo < n && set[lo] == v) continue;
            if (n >= CAP) { atomicExch(overflow, 1); return; }
            for (int j = n; j > lo; --j) set[j] = set[j - 1];
            set[lo] = v;
            ++n;
        }
    }
    if (!FILL) {
        count[a] = n;
    } else {
        int64_t o = nbr_ptr[a];
        for (int j = 0; j < n; ++j) nbr[o + j] = set[j];
    }
}

__global__ void k_node_free(const int32_t* __restrict__ eq, int dim, int64_t n_nodes, int* __restrict__ nfree) {
    int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (a >= n_nodes) return;
    int c = 0;
    for (int d = 0; d < dim; ++d) c += eq[a * dim + d] >= 0;
    nfree[a] = c;
}

// per node: dof offset of every neighbour inside the node's rows and the row length
__global__ void k_node_rowlen(const int64_t* __restrict__ nbr_ptr, const int32_t* __restrict__ nbr, const int* __restrict__ nfree,
                              const int32_t* __restrict__ eq, int dim, int64_t n_nodes, uint16_t* __restrict__ nbr_off,
                              uint8_t* __restrict__ nbr_free, int32_t* __restrict__ node_rl, int* __restrict__ overflow) {
    int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (a >= n_nodes) return;
    int off = 0;
    for (int64_t k = nbr_ptr[a]; k < nbr_ptr[a + 1]; ++k) {
        if (off > 65535) { atomicExch(overflow, 2); break; }
        nbr_off[k] = (uint16_t)off;
        const int b = nbr[k];
        int mask = 0;
        for (int j = 0; j < dim; ++j) mask |= (eq[(int64_t)b * dim + j] >= 0) << j;
        nbr_free[k] = (uint8_t)mask;
        off += nfree[b];
    }
    node_rl[a] = off;
}

// row length per equation: every free dof of node a gets node_rl[a]
__global__ void k_row_lengths(const int32_t* __restrict__ eq, int dim, const int32_t* __restrict__ node_rl, int64_t n_nodes,
                              int64_t* __restrict__ rowlen) {
    int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (a >= n_nodes) return;
    for (int d = 0; d < dim; ++d) {
        int r = eq[a * dim + d];
        if (r >= 0) rowlen[r] = node_rl[a];
    }
}

// one warp per node: lanes stride over the neighbours and write the free dofs of each into all rows of the node
__global__ void k_fill_columns(const int32_t* __restrict__ eq, int dim, const int64_t* __restrict__ nbr_ptr,
                               const int32_t* __restrict__ nbr, const uint16_t* __restrict__ nbr_off,
                               const int64_t* __restrict__ rowptr, int64_t n_nodes, int32_t* __restrict__ col) {
    int64_t a = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (a >= n_nodes) return;
    int64_t s = nbr_ptr[a], e = nbr_ptr[a + 1];
    if (s == e) return;
    for (int i = 0; i < dim; ++i) {
        int r = eq[a * dim + i];
        if (r < 0) continue;
        int64_t base = rowptr[r];
        for (int64_t k = s + lane; k < e; k += 32) {
            int b = nbr[k];
            int64_t o = base + nbr_off[k];
            for (int j = 0; j < dim; ++j) {
                int c = eq[(int64_t)b * dim + j];
                if (c >= 0) col[o++] = c;
            }
        }
    }
}

// node-blocked structure: one column list per node (shared by all its rows) + a 24-byte descriptor per node
__global__ void k_node_columns(const int32_t* __restrict__ eq, int dim, const int64_t* __restrict__ nbr_ptr,
                               const int32_t* __restrict__ nbr, const uint16_t* __restrict__ nbr_off,
                               const int64_t* __restrict__ ncol_ptr, int64_t n_nodes, int32_t* __restrict__ ncol) {
    int64_t a = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (a >= n_nodes) return;
    const int64_t s = nbr_ptr[a], e = nbr_ptr[a + 1];
    const int64_t base = ncol_ptr[a];
    for (int64_t k = s + lane; k < e; k += 32) {
        const int b = nbr[k];
        int64_t o = base + nbr_off[k];
        for (int j = 0; j < dim; ++j) {
            const int c = eq[(int64_t)b * dim + j];
            if (c >= 0) ncol[o++] = c;
        }
    }
}
__global__ void k_node_desc(const int64_t* __restrict__ node_row0, const int32_t* __restrict__ node_rl, const int64_t* __restrict__ rowptr,
                            const int64_t* __restrict__ ncol_ptr, int64_t n_nodes, int64_t n_pad, sc_ctx::NodeDesc* __restrict__ nd) {
    int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (a >= n_pad) return;
    sc_ctx::NodeDesc d;
    if (a < n_nodes) {
        const int64_t r0 = node_row0[a];
        const int nfree = (int)(node_row0[a + 1] - r0);
        d.row0 = (int32_t)r0;
        d.val_off = rowptr[r0];
        d.col_off = ncol_ptr[a];
        d.len_nfree = nfree > 0 ? (node_rl[a] | (nfree << 24)) : 0;     // fully fixed nodes own no rows: nothing to stream
    } else {   // padding: empty nodes at the end of the value / column arrays
        d.row0 = (int32_t)node_row0[n_nodes];
        d.val_off = rowptr[node_row0[n_nodes]];
        d.col_off = ncol_ptr[n_nodes];
        d.len_nfree = 0;
    }
    nd[a] = d;
}

// pair map for the assembly kernel: for pair k = (node a, element e) the position of every node of e in the (ascending)
// neighbour list of a, and the local index of a in e
__global__ void k_pair_map(const int32_t* __restrict__ conn, int nne, const int64_t* __restrict__ n2e_ptr, const int32_t* __restrict__ n2e,
                           const int64_t* __restrict__ nbr_ptr, const int32_t* __restrict__ nbr, int64_t n_nodes,
                           uint8_t* __restrict__ pos, uint8_t* __restrict__ al) {
    int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (a >= n_nodes) return;
    const int64_t nb0 = nbr_ptr[a];
    const int nn = (int)(nbr_ptr[a + 1] - nb0);
    for (int64_t k = n2e_ptr[a]; k < n2e_ptr[a + 1]; ++k) {
        const int32_t* c = conn + (int64_t)n2e[k] * nne;
        for (int b = 0; b < nne; ++b) {
            const int v = c[b];
            int lo = 0, hi = nn;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (nbr[nb0 + mid] < v) lo = mid + 1; else hi = mid;
            }
            pos[k * nne + b] = (uint8_t)lo;
            if (v == (int)a) al[k] = (uint8_t)b;
        }
    }
}

__global__ void k_i32_to_i64(const int* __restrict__ in, int64_t* __restrict__ out, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

// exclusive scan of n int64 values into out[0..n] (out[n] = total); in may alias a temp buffer
int scan64(sc_ctx* ctx, const int64_t* d_in, int64_t* d_out, int64_t n) {
    // CUB exclusive sum over n+1 items: the extra trailing input item is ignored by writing total at out[n]
    size_t bytes = 0;
    SC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, bytes, d_in, d_out, n + 1, ctx->stream));
    void* tmp = nullptr;
    SC_CUDA(ctx, cudaMalloc(&tmp, bytes ? bytes : 1));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, bytes, d_in, d_out, n + 1, ctx->stream);
    ctx->launches += 2;
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return sc_fail(ctx, SC_ERR_CUDA, "cub scan failed: %s", cudaGetErrorString(e));
    if (e2 != cudaSuccess) return sc_fail(ctx, SC_ERR_CUDA, "cub scan sync failed: %s", cudaGetErrorString(e2));
    return SC_OK;
}

int max_i32(sc_ctx* ctx, const int* d_in, int64_t n, int* h_out) {
    int* d_out = nullptr;
    SC_CUDA(ctx, cudaMalloc(&d_out, sizeof(int)));
    size_t bytes = 0;
    cub::DeviceReduce::Max(nullptr, bytes, d_in, d_out, n, ctx->stream);
    void* tmp = nullptr;
    SC_CUDA(ctx, cudaMalloc(&tmp, bytes ? bytes : 1));
    cudaError_t e = cub::DeviceReduce::Max(tmp, bytes, d_in, d_out, n, ctx->stream);
    ctx->launches += 1;
    cudaMemcpyAsync(h_out, d_out, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    cudaFree(d_out);
    if (e != cudaSuccess || e2 != cudaSuccess) return sc_fail(ctx, SC_ERR_CUDA, "cub max failed");
    return SC_OK;
}

}  // namespace

int sc_pattern_build(sc_ctx* ctx) {
    const int64_t nn = ctx->n_nodes, ne = ctx->n_elem;
    const int nne = ctx->nne, dim = ctx->dim;
    cudaStream_t st = ctx->stream;
    const int T = 256;

    int* d_cnt = nullptr;       // int32 [nn+1] scratch counters
    int64_t* d_tmp64 = nullptr; // int64 [max(nn, n_eq)+1] scan input
    int* d_flag = nullptr;
    SC_TRY(sc_alloc(ctx, &d_cnt, (size_t)nn + 1));
    SC_TRY(sc_alloc(ctx, &d_tmp64, (size_t)std::max(nn, ctx->n_eq) + 1));
    SC_TRY(sc_alloc(ctx, &d_flag, 1));
    SC_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), st));

    // ---- node -> elements -------------------------------------------------------------------------------------
    SC_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, sizeof(int) * (nn + 1), st));
    k_count_node_elems<<<nblk(ne * nne, T), T, 0, st>>>(ctx->d_conn, ne, nne, d_cnt);
    SC_CHECK_LAUNCH(ctx);
    SC_TRY(max_i32(ctx, d_cnt, nn, &ctx->max_valence));
    k_i32_to_i64<<<nblk(nn + 1, T), T, 0, st>>>(d_cnt, d_tmp64, nn + 1);
    SC_CHECK_LAUNCH(ctx);
    SC_TRY(sc_alloc(ctx, &ctx->d_n2e_ptr, (size_t)nn + 1));
    SC_TRY(scan64(ctx, d_tmp64, ctx->d_n2e_ptr, nn));
    SC_TRY(sc_alloc(ctx, &ctx->d_n2e, (size_t)ne * nne));
    SC_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, sizeof(int) * (nn + 1), st));
    k_fill_node_elems<<<nblk(ne * nne, T), T, 0, st>>>(ctx->d_conn, ne, nne, ctx->d_n2e_ptr, d_cnt, ctx->d_n2e);
    SC_CHECK_LAUNCH(ctx);
    k_sort_node_elems<<<nblk(nn, T), T, 0, st>>>(ctx->d_n2e_ptr, ctx->d_n2e, nn);
    SC_CHECK_LAUNCH(ctx);

    // ---- node -> nodes ------------------------------------------------------------------------------------------
    // capacity tiers for the per-thread sorted set
    int tier = 0;
    for (;; ++tier) {
        if (tier == 3) { sc_free(&d_cnt); sc_free(&d_tmp64); sc_free(&d_flag);
            return sc_fail(ctx, SC_ERR_UNSUPPORTED, "a node has more than 1024 neighbour nodes"); }
        SC_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), st));
        if (tier == 0) k_node_neighbours<64, false><<<nblk(nn, 128), 128, 0, st>>>(ctx->d_conn, nne, ctx->d_n2e_ptr, ctx->d_n2e, nn, ctx->d_active, d_cnt, nullptr, nullptr, d_flag);
        if (tier == 1) k_node_neighbours<256, false><<<nblk(nn, 128), 128, 0, st>>>(ctx->d_conn, nne, ctx->d_n2e_ptr, ctx->d_n2e, nn, ctx->d_active, d_cnt, nullptr, nullptr, d_flag);
        if (tier == 2) k_node_neighbours<1024, false><<<nblk(nn, 64), 64, 0, st>>>(ctx->d_conn, nne, ctx->d_n2e_ptr, ctx->d_n2e, nn, ctx->d_active, d_cnt, nullptr, nullptr, d_flag);
        SC_CHECK_LAUNCH(ctx);
        int flag = 0;
        SC_CUDA(ctx, cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        SC_CUDA(ctx, cudaStreamSynchronize(st));
        if (!flag) break;
    }
    SC_CUDA(ctx, cudaMemsetAsync(d_cnt + nn, 0, sizeof(int), st));
    SC_TRY(max_i32(ctx, d_cnt, nn, &ctx->max_nbr));
    k_i32_to_i64<<<nblk(nn + 1, T), T, 0, st>>>(d_cnt, d_tmp64, nn + 1);
    SC_CHECK_LAUNCH(ctx);
    SC_TRY(sc_alloc(ctx, &ctx->d_nbr_ptr, (size_t)nn + 1));
    SC_TRY(scan64(ctx, d_tmp64, ctx->d_nbr_ptr, nn));
    int64_t total_nbr = 0;
    SC_CUDA(ctx, cudaMemcpy(&total_nbr, ctx->d_nbr_ptr + nn, sizeof(int64_t), cudaMemcpyDeviceToHost));
    SC_TRY(sc_alloc(ctx, &ctx->d_nbr, (size_t)total_nbr));
    SC_TRY(sc_alloc(ctx, &ctx->d_nbr_off, (size_t)total_nbr));
    SC_TRY(sc_alloc(ctx, &ctx->d_nbr_free, (size_t)total_nbr));
    if (tier == 0) k_node_neighbours<64, true><<<nblk(nn, 128), 128, 0, st>>>(ctx->d_conn, nne, ctx->d_n2e_ptr, ctx->d_n2e, nn, ctx->d_active, nullptr, ctx->d_nbr_ptr, ctx->d_nbr, d_flag);
    if (tier == 1) k_node_neighbours<256, true><<<nblk(nn, 128), 128, 0, st>>>(ctx->d_conn, nne, ctx->d_n2e_ptr, ctx->d_n2e, nn, ctx->d_active, nullptr, ctx->d_nbr_ptr, ctx->d_nbr, d_flag);
    if (tier == 2) k_node_neighbours<1024, true><<<nblk(nn, 64), 64, 0, st>>>(ctx->d_conn, nne, ctx->d_n2e_ptr, ctx->d_n2e, nn, ctx->d_active, nullptr, ctx->d_nbr_ptr, ctx->d_nbr, d_flag);
    SC_CHECK_LAUNCH(ctx);

    // ---- dof-level CSR ----------------------------------------------------------------------------------------
    int* d_nfree = d_cnt;   // reuse
    k_node_free<<<nblk(nn, T), T, 0, st>>>(ctx->d_eq, dim, nn, d_nfree);
    SC_CHECK_LAUNCH(ctx);
    SC_CUDA(ctx, cudaMemsetAsync(d_nfree + nn, 0, sizeof(int), st));
    // first row of every node = exclusive scan of the free-dof counts
    k_i32_to_i64<<<nblk(nn + 1, T), T, 0, st>>>(d_nfree, d_tmp64, nn + 1);
    SC_CHECK_LAUNCH(ctx);
    SC_TRY(sc_alloc(ctx, &ctx->d_node_row0, (size_t)nn + 1));
    SC_TRY(scan64(ctx, d_tmp64, ctx->d_node_row0, nn));
    int64_t nfree_total = 0;
    SC_CUDA(ctx, cudaMemcpy(&nfree_total, ctx->d_node_row0 + nn, sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (nfree_total != ctx->n_eq) {
        sc_free(&d_cnt); sc_free(&d_tmp64); sc_free(&d_flag);
        return sc_fail(ctx, SC_ERR_ARG, "eq table has %lld free dofs but n_eq = %lld", (long long)nfree_total, (long long)ctx->n_eq);
    }
    SC_TRY(sc_alloc(ctx, &ctx->d_node_rl, (size_t)nn));
    SC_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    k_node_rowlen<<<nblk(nn, T), T, 0, st>>>(ctx->d_nbr_ptr, ctx->d_nbr, d_nfree, ctx->d_eq, dim, nn, ctx->d_nbr_off, ctx->d_nbr_free,
                                              ctx->d_node_rl, d_flag);
    SC_CHECK_LAUNCH(ctx);
    int flag = 0;
    SC_CUDA(ctx, cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    SC_CUDA(ctx, cudaStreamSynchronize(st));
    if (flag) { sc_free(&d_cnt); sc_free(&d_tmp64); sc_free(&d_flag);
        return sc_fail(ctx, SC_ERR_UNSUPPORTED, "a matrix row is longer than 65535 entries"); }
    SC_TRY(max_i32(ctx, ctx->d_node_rl, nn, &ctx->max_rl));
    SC_CUDA(ctx, cudaMemsetAsync(d_tmp64, 0, sizeof(int64_t) * (ctx->n_eq + 1), st));
    k_row_lengths<<<nblk(nn, T), T, 0, st>>>(ctx->d_eq, dim, ctx->d_node_rl, nn, d_tmp64);
    SC_CHECK_LAUNCH(ctx);
    SC_TRY(sc_alloc(ctx, &ctx->d_rowptr, (size_t)ctx->n_eq + 1));
    SC_TRY(scan64(ctx, d_tmp64, ctx->d_rowptr, ctx->n_eq));
    SC_CUDA(ctx, cudaMemcpy(&ctx->nnz, ctx->d_rowptr + ctx->n_eq, sizeof(int64_t), cudaMemcpyDeviceToHost));
    SC_TRY(sc_alloc(ctx, &ctx->d_col, (size_t)ctx->nnz));
    k_fill_columns<<<nblk(nn * 32, T), T, 0, st>>>(ctx->d_eq, dim, ctx->d_nbr_ptr, ctx->d_nbr, ctx->d_nbr_off, ctx->d_rowptr, nn, ctx->d_col);
    SC_CHECK_LAUNCH(ctx);
    // ---- pair map (assembly kernel) ------------------------------------------------------------------------------------
    sc_free(&ctx->d_pair_pos); sc_free(&ctx->d_pair_al);
    if (ctx->max_nbr <= 255) {
        SC_TRY(sc_alloc(ctx, &ctx->d_pair_pos, (size_t)ne * nne * nne));
        SC_TRY(sc_alloc(ctx, &ctx->d_pair_al, (size_t)ne * nne));
        k_pair_map<<<nblk(nn, 128), 128, 0, st>>>(ctx->d_conn, nne, ctx->d_n2e_ptr, ctx->d_n2e, ctx->d_nbr_ptr, ctx->d_nbr, nn,
                                                  ctx->d_pair_pos, ctx->d_pair_al);
        SC_CHECK_LAUNCH(ctx);
    }
    SC_TRY(asm_build_block_desc(ctx));
    // ---- node-blocked column lists + descriptors (time-loop kernels, spmv_node.cu) ------------------------------------
    {
        int64_t* d_ncol_ptr = nullptr;
        SC_TRY(sc_alloc(ctx, &d_ncol_ptr, (size_t)nn + 1));
        SC_CUDA(ctx, cudaMemsetAsync(d_tmp64, 0, sizeof(int64_t) * (nn + 1), st));
        k_i32_to_i64<<<nblk(nn, T), T, 0, st>>>(ctx->d_node_rl, d_tmp64, nn);
        SC_CHECK_LAUNCH(ctx);
        SC_TRY(scan64(ctx, d_tmp64, d_ncol_ptr, nn));
        SC_CUDA(ctx, cudaMemcpy(&ctx->ncol_total, d_ncol_ptr + nn, sizeof(int64_t), cudaMemcpyDeviceToHost));
        SC_TRY(sc_alloc(ctx, &ctx->d_ncol, (size_t)ctx->ncol_total));
        k_node_columns<<<nblk(nn * 32, T), T, 0, st>>>(ctx->d_eq, dim, ctx->d_nbr_ptr, ctx->d_nbr, ctx->d_nbr_off, d_ncol_ptr, nn, ctx->d_ncol);
        SC_CHECK_LAUNCH(ctx);
        const int64_t n_pad = nn + 32;
        SC_TRY(sc_alloc(ctx, &ctx->d_nd, (size_t)n_pad));
        k_node_desc<<<nblk(n_pad, T), T, 0, st>>>(ctx->d_node_row0, ctx->d_node_rl, ctx->d_rowptr, d_ncol_ptr, nn, n_pad, ctx->d_nd);
        SC_CHECK_LAUNCH(ctx);
        SC_CUDA(ctx, cudaStreamSynchronize(st));
        sc_free(&d_ncol_ptr);
    }
    SC_CUDA(ctx, cudaStreamSynchronize(st));
    sc_free(&d_cnt); sc_free(&d_tmp64); sc_free(&d_flag);
    SC_TRY(node_dict_build(ctx));     // frequent relative column lists -> dictionary, explicit lists only for the rest
    ctx->have_pattern = true;
    return SC_OK;
}
