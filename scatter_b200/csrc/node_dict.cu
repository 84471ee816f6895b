// Column-pattern dictionary for the node-blocked SpMV (spmv_node.cu).
//
// `pattern.cu` stores one column list per node.  On meshes with any regularity (the structured soil boxes of the
// benchmark, extruded or block-structured gmsh meshes) most nodes have the *same list relative to their own first row*:
// col[k] - row0 is identical for every node whose neighbourhood has the same shape and the same fixed dofs.  This file
// finds the most frequent relative lists (at most SC_DICT_MAX), keeps them in a small dictionary that the SpMV kernel
// holds in shared memory, and drops the explicit lists of the nodes they cover: those nodes then cost one byte of pattern
// id instead of 4 bytes per column -- for the 255^3 hexa8 box 98 % of the 5.3 GB of node column lists disappear from
// every SpMV (39.8 -> 34.5 GB per fused time step).  Nodes with any other list keep their explicit list, so nothing is
// assumed about the mesh; the result of the product is bit-identical (same values, same columns, same order).
//
// Steps (all on the device except picking the winners among the run counts):
//   k_node_hash     64-bit hash of (length, relative list) per node, one warp per node
//   radix sort + run-length encode of the hashes -> frequency of every distinct hash
//   host: the SC_DICT_MAX most frequent hashes with count >= 2 -> representative nodes -> dictionary rows
//   k_assign_pid    per node: dictionary row with the same hash AND exactly the same list (no trust in the hash), or 0
//   compaction of the explicit lists, new node descriptors (pattern id in bits 16..23 of len_nfree)
#include <algorithm>
#include <numeric>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>
#include "common.h"

namespace {

using NodeDesc = sc_ctx::NodeDesc;

__host__ __device__ inline uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

__global__ void k_node_hash(const NodeDesc* __restrict__ nd, const int32_t* __restrict__ ncol, int64_t n_nodes, uint64_t* __restrict__ hash) {
    const int64_t a = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (a >= n_nodes) return;
    const NodeDesc d = nd[a];
    const int nfree = d.len_nfree >> 24;
    const int L = nfree > 0 ? (d.len_nfree & 0xffff) : 0;
    uint64_t h = 0;
    for (int k = lane; k < L; k += 32) {
        const uint64_t v = (uint64_t)(uint32_t)(ncol[d.col_off + k] - d.row0);
        h += mix64(v ^ ((uint64_t)(k + 1) << 32));          // position-dependent terms, order-independent sum
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if (lane == 0) {
        h = mix64(h ^ (uint64_t)L);
        hash[a] = L > 0 ? (h | 1ULL) : 0ULL;                // 0 marks nodes without a list
    }
}

__global__ void k_iota32(int32_t* v, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int32_t)i;
}

// pattern id of every node: dictionary row with equal hash, equal length and equal relative list; 0 otherwise
__global__ void k_assign_pid(const NodeDesc* __restrict__ nd, const int32_t* __restrict__ ncol, const uint64_t* __restrict__ hash,
                             int64_t n_nodes, int n_dict, const uint64_t* __restrict__ dict_hash, const int32_t* __restrict__ dict_len,
                             const int32_t* __restrict__ dict, int stride, uint8_t* __restrict__ pid, int64_t* __restrict__ len2) {
    const int64_t a = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (a >= n_nodes) return;
    const NodeDesc d = nd[a];
    const int nfree = d.len_nfree >> 24;
    const int L = nfree > 0 ? (d.len_nfree & 0xffff) : 0;
    const uint64_t h = hash[a];
    int p = -1;
    if (h != 0)
        for (int q = 0; q < n_dict; ++q)
            if (dict_hash[q] == h && dict_len[q] == L) { p = q; break; }
    bool same = p >= 0;
    if (p >= 0)
        for (int k = lane; k < L; k += 32) same &= (ncol[d.col_off + k] - d.row0) == dict[(size_t)p * stride + k];
    same = __all_sync(0xffffffffu, same);
    if (lane == 0) {
        pid[a] = same ? (uint8_t)(p + 1) : 0;
        len2[a] = same ? 0 : L;
    }
}

__global__ void k_compact(const NodeDesc* __restrict__ nd, const int32_t* __restrict__ ncol, const uint8_t* __restrict__ pid,
                          const int64_t* __restrict__ off2, int64_t n_nodes, int64_t n_pad, int32_t* __restrict__ ncol2,
                          NodeDesc* __restrict__ nd2) {
    const int64_t a = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (a >= n_pad) return;
    NodeDesc d = nd[a];
    if (a < n_nodes) {
        const int nfree = d.len_nfree >> 24;
        const int L = nfree > 0 ? (d.len_nfree & 0xffff) : 0;
        const int p = pid[a];
        const int64_t o = off2[a];
        if (p == 0)
            for (int k = lane; k < L; k += 32) ncol2[o + k] = ncol[d.col_off + k];
        d.col_off = o;
        d.len_nfree |= p << 16;
    } else {
        d.col_off = off2[n_nodes];
    }
    if (lane == 0) nd2[a] = d;
}

unsigned nblk64(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

// Replaces ctx->d_ncol / ctx->d_nd by their dictionary-compressed form when that saves anything.  Leaves everything as it
// is (n_dict = 0) on meshes without repeated lists, with rows longer than the shared-memory dictionary allows, or when
// SCATTER_B200_NO_DICT=1.
int node_dict_build(sc_ctx* ctx) {
    sc_free(&ctx->d_dict);
    ctx->n_dict = 0; ctx->dict_stride = 0;
    const int64_t nn = ctx->n_nodes;
    if (ctx->no_dict || !ctx->d_nd || !ctx->d_ncol || nn < 64 || ctx->max_rl <= 0) return SC_OK;
    // the dictionary lives in the shared memory the SpMV ring leaves free (none for the long rows of hexa20 / tetra10)
    const int max_entries = (int)std::min<int64_t>(SC_DICT_MAX, node_dict_room(ctx) / ((int64_t)ctx->max_rl * (int64_t)sizeof(int32_t)));
    if (max_entries < 1) return SC_OK;
    cudaStream_t st = ctx->stream;
    const int T = 256;
    uint64_t *d_hash = nullptr, *d_keys = nullptr, *d_uniq = nullptr, *d_dhash = nullptr;
    int32_t *d_ids = nullptr, *d_ids2 = nullptr, *d_cnt = nullptr, *d_nruns = nullptr, *d_dlen = nullptr, *d_dict = nullptr, *d_ncol2 = nullptr;
    uint8_t* d_pid = nullptr;
    int64_t *d_len2 = nullptr, *d_off2 = nullptr;
    NodeDesc* d_nd2 = nullptr;
    void* d_tmp = nullptr;
    auto body = [&]() -> int {
        SC_TRY(sc_alloc(ctx, &d_hash, (size_t)nn));
        SC_TRY(sc_alloc(ctx, &d_keys, (size_t)nn));
        SC_TRY(sc_alloc(ctx, &d_ids, (size_t)nn));
        SC_TRY(sc_alloc(ctx, &d_ids2, (size_t)nn));
        SC_TRY(sc_alloc(ctx, &d_uniq, (size_t)nn));
        SC_TRY(sc_alloc(ctx, &d_cnt, (size_t)nn));
        SC_TRY(sc_alloc(ctx, &d_nruns, 1));
        k_node_hash<<<nblk64(nn * 32, T), T, 0, st>>>(ctx->d_nd, ctx->d_ncol, nn, d_hash);
        SC_CHECK_LAUNCH(ctx);
        k_iota32<<<nblk64(nn, T), T, 0, st>>>(d_ids, nn);
        SC_CHECK_LAUNCH(ctx);
        size_t b1 = 0, b2 = 0;
        SC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, b1, d_hash, d_keys, d_ids, d_ids2, (int)nn, 0, 64, st));
        SC_CUDA(ctx, cub::DeviceRunLengthEncode::Encode(nullptr, b2, d_keys, d_uniq, d_cnt, d_nruns, (int)nn, st));
        const size_t tmp_bytes = std::max(b1, b2);
        SC_CUDA(ctx, cudaMalloc(&d_tmp, tmp_bytes + 16));
        size_t b = tmp_bytes;
        SC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, b, d_hash, d_keys, d_ids, d_ids2, (int)nn, 0, 64, st));
        b = tmp_bytes;
        SC_CUDA(ctx, cub::DeviceRunLengthEncode::Encode(d_tmp, b, d_keys, d_uniq, d_cnt, d_nruns, (int)nn, st));
        int n_runs = 0;
        SC_CUDA(ctx, cudaMemcpyAsync(&n_runs, d_nruns, sizeof(int), cudaMemcpyDeviceToHost, st));
        SC_CUDA(ctx, cudaStreamSynchronize(st));
        if (n_runs <= 0 || n_runs > (4 << 20)) return SC_OK;           // no structure worth a dictionary
        std::vector<uint64_t> uniq((size_t)n_runs);
        std::vector<int32_t> cnt((size_t)n_runs);
        SC_CUDA(ctx, cudaMemcpy(uniq.data(), d_uniq, sizeof(uint64_t) * n_runs, cudaMemcpyDeviceToHost));
        SC_CUDA(ctx, cudaMemcpy(cnt.data(), d_cnt, sizeof(int32_t) * n_runs, cudaMemcpyDeviceToHost));
        std::vector<int64_t> start((size_t)n_runs + 1, 0);
        for (int r = 0; r < n_runs; ++r) start[r + 1] = start[r] + cnt[r];
        std::vector<int> order((size_t)n_runs);
        std::iota(order.begin(), order.end(), 0);
        const int want = std::min(n_runs, max_entries);
        // most frequent first; ties by hash so that the choice does not depend on the sort implementation
        std::partial_sort(order.begin(), order.begin() + want, order.end(), [&](int x, int y) {
            return cnt[x] != cnt[y] ? cnt[x] > cnt[y] : uniq[x] < uniq[y];
        });
        std::vector<int> sel;
        int64_t covered = 0;
        for (int i = 0; i < want; ++i) {
            const int r = order[i];
            if (uniq[r] == 0 || cnt[r] < 2) continue;
            sel.push_back(r);
            covered += cnt[r];
        }
        if (sel.empty() || covered * 8 < nn) return SC_OK;              // covers less than 1/8 of the nodes: not worth it
        const int nd_sel = (int)sel.size();
        const int stride = ctx->max_rl;
        std::vector<int32_t> h_dict((size_t)nd_sel * stride, 0), h_len((size_t)nd_sel, 0);
        std::vector<uint64_t> h_hash((size_t)nd_sel, 0);
        std::vector<int32_t> tmp_cols((size_t)stride);
        for (int i = 0; i < nd_sel; ++i) {
            int32_t node = 0;
            SC_CUDA(ctx, cudaMemcpy(&node, d_ids2 + start[sel[i]], sizeof(int32_t), cudaMemcpyDeviceToHost));
            NodeDesc d;
            SC_CUDA(ctx, cudaMemcpy(&d, ctx->d_nd + node, sizeof(NodeDesc), cudaMemcpyDeviceToHost));
            const int L = d.len_nfree & 0xffff;
            if (L <= 0 || L > stride) return sc_fail(ctx, SC_ERR_STATE, "inconsistent node descriptor while building the column dictionary");
            SC_CUDA(ctx, cudaMemcpy(tmp_cols.data(), ctx->d_ncol + d.col_off, sizeof(int32_t) * L, cudaMemcpyDeviceToHost));
            for (int k = 0; k < L; ++k) h_dict[(size_t)i * stride + k] = tmp_cols[k] - d.row0;
            h_len[i] = L;
            h_hash[i] = uniq[sel[i]];
        }
        SC_TRY(sc_alloc(ctx, &d_dict, h_dict.size()));
        SC_TRY(sc_alloc(ctx, &d_dlen, h_len.size()));
        SC_TRY(sc_alloc(ctx, &d_dhash, h_hash.size()));
        SC_CUDA(ctx, cudaMemcpy(d_dict, h_dict.data(), sizeof(int32_t) * h_dict.size(), cudaMemcpyHostToDevice));
        SC_CUDA(ctx, cudaMemcpy(d_dlen, h_len.data(), sizeof(int32_t) * h_len.size(), cudaMemcpyHostToDevice));
        SC_CUDA(ctx, cudaMemcpy(d_dhash, h_hash.data(), sizeof(uint64_t) * h_hash.size(), cudaMemcpyHostToDevice));
        SC_TRY(sc_alloc(ctx, &d_pid, (size_t)nn));
        SC_TRY(sc_alloc(ctx, &d_len2, (size_t)nn + 1));
        SC_TRY(sc_alloc(ctx, &d_off2, (size_t)nn + 1));
        SC_CUDA(ctx, cudaMemsetAsync(d_len2, 0, sizeof(int64_t) * (nn + 1), st));
        k_assign_pid<<<nblk64(nn * 32, T), T, 0, st>>>(ctx->d_nd, ctx->d_ncol, d_hash, nn, nd_sel, d_dhash, d_dlen, d_dict, stride, d_pid, d_len2);
        SC_CHECK_LAUNCH(ctx);
        size_t b3 = 0;
        SC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, b3, d_len2, d_off2, (int)(nn + 1), st));
        if (b3 > tmp_bytes) { cudaFree(d_tmp); d_tmp = nullptr; SC_CUDA(ctx, cudaMalloc(&d_tmp, b3 + 16)); }
        SC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp, b3, d_len2, d_off2, (int)(nn + 1), st));
        int64_t total2 = 0;
        SC_CUDA(ctx, cudaMemcpyAsync(&total2, d_off2 + nn, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        SC_CUDA(ctx, cudaStreamSynchronize(st));
        if (total2 < 0 || total2 > ctx->ncol_total) return sc_fail(ctx, SC_ERR_STATE, "column dictionary compaction went wrong");
        const int64_t n_pad = nn + 32;
        SC_TRY(sc_alloc(ctx, &d_ncol2, (size_t)total2 + 8));
        SC_TRY(sc_alloc(ctx, &d_nd2, (size_t)n_pad));
        k_compact<<<nblk64(n_pad * 32, T), T, 0, st>>>(ctx->d_nd, ctx->d_ncol, d_pid, d_off2, nn, n_pad, d_ncol2, d_nd2);
        SC_CHECK_LAUNCH(ctx);
        SC_CUDA(ctx, cudaStreamSynchronize(st));
        // adopt
        sc_free(&ctx->d_ncol); sc_free(&ctx->d_nd);
        ctx->d_ncol = d_ncol2; d_ncol2 = nullptr;
        ctx->d_nd = d_nd2; d_nd2 = nullptr;
        ctx->d_dict = d_dict; d_dict = nullptr;
        ctx->n_dict = nd_sel;
        ctx->dict_stride = stride;
        ctx->ncol_total = total2;
        return SC_OK;
    };
    const int rc = body();
    sc_free(&d_hash); sc_free(&d_keys); sc_free(&d_uniq); sc_free(&d_dhash); sc_free(&d_ids); sc_free(&d_ids2); sc_free(&d_cnt);
    sc_free(&d_nruns); sc_free(&d_dlen); sc_free(&d_dict); sc_free(&d_ncol2); sc_free(&d_pid); sc_free(&d_len2); sc_free(&d_off2);
    sc_free(&d_nd2);
    if (d_tmp) cudaFree(d_tmp);
    return rc;
}
