// Factorised sparse approximate inverse (FSAI) preconditioner for the PCG of the implicit integrators (sm_100a).
//
// The reference hands the effective matrix of Newmark / Bathe / the static solver to a sparse direct solver (PuggleSolvers
// called at scatter/scatter.py:156-159: `splu` once, two triangular solves per step).  On the device the solve is a
// preconditioned CG (timeloop.cu); with the Jacobi preconditioner of round 1 the 2x2x2-integrated serendipity mass of hexa20
// -- singular, the effective matrix is only regularised by K -- needs ~350 iterations per step at the parity tolerance.
// FSAI is the preconditioner that fits the device: set-up and application are embarrassingly parallel (no triangular
// solves), and the application is two short gather-form SpMVs.
//
//   A^-1 ~ G^T G,  G lower triangular with the pattern  P_i = { j < i : |a_ij| >= tau sqrt(a_ii a_jj) } + {i}   per row i
//   (at most FSAI_CAP entries: rows that exceed it raise their own tau by 30 % until they fit),
//   row i of G = g / sqrt(g_i) with  A[P_i, P_i] g = e_i   -- one dense SPD solve of order |P_i| <= 48 per row,
//   which is  L^T g' = e_m  after the Cholesky factorisation  A[P_i, P_i] = L L^T  (the scaling cancels).
//
// Mass-dominated effective matrices are almost block diagonal by displacement component, so the filter keeps ~1/6 of A's
// entries (hexa20 94^3: 26 of 172 per row; hexa8: 9 of 81) and one application moves ~0.3 of the bytes of one product
// with A, for 3x (hexa20) to 3.6x (hexa8) fewer iterations than Jacobi (scripts/probes).
//
// Domain decomposition: ghost rows are empty in the local CSR; columns of ghost dofs are left out of every P_i, which
// makes the preconditioner block diagonal by rank (still SPD, no extra halo exchange).  Deterministic: fixed patterns,
// fixed summation orders, no atomics.
#include <cub/device/device_scan.cuh>
#include "common.h"

namespace {

constexpr int FSAI_CAP = 48;                       // largest local system
constexpr int FSAI_TRI = FSAI_CAP * (FSAI_CAP + 1) / 2;
constexpr int FSAI_WARPS = 4;                      // warps (rows) per block of the set-up kernels
constexpr unsigned FULL = 0xffffffffu;

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

__device__ __forceinline__ bool fsai_keep(double v, double dinv_i, double dinv_c, double tau) {
    return fabs(v) * sqrt(dinv_i * dinv_c) >= tau;
}

// entries of row i that pass the filter (strictly lower part, non-ghost columns); the row's own tau is raised until they fit
__global__ void __launch_bounds__(32 * FSAI_WARPS)
k_fsai_count(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ vals,
             const double* __restrict__ dinv, double tau0, int64_t n, int64_t* __restrict__ cnt, double* __restrict__ row_tau) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * FSAI_WARPS + (threadIdx.x >> 5);
    if (i >= n) return;
    const int64_t k0 = rowptr[i], k1 = rowptr[i + 1];
    if (k1 == k0 || !(dinv[i] > 0.0)) {            // ghost row (or a row without a positive diagonal): no row in G
        if (lane == 0) { cnt[i] = 0; row_tau[i] = 0.0; }
        return;
    }
    const double di = dinv[i];
    double tau = tau0;
    int c = 0;
    for (;;) {
        c = 0;
        for (int64_t k = k0 + lane; k < k1; k += 32) {
            const int j = col[k];
            if (j < i && dinv[j] > 0.0 && fsai_keep(vals[k], di, dinv[j], tau)) ++c;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
        if (c <= FSAI_CAP - 1) break;
        tau *= 1.3;
    }
    if (lane == 0) { cnt[i] = c + 1; row_tau[i] = tau; }
}

__device__ __forceinline__ int tri(int a, int b) { return a * (a + 1) / 2 + b; }      // b <= a

// one warp per row: gather A[P,P], Cholesky, back substitution, write the row of G
__global__ void __launch_bounds__(32 * FSAI_WARPS)
k_fsai_fill(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ vals,
            const double* __restrict__ dinv, int64_t n, const int64_t* __restrict__ g_rowptr,
            const double* __restrict__ row_tau, int2* __restrict__ g_cv, int* __restrict__ n_fail) {
    __shared__ double sL[FSAI_WARPS][FSAI_TRI];
    __shared__ int sP[FSAI_WARPS][FSAI_CAP];
    __shared__ double sg[FSAI_WARPS][FSAI_CAP];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * FSAI_WARPS + w;
    if (i >= n) return;
    const int64_t g0 = g_rowptr[i];
    const int m = (int)(g_rowptr[i + 1] - g0);
    if (m == 0) return;
    double* L = sL[w];
    int* P = sP[w];
    double* g = sg[w];
    const int64_t k0 = rowptr[i], k1 = rowptr[i + 1];
    const double di = dinv[i];
    const double tau = row_tau[i];                 // the row's own threshold (k_fsai_count)
    // 1. pattern, in column order
    int filled = 0;
    for (int64_t kb = k0; kb < k1; kb += 32) {
        const int64_t k = kb + lane;
        bool keep = false;
        int j = 0;
        if (k < k1) {
            j = col[k];
            keep = j < i && dinv[j] > 0.0 && fsai_keep(vals[k], di, dinv[j], tau);
        }
        const unsigned bal = __ballot_sync(FULL, keep);
        if (keep) {
            const int pos = filled + __popc(bal & ((1u << lane) - 1u));
            if (pos < m - 1) P[pos] = j;
        }
        filled += __popc(bal);
    }
    if (lane == 0) P[m - 1] = (int)i;
    for (int t = lane; t < m * (m + 1) / 2; t += 32) L[t] = 0.0;
    __syncwarp();
    // 2. A[P, P], lower part: row P[a] of the CSR is scanned once, every entry looks its column up in P[0..a]
    for (int a = 0; a < m; ++a) {
        const int r = P[a];
        const int64_t r0 = rowptr[r], r1 = rowptr[r + 1];
        for (int64_t k = r0 + lane; k < r1; k += 32) {
            const int c = col[k];
            if (c > r) break;                       // columns ascend: the rest of this lane's entries lie in the upper part
            int lo = 0, hi = a + 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (P[mid] < c) lo = mid + 1; else hi = mid;
            }
            if (lo <= a && P[lo] == c) L[tri(a, lo)] = vals[k];
        }
    }
    __syncwarp();
    // 3. Cholesky, right-looking
    bool ok = true;
    for (int k = 0; k < m; ++k) {
        const double d = L[tri(k, k)];
        if (!(d > 0.0)) { ok = false; break; }
        const double sd = sqrt(d);
        __syncwarp();
        if (lane == 0) L[tri(k, k)] = sd;
        for (int a = k + 1 + lane; a < m; a += 32) L[tri(a, k)] /= sd;
        __syncwarp();
        for (int a = k + 1 + lane; a < m; a += 32) {
            const double lak = L[tri(a, k)];
            for (int b = k + 1; b <= a; ++b) L[tri(a, b)] -= lak * L[tri(b, k)];
        }
        __syncwarp();
    }
    // 4. L^T g = e_m
    if (ok) {
        if (lane == 0) g[m - 1] = 1.0 / L[tri(m - 1, m - 1)];
        __syncwarp();
        for (int a = m - 2; a >= 0; --a) {
            double s = 0.0;
            for (int b = a + 1 + lane; b < m; b += 32) s += L[tri(b, a)] * g[b];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
            if (lane == 0) g[a] = -s / L[tri(a, a)];
            __syncwarp();
        }
        for (int a = lane; a < m; a += 32) ok = ok && isfinite(g[a]);
        ok = __all_sync(FULL, ok);
    }
    if (!ok) {                                      // numerically indefinite local block: this row falls back to Jacobi
        for (int a = lane; a < m; a += 32) g[a] = a == m - 1 ? sqrt(di) : 0.0;
        if (lane == 0) atomicAdd(n_fail, 1);        // a counter, not a float sum
        __syncwarp();
    }
    for (int a = lane; a < m; a += 32) g_cv[g0 + a] = make_int2(P[a], __float_as_int((float)g[a]));
}

// G^T row j = { (i, g_ij) : i >= j, j in P_i }: walk the upper part of row j of A (structurally symmetric) in column order
// and look j up in the row of G of every candidate.  FILL = false counts, FILL = true writes.
template <bool FILL>
__global__ void __launch_bounds__(32 * FSAI_WARPS)
k_fsai_transpose(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n, const int64_t* __restrict__ g_rowptr,
                 const int2* __restrict__ g_cv, int64_t* __restrict__ cnt, const int64_t* __restrict__ t_rowptr,
                 int2* __restrict__ t_cv) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * FSAI_WARPS + (threadIdx.x >> 5);
    if (j >= n) return;
    const int64_t k0 = rowptr[j], k1 = rowptr[j + 1];
    int64_t out = FILL ? t_rowptr[j] : 0;
    int total = 0;
    for (int64_t kb = k0; kb < k1; kb += 32) {
        const int64_t k = kb + lane;
        bool hit = false;
        int i = 0, v = 0;
        if (k < k1) {
            i = col[k];
            if (i >= j) {
                int64_t lo = g_rowptr[i], hi = g_rowptr[i + 1];
                const int64_t end = hi;
                while (lo < hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (g_cv[mid].x < j) lo = mid + 1; else hi = mid;
                }
                if (lo < end) {
                    const int2 cv = g_cv[lo];
                    if (cv.x == j) { hit = true; v = cv.y; }
                }
            }
        }
        const unsigned bal = __ballot_sync(FULL, hit);
        if (FILL && hit) {
            const int64_t pos = out + __popc(bal & ((1u << lane) - 1u));
            t_cv[pos] = make_int2(i, v);
        }
        out += __popc(bal);
        total += __popc(bal);
    }
    if (!FILL && lane == 0) cnt[j] = total;
}

// y = B x for a CSR matrix with (column, FP32 value) pairs (B = G or G^T): LPR lanes per row, U entries per lane in flight;
// DOT: per-block partial of w.y.  Every block owns one contiguous chunk of rows, so the partials do not depend on the launch
// geometry.  The three dependent loads of a row (row pointers -> entries -> gathered x) are software-pipelined across the
// block's row groups: while the gathers of group j are in flight, the entries of group j+1 and the row pointers of group
// j+2 are already requested (the unpipelined form was latency-bound at 1.9 TB/s).
template <int LPR, int U, bool DOT>
__global__ void __launch_bounds__(256)
k_csr32_spmv(const int64_t* __restrict__ rowptr, const int2* __restrict__ cv, const double* __restrict__ x, double* __restrict__ y,
             int64_t n, const double* __restrict__ w, double* __restrict__ partial) {
    constexpr int RPB = 256 / LPR;
    const int sub = threadIdx.x / LPR, l = threadIdx.x % LPR;
    const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
    const int64_t s = blockIdx.x * chunk, e = min(s + chunk, n);
    const int64_t niter = e > s ? (e - s + RPB - 1) / RPB : 0;
    double dot = 0.0;
    auto load_rp = [&](int64_t it, int64_t& k0, int64_t& k1) {
        const int64_t row = s + it * RPB + sub;
        k0 = 0; k1 = 0;
        if (it < niter && row < e) { k0 = rowptr[row]; k1 = rowptr[row + 1]; }
    };
    auto load_cv = [&](int64_t k0, int64_t k1, int2 (&c)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t k = k0 + u * LPR + l;
            c[u] = k < k1 ? __ldg(cv + k) : make_int2(-1, 0);
        }
    };
    int64_t k0c, k1c, k0n, k1n, k0nn, k1nn;
    int2 c[U], cn[U];
    load_rp(0, k0c, k1c);
    load_rp(1, k0n, k1n);
    load_cv(k0c, k1c, c);
    for (int64_t j = 0; j < niter; ++j) {
        load_rp(j + 2, k0nn, k1nn);
        load_cv(k0n, k1n, cn);
        double xg[U];
#pragma unroll
        for (int u = 0; u < U; ++u) xg[u] = c[u].x >= 0 ? x[c[u].x] : 0.0;
        double sum = 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) sum += (double)__int_as_float(c[u].y) * xg[u];
        for (int64_t k = k0c + LPR * U + l; k < k1c; k += LPR) {           // rows longer than one pass
            const int2 t = __ldg(cv + k);
            sum += (double)__int_as_float(t.y) * x[t.x];
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
        const int64_t row = s + j * RPB + sub;
        if (l == 0 && row < e) {
            y[row] = sum;
            if (DOT) dot += w[row] * sum;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = cn[u];
        k0c = k0n; k1c = k1n; k0n = k0nn; k1n = k1nn;
    }
    if (DOT) {
        __shared__ double red[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_down_sync(FULL, dot, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int k = 0; k < 8; ++k) t += red[k];
            partial[blockIdx.x] = t;
        }
    }
}

int scan_counts(sc_ctx* ctx, int64_t* d_cnt, int64_t* d_ptr, int64_t n, int64_t* total) {
    size_t bytes = 0;
    SC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, bytes, d_cnt, d_ptr, n + 1, ctx->stream));
    void* tmp = nullptr;
    SC_CUDA(ctx, cudaMalloc(&tmp, bytes ? bytes : 1));
    const cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, bytes, d_cnt, d_ptr, n + 1, ctx->stream);
    ctx->launches += 2;
    cudaError_t e2 = cudaMemcpyAsync(total, d_ptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    if (e != cudaSuccess || e2 != cudaSuccess)
        return sc_fail(ctx, SC_ERR_CUDA, "FSAI scan failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    return SC_OK;
}

template <bool DOT>
int csr32_launch(sc_ctx* ctx, int lanes, const int64_t* rowptr, const int2* cv, const double* x, double* y, const double* w,
                 double* partial, int nb) {
    const int64_t n = ctx->n_eq;
    cudaStream_t st = ctx->stream;
    if (lanes <= 4) k_csr32_spmv<4, 4, DOT><<<nb, 256, 0, st>>>(rowptr, cv, x, y, n, w, partial);      // short rows (hexa8: ~9 entries)
    else k_csr32_spmv<8, 6, DOT><<<nb, 256, 0, st>>>(rowptr, cv, x, y, n, w, partial);                  // one pass up to FSAI_CAP
    SC_CHECK_LAUNCH(ctx);
    return SC_OK;
}

}  // namespace

void fsai_free(sc_fsai* f) {
    sc_free(&f->rowptr); sc_free(&f->cv);
    sc_free(&f->t_rowptr); sc_free(&f->t_cv);
    f->for_vals = nullptr;
    f->nnz = 0;
}

int fsai_build(sc_ctx* ctx, sc_fsai* f, const double* vals) {
    fsai_free(f);
    const int64_t n = ctx->n_eq;
    cudaStream_t st = ctx->stream;
    sc_gpu_timer timer(st);
    timer.start();
    double* dinv = nullptr;
    int64_t* cnt = nullptr;
    double* row_tau = nullptr;
    int* n_fail = nullptr;
    int rc = SC_OK;
    auto body = [&]() -> int {
        SC_TRY(sc_alloc(ctx, &dinv, (size_t)n));
        SC_TRY(sc_alloc(ctx, &cnt, (size_t)n + 1));
        SC_TRY(sc_alloc(ctx, &row_tau, (size_t)n));
        SC_TRY(sc_alloc(ctx, &n_fail, 1));
        SC_CUDA(ctx, cudaMemsetAsync(n_fail, 0, sizeof(int), st));
        SC_CUDA(ctx, cudaMemsetAsync(cnt + n, 0, sizeof(int64_t), st));
        SC_TRY(la_extract_diag(ctx, vals, dinv, true));
        const unsigned nb = nblk(n, FSAI_WARPS);
        k_fsai_count<<<nb, 32 * FSAI_WARPS, 0, st>>>(ctx->d_rowptr, ctx->d_col, vals, dinv, ctx->fsai_tau, n, cnt, row_tau);
        SC_CHECK_LAUNCH(ctx);
        SC_TRY(sc_alloc(ctx, &f->rowptr, (size_t)n + 1));
        SC_TRY(scan_counts(ctx, cnt, f->rowptr, n, &f->nnz));
        SC_TRY(sc_alloc(ctx, &f->cv, (size_t)f->nnz));
        k_fsai_fill<<<nb, 32 * FSAI_WARPS, 0, st>>>(ctx->d_rowptr, ctx->d_col, vals, dinv, n, f->rowptr, row_tau, f->cv, n_fail);
        SC_CHECK_LAUNCH(ctx);
        // transpose (same number of entries)
        SC_CUDA(ctx, cudaMemsetAsync(cnt + n, 0, sizeof(int64_t), st));
        k_fsai_transpose<false><<<nb, 32 * FSAI_WARPS, 0, st>>>(ctx->d_rowptr, ctx->d_col, n, f->rowptr, f->cv, cnt, nullptr, nullptr);
        SC_CHECK_LAUNCH(ctx);
        SC_TRY(sc_alloc(ctx, &f->t_rowptr, (size_t)n + 1));
        int64_t tn = 0;
        SC_TRY(scan_counts(ctx, cnt, f->t_rowptr, n, &tn));
        if (tn != f->nnz) return sc_fail(ctx, SC_ERR_STATE, "FSAI transpose found %lld of %lld entries: the CSR pattern is not "
                                         "structurally symmetric", (long long)tn, (long long)f->nnz);
        SC_TRY(sc_alloc(ctx, &f->t_cv, (size_t)f->nnz));
        k_fsai_transpose<true><<<nb, 32 * FSAI_WARPS, 0, st>>>(ctx->d_rowptr, ctx->d_col, n, f->rowptr, f->cv, nullptr, f->t_rowptr,
                                                               f->t_cv);
        SC_CHECK_LAUNCH(ctx);
        return SC_OK;
    };
    rc = body();
    timer.stop();
    if (rc == SC_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = sc_fail(ctx, SC_ERR_CUDA, "FSAI set-up failed: %s", cudaGetErrorString(cudaGetLastError()));
    sc_free(&dinv); sc_free(&cnt); sc_free(&row_tau); sc_free(&n_fail);
    if (rc != SC_OK) { fsai_free(f); return rc; }
    f->seconds = timer.ms() * 1e-3;
    const double avg = n > 0 ? (double)f->nnz / (double)n : 1.0;
    f->lanes = avg <= 12.0 ? 4 : 8;
    f->for_vals = vals;
    return SC_OK;
}

// t = G r, z = G^T t, partial[b] = sum over the rows of block b of r.z   (nb blocks)
int fsai_apply(sc_ctx* ctx, const sc_fsai* f, const double* r, double* t, double* z, double* partial, int nb) {
    SC_TRY(csr32_launch<false>(ctx, f->lanes, f->rowptr, f->cv, r, t, nullptr, nullptr, nb));
    SC_TRY(csr32_launch<true>(ctx, f->lanes, f->t_rowptr, f->t_cv, t, z, r, partial, nb));
    return SC_OK;
}
