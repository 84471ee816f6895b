// Sparse / dense vector kernels of the time loop: streaming CSR SpMV, fused vector updates, deterministic dots.
//
// These replace the scipy.sparse products the reference's external solver performs every step
// (solvers.NewmarkExplicit.calculate called at scatter/scatter.py:159; recurrence in SURVEY.md 3.3).
//
// SpMV layout: rows are consecutive slices of the value / column arrays; each warp streams the slices of a few
// consecutive rows with coalesced loads and multiplies by gathered x entries (L2-resident for FEM orderings).
// No atomics, fixed summation order => bit-reproducible.
#include <algorithm>
#include "common.h"

namespace {

constexpr int SPMV_THREADS = 256;
constexpr int SPMV_WARPS = SPMV_THREADS / 32;

// CSR SpMV, one warp per RPW consecutive rows, U chunks of 32 entries in flight per row and pass.
// There is no shared-memory staging and no block barrier in the streaming part: every warp keeps RPW*U independent
// (value, column) loads plus their gathers in flight, which is what hides the HBM latency (ncu of the first version --
// products staged in shared memory, one thread per row adding them -- showed 51 long-scoreboard stalls per issue and
// 40 % DRAM throughput).  Lane partials are added in entry order, then a fixed xor butterfly: deterministic.
// MODE 0: y = A xa            MODE 1: y = A xa + B xb
// MODE 2: central-difference step  un[i] = inv_d[i]*(-sum) + alpha[i]*xe[i] - (alpha[i]-1)*un[i]   (un holds u_prev; xa is the
//         gathered vector w = (1+g) u - g u_prev that carries the lagged stiffness-proportional damping, xe = u; with
//         y2 != null the next gather vector y2[i] = (1+g) un[i] - g xe[i] is written by the same epilogue)
// MODE 3: PCG product  y = A xa  and per-block partial of xa.y  (partial[blockIdx])
template <int MODE, int RPW, int U>
__global__ void __launch_bounds__(SPMV_THREADS)
k_spmv(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ va,
       const double* __restrict__ xa, const double* __restrict__ vb, const double* __restrict__ xb,
       double* __restrict__ y, const double* __restrict__ inv_d, const double* __restrict__ alpha,
       double* __restrict__ partial, int64_t n_rows, const double* __restrict__ xe, double* __restrict__ y2, double lag) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = ((int64_t)blockIdx.x * SPMV_WARPS + warp) * RPW;
    // row pointers of the warp's rows: lanes 0..RPW load, everybody reads them through shuffles
    int64_t rp = 0;
    if (lane <= RPW && row0 + lane <= n_rows) rp = rowptr[row0 + lane];
    int64_t rs[RPW], re[RPW];
    int niter = 0;
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        rs[r] = __shfl_sync(0xffffffffu, rp, r);
        re[r] = __shfl_sync(0xffffffffu, rp, r + 1);
        if (row0 + r >= n_rows) re[r] = rs[r];
        niter = max(niter, (int)((re[r] - rs[r] + 31) >> 5));
    }
    // epilogue operands do not depend on the sums: issue their loads first
    const int64_t myrow = row0 + lane;
    const bool owner = lane < RPW && myrow < n_rows;
    double e_al = 0.0, e_id = 0.0, e_x = 0.0, e_y = 0.0;
    if (owner) {
        if (MODE == 2) { e_al = alpha[myrow]; e_id = inv_d[myrow]; e_x = xe[myrow]; e_y = y[myrow]; }
        if (MODE == 3) e_x = xa[myrow];
    }
    double sum[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) sum[r] = 0.0;
    for (int j = 0; j < niter; j += U) {
        // all loads of the pass are unconditional (inactive lanes read entry 0 and are masked afterwards), so the
        // compiler issues the RPW*U (value, column) loads back to back, then the gathers, then the FMAs
        double v[RPW][U], w[RPW][U];
        int c[RPW][U];
#pragma unroll
        for (int r = 0; r < RPW; ++r)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t k = rs[r] + lane + 32 * (j + u);
                const bool ok = k < re[r];
                const int64_t ks = ok ? k : 0;
                c[r][u] = __ldg(col + ks);
                const double vv = __ldg(va + ks);
                v[r][u] = ok ? vv : 0.0;
                if (MODE == 1) { const double ww = __ldg(vb + ks); w[r][u] = ok ? ww : 0.0; }
            }
        __syncwarp();      // scheduling fence: every (value, column) load of the pass is issued before the first gather
        double xg[RPW][U], xh[RPW][U];
#pragma unroll
        for (int r = 0; r < RPW; ++r)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                xg[r][u] = xa[c[r][u]];
                if (MODE == 1) xh[r][u] = xb[c[r][u]];
            }
        __syncwarp();      // ... and every gather before the first FMA
#pragma unroll
        for (int r = 0; r < RPW; ++r)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                double t = v[r][u] * xg[r][u];
                if (MODE == 1) t += w[r][u] * xh[r][u];
                sum[r] += t;
            }
    }
    double mine = 0.0;
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        double s = sum[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == r) mine = s;
    }
    double dotv = 0.0;
    // lane r owns row r in the epilogue; empty rows (ghost dofs of a domain decomposition) are left untouched
    const int64_t rp_next = __shfl_down_sync(0xffffffffu, rp, 1);
    if (owner) {
        if (MODE == 2) {
            if (rp_next > rp) {
                const double un = e_id * (-mine) + e_al * e_x - (e_al - 1.0) * e_y;
                y[myrow] = un;
                if (y2) y2[myrow] = (1.0 + lag) * un - lag * e_x;
            }
        } else {
            y[myrow] = mine;
            if (MODE == 3) dotv = e_x * mine;
        }
    }
    if (MODE == 3) {
        __shared__ double red[SPMV_WARPS];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dotv += __shfl_down_sync(0xffffffffu, dotv, o);
        if (lane == 0) red[warp] = dotv;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w2 = 0; w2 < SPMV_WARPS; ++w2) s += red[w2];
            partial[blockIdx.x] = s;
        }
    }
}

// ---- deterministic reductions ------------------------------------------------------------------------------
constexpr int RED_BLOCKS = 1024;
constexpr int RED_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < RED_THREADS / 32; ++w) s += sh[w];
    __syncthreads();
    return s;   // valid on thread 0
}

__global__ void __launch_bounds__(RED_THREADS)
k_dot(const double* __restrict__ x, const double* __restrict__ y, int64_t n, double* __restrict__ partial) {
    __shared__ double sh[RED_THREADS / 32];
    const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
    const int64_t b = blockIdx.x * chunk, e = min(b + chunk, n);
    double s = 0.0;
    for (int64_t i = b + threadIdx.x; i < e; i += RED_THREADS) s += x[i] * y[i];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// out[j] = sum_k partial[j*stride + k], k < count   (single block; fixed order)
__global__ void __launch_bounds__(RED_THREADS)
k_reduce_final(const double* __restrict__ partial, int64_t count, int nout, int64_t stride, double* __restrict__ out) {
    __shared__ double sh[RED_THREADS / 32];
    for (int j = 0; j < nout; ++j) {
        double s = 0.0;
        for (int64_t i = threadIdx.x; i < count; i += RED_THREADS) s += partial[j * stride + i];
        s = block_sum(s, sh);
        if (threadIdx.x == 0) out[j] = s;
    }
}

// ---- vector kernels -------------------------------------------------------------------------------------------
__global__ void k_fill(double* __restrict__ x, double v, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}
__global__ void k_axpby(double* out, double a, const double* x, double b, const double* __restrict__ y, int64_t n) {   // out may alias x
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = a * x[i] + (y ? b * y[i] : 0.0);
}
__global__ void k_extract_diag(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ vals,
                               double* __restrict__ diag, int64_t n, int invert) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double d = 0.0;
    int64_t lo = rowptr[i], hi = rowptr[i + 1];
    while (lo < hi) {   // columns are sorted
        int64_t mid = (lo + hi) >> 1;
        int c = col[mid];
        if (c == i) { d = vals[mid]; break; }
        if (c < i) lo = mid + 1; else hi = mid;
    }
    diag[i] = invert ? (d != 0.0 ? 1.0 / d : 0.0) : d;
}
// y[row] += scale * sum_k val[k] x[col[k]] over the row-compressed C_abs list (one thread per listed row)
__global__ void k_cabs_spmv_add(const int64_t* __restrict__ rowid, const int64_t* __restrict__ rptr, const int32_t* __restrict__ col,
                                const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y,
                                double scale, int64_t nrows) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nrows) return;
    double s = 0.0;
    for (int64_t k = rptr[t]; k < rptr[t + 1]; ++k) s += val[k] * x[col[k]];
    y[rowid[t]] += scale * s;
}
__global__ void k_cabs_add_values(const int64_t* __restrict__ slot, const double* __restrict__ val, double* __restrict__ vals,
                                  double scale, int64_t n) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) vals[slot[t]] += scale * val[t];
}

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

// launch configuration from the average row length: U chunks of 32 entries per pass, RPW rows per warp
template <int MODE>
int spmv_launch(sc_ctx* ctx, const double* va, const double* xa, const double* vb, const double* xb, double* y,
                const double* inv_d, const double* alpha, double* partial, unsigned* nblocks_out,
                const double* xe = nullptr, double* y2 = nullptr, double g = 0.0) {
    const double avg = ctx->n_eq > 0 ? (double)ctx->nnz / (double)ctx->n_eq : 1.0;
    const int64_t n = ctx->n_eq;
    cudaStream_t st = ctx->stream;
#define SC_SPMV_GO(RPW, U)                                                                                              \
    {                                                                                                                   \
        const unsigned nb = (unsigned)((n + (int64_t)SPMV_WARPS * RPW - 1) / ((int64_t)SPMV_WARPS * RPW));             \
        if (nblocks_out) *nblocks_out = nb;                                                                             \
        k_spmv<MODE, RPW, U><<<nb, SPMV_THREADS, 0, st>>>(ctx->d_rowptr, ctx->d_col, va, xa, vb, xb, y, inv_d, alpha,   \
                                                           partial, n, xe, y2, g);                                      \
    }
    if (avg <= 32.0) SC_SPMV_GO(4, 1)
    else if (avg <= 64.0) SC_SPMV_GO(2, 2)
    else if (avg <= 96.0) SC_SPMV_GO(2, 3)
    else SC_SPMV_GO(2, 4)
#undef SC_SPMV_GO
    SC_CHECK_LAUNCH(ctx);
    return SC_OK;
}

}  // namespace

int sc_work(sc_ctx* ctx, int idx, double** out) {
    if ((int)ctx->work.size() <= idx) ctx->work.resize(idx + 1, nullptr);
    if (!ctx->work[idx]) {
        SC_TRY(sc_alloc(ctx, &ctx->work[idx], (size_t)ctx->n_eq));
        SC_CUDA(ctx, cudaMemsetAsync(ctx->work[idx], 0, sizeof(double) * ctx->n_eq, ctx->stream));
    }
    *out = ctx->work[idx];
    return SC_OK;
}

int la_scratch(sc_ctx* ctx) {
    if (!ctx->d_scal) SC_TRY(sc_alloc(ctx, &ctx->d_scal, 64));
    if (!ctx->d_partial) {
        size_t nb = (size_t)std::max<int64_t>(4 * RED_BLOCKS, (ctx->n_eq + 7) / 8 + 16);
        SC_TRY(sc_alloc(ctx, &ctx->d_partial, nb));
    }
    if (!ctx->h_pinned) SC_CUDA(ctx, cudaMallocHost((void**)&ctx->h_pinned, 64 * sizeof(double)));
    return SC_OK;
}

int la_spmv(sc_ctx* ctx, const double* vals, const double* x, double* y) {
    if (la_node_usable(ctx)) return la_node_spmv(ctx, vals, x, y);
    if (la_tma_usable(ctx)) return la_tma_spmv(ctx, vals, x, y);
    return spmv_launch<0>(ctx, vals, x, nullptr, nullptr, y, nullptr, nullptr, nullptr, nullptr);
}

int la_spmv2(sc_ctx* ctx, const double* va, const double* xa, const double* vb, const double* xb, double* y) {
    return spmv_launch<1>(ctx, va, xa, vb, xb, y, nullptr, nullptr, nullptr, nullptr);
}

// u_next (in place over u_prev) = inv_d*(-K w) + alpha*u - (alpha-1)*u_prev;  w_next = (1+g) u_next - g u when w_next != null
// (g = c1/dt: lagged stiffness-proportional damping; with g = 0 the caller passes w = u and w_next = null)
int la_cd_step(sc_ctx* ctx, const double* K, const double* w, const double* u, double* uprev_next, const double* inv_d,
               const double* alpha, double g, double* w_next) {
    if (la_node_usable(ctx)) return la_node_cd_step(ctx, K, w, u, uprev_next, inv_d, alpha, g, w_next);
    if (la_tma_usable(ctx)) return la_tma_cd_step(ctx, K, w, u, uprev_next, inv_d, alpha, g, w_next);
    return spmv_launch<2>(ctx, K, w, nullptr, nullptr, uprev_next, inv_d, alpha, nullptr, nullptr, u, w_next, g);
}

// q = A p and d_out[0] = p.q (device scalar)
int la_spmv_dot(sc_ctx* ctx, const double* vals, const double* p, double* q, double* d_out) {
    SC_TRY(la_scratch(ctx));
    unsigned nb = 0;
    if (la_node_usable(ctx)) SC_TRY(la_node_spmv_dot(ctx, vals, p, q, ctx->d_partial, &nb));
    else if (la_tma_usable(ctx)) SC_TRY(la_tma_spmv_dot(ctx, vals, p, q, ctx->d_partial, &nb));
    else SC_TRY(spmv_launch<3>(ctx, vals, p, nullptr, nullptr, q, nullptr, nullptr, ctx->d_partial, &nb));
    k_reduce_final<<<1, RED_THREADS, 0, ctx->stream>>>(ctx->d_partial, nb, 1, 0, d_out);
    SC_CHECK_LAUNCH(ctx);
    if (ctx->world > 1) SC_TRY(dist_allreduce_sum(ctx, d_out, 1, ctx->stream));
    return SC_OK;
}

int la_dot(sc_ctx* ctx, const double* x, const double* y, double* d_out) {
    SC_TRY(la_scratch(ctx));
    k_dot<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(x, y, ctx->n_eq, ctx->d_partial);
    SC_CHECK_LAUNCH(ctx);
    k_reduce_final<<<1, RED_THREADS, 0, ctx->stream>>>(ctx->d_partial, RED_BLOCKS, 1, 0, d_out);
    SC_CHECK_LAUNCH(ctx);
    if (ctx->world > 1) SC_TRY(dist_allreduce_sum(ctx, d_out, 1, ctx->stream));
    return SC_OK;
}

int la_cabs_spmv_add(sc_ctx* ctx, const double* x, double* y, double scale) {
    if (ctx->cabs_rows == 0) return SC_OK;
    k_cabs_spmv_add<<<nblk(ctx->cabs_rows, 128), 128, 0, ctx->stream>>>(ctx->d_cabs_rowid, ctx->d_cabs_rptr, ctx->d_cabs_col,
                                                                        ctx->d_cabs_val, x, y, scale, ctx->cabs_rows);
    SC_CHECK_LAUNCH(ctx);
    return SC_OK;
}

int la_cabs_add_values(sc_ctx* ctx, double* vals, double scale) {
    if (ctx->cabs_n == 0) return SC_OK;
    k_cabs_add_values<<<nblk(ctx->cabs_n, 128), 128, 0, ctx->stream>>>(ctx->d_cabs_slot, ctx->d_cabs_val, vals, scale, ctx->cabs_n);
    SC_CHECK_LAUNCH(ctx);
    return SC_OK;
}

int la_axpby_vals(sc_ctx* ctx, double* out, double a, const double* x, double b, const double* y, int64_t n) {
    if (n == 0) return SC_OK;
    k_axpby<<<nblk(n, 256), 256, 0, ctx->stream>>>(out, a, x, b, y, n);
    SC_CHECK_LAUNCH(ctx);
    return SC_OK;
}

int la_extract_diag(sc_ctx* ctx, const double* vals, double* diag, bool invert) {
    k_extract_diag<<<nblk(ctx->n_eq, 256), 256, 0, ctx->stream>>>(ctx->d_rowptr, ctx->d_col, vals, diag, ctx->n_eq, invert ? 1 : 0);
    SC_CHECK_LAUNCH(ctx);
    return SC_OK;
}

int la_fill(sc_ctx* ctx, double* x, double v, int64_t n) {
    if (n == 0) return SC_OK;
    k_fill<<<nblk(n, 256), 256, 0, ctx->stream>>>(x, v, n);
    SC_CHECK_LAUNCH(ctx);
    return SC_OK;
}
