// Sparse / dense vector kernels of the time loop: streaming CSR SpMV, fused vector updates, deterministic dots.
//
// These replace the scipy.sparse products the reference's external solver performs every step
// (solvers.NewmarkExplicit.calculate called at scatter/scatter.py:159; recurrence in SURVEY.md 3.3).
//
// SpMV layout ("CSR stream"): a thread block owns R consecutive rows = one contiguous slice of the value / column
// arrays.  All threads stream that slice with fully coalesced loads, multiply by the gathered x entries (L2-resident
// for FEM orderings) and park the products in shared memory; then one thread per row adds its products in column
// order.  No atomics, fixed order => bit-reproducible, and identical to a sequential CSR row sum.
#include <algorithm>
#include "common.h"

namespace {

constexpr int SPMV_THREADS = 256;
constexpr int SPMV_TILE = 4096;      // products staged per pass (32 KB)
constexpr int SPMV_MAX_ROWS = 256;

// MODE 0: y = A xa            MODE 1: y = A xa + B xb
// MODE 2: central-difference step  un[i] = inv_d[i]*(-sum) + alpha[i]*xa[i] - (alpha[i]-1)*un[i]   (un holds u_prev)
// MODE 3: PCG product  y = A xa  and per-block partial of xa.y  (partial[blockIdx])
template <int MODE>
__global__ void __launch_bounds__(SPMV_THREADS)
k_spmv(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ va,
       const double* __restrict__ xa, const double* __restrict__ vb, const double* __restrict__ xb,
       double* __restrict__ y, const double* __restrict__ inv_d, const double* __restrict__ alpha,
       double* __restrict__ partial, int64_t n_rows, int R) {
    __shared__ double prod[SPMV_TILE];
    __shared__ int64_t rp[SPMV_MAX_ROWS + 1];
    __shared__ double red[SPMV_THREADS / 32];
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * R;
    const int nr = (int)min((int64_t)R, n_rows - r0);
    for (int t = tid; t <= nr; t += SPMV_THREADS) rp[t] = rowptr[r0 + t];
    __syncthreads();
    const int64_t s0 = rp[0], s1 = rp[nr];
    double acc = 0.0;
    int64_t rs = 0, re = 0;
    if (tid < nr) { rs = rp[tid]; re = rp[tid + 1]; }
    for (int64_t c0 = s0; c0 < s1; c0 += SPMV_TILE) {
        const int cnt = (int)min((int64_t)SPMV_TILE, s1 - c0);
#pragma unroll 4
        for (int k = tid; k < cnt; k += SPMV_THREADS) {
            const int64_t g = c0 + k;
            const int c = col[g];
            double v = va[g] * xa[c];
            if (MODE == 1) v += vb[g] * xb[c];
            prod[k] = v;
        }
        __syncthreads();
        if (tid < nr) {
            const int b = (int)(max(rs, c0) - c0), e = (int)(min(re, c0 + cnt) - c0);
            for (int k = b; k < e; ++k) acc += prod[k];
        }
        __syncthreads();
    }
    double dotv = 0.0;
    if (tid < nr) {
        const int64_t i = r0 + tid;
        if (MODE == 2) {
            if (re > rs) {
                const double al = alpha[i];
                y[i] = inv_d[i] * (-acc) + al * xa[i] - (al - 1.0) * y[i];
            }
        } else {
            y[i] = acc;
            if (MODE == 3) dotv = xa[i] * acc;
        }
    }
    if (MODE == 3) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dotv += __shfl_down_sync(0xffffffffu, dotv, o);
        if ((tid & 31) == 0) red[tid >> 5] = dotv;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < SPMV_THREADS / 32; ++w) s += red[w];
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
__global__ void k_axpby(double* __restrict__ out, double a, const double* __restrict__ x, double b, const double* __restrict__ y, int64_t n) {
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

int spmv_rows_per_block(sc_ctx* ctx) {
    double avg = ctx->n_eq > 0 ? (double)ctx->nnz / (double)ctx->n_eq : 1.0;
    int R = (int)(SPMV_TILE / std::max(avg, 1.0));
    return std::max(8, std::min(R, SPMV_MAX_ROWS));
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
    const int R = spmv_rows_per_block(ctx);
    k_spmv<0><<<nblk(ctx->n_eq, R), SPMV_THREADS, 0, ctx->stream>>>(ctx->d_rowptr, ctx->d_col, vals, x, nullptr, nullptr, y,
                                                                     nullptr, nullptr, nullptr, ctx->n_eq, R);
    SC_CHECK_LAUNCH(ctx);
    return SC_OK;
}

int la_spmv2(sc_ctx* ctx, const double* va, const double* xa, const double* vb, const double* xb, double* y) {
    const int R = spmv_rows_per_block(ctx);
    k_spmv<1><<<nblk(ctx->n_eq, R), SPMV_THREADS, 0, ctx->stream>>>(ctx->d_rowptr, ctx->d_col, va, xa, vb, xb, y, nullptr,
                                                                     nullptr, nullptr, ctx->n_eq, R);
    SC_CHECK_LAUNCH(ctx);
    return SC_OK;
}

// u_next (in place over u_prev) = inv_d*(-K u) + alpha*u - (alpha-1)*u_prev
int la_cd_step(sc_ctx* ctx, const double* K, const double* u, double* uprev_next, const double* inv_d, const double* alpha) {
    const int R = spmv_rows_per_block(ctx);
    k_spmv<2><<<nblk(ctx->n_eq, R), SPMV_THREADS, 0, ctx->stream>>>(ctx->d_rowptr, ctx->d_col, K, u, nullptr, nullptr,
                                                                     uprev_next, inv_d, alpha, nullptr, ctx->n_eq, R);
    SC_CHECK_LAUNCH(ctx);
    return SC_OK;
}

// q = A p and d_out[0] = p.q (device scalar)
int la_spmv_dot(sc_ctx* ctx, const double* vals, const double* p, double* q, double* d_out) {
    SC_TRY(la_scratch(ctx));
    const int R = spmv_rows_per_block(ctx);
    const unsigned nb = nblk(ctx->n_eq, R);
    k_spmv<3><<<nb, SPMV_THREADS, 0, ctx->stream>>>(ctx->d_rowptr, ctx->d_col, vals, p, nullptr, nullptr, q, nullptr, nullptr,
                                                    ctx->d_partial, ctx->n_eq, R);
    SC_CHECK_LAUNCH(ctx);
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
