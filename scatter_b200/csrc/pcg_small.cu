// Whole Jacobi-PCG solve in ONE cooperative kernel, for systems whose iteration is launch-bound.
//
// The reference integrates its own test meshes (1e3 .. 1e5 equations) with a sparse direct solver
// (scatter/scatter.py:120-159 -> solvers.newmark_solver); on such systems the stream-ordered PCG of timeloop.cu spends its
// time in launches and in the host's stopping test (6 kernels + one synchronisation per iteration, ~30 us for a 1 656-dof
// column).  Here all iterations run on the device: a co-resident grid, three grid barriers per iteration, reductions in a
// fixed order (warp partial -> block partial -> every block sums the block partials in the same order), so the result
// is deterministic and every block takes the same stopping decision.  Same recurrences and stopping rule as pcg() in
// timeloop.cu.  Used on single-GPU contexts up to SMALL_PCG_MAX_N equations.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>

#include "common.h"

namespace cg = cooperative_groups;

namespace {

constexpr int TPB = 256;
constexpr int WPB = TPB / 32;

struct SmallPcgArgs {
    const int64_t* rowptr; const int32_t* col; const double* vals; const double* dinv; const double* b;
    double *x, *r, *p, *q;
    double* partial;      // [3 * gridDim.x]
    double* out;          // [0] iterations  [1] r.r  [2] reference norm^2 used by the stopping test  [3] b.b
    int64_t n;
    double rtol2, ref_norm2;
    int maxit;
};

// sums of partial[0 .. m) and partial[m .. 2m) in a fixed order, same values in every thread of the block
__device__ __forceinline__ void block_sum_of_partials2(const double* partial, int m, double* sh, double& s0, double& s1) {
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
        double a = 0.0, b = 0.0;
        for (int k = lane; k < m; k += 32) { a += __ldcg(partial + k); b += __ldcg(partial + m + k); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        if (lane == 0) { sh[0] = a; sh[1] = b; }
    }
    __syncthreads();
    s0 = sh[0]; s1 = sh[1];
    __syncthreads();
}

// block-level sums of two values in a fixed order; results valid in thread 0
__device__ __forceinline__ void block_reduce2(double& v0, double& v1, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v0 += __shfl_down_sync(0xffffffffu, v0, o); v1 += __shfl_down_sync(0xffffffffu, v1, o); }
    if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = v0; sh[WPB + (threadIdx.x >> 5)] = v1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < WPB; ++w) { a += sh[w]; b += sh[WPB + w]; }
        v0 = a; v1 = b;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TPB, 4) k_pcg_small(SmallPcgArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[2 * WPB];
    __shared__ double sh1[2];
    const int G = gridDim.x;
    const int64_t tid = (int64_t)blockIdx.x * TPB + threadIdx.x, nthreads = (int64_t)G * TPB;
    constexpr int LPR = 8;                               // lanes per row: four rows of a warp in flight
    const int sub = threadIdx.x & (LPR - 1);
    const int64_t grp = tid / LPR, ngrp = nthreads / LPR;

    // x = 0, r = b, p = z = dinv b
    double rz_p = 0.0, bb_p = 0.0;
    for (int64_t i = tid; i < a.n; i += nthreads) {
        const double bi = a.b[i], z = a.dinv[i] * bi;
        a.x[i] = 0.0; a.r[i] = bi; a.p[i] = z;
        rz_p += bi * z; bb_p += bi * bi;
    }
    block_reduce2(rz_p, bb_p, sh);
    if (threadIdx.x == 0) { a.partial[blockIdx.x] = rz_p; a.partial[G + blockIdx.x] = bb_p; }
    grid.sync();
    double rz, bb;
    block_sum_of_partials2(a.partial, G, sh1, rz, bb);
    double ref = bb;
    int it = 0;
    double rr = bb;
    bool done = !(bb > 0.0);
    if (!done && a.ref_norm2 > 0.0) {
        if (bb <= a.rtol2 * a.ref_norm2) done = true;
        ref = a.ref_norm2;
    }
    const double target = a.rtol2 * ref;
    // partial[] has three slices: [0, 2G) for (r.z, r.r) and [2G, 3G) for p.q, so a slow block still reading one slice
    // never meets a fast block writing the other
    while (!done && it < a.maxit) {
        ++it;
        // q = A p (eight lanes per row), partial of p.q
        double pq_p = 0.0, zero = 0.0;
        for (int64_t row0 = 0; row0 < a.n; row0 += ngrp) {          // uniform trip count: the shuffles below are full-warp
            const int64_t row = row0 + grp;
            double s = 0.0;
            if (row < a.n) {
                const int64_t lo = a.rowptr[row], hi = a.rowptr[row + 1];
#pragma unroll 4
                for (int64_t k = lo + sub; k < hi; k += LPR) s += a.vals[k] * a.p[a.col[k]];
            }
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            if (sub == 0 && row < a.n) { a.q[row] = s; pq_p += a.p[row] * s; }
        }
        block_reduce2(pq_p, zero, sh);
        if (threadIdx.x == 0) a.partial[2 * G + blockIdx.x] = pq_p;
        grid.sync();
        double pq;
        {
            const int lane = threadIdx.x & 31;
            if (threadIdx.x < 32) {
                double v = 0.0;
                for (int k = lane; k < G; k += 32) v += __ldcg(a.partial + 2 * G + k);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) sh1[0] = v;
            }
            __syncthreads();
            pq = sh1[0];
            __syncthreads();
        }
        const double alpha = pq != 0.0 ? rz / pq : 0.0;
        // x += alpha p ; r -= alpha q ; partials of r.z and r.r
        double rz_n = 0.0, rr_n = 0.0;
        for (int64_t i = tid; i < a.n; i += nthreads) {
            a.x[i] += alpha * a.p[i];
            const double ri = a.r[i] - alpha * a.q[i];
            a.r[i] = ri;
            rz_n += ri * (a.dinv[i] * ri); rr_n += ri * ri;
        }
        block_reduce2(rz_n, rr_n, sh);
        if (threadIdx.x == 0) { a.partial[blockIdx.x] = rz_n; a.partial[G + blockIdx.x] = rr_n; }
        grid.sync();
        double rz_new;
        block_sum_of_partials2(a.partial, G, sh1, rz_new, rr);
        if (!(rr == rr) || rr <= target) { done = true; break; }
        const double beta = rz != 0.0 ? rz_new / rz : 0.0;
        rz = rz_new;
        for (int64_t i = tid; i < a.n; i += nthreads) a.p[i] = a.dinv[i] * a.r[i] + beta * a.p[i];
        grid.sync();
    }
    if (tid == 0) { a.out[0] = (double)it; a.out[1] = rr; a.out[2] = ref; a.out[3] = bb; }
}

}  // namespace

bool pcg_small_usable(sc_ctx* ctx) {
    return ctx->world == 1 && !ctx->no_small_pcg && ctx->n_eq <= SMALL_PCG_MAX_N && ctx->d_rowptr && ctx->d_col;
}

// Same contract as pcg() in timeloop.cu.
int pcg_small(sc_ctx* ctx, const double* vals, const double* dinv, const double* b, double* x, double* r, double* p, double* q,
              double rtol, int maxit, int* iters, double* relres, double ref_norm2) {
    SC_TRY(la_scratch(ctx));
    if (ctx->small_pcg_grid == 0) {
        int per_sm = 0, coop = 0;
        SC_CUDA(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
        if (!coop) return sc_fail(ctx, SC_ERR_UNSUPPORTED, "device does not support cooperative launches");
        SC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_small, TPB, 0));
        ctx->small_pcg_grid = std::max(1, std::min(per_sm, 4) * ctx->sm_count);
    }
    const int64_t want = (ctx->n_eq + 63) / 64;           // about 8 rows per warp on the smallest systems
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, ctx->small_pcg_grid));
    if (3 * (size_t)grid > SC_PARTIAL_DOUBLES) return sc_fail(ctx, SC_ERR_UNSUPPORTED, "partial buffer too small for the cooperative PCG");
    SmallPcgArgs a;
    a.rowptr = ctx->d_rowptr; a.col = ctx->d_col; a.vals = vals; a.dinv = dinv; a.b = b;
    a.x = x; a.r = r; a.p = p; a.q = q;
    a.partial = ctx->d_partial; a.out = ctx->d_scal + 16;
    a.n = ctx->n_eq; a.rtol2 = rtol * rtol; a.ref_norm2 = ref_norm2; a.maxit = maxit;
    void* args[] = {&a};
    SC_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)k_pcg_small, dim3(grid), dim3(TPB), args, 0, ctx->stream));
    ctx->launches++;
    SC_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 16, ctx->d_scal + 16, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double* o = ctx->h_pinned + 16;
    *iters = (int)o[0];
    *relres = 0.0;
    if (!(o[3] > 0.0) || *iters == 0) return SC_OK;      // zero right-hand side (x = 0) or already below the caller's scale
    const double rr = o[1];
    *relres = std::sqrt(rr / o[2]);
    if (!(rr == rr)) return sc_fail(ctx, SC_ERR_NOCONV, "PCG produced NaN at iteration %d", *iters);
    if (rr <= rtol * rtol * o[2]) return SC_OK;
    return sc_fail(ctx, SC_ERR_NOCONV, "PCG did not converge in %d iterations (relative residual %.3e, target %.3e)", maxit, *relres, rtol);
}
