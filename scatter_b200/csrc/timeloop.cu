// Time integration on the device.
//
// Replaces the external solver the reference calls at scatter/scatter.py:156-159
// (PuggleSolvers 1.0.1: NewmarkExplicit / CentralDifferenceSolver .calculate(M, C, K, F, t0, t1)); protocol and
// recurrence as written out in SURVEY.md 3.3 (validated against every golden history of the reference).
//
//   Newmark (beta, gamma), incremental form, one Jacobi-PCG solve with the fixed effective matrix per step:
//       Khat = K + a4 C + a1 M,  a1 = 1/(beta dt^2), a4 = gamma/(beta dt),  C = C_abs + c0 M + c1 K (never stored)
//       rhs  = dF + M (v/(beta dt) + a/(2 beta)) + C ((gamma/beta) v + dt (gamma/(2 beta) - 1) a)
//            = dF + M x1 + K x2 + C_abs q,   x1 = p + c0 q, x2 = c1 q            (one fused two-matrix SpMV)
//   Central difference with row-sum lumped M (diagonal system, one fused SpMV+update kernel per step).  The damping
//   C = C_abs + c0 M + c1 K is split: c_d = c0 m + rowsum(C_abs) is diagonal and centred, the stiffness-proportional
//   part acts on the lagged velocity (u - u-)/dt and rides the same SpMV:
//       u+ = inv_d (F - K w) + alpha u - (alpha - 1) u-,   w = (1 + c1/dt) u - (c1/dt) u-,
//       inv_d = 1/(m/dt^2 + c_d/(2dt)),  alpha = 2 m inv_d / dt^2
#include <chrono>
#include <cmath>
#include "common.h"

namespace {

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

// scalars on the device: [0] rz  [1] pq  [2] rz_new  [3] rr  [4] bb
__global__ void k_pcg_init(const double* __restrict__ b, const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r,
                           double* __restrict__ p, int64_t n, double* __restrict__ partial, int nb) {
    // x = 0, r = b, p = z = dinv*r ; partials of r.z and r.r
    __shared__ double sh[2][8];
    const int64_t chunk = (n + nb - 1) / nb;
    const int64_t s = blockIdx.x * chunk, e = min(s + chunk, n);
    double rz = 0.0, rr = 0.0;
    for (int64_t i = s + threadIdx.x; i < e; i += blockDim.x) {
        const double bi = b[i], z = dinv[i] * bi;
        x[i] = 0.0; r[i] = bi; p[i] = z;
        rz += bi * z; rr += bi * bi;
    }
    for (int o = 16; o > 0; o >>= 1) { rz += __shfl_down_sync(0xffffffffu, rz, o); rr += __shfl_down_sync(0xffffffffu, rr, o); }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = rz; sh[1][threadIdx.x >> 5] = rr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sh[0][w]; c += sh[1][w]; }
        partial[blockIdx.x] = a; partial[nb + blockIdx.x] = c;
    }
}

// x += alpha p ; r -= alpha q ; partials of r.(dinv r) and r.r          alpha = scal[0]/scal[1]
__global__ void k_pcg_update(const double* __restrict__ scal, const double* __restrict__ p, const double* __restrict__ q,
                             const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r, int64_t n,
                             double* __restrict__ partial, int nb) {
    __shared__ double sh[2][8];
    const double alpha = scal[1] != 0.0 ? scal[0] / scal[1] : 0.0;   // p.Ap = 0 only once r = 0 (iterations past convergence)
    const int64_t chunk = (n + nb - 1) / nb;
    const int64_t s = blockIdx.x * chunk, e = min(s + chunk, n);
    double rz = 0.0, rr = 0.0;
    for (int64_t i = s + threadIdx.x; i < e; i += blockDim.x) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        rz += ri * (dinv[i] * ri); rr += ri * ri;
    }
    for (int o = 16; o > 0; o >>= 1) { rz += __shfl_down_sync(0xffffffffu, rz, o); rr += __shfl_down_sync(0xffffffffu, rr, o); }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = rz; sh[1][threadIdx.x >> 5] = rr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sh[0][w]; c += sh[1][w]; }
        partial[blockIdx.x] = a; partial[nb + blockIdx.x] = c;
    }
}

// out[j] = sum partial[j*nb + k]  (one block, fixed order)
__global__ void k_reduce2(const double* __restrict__ partial, int nb, double* __restrict__ out0, double* __restrict__ out1) {
    __shared__ double sh[8];
    for (int j = 0; j < 2; ++j) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nb; i += blockDim.x) s += partial[j * nb + i];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
            *(j == 0 ? out0 : out1) = t;
        }
        __syncthreads();
    }
}

// p = dinv r + beta p, beta = scal[2]/scal[0]; then rz <- rz_new is done by k_shift
__global__ void k_pcg_p(const double* __restrict__ scal, const double* __restrict__ r, const double* __restrict__ dinv,
                        double* __restrict__ p, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double beta = scal[0] != 0.0 ? scal[2] / scal[0] : 0.0;
    p[i] = dinv[i] * r[i] + beta * p[i];
}
__global__ void k_shift(double* scal) { scal[0] = scal[2]; }

// ---- PCG with a general preconditioner z = P r (fsai.cu): the vector kernels without the Jacobi scaling ---------------
// x = 0, r = b ; partials of r.r in partial[nb + block]
__global__ void k_pcg_init_np(const double* __restrict__ b, double* __restrict__ x, double* __restrict__ r, int64_t n,
                              double* __restrict__ partial, int nb) {
    __shared__ double sh[8];
    const int64_t chunk = (n + nb - 1) / nb;
    const int64_t s = blockIdx.x * chunk, e = min(s + chunk, n);
    double rr = 0.0;
    for (int64_t i = s + threadIdx.x; i < e; i += blockDim.x) {
        const double bi = b[i];
        x[i] = 0.0; r[i] = bi;
        rr += bi * bi;
    }
    for (int o = 16; o > 0; o >>= 1) rr += __shfl_down_sync(0xffffffffu, rr, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = rr;
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) c += sh[w];
        partial[nb + blockIdx.x] = c;
    }
}
// x += alpha p ; r -= alpha q ; partials of r.r in partial[nb + block]          alpha = scal[0]/scal[1]
__global__ void k_pcg_update_np(const double* __restrict__ scal, const double* __restrict__ p, const double* __restrict__ q,
                                double* __restrict__ x, double* __restrict__ r, int64_t n, double* __restrict__ partial, int nb) {
    __shared__ double sh[8];
    const double alpha = scal[1] != 0.0 ? scal[0] / scal[1] : 0.0;
    const int64_t chunk = (n + nb - 1) / nb;
    const int64_t s = blockIdx.x * chunk, e = min(s + chunk, n);
    double rr = 0.0;
    for (int64_t i = s + threadIdx.x; i < e; i += blockDim.x) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        rr += ri * ri;
    }
    for (int o = 16; o > 0; o >>= 1) rr += __shfl_down_sync(0xffffffffu, rr, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = rr;
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) c += sh[w];
        partial[nb + blockIdx.x] = c;
    }
}
// p = z + beta p, beta = scal[2]/scal[0]; z is read through the factor's renumbering when there is one
__global__ void k_pcg_pz(const double* __restrict__ scal, const double* __restrict__ z, const int32_t* __restrict__ perm,
                         double* __restrict__ p, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double beta = scal[0] != 0.0 ? scal[2] / scal[0] : 0.0;
    p[i] = z[perm ? perm[i] : i] + beta * p[i];
}

// ---- projection onto previous solutions (Fischer 1998): dots with / combinations of up to proj_k vectors ---------------
// partial[v * nb + block] = sum over the block's chunk of V[v][i] y[i], v < nv; with `self` one more entry: y.y
__global__ void k_multidot(const double* const* __restrict__ V, int nv, int self, const double* __restrict__ y, int64_t n,
                           double* __restrict__ partial, int nb) {
    __shared__ double sh[8];
    const int64_t chunk = (n + nb - 1) / nb;
    const int64_t s = blockIdx.x * chunk, e = min(s + chunk, n);
    for (int v = 0; v < nv + self; ++v) {
        const double* x = v < nv ? V[v] : y;
        double d = 0.0;
        for (int64_t i = s + threadIdx.x; i < e; i += blockDim.x) d += x[i] * y[i];
        for (int o = 16; o > 0; o >>= 1) d += __shfl_down_sync(0xffffffffu, d, o);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = d;
        __syncthreads();
        if (threadIdx.x == 0) {
            double c = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) c += sh[w];
            partial[(int64_t)v * nb + blockIdx.x] = c;
        }
        __syncthreads();
    }
}
// out[v] = sum_k partial[v * nb + k]   (one block per v, fixed order)
__global__ void k_reduce_multi(const double* __restrict__ partial, int nb, double* __restrict__ out) {
    __shared__ double sh[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s += partial[(int64_t)blockIdx.x * nb + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        out[blockIdx.x] = t;
    }
}
// out = base + sign * sum_v coef[v] V[v]      (out may alias base)
__global__ void k_multiaxpy(double* out, const double* base, double sign, const double* __restrict__ coef,
                            const double* const* __restrict__ V, int nv, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int v = 0; v < nv; ++v) s += coef[v] * V[v][i];
    out[i] = base[i] + sign * s;
}
// x *= 1/sqrt(*nrm2), y likewise (both zeroed when the norm is not positive: a null vector in the basis is harmless)
__global__ void k_scale_pair(double* __restrict__ x, double* __restrict__ y, const double* __restrict__ nrm2, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double f = *nrm2 > 0.0 ? rsqrt(*nrm2) : 0.0;
    x[i] *= f; y[i] *= f;
}
__global__ void k_copy1(const double* src, double* dst) { *dst = *src; }

// Newmark helper vectors
__global__ void k_nm_inputs(const double* __restrict__ v, const double* __restrict__ a, double* __restrict__ x1, double* __restrict__ x2,
                            double* __restrict__ q, double pv, double pa, double qv, double qa, double c0, double c1, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double vi = v[i], ai = a[i];
    const double pp = pv * vi + pa * ai, qq = qv * vi + qa * ai;
    x1[i] = pp + c0 * qq;
    x2[i] = c1 * qq;
    if (q) q[i] = qq;
}
__global__ void k_nm_update(const double* __restrict__ du, double* __restrict__ u, double* __restrict__ v, double* __restrict__ a,
                            double a4, double gb, double dvc, double a1, double ivb, double i2b, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d = du[i], vi = v[i], ai = a[i];
    const double dv = a4 * d - gb * vi + dvc * ai;
    const double da = a1 * d - ivb * vi - i2b * ai;
    u[i] += d; v[i] = vi + dv; a[i] = ai + da;
}
// sparse load entries: y[dof] += scale * val   (dofs are unique within a step)
// sel 0: every entry; 1: dofs outside [lo, hi); 2: dofs inside (the two launches of a step split for the halo overlap)
__global__ void k_load_add(const int32_t* __restrict__ dof, const double* __restrict__ val, int64_t n, double scale,
                           const double* __restrict__ mult, double* __restrict__ y, int sel, int64_t lo, int64_t hi) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int d = dof[t];
    const bool inside = d >= lo && d < hi;
    if ((sel == 1 && inside) || (sel == 2 && !inside)) return;
    y[d] += scale * val[t] * (mult ? mult[d] : 1.0);
}
__global__ void k_cd_coeffs(const double* __restrict__ m, const double* __restrict__ c, double a0, double a1, double* __restrict__ inv_d,
                            double* __restrict__ alpha, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d = a0 * m[i] + a1 * c[i];
    const double id = d != 0.0 ? 1.0 / d : 0.0;
    inv_d[i] = id;
    alpha[i] = 2.0 * a0 * m[i] * id;
}
// start-up: u_prev = u - dt v + dt^2/2 * (f - Ku - c v)/m       (rhs holds f - Ku on entry)
__global__ void k_cd_start(const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ rhs,
                           const double* __restrict__ m, const double* __restrict__ c, double dt, double* __restrict__ uprev, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double mi = m[i];
    const double acc = mi != 0.0 ? (rhs[i] - c[i] * v[i]) / mi : 0.0;
    uprev[i] = u[i] - dt * v[i] + 0.5 * dt * dt * acc;
}
__global__ void k_cd_va(const double* __restrict__ unext, const double* __restrict__ u, const double* __restrict__ uprev, double a0, double a1,
                        double* __restrict__ v, double* __restrict__ a, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    v[i] = a1 * (unext[i] - uprev[i]);
    a[i] = a0 * (unext[i] - 2.0 * u[i] + uprev[i]);
}
__global__ void k_neg_add(double* __restrict__ y, int64_t n) {   // y = -y
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) y[i] = -y[i];
}


// out = c0 x0 + c1 x1 + c2 x2 + c3 x3 (null pointers are skipped; out may alias x0: every thread reads its entry before writing it)
__global__ void k_lincomb(double* out, double c0, const double* x0, double c1, const double* __restrict__ x1,
                          double c2, const double* __restrict__ x2, double c3, const double* __restrict__ x3, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = c0 * x0[i];
    if (x1) s += c1 * x1[i];
    if (x2) s += c2 * x2[i];
    if (x3) s += c3 * x3[i];
    out[i] = s;
}
int lincomb(sc_ctx* ctx, double* out, double c0, const double* x0, double c1, const double* x1, double c2 = 0.0,
            const double* x2 = nullptr, double c3 = 0.0, const double* x3 = nullptr) {
    k_lincomb<<<nblk(ctx->n_eq, 256), 256, 0, ctx->stream>>>(out, c0, x0, c1, x1, c2, x2, c3, x3, ctx->n_eq);
    SC_CHECK_LAUNCH(ctx);
    return SC_OK;
}
// rhs = M xm + C xc with C = C_abs + c0 M + c1 K  ->  M (xm + c0 xc) + K (c1 xc) + C_abs xc   (x1, x2 are scratch)
int apply_M_C(sc_ctx* ctx, const double* xm, const double* xc, double* x1, double* x2, double* rhs) {
    if (ctx->d_C) return la_spmv2(ctx, ctx->d_M, xm, ctx->d_C, xc, rhs);   // explicit damping matrix (sc_set_csr)
    SC_TRY(lincomb(ctx, x1, 1.0, xm, ctx->c0, xc));
    SC_TRY(lincomb(ctx, x2, ctx->c1, xc, 0.0, nullptr));
    if (ctx->world > 1) { SC_TRY(dist_halo(ctx, x1, ctx->stream)); SC_TRY(dist_halo(ctx, x2, ctx->stream)); }
    SC_TRY(la_spmv2(ctx, ctx->d_M, x1, ctx->d_K, x2, rhs));
    SC_TRY(la_cabs_spmv_add(ctx, xc, rhs, 1.0));
    return SC_OK;
}

constexpr int PCG_NB = 1024;

// Jacobi-preconditioned CG for A x = b, x0 = 0.  Returns iterations and the relative residual.
// non-finite detector for the explicit schemes: one pass over u at the end of a stage (0.07 ms for 50 M dofs)
__global__ void k_flag_nonfinite(const double* __restrict__ u, int64_t n, double* __restrict__ flag) {
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) bad |= !isfinite(u[i]);
    if (__syncthreads_or(bad) && threadIdx.x == 0) *flag = 1.0;
}

}  // namespace

void pcg_graph_drop(sc_ctx* ctx) {
    for (auto& g : ctx->pcg_graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g.exec = nullptr;
    }
}

namespace {

// `F` != null: FSAI preconditioner z = G^T G r instead of Jacobi (needs the two extra vectors z, tv).  `ref_dev`: the
// reference norm^2 of the stopping test was put into d_scal[4] by the caller (projection onto previous solutions).
int pcg(sc_ctx* ctx, const double* vals, const double* dinv, const double* b, double* x, double* r, double* p, double* q,
        double rtol, int maxit, int* iters, double* relres, double ref_norm2 = -1.0, const sc_fsai* F = nullptr, double* z = nullptr,
        double* tv = nullptr, bool ref_dev = false) {
    if (pcg_small_usable(ctx) && !ref_dev) return pcg_small(ctx, vals, dinv, b, x, r, p, q, rtol, maxit, iters, relres, ref_norm2);
    const int64_t n = ctx->n_eq;
    cudaStream_t st = ctx->stream;
    double* sc = ctx->d_scal;
    if (F) {
        k_pcg_init_np<<<PCG_NB, 256, 0, st>>>(b, x, r, n, ctx->d_partial, PCG_NB);
        SC_CHECK_LAUNCH(ctx);
        SC_TRY(fsai_apply(ctx, F, r, tv, z, ctx->d_partial, PCG_NB, q));       // z = G^T G r (factor numbering), partials of r.z; q: scratch here
        SC_TRY(fsai_unpermute(ctx, F, z, p));
    } else {
        k_pcg_init<<<PCG_NB, 256, 0, st>>>(b, dinv, x, r, p, n, ctx->d_partial, PCG_NB);
        SC_CHECK_LAUNCH(ctx);
    }
    // r.z and r.r land next to each other ([2], [3]): one all-reduce for both, then [0] <- [2]
    k_reduce2<<<1, 256, 0, st>>>(ctx->d_partial, PCG_NB, sc + 2, sc + 3);
    SC_CHECK_LAUNCH(ctx);
    if (ctx->world > 1) SC_TRY(dist_allreduce_sum(ctx, sc + 2, 2, st));
    k_shift<<<1, 1, 0, st>>>(sc);
    SC_CHECK_LAUNCH(ctx);
    SC_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, sc, 5 * sizeof(double), cudaMemcpyDeviceToHost, st));
    SC_CUDA(ctx, cudaStreamSynchronize(st));
    double bb = ctx->h_pinned[3];
    *iters = 0;
    *relres = 0.0;
    if (ref_dev) ref_norm2 = ctx->h_pinned[4];
    if (!(bb > 0.0)) return SC_OK;   // zero right-hand side: x = 0
    if (ref_norm2 > 0.0) {           // stopping test relative to a caller-supplied scale (incremental static solve)
        if (bb <= rtol * rtol * ref_norm2) return SC_OK;
        bb = ref_norm2;
    }
    const double target = rtol * rtol * bb;
    // one iteration = six small launches + a 40-byte read-back; on small and mid-size meshes the launch overhead is
    // the iteration time, so the sequence is captured once into a CUDA graph (single-GPU runs; NCCL calls stay eager)
    auto iteration = [&]() -> int {
        if (ctx->world > 1) SC_TRY(dist_halo(ctx, p, st));
        SC_TRY(la_spmv_dot(ctx, vals, p, q, sc + 1));
        if (F) {
            k_pcg_update_np<<<PCG_NB, 256, 0, st>>>(sc, p, q, x, r, n, ctx->d_partial, PCG_NB);
            SC_CHECK_LAUNCH(ctx);
            SC_TRY(fsai_apply(ctx, F, r, tv, z, ctx->d_partial, PCG_NB, q));    // q is free between the update and the next product
        } else {
            k_pcg_update<<<PCG_NB, 256, 0, st>>>(sc, p, q, dinv, x, r, n, ctx->d_partial, PCG_NB);
            SC_CHECK_LAUNCH(ctx);
        }
        k_reduce2<<<1, 256, 0, st>>>(ctx->d_partial, PCG_NB, sc + 2, sc + 3);
        SC_CHECK_LAUNCH(ctx);
        if (ctx->world > 1) SC_TRY(dist_allreduce_sum(ctx, sc + 2, 2, st));
        SC_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, sc, 5 * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (F) k_pcg_pz<<<nblk(n, 256), 256, 0, st>>>(sc, z, F->perm, p, n);
        else k_pcg_p<<<nblk(n, 256), 256, 0, st>>>(sc, r, dinv, p, n);
        SC_CHECK_LAUNCH(ctx);
        k_shift<<<1, 1, 0, st>>>(sc);
        SC_CHECK_LAUNCH(ctx);
        return SC_OK;
    };
    const bool use_graph = ctx->world == 1 && !ctx->no_graph;
    sc_ctx::PcgGraph* gr = nullptr;
    if (use_graph) {
        const void* key[8] = {vals, F ? (const void*)F->cv : (const void*)dinv, x, r, p, q, ctx->d_nd, ctx->d_partial};
        for (auto& g : ctx->pcg_graphs) {
            bool same = g.exec != nullptr && g.n == n;
            for (int k = 0; k < 8 && same; ++k) same = g.key[k] == key[k];
            if (same) gr = &g;
        }
        if (!gr) {
            gr = ctx->pcg_graphs[0].used <= ctx->pcg_graphs[1].used ? &ctx->pcg_graphs[0] : &ctx->pcg_graphs[1];   // least recently used
            if (gr->exec) { cudaGraphExecDestroy(gr->exec); gr->exec = nullptr; }
            const int64_t l0 = ctx->launches;
            cudaGraph_t graph = nullptr;
            SC_CUDA(ctx, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            const int rc = iteration();
            const cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc != SC_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            SC_CUDA(ctx, ce);
            const cudaError_t ie = cudaGraphInstantiate(&gr->exec, graph, 0);
            cudaGraphDestroy(graph);
            SC_CUDA(ctx, ie);
            gr->launches = (int)(ctx->launches - l0);
            ctx->launches = l0;
            gr->n = n;
            for (int k = 0; k < 8; ++k) gr->key[k] = key[k];
        }
        gr->used = ++ctx->pcg_graph_clock;
    }
    // The stopping test needs a stream synchronisation.  Small systems (graph replay): every fourth iteration, the up to
    // three extra iterations only tighten the solution (alpha, beta are guarded against r = 0).  Otherwise the next test is
    // scheduled from the observed convergence rate -- half-way to the predicted end, at most 16 iterations ahead -- so that
    // the host (and, on several GPUs, the eager NCCL launches) run ahead of the device most of the time.  Every rank sees
    // the same all-reduced scalars, hence takes the same decisions.
    const bool fixed4 = use_graph && n < 100000;
    int next_check = fixed4 ? 4 : 1, last_it = 0;
    double last_rr = ctx->h_pinned[3], best_rr = last_rr;
    int best_it = 0;
    for (int it = 1; it <= maxit; ++it) {
        if (use_graph) {
            SC_CUDA(ctx, cudaGraphLaunch(gr->exec, st));
            ctx->launches += gr->launches;
        } else {
            SC_TRY(iteration());
        }
        if (it < next_check && it != maxit) continue;
        SC_CUDA(ctx, cudaStreamSynchronize(st));
        const double rr = ctx->h_pinned[3];
        *iters = it;
        *relres = std::sqrt(rr / bb);
        if (!(rr == rr)) return sc_fail(ctx, SC_ERR_NOCONV, "PCG produced NaN at iteration %d", it);
        if (rr <= target) return SC_OK;
        if (rr < 0.9 * best_rr) { best_rr = rr; best_it = it; }
        // stagnation at the round-off floor of the recursive residual: the reference's direct solve cannot fail this way,
        // so a residual that no longer decreases is accepted once it is below 1e-9 (counted in the stage statistics)
        if (it - best_it >= 100 && *relres <= 1e-9) { ctx->pcg_stagnations++; return SC_OK; }
        if (fixed4) { next_check = it + 4; continue; }
        int ahead = 1;
        if (rr < last_rr && it > last_it) {
            const double rate = (std::log(last_rr) - std::log(rr)) / (double)(it - last_it);      // per iteration
            const double remaining = (std::log(rr) - std::log(target)) / rate;
            ahead = (int)std::min(16.0, std::max(1.0, std::floor(0.5 * remaining)));
        }
        last_rr = rr; last_it = it;
        next_check = it + ahead;
    }
    return sc_fail(ctx, SC_ERR_NOCONV, "PCG did not converge in %d iterations (relative residual %.3e, target %.3e)", maxit, *relres, rtol);
}


// ---- preconditioner and initial guess of the implicit solves --------------------------------------------------------
void proj_free(sc_ctx* ctx) {
    for (auto& v : ctx->proj_x) sc_free(&v);
    for (auto& v : ctx->proj_ax) sc_free(&v);
    ctx->proj_x.clear(); ctx->proj_ax.clear();
    sc_free(&ctx->d_proj_ptr); sc_free(&ctx->d_proj_coef);
    ctx->proj_n = 0; ctx->proj_for = nullptr;
}
}  // namespace

// forget everything derived from the effective matrices (their values changed, or an option that shapes it did)
void precond_drop(sc_ctx* ctx) {
    fsai_free(&ctx->fsai[0]); fsai_free(&ctx->fsai[1]);
    ctx->proj_n = 0; ctx->proj_for = nullptr;
    if ((int)ctx->proj_x.size() != ctx->proj_k) proj_free(ctx);
    pcg_graph_drop(ctx);
}
void precond_destroy(sc_ctx* ctx) { precond_drop(ctx); proj_free(ctx); }

namespace {

// FSAI factor of `vals` in slot `slot`, built on first use; *out stays null where Jacobi is used (option, small systems
// solved by the cooperative kernel, caller-supplied patterns that are not structurally symmetric)
int precond_for(sc_ctx* ctx, int slot, const double* vals, const sc_fsai** out, double* seconds) {
    *out = nullptr;
    if (ctx->no_fsai || pcg_small_usable(ctx)) return SC_OK;
    sc_fsai* f = &ctx->fsai[slot];
    if (f->for_vals != vals) {
        const int rc = fsai_build(ctx, f, vals);
        if (rc == SC_ERR_STATE && ctx->csr_only) { ctx->err.clear(); return SC_OK; }
        SC_TRY(rc);
        if (seconds) *seconds += f->seconds;
        pcg_graph_drop(ctx);
    }
    *out = f;
    return SC_OK;
}

constexpr int PROJ_COEF = 64;      // coefficient slots in front of the partials in d_proj_coef

int proj_slot(sc_ctx* ctx, int k) {          // k = -1: only the coefficient block (no basis)
    const int K = ctx->proj_k;
    if ((int)ctx->proj_x.size() != K) { proj_free(ctx); ctx->proj_x.assign(K, nullptr); ctx->proj_ax.assign(K, nullptr); }
    if (!ctx->d_proj_ptr) SC_TRY(sc_alloc(ctx, &ctx->d_proj_ptr, (size_t)2 * K + 1));
    if (!ctx->d_proj_coef) SC_TRY(sc_alloc(ctx, &ctx->d_proj_coef, (size_t)PROJ_COEF + (size_t)(K + 1) * PCG_NB));
    if (k >= 0 && !ctx->proj_x[k]) {
        SC_TRY(sc_alloc(ctx, &ctx->proj_x[k], (size_t)ctx->n_eq));
        SC_TRY(sc_alloc(ctx, &ctx->proj_ax[k], (size_t)ctx->n_eq));
        std::vector<double*> tab(2 * K);
        for (int i = 0; i < K; ++i) { tab[i] = ctx->proj_x[i]; tab[K + i] = ctx->proj_ax[i]; }
        SC_CUDA(ctx, cudaMemcpyAsync(ctx->d_proj_ptr, tab.data(), sizeof(double*) * 2 * K, cudaMemcpyHostToDevice, ctx->stream));
        SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // `tab` goes out of scope
    }
    return SC_OK;
}

// coef[0..nv+self) = V[v].y (+ y.y), all-reduced
int proj_dots(sc_ctx* ctx, const double* const* V, int nv, int self, const double* y) {
    double* coef = ctx->d_proj_coef;
    k_multidot<<<PCG_NB, 256, 0, ctx->stream>>>(V, nv, self, y, ctx->n_eq, coef + PROJ_COEF, PCG_NB);
    SC_CHECK_LAUNCH(ctx);
    k_reduce_multi<<<nv + self, 256, 0, ctx->stream>>>(coef + PROJ_COEF, PCG_NB, coef);
    SC_CHECK_LAUNCH(ctx);
    if (ctx->world > 1) SC_TRY(dist_allreduce_sum(ctx, coef, nv + self, ctx->stream));
    return SC_OK;
}

// One implicit solve  A du = rhs  of a time loop whose matrix stays fixed (stream-ordered PCG driver).
//  1. The right-hand side is projected onto the A-orthonormal basis of the previous solutions (Fischer, "Projection
//     techniques for iterative solution of Ax = b with successive right-hand sides", 1998); PCG only solves for the
//     remainder, against the same absolute tolerance rtol ||rhs||.  Smooth histories lose 30-45 % of their iterations.
//  2. The accepted solution is checked against the TRUE residual rhs - A du (one more product): the recursive residual of
//     CG -- and the stored products A x_i of the basis, which are built from it -- drift from the true one by round-off,
//     and Newmark feeds every increment back through a1 = 4/dt^2 (the histories of the reference's 2000-step goldens
//     moved by 3e-8 without this step).  Up to two short correction solves follow while the true residual is above the
//     target and still decreasing.
//  3. The solution joins the basis (modified Gram-Schmidt in the A inner product, A du = rhs - r_true); a full basis
//     restarts from the latest solution.
int proj_solve(sc_ctx* ctx, const double* vals, const double* dinv, const sc_fsai* F, const double* rhs, double* du, double* r,
               double* p, double* q, double* zv, double* tv, double* r0, double* d2, double rtol, int maxit, int* iters,
               double* relres) {
    const int64_t n = ctx->n_eq;
    cudaStream_t st = ctx->stream;
    const int K = ctx->proj_k;
    if (ctx->proj_for != vals) { ctx->proj_n = 0; ctx->proj_for = vals; }
    int nv = K > 0 ? ctx->proj_n : 0;
    SC_TRY(proj_slot(ctx, K > 0 ? (nv == K ? 0 : nv) : -1));
    double* coef = ctx->d_proj_coef;
    const double* const* PX = ctx->d_proj_ptr;
    const double* const* PAX = ctx->d_proj_ptr + K;
    // c_i = x_i . rhs and ||rhs||^2 (the reference norm of every stopping test below)
    SC_TRY(proj_dots(ctx, PX, nv, 1, rhs));
    k_copy1<<<1, 1, 0, st>>>(coef + nv, ctx->d_scal + 4);
    SC_CHECK_LAUNCH(ctx);
    const double* b = rhs;
    if (nv > 0) {
        k_multiaxpy<<<nblk(n, 256), 256, 0, st>>>(r0, rhs, -1.0, coef, PAX, nv, n);
        SC_CHECK_LAUNCH(ctx);
        b = r0;
    }
    SC_TRY(pcg(ctx, vals, dinv, b, du, r, p, q, rtol, maxit, iters, relres, -1.0, F, zv, tv, true));
    const double bb = ctx->h_pinned[4];
    if (nv > 0) {
        k_multiaxpy<<<nblk(n, 256), 256, 0, st>>>(du, du, 1.0, coef, PX, nv, n);
        SC_CHECK_LAUNCH(ctx);
    }
    // true residual r0 = rhs - A du, correction solves
    double last = -1.0;
    for (int round = 0; bb > 0.0; ++round) {
        if (ctx->world > 1) SC_TRY(dist_halo(ctx, du, st));
        SC_TRY(la_spmv(ctx, vals, du, r0));
        SC_TRY(lincomb(ctx, r0, 1.0, rhs, -1.0, r0));
        SC_TRY(proj_dots(ctx, PX, 0, 1, r0));
        SC_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 10, coef, sizeof(double), cudaMemcpyDeviceToHost, st));
        SC_CUDA(ctx, cudaStreamSynchronize(st));
        const double rr = ctx->h_pinned[10];
        *relres = std::sqrt(rr / bb);
        if (!(rr == rr)) return sc_fail(ctx, SC_ERR_NOCONV, "implicit solve produced NaN");
        if (rr <= rtol * rtol * bb || round == 2 || (last >= 0.0 && rr > 0.0625 * last)) break;
        last = rr;
        int it2 = 0;
        double rel2 = 0.0;
        SC_TRY(pcg(ctx, vals, dinv, r0, d2, r, p, q, rtol, maxit, &it2, &rel2, -1.0, F, zv, tv, true));   // d_scal[4] still holds ||rhs||^2
        *iters += it2;
        SC_TRY(lincomb(ctx, du, 1.0, du, 1.0, d2));
    }
    if (K == 0) return SC_OK;
    // basis update: w = du - sum (A x_i . du) x_i,  A w likewise from A du = rhs - r0
    SC_TRY(lincomb(ctx, r0, 1.0, rhs, -1.0, r0));
    if (nv == K) nv = 0;
    double* xs = ctx->proj_x[nv];
    double* axs = ctx->proj_ax[nv];
    if (nv > 0) {
        SC_TRY(proj_dots(ctx, PAX, nv, 0, du));
        k_multiaxpy<<<nblk(n, 256), 256, 0, st>>>(xs, du, -1.0, coef, PX, nv, n);
        SC_CHECK_LAUNCH(ctx);
        k_multiaxpy<<<nblk(n, 256), 256, 0, st>>>(axs, r0, -1.0, coef, PAX, nv, n);
        SC_CHECK_LAUNCH(ctx);
    } else {
        SC_CUDA(ctx, cudaMemcpyAsync(xs, du, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
        SC_CUDA(ctx, cudaMemcpyAsync(axs, r0, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    }
    SC_TRY(la_dot(ctx, xs, axs, coef + PROJ_COEF - 1));
    k_scale_pair<<<nblk(n, 256), 256, 0, st>>>(xs, axs, coef + PROJ_COEF - 1, n);
    SC_CHECK_LAUNCH(ctx);
    ctx->proj_n = nv + 1;
    return SC_OK;
}

int apply_load(sc_ctx* ctx, int64_t t, double scale, const double* mult, double* y, int sel = 0) {
    if (t < 0 || t >= ctx->load_steps) return SC_OK;   // outside the schedule: zero force
    const int64_t s = ctx->h_load_ptr[t], e = ctx->h_load_ptr[t + 1];
    if (e == s) return SC_OK;
    k_load_add<<<nblk(e - s, 128), 128, 0, ctx->stream>>>(ctx->d_load_dof + s, ctx->d_load_val + s, e - s, scale, mult, y, sel,
                                                          ctx->ov_row_lo, ctx->ov_row_hi);
    SC_CHECK_LAUNCH(ctx);
    return SC_OK;
}

__global__ void k_gather_sel(const double* __restrict__ src, const int32_t* __restrict__ sel, int64_t n, double* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[sel[i]];
}
inline int64_t row_len(const sc_ctx* ctx) { return ctx->n_sel >= 0 ? ctx->n_sel : ctx->n_eq; }

// synchronous-in-stream copy of one output row (Bathe / static loops)
int store_row(sc_ctx* ctx, double* host, int64_t row, const double* dev) {
    if (!host) return SC_OK;
    const int64_t len = row_len(ctx);
    if (len == 0) return SC_OK;
    if (ctx->n_sel >= 0) {
        if (!ctx->d_selbuf[0]) SC_TRY(sc_alloc(ctx, &ctx->d_selbuf[0], (size_t)len));
        k_gather_sel<<<nblk(len, 256), 256, 0, ctx->stream>>>(dev, ctx->d_sel, len, ctx->d_selbuf[0]);
        SC_CHECK_LAUNCH(ctx);
        dev = ctx->d_selbuf[0];
    }
    SC_CUDA(ctx, cudaMemcpyAsync(host + row * len, dev, sizeof(double) * len, cudaMemcpyDeviceToHost, ctx->stream));
    return SC_OK;
}

// Output rows leave the device on the copy stream while the time loop keeps running: the compute stream only waits when
// it is about to overwrite a buffer whose copy is still in flight.  `snap[k]` says whether source k changes before the
// next output step (then it is snapshotted device-to-device first, 0.1 ms for 50 M dofs).  With an output selection
// (sc_set_output_dofs) the gather into the compact row buffer is the snapshot.
int store_rows_async(sc_ctx* ctx, double* hu, double* hv, double* ha, int64_t row, const double* su, const double* sv, const double* sa,
                     bool snap_u, bool snap_v, bool snap_a) {
    if (!hu && !hv && !ha) return SC_OK;
    const int64_t len = row_len(ctx);
    if (len == 0) return SC_OK;
    const size_t bytes = sizeof(double) * len;
    if (!ctx->ev_rows_ready) {
        SC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_rows_ready, cudaEventDisableTiming));
        SC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_rows_done, cudaEventDisableTiming));
    }
    if (ctx->rows_pending) SC_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_rows_done, 0));
    double* host[3] = {hu, hv, ha};
    const double* src[3] = {su, sv, sa};
    const bool snap[3] = {snap_u, snap_v, snap_a};
    for (int k = 0; k < 3; ++k) {
        if (!host[k]) continue;
        if (ctx->n_sel >= 0) {
            if (!ctx->d_selbuf[k]) SC_TRY(sc_alloc(ctx, &ctx->d_selbuf[k], (size_t)len));
            k_gather_sel<<<nblk(len, 256), 256, 0, ctx->stream>>>(src[k], ctx->d_sel, len, ctx->d_selbuf[k]);
            SC_CHECK_LAUNCH(ctx);
            src[k] = ctx->d_selbuf[k];
        } else if (snap[k]) {
            if (!ctx->d_snap[k]) SC_TRY(sc_alloc(ctx, &ctx->d_snap[k], (size_t)ctx->n_eq));
            SC_CUDA(ctx, cudaMemcpyAsync(ctx->d_snap[k], src[k], bytes, cudaMemcpyDeviceToDevice, ctx->stream));
            src[k] = ctx->d_snap[k];
        }
    }
    SC_CUDA(ctx, cudaEventRecord(ctx->ev_rows_ready, ctx->stream));
    SC_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_rows_ready, 0));
    for (int k = 0; k < 3; ++k)
        if (host[k]) SC_CUDA(ctx, cudaMemcpyAsync(host[k] + row * len, src[k], bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
    SC_CUDA(ctx, cudaEventRecord(ctx->ev_rows_done, ctx->copy_stream));
    ctx->rows_pending = true;
    return SC_OK;
}
int finish_rows(sc_ctx* ctx) {
    if (ctx->rows_pending) SC_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    ctx->rows_pending = false;
    return SC_OK;
}

// output steps: multiples of the interval, plus the one extra step of sc_set_final_output_step
inline bool is_out_step(const sc_ctx* ctx, int64_t t, int64_t oi) { return t % oi == 0 || t == ctx->extra_out_step; }

int ensure_state(sc_ctx* ctx) {
    const int64_t n = ctx->n_eq;
    for (double** v : {&ctx->d_u, &ctx->d_v, &ctx->d_a}) {
        if (!*v) {
            SC_TRY(sc_alloc(ctx, v, (size_t)n));
            SC_CUDA(ctx, cudaMemsetAsync(*v, 0, sizeof(double) * n, ctx->stream));
        }
    }
    return SC_OK;
}

}  // namespace

int tl_newmark(sc_ctx* ctx, double dt, int64_t t0, int64_t n_steps, int64_t oi, double beta, double gamma, double rtol,
               int maxit, int64_t n_out, double* u_out, double* v_out, double* a_out, sc_stats* stats) {
    asm_release_scratch(ctx);                 // the assembly's element records are not needed while stepping
    auto wall0 = std::chrono::steady_clock::now();
    const int64_t n = ctx->n_eq;
    cudaStream_t st = ctx->stream;
    const int64_t launches0 = ctx->launches, stag0 = ctx->pcg_stagnations;
    SC_TRY(la_scratch(ctx));
    SC_TRY(ensure_state(ctx));
    double *x1, *x2, *qv, *rhs, *du, *r, *p, *q, *dinv, *dinvM;
    SC_TRY(sc_work(ctx, 0, &x1)); SC_TRY(sc_work(ctx, 1, &x2)); SC_TRY(sc_work(ctx, 2, &qv)); SC_TRY(sc_work(ctx, 3, &rhs));
    SC_TRY(sc_work(ctx, 4, &du)); SC_TRY(sc_work(ctx, 5, &r)); SC_TRY(sc_work(ctx, 6, &p)); SC_TRY(sc_work(ctx, 7, &q));
    SC_TRY(sc_work(ctx, 8, &dinv)); SC_TRY(sc_work(ctx, 9, &dinvM));
    const double a1 = 1.0 / (beta * dt * dt), a4 = gamma / (beta * dt);
    const double c0 = ctx->c0, c1 = ctx->c1;
    const bool cabs = ctx->cabs_n > 0;

    // effective matrix Khat = (1 + a4 c1) K + (a1 + a4 c0) M + a4 C_abs (kept across stages with the same dt)
    if (!ctx->d_Khat || ctx->khat_a1 != a1 || ctx->khat_a4 != a4) {
        SC_TRY(sc_alloc(ctx, &ctx->d_Khat, (size_t)ctx->nnz));
        SC_TRY(la_axpby_vals(ctx, ctx->d_Khat, 1.0 + a4 * c1, ctx->d_K, a1 + a4 * c0, ctx->d_M, ctx->nnz));
        if (ctx->d_C) SC_TRY(la_axpby_vals(ctx, ctx->d_Khat, 1.0, ctx->d_Khat, a4, ctx->d_C, ctx->nnz));
        SC_TRY(la_cabs_add_values(ctx, ctx->d_Khat, a4));
        ctx->khat_a1 = a1; ctx->khat_a4 = a4;
        precond_drop(ctx);
    }
    SC_TRY(la_extract_diag(ctx, ctx->d_Khat, dinv, true));
    // preconditioner (FSAI, fsai.cu) and initial guesses (projection onto previous solutions) of the per-step solve
    const sc_fsai* F = nullptr;
    double fsai_seconds = 0.0;
    SC_TRY(precond_for(ctx, 0, ctx->d_Khat, &F, &fsai_seconds));
    double *zv = nullptr, *tv = nullptr, *r0 = nullptr;
    if (F) { SC_TRY(sc_work(ctx, 14, &zv)); SC_TRY(sc_work(ctx, 15, &tv)); }
    // large systems go through proj_solve (projection, true-residual check); small ones through the cooperative kernel
    const bool use_proj = !pcg_small_usable(ctx);
    double* d2 = nullptr;
    if (use_proj) { SC_TRY(sc_work(ctx, 16, &r0)); SC_TRY(sc_work(ctx, 17, &d2)); }

    int64_t pcg_total = 0;
    int iters = 0;
    double relres = 0.0;
    int64_t row = 0;

    const bool resume = ctx->nm_resume_valid && ctx->nm_resume_t == t0;
    ctx->nm_resume_valid = false;
    if (!resume) {
        // initial acceleration a = M^-1 (F(t0) - C v - K u) = M^-1 (F - M (c0 v) - K (c1 v + u) - C_abs v)
        SC_TRY(la_extract_diag(ctx, ctx->d_M, dinvM, true));
        if (ctx->d_C) {
            SC_TRY(la_spmv2(ctx, ctx->d_C, ctx->d_v, ctx->d_K, ctx->d_u, rhs));
        } else {
            SC_TRY(la_axpby_vals(ctx, x1, c0, ctx->d_v, 0.0, nullptr, n));
            SC_TRY(la_axpby_vals(ctx, x2, c1, ctx->d_v, 1.0, ctx->d_u, n));
            if (ctx->world > 1) { SC_TRY(dist_halo(ctx, x1, st)); SC_TRY(dist_halo(ctx, x2, st)); }
            SC_TRY(la_spmv2(ctx, ctx->d_M, x1, ctx->d_K, x2, rhs));
        }
        SC_TRY(la_cabs_spmv_add(ctx, ctx->d_v, rhs, 1.0));
        k_neg_add<<<nblk(n, 256), 256, 0, st>>>(rhs, n);
        SC_CHECK_LAUNCH(ctx);
        SC_TRY(apply_load(ctx, t0, 1.0, nullptr, rhs));
        SC_TRY(pcg(ctx, ctx->d_M, dinvM, rhs, ctx->d_a, r, p, q, rtol, maxit, &iters, &relres));
        pcg_total += iters;
        if (ctx->world > 1) SC_TRY(dist_halo(ctx, ctx->d_a, st));
    }

    if (is_out_step(ctx, t0, oi) && row < n_out) {
        SC_TRY(store_rows_async(ctx, u_out, v_out, a_out, row, ctx->d_u, ctx->d_v, ctx->d_a, true, true, true));
        ++row;
    }
    sc_gpu_timer timer(st);
    timer.start();
    const double pv = 1.0 / (beta * dt), pa = 1.0 / (2.0 * beta);
    const double qvv = gamma / beta, qa = dt * (gamma / (2.0 * beta) - 1.0);
    for (int64_t t = t0 + 1; t <= t0 + n_steps; ++t) {
        k_nm_inputs<<<nblk(n, 256), 256, 0, st>>>(ctx->d_v, ctx->d_a, x1, x2, (cabs || ctx->d_C) ? qv : nullptr, pv, pa, qvv, qa, c0, c1, n);
        SC_CHECK_LAUNCH(ctx);
        if (ctx->d_C) SC_TRY(la_spmv2(ctx, ctx->d_M, x1, ctx->d_C, qv, rhs));      // c0 = c1 = 0: x1 = p, qv = q
        else SC_TRY(la_spmv2(ctx, ctx->d_M, x1, ctx->d_K, x2, rhs));
        if (cabs) SC_TRY(la_cabs_spmv_add(ctx, qv, rhs, 1.0));
        SC_TRY(apply_load(ctx, t, 1.0, nullptr, rhs));
        SC_TRY(apply_load(ctx, t - 1, -1.0, nullptr, rhs));
        if (use_proj) SC_TRY(proj_solve(ctx, ctx->d_Khat, dinv, F, rhs, du, r, p, q, zv, tv, r0, d2, rtol, maxit, &iters, &relres));
        else SC_TRY(pcg(ctx, ctx->d_Khat, dinv, rhs, du, r, p, q, rtol, maxit, &iters, &relres, -1.0, F, zv, tv));
        pcg_total += iters;
        if (ctx->world > 1) SC_TRY(dist_halo(ctx, du, st));
        k_nm_update<<<nblk(n, 256), 256, 0, st>>>(du, ctx->d_u, ctx->d_v, ctx->d_a, a4, gamma / beta, dt * (1.0 - gamma / (2.0 * beta)), a1,
                                                   1.0 / (beta * dt), 1.0 / (2.0 * beta), n);
        SC_CHECK_LAUNCH(ctx);
        if (is_out_step(ctx, t, oi) && row < n_out) {
            SC_TRY(store_rows_async(ctx, u_out, v_out, a_out, row, ctx->d_u, ctx->d_v, ctx->d_a, true, true, true));
            ++row;
        }
    }
    timer.stop();
    SC_CUDA(ctx, cudaStreamSynchronize(st));
    SC_TRY(finish_rows(ctx));
    if (ctx->world > 1) SC_TRY(dist_peer_check(ctx));
    const float ms = timer.ms();
    if (stats) {
        stats->seconds_device = ms * 1e-3;
        stats->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
        stats->steps = n_steps;
        stats->pcg_iterations = pcg_total;
        stats->kernel_launches = ctx->launches - launches0;
        stats->last_residual = relres;
        stats->seconds_halo = 0.0;
        stats->reserved[2] = (double)(ctx->pcg_stagnations - stag0);
        stats->reserved[3] = fsai_seconds;       // FSAI set-up done inside this call (0 when the factor was reused)
    }
    ctx->nm_resume_valid = true;
    ctx->nm_resume_t = t0 + n_steps;
    return SC_OK;
}

int tl_central_difference(sc_ctx* ctx, double dt, int64_t t0, int64_t n_steps, int64_t oi, int64_t n_out, double* u_out,
                          double* v_out, double* a_out, sc_stats* stats) {
    asm_release_scratch(ctx);                 // the assembly's element records are not needed while stepping
    auto wall0 = std::chrono::steady_clock::now();
    const int64_t n = ctx->n_eq;
    cudaStream_t st = ctx->stream;
    const int64_t launches0 = ctx->launches;
    SC_TRY(la_scratch(ctx));
    SC_TRY(ensure_state(ctx));
    double *ub, *uc, *inv_d, *alpha, *cl, *tmp, *vv, *aa, *wa = nullptr, *wb = nullptr;
    SC_TRY(sc_work(ctx, 0, &ub)); SC_TRY(sc_work(ctx, 1, &uc)); SC_TRY(sc_work(ctx, 2, &inv_d)); SC_TRY(sc_work(ctx, 3, &alpha));
    SC_TRY(sc_work(ctx, 4, &cl)); SC_TRY(sc_work(ctx, 5, &tmp));
    const double a0 = 1.0 / (dt * dt), a1 = 1.0 / (2.0 * dt);
    // stiffness-proportional Rayleigh damping c1 K acts on the lagged velocity (u(t) - u(t-dt))/dt through the SpMV:
    // the kernels gather w = (1+g) u(t) - g u(t-dt), g = c1/dt, instead of u (one more vector written per step)
    const double g = ctx->c1 / dt;
    const bool lagged = g != 0.0;
    if (lagged) { SC_TRY(sc_work(ctx, 8, &wa)); SC_TRY(sc_work(ctx, 9, &wb)); }

    // diagonal part of the damping: c_d = c0 m + rowsum(C_abs)   (work[1] is the vector of ones here)
    // inv_d, alpha, c_d only depend on (m, C_abs, c0, dt): a stage that follows another one with the same dt reuses them
    // (`cd_coef_dt` is reset by every call that changes a matrix or borrows work[2..4]).
    if (ctx->cd_coef_dt != dt) {
        SC_TRY(la_axpby_vals(ctx, cl, ctx->c0, ctx->d_Ml, 0.0, nullptr, n));
        if (ctx->d_C) {                              // caller-supplied damping matrix: all of it lumped by row sums
            SC_TRY(la_fill(ctx, uc, 1.0, n));
            SC_TRY(la_spmv(ctx, ctx->d_C, uc, cl));
        }
        if (ctx->cabs_rows > 0) {
            SC_TRY(la_fill(ctx, uc, 1.0, n));
            SC_TRY(la_cabs_spmv_add(ctx, uc, cl, 1.0));
        }
        k_cd_coeffs<<<nblk(n, 256), 256, 0, st>>>(ctx->d_Ml, cl, a0, a1, inv_d, alpha, n);
        SC_CHECK_LAUNCH(ctx);
        ctx->cd_coef_dt = dt;
    }

    // rotating buffers: cur = u(t), prev = u(t-dt) (overwritten by u(t+dt) each step); w_cur = w(t), w_nxt receives w(t+dt)
    double* cur = ctx->d_u;
    double* prev = ub;
    double* w_cur = lagged ? wa : cur;
    double* w_nxt = lagged ? wb : nullptr;
    const bool resume = ctx->cd_resume_valid && ctx->cd_resume_t == t0 && ctx->cd_resume_dt == dt;
    if (!resume) {
        // start-up value u(t0 - dt) = u - dt v + dt^2/2 a(t0),  a(t0) = (F - K (u + c1 v) - c_d v)/m
        const double* x0 = cur;
        if (lagged) { SC_TRY(lincomb(ctx, uc, 1.0, cur, ctx->c1, ctx->d_v)); x0 = uc; }
        if (ctx->world > 1) { SC_TRY(dist_halo(ctx, cur, st)); if (lagged) SC_TRY(dist_halo(ctx, uc, st)); }
        SC_TRY(la_spmv(ctx, ctx->d_K, x0, tmp));
        k_neg_add<<<nblk(n, 256), 256, 0, st>>>(tmp, n);
        SC_CHECK_LAUNCH(ctx);
        SC_TRY(apply_load(ctx, t0, 1.0, nullptr, tmp));
        k_cd_start<<<nblk(n, 256), 256, 0, st>>>(cur, ctx->d_v, tmp, ctx->d_Ml, cl, dt, prev, n);
        SC_CHECK_LAUNCH(ctx);
        if (ctx->world > 1) SC_TRY(dist_halo(ctx, prev, st));
        if (lagged) SC_TRY(lincomb(ctx, w_cur, 1.0 + g, cur, -g, prev));
    }
    ctx->cd_resume_valid = false;

    const bool want_out = (u_out || v_out || a_out) && n_out > 0;
    if (want_out) { SC_TRY(sc_work(ctx, 6, &vv)); SC_TRY(sc_work(ctx, 7, &aa)); }
    // Several GPUs: the tiles that hold the interface rows step first, their values leave on the communication stream, and
    // the interior tiles (no row sent, no ghost column read) step while the exchange is in flight.
    bool overlap = false;
    if (ctx->world > 1 && la_node_usable(ctx)) {
        SC_TRY(la_node_overlap_plan(ctx));
        overlap = ctx->ov_ok;
        if (overlap && !ctx->comm_stream) {
            SC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
            SC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_bnd, cudaEventDisableTiming));
            SC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming));
        }
    }
    // device time of the exchange itself, sampled on the first steps of the stage (pack, NCCL send/recv, unpack)
    constexpr int HALO_SAMPLES = 16;
    cudaEvent_t hs[HALO_SAMPLES][2];
    int n_hs = 0;
    int64_t row = 0;
    sc_gpu_timer timer(st);
    timer.start();
    int64_t steps_done = 0;
    const int64_t t_end = t0 + n_steps;
    for (int64_t t = t0; t <= t_end; ++t) {
        const bool out_now = want_out && is_out_step(ctx, t, oi) && row < n_out;
        if (t == t_end && !out_now) break;       // the extra half step is only needed for v/a of an output row
        if (out_now) {
            // keep u(t-dt): the step overwrites it
            SC_CUDA(ctx, cudaMemcpyAsync(uc, prev, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
        }
        // prev <- u(t+dt), w_nxt <- w(t+dt)   (without stiffness-proportional damping the gathered vector is u(t) itself)
        if (!lagged) w_cur = cur;
        // ghost values: only the gathered vector needs them (u itself is read on owned rows only)
        double* hx = lagged ? w_nxt : prev;
        const bool sample = ctx->world > 1 && n_hs < HALO_SAMPLES;
        if (sample) { cudaEventCreate(&hs[n_hs][0]); cudaEventCreate(&hs[n_hs][1]); }
        if (overlap) {
            cudaStream_t cs = ctx->comm_stream;
            SC_TRY(la_node_cd_step(ctx, ctx->d_K, w_cur, cur, prev, inv_d, alpha, g, w_nxt, 1));
            SC_TRY(apply_load(ctx, t, 1.0, inv_d, prev, 1));
            if (lagged) SC_TRY(apply_load(ctx, t, 1.0 + g, inv_d, w_nxt, 1));
            SC_CUDA(ctx, cudaEventRecord(ctx->ev_bnd, st));
            SC_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_bnd, 0));
            if (sample) cudaEventRecord(hs[n_hs][0], cs);
            SC_TRY(dist_halo(ctx, hx, cs));
            if (sample) cudaEventRecord(hs[n_hs][1], cs);
            SC_CUDA(ctx, cudaEventRecord(ctx->ev_halo, cs));
            SC_TRY(la_node_cd_step(ctx, ctx->d_K, w_cur, cur, prev, inv_d, alpha, g, w_nxt, 2));
            SC_TRY(apply_load(ctx, t, 1.0, inv_d, prev, 2));
            if (lagged) SC_TRY(apply_load(ctx, t, 1.0 + g, inv_d, w_nxt, 2));
            SC_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_halo, 0));
        } else {
            SC_TRY(la_cd_step(ctx, ctx->d_K, w_cur, cur, prev, inv_d, alpha, g, w_nxt));
            SC_TRY(apply_load(ctx, t, 1.0, inv_d, prev));
            if (lagged) SC_TRY(apply_load(ctx, t, 1.0 + g, inv_d, w_nxt));
            if (ctx->world > 1) {
                if (sample) cudaEventRecord(hs[n_hs][0], st);
                SC_TRY(dist_halo(ctx, hx, st));
                if (sample) cudaEventRecord(hs[n_hs][1], st);
            }
        }
        if (sample) ++n_hs;
        if (out_now) {
            if (ctx->rows_pending) SC_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_rows_done, 0));   // vv / aa are still being copied
            k_cd_va<<<nblk(n, 256), 256, 0, st>>>(prev, cur, uc, a0, a1, vv, aa, n);
            SC_CHECK_LAUNCH(ctx);
            SC_TRY(store_rows_async(ctx, u_out, v_out, a_out, row, cur, vv, aa, true, false, false));
            ++row;
        }
        if (t == t_end) {
            // state stays at t_end: put u(t_end - dt) back so that a following stage continues seamlessly (w_cur still is w(t_end))
            SC_CUDA(ctx, cudaMemcpyAsync(prev, uc, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
            SC_CUDA(ctx, cudaMemcpyAsync(ctx->d_v, vv, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
            SC_CUDA(ctx, cudaMemcpyAsync(ctx->d_a, aa, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
            break;
        }
        double* t_ = cur; cur = prev; prev = t_;
        if (lagged) { t_ = w_cur; w_cur = w_nxt; w_nxt = t_; }
        ++steps_done;
    }
    timer.stop();
    // normalise the buffers: d_u = u(t_end), work[0] = u(t_end - dt), work[8] = w(t_end)
    if (cur != ctx->d_u) {
        SC_CUDA(ctx, cudaMemcpyAsync(uc, prev, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));     // prev lives in d_u
        SC_CUDA(ctx, cudaMemcpyAsync(ctx->d_u, cur, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
        SC_CUDA(ctx, cudaMemcpyAsync(ub, uc, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    }
    if (lagged && w_cur != wa) SC_CUDA(ctx, cudaMemcpyAsync(wa, w_cur, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    ctx->cd_resume_valid = true;
    ctx->cd_resume_t = t_end;
    ctx->cd_resume_dt = dt;
    // the scheme is only conditionally stable: report a blown-up state instead of handing back NaN/Inf histories
    SC_CUDA(ctx, cudaMemsetAsync(ctx->d_scal + 9, 0, sizeof(double), st));
    k_flag_nonfinite<<<592, 256, 0, st>>>(ctx->d_u, n, ctx->d_scal + 9);
    SC_CHECK_LAUNCH(ctx);
    SC_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 9, ctx->d_scal + 9, sizeof(double), cudaMemcpyDeviceToHost, st));
    SC_TRY(finish_rows(ctx));
    SC_CUDA(ctx, cudaStreamSynchronize(st));
    if (ctx->world > 1) SC_TRY(dist_peer_check(ctx));
    const float ms = timer.ms();
    double halo_ms = 0.0;
    for (int k = 0; k < n_hs; ++k) {
        float v = 0.f;
        if (cudaEventElapsedTime(&v, hs[k][0], hs[k][1]) == cudaSuccess) halo_ms += v;
        cudaEventDestroy(hs[k][0]); cudaEventDestroy(hs[k][1]);
    }
    const double halo_per_step = n_hs > 0 ? halo_ms * 1e-3 / n_hs : 0.0;
    double diverged = ctx->h_pinned[9];
    if (ctx->world > 1) {                                // every rank must take the same exit
        SC_TRY(dist_allreduce_sum(ctx, ctx->d_scal + 9, 1, st));
        SC_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 9, ctx->d_scal + 9, sizeof(double), cudaMemcpyDeviceToHost, st));
        SC_CUDA(ctx, cudaStreamSynchronize(st));
        diverged = ctx->h_pinned[9];
    }
    if (diverged != 0.0) {
        ctx->cd_resume_valid = false;
        return sc_fail(ctx, SC_ERR_NOCONV, "central difference diverged (non-finite displacements after step %lld): the time step %g "
                       "exceeds the stability limit of the mesh", (long long)t_end, dt);
    }
    if (stats) {
        stats->seconds_device = ms * 1e-3;
        stats->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
        stats->steps = steps_done;
        // algorithmic bytes of one fused step launch for the index format in use (reported by bench.py)
        stats->reserved[0] = la_node_usable(ctx) ? (double)la_node_step_bytes(ctx, lagged)
                                                 : (double)(ctx->nnz * 12 + ctx->n_eq * (lagged ? 64 : 48));
        stats->reserved[1] = la_node_usable(ctx) ? 2.0 : (la_tma_usable(ctx) ? 1.0 : 0.0);
        stats->pcg_iterations = 0;
        stats->kernel_launches = ctx->launches - launches0;
        stats->last_residual = 0.0;
        // exchange time extrapolated from the sampled steps; with the overlap it runs beside the interior tiles and is
        // not part of seconds_device any more (reserved[3] = 1 says so)
        stats->seconds_halo = halo_per_step * (double)steps_done;
        stats->reserved[3] = overlap ? 1.0 : 0.0;
    }
    return SC_OK;
}

// Bathe composite scheme (trapezoidal rule over dt/2, then 3-point backward Euler), two PCG solves per step.
//   sub-step 1:  (K + 4/dt C + 16/dt^2 M) u1 = F(t+dt/2) + M (16 u/dt^2 + 8 v/dt + a) + C (4 u/dt + v)
//                v1 = 4 (u1 - u)/dt - v
//   sub-step 2:  (K + 3/dt C + 9/dt^2 M) u2 = F(t+dt) + M (12 u1/dt^2 - 3 u/dt^2 + 4 v1/dt - v/dt) + C (4 u1/dt - u/dt)
//                v2 = (u - 4 u1 + 3 u2)/dt,  a2 = (v - 4 v1 + 3 v2)/dt
// The half-step force is the mean of the two step forces (the load schedule is defined on whole steps).
int tl_bathe(sc_ctx* ctx, double dt, int64_t t0, int64_t n_steps, int64_t oi, double rtol, int maxit, int64_t n_out, double* u_out,
             double* v_out, double* a_out, sc_stats* stats) {
    asm_release_scratch(ctx);                 // the assembly's element records are not needed while stepping
    auto wall0 = std::chrono::steady_clock::now();
    const int64_t n = ctx->n_eq;
    cudaStream_t st = ctx->stream;
    const int64_t launches0 = ctx->launches;
    SC_TRY(la_scratch(ctx));
    SC_TRY(ensure_state(ctx));
    double *x1, *x2, *xm, *xc, *rhs, *u1, *v1, *r, *p, *q, *dinv1, *dinv2, *dinvM, *u2;
    SC_TRY(sc_work(ctx, 0, &x1)); SC_TRY(sc_work(ctx, 1, &x2)); SC_TRY(sc_work(ctx, 2, &xm)); SC_TRY(sc_work(ctx, 3, &rhs));
    SC_TRY(sc_work(ctx, 4, &u1)); SC_TRY(sc_work(ctx, 5, &r)); SC_TRY(sc_work(ctx, 6, &p)); SC_TRY(sc_work(ctx, 7, &q));
    SC_TRY(sc_work(ctx, 8, &dinv1)); SC_TRY(sc_work(ctx, 9, &dinvM)); SC_TRY(sc_work(ctx, 10, &xc)); SC_TRY(sc_work(ctx, 11, &v1));
    SC_TRY(sc_work(ctx, 12, &dinv2)); SC_TRY(sc_work(ctx, 13, &u2));
    const double c0 = ctx->c0, c1 = ctx->c1;
    SC_TRY(sc_alloc(ctx, &ctx->d_Khat, (size_t)ctx->nnz));
    SC_TRY(sc_alloc(ctx, &ctx->d_Khat2, (size_t)ctx->nnz));
    SC_TRY(la_axpby_vals(ctx, ctx->d_Khat, 1.0 + 4.0 / dt * c1, ctx->d_K, 16.0 / (dt * dt) + 4.0 / dt * c0, ctx->d_M, ctx->nnz));
    if (ctx->d_C) SC_TRY(la_axpby_vals(ctx, ctx->d_Khat, 1.0, ctx->d_Khat, 4.0 / dt, ctx->d_C, ctx->nnz));
    SC_TRY(la_cabs_add_values(ctx, ctx->d_Khat, 4.0 / dt));
    SC_TRY(la_axpby_vals(ctx, ctx->d_Khat2, 1.0 + 3.0 / dt * c1, ctx->d_K, 9.0 / (dt * dt) + 3.0 / dt * c0, ctx->d_M, ctx->nnz));
    if (ctx->d_C) SC_TRY(la_axpby_vals(ctx, ctx->d_Khat2, 1.0, ctx->d_Khat2, 3.0 / dt, ctx->d_C, ctx->nnz));
    SC_TRY(la_cabs_add_values(ctx, ctx->d_Khat2, 3.0 / dt));
    SC_TRY(la_extract_diag(ctx, ctx->d_Khat, dinv1, true));
    SC_TRY(la_extract_diag(ctx, ctx->d_Khat2, dinv2, true));
    precond_drop(ctx);                           // both effective matrices were just rebuilt
    ctx->khat_a1 = ctx->khat_a4 = -1.0;          // d_Khat no longer is Newmark's
    const sc_fsai *F1 = nullptr, *F2 = nullptr;
    SC_TRY(precond_for(ctx, 0, ctx->d_Khat, &F1, nullptr));
    SC_TRY(precond_for(ctx, 1, ctx->d_Khat2, &F2, nullptr));
    double *zv = nullptr, *tv = nullptr;
    if (F1 || F2) { SC_TRY(sc_work(ctx, 14, &zv)); SC_TRY(sc_work(ctx, 15, &tv)); }
    SC_TRY(la_extract_diag(ctx, ctx->d_M, dinvM, true));
    int64_t pcg_total = 0, row = 0;
    int iters = 0;
    double relres = 0.0;
    // initial acceleration a = M^-1 (F(t0) - C v - K u)
    SC_TRY(lincomb(ctx, xm, 0.0, ctx->d_v, 0.0, nullptr));
    SC_TRY(apply_M_C(ctx, xm, ctx->d_v, x1, x2, rhs));
    SC_TRY(la_spmv(ctx, ctx->d_K, ctx->d_u, x1));
    SC_TRY(lincomb(ctx, rhs, -1.0, rhs, -1.0, x1));
    SC_TRY(apply_load(ctx, t0, 1.0, nullptr, rhs));
    SC_TRY(pcg(ctx, ctx->d_M, dinvM, rhs, ctx->d_a, r, p, q, rtol, maxit, &iters, &relres));
    pcg_total += iters;
    if (is_out_step(ctx, t0, oi) && row < n_out) {
        SC_TRY(store_row(ctx, u_out, row, ctx->d_u)); SC_TRY(store_row(ctx, v_out, row, ctx->d_v)); SC_TRY(store_row(ctx, a_out, row, ctx->d_a));
        ++row;
    }
    sc_gpu_timer timer(st);
    timer.start();
    double *u = ctx->d_u, *v = ctx->d_v, *a = ctx->d_a;
    for (int64_t t = t0 + 1; t <= t0 + n_steps; ++t) {
        // ---- sub-step 1
        SC_TRY(lincomb(ctx, xm, 16.0 / (dt * dt), u, 8.0 / dt, v, 1.0, a));
        SC_TRY(lincomb(ctx, xc, 4.0 / dt, u, 1.0, v));
        SC_TRY(apply_M_C(ctx, xm, xc, x1, x2, rhs));
        SC_TRY(apply_load(ctx, t - 1, 0.5, nullptr, rhs));
        SC_TRY(apply_load(ctx, t, 0.5, nullptr, rhs));
        SC_TRY(pcg(ctx, ctx->d_Khat, dinv1, rhs, u1, r, p, q, rtol, maxit, &iters, &relres, -1.0, F1, zv, tv));
        pcg_total += iters;
        if (ctx->world > 1) SC_TRY(dist_halo(ctx, u1, st));
        SC_TRY(lincomb(ctx, v1, 4.0 / dt, u1, -4.0 / dt, u, -1.0, v));
        // ---- sub-step 2
        SC_TRY(lincomb(ctx, xm, 12.0 / (dt * dt), u1, -3.0 / (dt * dt), u, 4.0 / dt, v1, -1.0 / dt, v));
        SC_TRY(lincomb(ctx, xc, 4.0 / dt, u1, -1.0 / dt, u));
        SC_TRY(apply_M_C(ctx, xm, xc, x1, x2, rhs));
        SC_TRY(apply_load(ctx, t, 1.0, nullptr, rhs));
        SC_TRY(pcg(ctx, ctx->d_Khat2, dinv2, rhs, u2, r, p, q, rtol, maxit, &iters, &relres, -1.0, F2, zv, tv));
        pcg_total += iters;
        if (ctx->world > 1) SC_TRY(dist_halo(ctx, u2, st));
        // v2 (into x1), a2, then commit
        SC_TRY(lincomb(ctx, x1, 1.0 / dt, u, -4.0 / dt, u1, 3.0 / dt, u2));
        SC_TRY(lincomb(ctx, a, 1.0 / dt, v, -4.0 / dt, v1, 3.0 / dt, x1));
        SC_CUDA(ctx, cudaMemcpyAsync(v, x1, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
        SC_CUDA(ctx, cudaMemcpyAsync(u, u2, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
        if (is_out_step(ctx, t, oi) && row < n_out) {
            SC_TRY(store_row(ctx, u_out, row, u)); SC_TRY(store_row(ctx, v_out, row, v)); SC_TRY(store_row(ctx, a_out, row, a));
            ++row;
        }
    }
    timer.stop();
    SC_CUDA(ctx, cudaStreamSynchronize(st));
    const float ms = timer.ms();
    if (stats) {
        stats->seconds_device = ms * 1e-3;
        stats->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
        stats->steps = n_steps;
        stats->pcg_iterations = pcg_total;
        stats->kernel_launches = ctx->launches - launches0;
        stats->last_residual = relres;
    }
    return SC_OK;
}

// Static solver: K u(t) = F(t) for every step, solved incrementally (du = K^-1 (F(t) - K u)) with Jacobi-PCG.
int tl_static(sc_ctx* ctx, int64_t t0, int64_t n_steps, int64_t oi, double rtol, int maxit, int64_t n_out, double* u_out, sc_stats* stats) {
    asm_release_scratch(ctx);                 // the assembly's element records are not needed while stepping
    auto wall0 = std::chrono::steady_clock::now();
    const int64_t n = ctx->n_eq;
    cudaStream_t st = ctx->stream;
    const int64_t launches0 = ctx->launches;
    SC_TRY(la_scratch(ctx));
    SC_TRY(ensure_state(ctx));
    double *rhs, *du, *r, *p, *q, *dinv, *f;
    SC_TRY(sc_work(ctx, 3, &rhs)); SC_TRY(sc_work(ctx, 4, &du)); SC_TRY(sc_work(ctx, 5, &r)); SC_TRY(sc_work(ctx, 6, &p));
    SC_TRY(sc_work(ctx, 7, &q)); SC_TRY(sc_work(ctx, 8, &dinv)); SC_TRY(sc_work(ctx, 0, &f));
    SC_TRY(la_extract_diag(ctx, ctx->d_K, dinv, true));
    const sc_fsai* F = nullptr;
    SC_TRY(precond_for(ctx, 0, ctx->d_K, &F, nullptr));
    double *zv = nullptr, *tv = nullptr;
    if (F) { SC_TRY(sc_work(ctx, 14, &zv)); SC_TRY(sc_work(ctx, 15, &tv)); }
    SC_CUDA(ctx, cudaMemsetAsync(ctx->d_v, 0, sizeof(double) * n, st));
    SC_CUDA(ctx, cudaMemsetAsync(ctx->d_a, 0, sizeof(double) * n, st));
    int64_t pcg_total = 0, row = 0;
    int iters = 0;
    double relres = 0.0;
    sc_gpu_timer timer(st);
    timer.start();
    for (int64_t t = t0; t <= t0 + n_steps; ++t) {
        if (ctx->world > 1) SC_TRY(dist_halo(ctx, ctx->d_u, st));
        SC_TRY(la_spmv(ctx, ctx->d_K, ctx->d_u, rhs));
        SC_TRY(la_fill(ctx, f, 0.0, n));
        SC_TRY(apply_load(ctx, t, 1.0, nullptr, f));
        SC_TRY(la_dot(ctx, f, f, ctx->d_scal + 8));
        SC_TRY(lincomb(ctx, rhs, -1.0, rhs, 1.0, f));
        SC_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 8, ctx->d_scal + 8, sizeof(double), cudaMemcpyDeviceToHost, st));
        SC_CUDA(ctx, cudaStreamSynchronize(st));
        const double ff = ctx->h_pinned[8];
        SC_TRY(pcg(ctx, ctx->d_K, dinv, rhs, du, r, p, q, rtol, maxit, &iters, &relres, ff > 0.0 ? ff : -1.0, F, zv, tv));
        pcg_total += iters;
        SC_TRY(lincomb(ctx, ctx->d_u, 1.0, ctx->d_u, 1.0, du));
        if (is_out_step(ctx, t, oi) && row < n_out) { SC_TRY(store_row(ctx, u_out, row, ctx->d_u)); ++row; }
    }
    timer.stop();
    SC_CUDA(ctx, cudaStreamSynchronize(st));
    const float ms = timer.ms();
    if (stats) {
        stats->seconds_device = ms * 1e-3;
        stats->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
        stats->steps = n_steps;
        stats->pcg_iterations = pcg_total;
        stats->kernel_launches = ctx->launches - launches0;
        stats->last_residual = relres;
    }
    return SC_OK;
}
