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
#include <algorithm>
#include <cub/device/device_scan.cuh>
#include "common.h"
#include "tma.h"

namespace {

constexpr int FSAI_CAP = 48;                       // largest local system
constexpr int FSAI_TRI = FSAI_CAP * (FSAI_CAP + 1) / 2;
constexpr int FSAI_WARPS = 4;                      // warps (rows) per block of the set-up kernels
constexpr unsigned FULL = 0xffffffffu;

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

__device__ __forceinline__ bool fsai_keep(double v, double dinv_i, double dinv_c, double tau) {
    return fabs(v) * sqrt(dinv_i * dinv_c) >= tau;
}

// Elimination order of the triangle.  "Lower" means "earlier in the order", and the order lists the vertex equations of a
// quadratic mesh (class 0) before its mid-side equations (class 1), each class by equation number: for the serendipity
// elements the vertex block is the badly conditioned one (negative lumped vertex mass), and a mid-side row that sees all
// its vertex neighbours in its pattern acts like a two-level factorisation -- hexa20 94^3: 104 PCG iterations per step
// with this order on either node numbering of the box, 250 with plain equation order on the cell-by-cell numbering
// (gpurun_out/r2_precond9.log).  cls == nullptr (linear elements, caller-supplied CSR): plain equation order.
__device__ __forceinline__ bool fsai_before(const unsigned char* __restrict__ cls, int j, int64_t i) {
    if (!cls) return j < i;
    const unsigned char cj = cls[j], ci = cls[i];
    return cj < ci || (cj == ci && j < i);
}

// entries of row i that pass the filter (strictly earlier equations, non-ghost columns); the row's own tau is raised until they fit
__global__ void __launch_bounds__(32 * FSAI_WARPS)
k_fsai_count(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ vals,
             const double* __restrict__ dinv, const unsigned char* __restrict__ cls, double tau0, int64_t n,
             int64_t* __restrict__ cnt, double* __restrict__ row_tau, int* __restrict__ max_row) {
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
            if (fsai_before(cls, j, i) && dinv[j] > 0.0 && fsai_keep(vals[k], di, dinv[j], tau)) ++c;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
        if (c <= FSAI_CAP - 1) break;
        tau *= 1.3;
    }
    if (lane == 0) { cnt[i] = c + 1; row_tau[i] = tau; atomicMax(max_row, c + 1); }
}

__device__ __forceinline__ int tri(int a, int b) { return a * (a + 1) / 2 + b; }      // b <= a

// one warp per row: gather A[P,P], Cholesky, back substitution, write the row of G
__global__ void __launch_bounds__(32 * FSAI_WARPS)
k_fsai_fill(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ vals,
            const double* __restrict__ dinv, const unsigned char* __restrict__ cls, int64_t n,
            const int64_t* __restrict__ g_rowptr, const double* __restrict__ row_tau, int2* __restrict__ g_cv,
            int* __restrict__ n_fail) {
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
            keep = fsai_before(cls, j, i) && dinv[j] > 0.0 && fsai_keep(vals[k], di, dinv[j], tau);
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
    // 2. A[P, P], lower part: row P[a] of the CSR is scanned once, every entry looks its column up in P[0..a].  P[0..m-2]
    //    ascend in equation number; the row's own equation i stands last whatever its number (elimination order above), so
    //    its CSR row is scanned completely and the entries A[P[a], i] of the other rows are taken from it by symmetry.
    for (int a = 0; a < m; ++a) {
        const int r = P[a];
        const bool own = a == m - 1;
        const int64_t r0 = rowptr[r], r1 = rowptr[r + 1];
        for (int64_t k = r0 + lane; k < r1; k += 32) {
            const int c = col[k];
            if (own) {
                if (c == r) { L[tri(a, a)] = vals[k]; continue; }
            } else if (c > r) break;                // columns ascend: the rest of this lane's entries lie in the upper part
            int lo = 0, hi = own ? a : a + 1;
            const int top = hi;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (P[mid] < c) lo = mid + 1; else hi = mid;
            }
            if (lo < top && P[lo] == c) L[tri(a, lo)] = vals[k];
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

// G^T row j = { (i, g_ij) : i = j or j before i, j in P_i }: walk the later part of row j of A (structurally symmetric) in
// column order and look j up in the row of G of every candidate (ascending columns, then the diagonal).  FILL = false
// counts, FILL = true writes.
template <bool FILL>
__global__ void __launch_bounds__(32 * FSAI_WARPS)
k_fsai_transpose(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const unsigned char* __restrict__ cls, int64_t n,
                 const int64_t* __restrict__ g_rowptr, const int2* __restrict__ g_cv, int64_t* __restrict__ cnt, const int64_t* __restrict__ t_rowptr,
                 int2* __restrict__ t_cv, int* __restrict__ max_row) {
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
            if (i == j) {
                const int64_t g1 = g_rowptr[i + 1];
                if (g1 > g_rowptr[i]) { hit = true; v = g_cv[g1 - 1].y; }      // the diagonal stands last in its row
            } else if (fsai_before(cls, (int)j, i)) {
                int64_t lo = g_rowptr[i], hi = g_rowptr[i + 1] - 1;            // off-diagonal part, ascending
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
    if (!FILL && lane == 0) { cnt[j] = total; atomicMax(max_row, total); }
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

// The same product fed by a TMA ring (default): persistent CTAs, a producer thread streams the contiguous (column, value)
// slice and the row pointers of a tile of RT rows with `cp.async.bulk` into a ring of shared-memory stages, eight
// consumer warps reduce NGRP groups of 32 / LPR rows at a time from shared memory -- the only long-latency access left
// to a consumer is the gather of x.  (The register-pipelined kernel above stays at 40-44 % of the DRAM peak: 12-14
// long-scoreboard stalls per issue, and its 1024 blocks leave a 38 % second wave; profiles/r2_hexa20_pcg_kernels_94cube.txt.)
constexpr int CT_WARPS = 8;
template <int LPR, int U, int NGRP, int RT, int STAGES, bool DOT>
__global__ void __launch_bounds__(32 * (CT_WARPS + 1), 2)
k_csr32_tma(const int64_t* __restrict__ rowptr, const int2* __restrict__ cv, const double* __restrict__ x, double* __restrict__ y,
            int64_t n, const double* __restrict__ w, double* __restrict__ partial, int n_partial, int64_t n_tiles, int cap) {
    constexpr int RPW = (32 / LPR) * NGRP;               // rows a warp reduces per pass
    constexpr int PASSES = RT / (CT_WARPS * RPW);
    constexpr int RPS = RT + 2;                          // row pointers per stage
    static_assert(RT % (CT_WARPS * RPW) == 0, "tile rows must split evenly over the consumer warps");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int2* s_cv = reinterpret_cast<int2*>(smem_raw);                                     // [STAGES][cap]
    int64_t* s_rp = reinterpret_cast<int64_t*>(s_cv + (size_t)STAGES * cap);           // [STAGES][RPS]
    double* s_w = reinterpret_cast<double*>(s_rp + (size_t)STAGES * RPS);              // [STAGES][RT] (DOT only)
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(s_w + (DOT ? (size_t)STAGES * RT : 0));
    uint64_t* bar_empty = bar_full + STAGES;
    __shared__ double red[CT_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { tma_mbar_init(&bar_full[s], 1); tma_mbar_init(&bar_empty[s], CT_WARPS); }
        tma_mbar_fence_init();
    }
    __syncthreads();
    const int64_t G = gridDim.x;
    double dot = 0.0;
    if (warp == CT_WARPS) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int64_t tj = blockIdx.x;
            int64_t k0 = 0, k1 = 0;
            if (tj < n_tiles) { k0 = rowptr[tj * RT]; k1 = rowptr[min((tj + 1) * RT, n)]; }
            for (; tj < n_tiles; tj += G) {
                const int64_t a0 = k0, a1 = k1;
                if (tj + G < n_tiles) { k0 = rowptr[(tj + G) * RT]; k1 = rowptr[min((tj + G + 1) * RT, n)]; }
                tma_mbar_wait(&bar_empty[stage], phase ^ 1u);
                const int64_t ks = a0 & ~(int64_t)1;
                const uint32_t cb = (uint32_t)(((a1 - ks + 1) & ~(int64_t)1) * 8);
                const int rows = (int)(min((tj + 1) * RT, n) - tj * RT);
                const uint32_t rb = (uint32_t)(((rows + 1 + 1) & ~1) * 8);
                const uint32_t wb = DOT ? (uint32_t)(((rows + 1) & ~1) * 8) : 0u;
                tma_mbar_expect_tx(&bar_full[stage], cb + rb + wb);
                tma_bulk_load(s_rp + (size_t)stage * RPS, rowptr + tj * RT, rb, &bar_full[stage]);
                if (DOT) tma_bulk_load(s_w + (size_t)stage * RT, w + tj * RT, wb, &bar_full[stage]);
                if (cb) tma_bulk_load_stream(s_cv + (size_t)stage * cap, cv + ks, cb, &bar_full[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        const int sub = lane / LPR, l = lane % LPR;
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t tj = blockIdx.x; tj < n_tiles; tj += G) {
            tma_mbar_wait(&bar_full[stage], phase);
            const int64_t* rp = s_rp + (size_t)stage * RPS;
            const int2* sc = s_cv + (size_t)stage * cap;
            const int64_t ks = rp[0] & ~(int64_t)1;
            const int rows = (int)(min((tj + 1) * RT, n) - tj * RT);
            const double* sw = s_w + (size_t)stage * RT;
#pragma unroll
            for (int pass = 0; pass < PASSES; ++pass) {
                int k0[NGRP], len[NGRP];
                double sum[NGRP], wr[NGRP];
                int maxlen = 0;
#pragma unroll
                for (int g = 0; g < NGRP; ++g) {
                    const int r = (pass * CT_WARPS + warp) * RPW + g * (32 / LPR) + sub;
                    k0[g] = 0; len[g] = 0; sum[g] = 0.0; wr[g] = 0.0;
                    if (r < rows) {
                        k0[g] = (int)(rp[r] - ks); len[g] = (int)(rp[r + 1] - rp[r]);
                        if (DOT) wr[g] = sw[r];
                    }
                    maxlen = max(maxlen, len[g]);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(FULL, maxlen, o));
                for (int kb = 0; kb < maxlen || kb == 0; kb += LPR * U) {
                    int2 c[NGRP][U];
#pragma unroll
                    for (int g = 0; g < NGRP; ++g)
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int k = kb + u * LPR + l;
                            c[g][u] = k < len[g] ? sc[k0[g] + k] : make_int2(-1, 0);
                        }
                    double xg[NGRP][U];
#pragma unroll
                    for (int g = 0; g < NGRP; ++g)
#pragma unroll
                        for (int u = 0; u < U; ++u) xg[g][u] = c[g][u].x >= 0 ? x[c[g][u].x] : 0.0;
                    // last read of the stage by this warp: hand it back while the gathers are in flight
                    if (pass == PASSES - 1 && kb + LPR * U >= maxlen) {
                        __syncwarp();
                        if (lane == 0) tma_mbar_arrive(&bar_empty[stage]);
                    }
#pragma unroll
                    for (int g = 0; g < NGRP; ++g)
#pragma unroll
                        for (int u = 0; u < U; ++u) sum[g] += (double)__int_as_float(c[g][u].y) * xg[g][u];
                }
#pragma unroll
                for (int g = 0; g < NGRP; ++g) {
                    double sv = sum[g];
#pragma unroll
                    for (int o = LPR / 2; o > 0; o >>= 1) sv += __shfl_xor_sync(FULL, sv, o);
                    const int r = (pass * CT_WARPS + warp) * RPW + g * (32 / LPR) + sub;
                    if (l == 0 && r < rows) {
                        const int64_t row = tj * RT + r;
                        y[row] = sv;
                        if (DOT) dot += wr[g] * sv;
                    }
                }
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
    }
    if (DOT) {
        if (warp < CT_WARPS) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dot += __shfl_down_sync(FULL, dot, o);
            if (lane == 0) red[warp] = dot;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int k = 0; k < CT_WARPS; ++k) t += red[k];
            partial[blockIdx.x] = t;
        }
        if (blockIdx.x == 0)                             // the reduction that follows sums n_partial entries
            for (int t = (int)G + (int)threadIdx.x; t < n_partial; t += blockDim.x) partial[t] = 0.0;
    }
}

// ---- component-major renumbering of the factors -----------------------------------------------------------------------
// The filter keeps, for mass-dominated matrices, almost only couplings between equal displacement components.  In the
// node-interleaved equation numbering of the reference (x, y, z of a node are consecutive) the gathers of such a row use one
// double of every 32-byte sector they touch, and the products were bound by the L2 -> SM sector traffic (18.7 GB per
// application for hexa20 94^3 against 4.7 GB of factor data).  The factors are therefore stored in a numbering that lists
// all x equations first, then y, then z (node order kept inside a component): neighbours along the mesh's fastest
// direction share sectors again.  r enters and z leaves through one permutation pass each.
__global__ void k_comp_of_eq(const int32_t* __restrict__ eq, int64_t n_nodes, int dim, unsigned char* __restrict__ comp) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_nodes * dim) return;
    const int e = eq[t];
    if (e >= 0) comp[e] = (unsigned char)(t % dim);
}
__global__ void k_comp_flag(const unsigned char* __restrict__ comp, int c, int64_t n, int64_t* __restrict__ flag) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i <= n) flag[i] = (i < n && comp[i] == c) ? 1 : 0;
}
__global__ void k_comp_assign(const unsigned char* __restrict__ comp, int c, int64_t n, const int64_t* __restrict__ rank, int64_t offset,
                              int32_t* __restrict__ perm) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && comp[i] == c) perm[i] = (int32_t)(offset + rank[i]);
}
__global__ void k_perm_count(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ perm, int64_t n, int64_t* __restrict__ cnt) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) cnt[perm[i]] = rowptr[i + 1] - rowptr[i];
    if (i == n) cnt[n] = 0;
}
// row perm[i] of the output = row i of the input with renumbered columns (entry order kept); 8 lanes per row
__global__ void k_perm_fill(const int64_t* __restrict__ rowptr, const int2* __restrict__ cv, const int32_t* __restrict__ perm, int64_t n,
                            const int64_t* __restrict__ out_rowptr, int2* __restrict__ out) {
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3;
    const int l = threadIdx.x & 7;
    if (i >= n) return;
    const int64_t k0 = rowptr[i], len = rowptr[i + 1] - k0, o0 = out_rowptr[perm[i]];
    for (int64_t k = l; k < len; k += 8) {
        const int2 e = cv[k0 + k];
        out[o0 + k] = make_int2(perm[e.x], e.y);
    }
}
__global__ void k_perm_scatter(const double* __restrict__ x, const int32_t* __restrict__ perm, int64_t n, double* __restrict__ xp) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) xp[perm[i]] = x[i];
}
__global__ void k_perm_gather(const double* __restrict__ xp, const int32_t* __restrict__ perm, int64_t n, double* __restrict__ x) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) x[i] = xp[perm[i]];
}

// elimination class of every equation: 1 for the equations of mid-side nodes (local nodes >= n_vertex of any element)
__global__ void k_mark_midside(const int32_t* __restrict__ conn, int64_t n_elem, int nne, int n_vertex, unsigned char* __restrict__ node_cls) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int nm = nne - n_vertex;
    if (t >= n_elem * nm) return;
    node_cls[conn[(t / nm) * nne + n_vertex + t % nm]] = 1;           // every writer stores the same value
}
__global__ void k_cls_of_eq(const int32_t* __restrict__ eq, const unsigned char* __restrict__ node_cls, int64_t n_nodes, int dim,
                            unsigned char* __restrict__ cls) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_nodes * dim) return;
    const int e = eq[t];
    if (e >= 0) cls[e] = node_cls[t / dim];
}

// largest number of entries in a tile of rt consecutive rows (ring stage size of k_csr32_tma)
__global__ void k_tile_max(const int64_t* __restrict__ rowptr, int64_t n, int rt, int* __restrict__ out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t * rt >= n) return;
    const int64_t e = min((t + 1) * rt, n);
    atomicMax(out, (int)(rowptr[e] - rowptr[t * rt]));
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

constexpr size_t CT_SMEM_2CTA = 110 * 1024;

template <int LPR, int U, int NGRP, int RT, bool DOT>
int csr32_tma_go(sc_ctx* ctx, int tile_max, const int64_t* rowptr, const int2* cv, const double* x, double* y, const double* w,
                 double* partial, int nb, bool* done) {
    const int64_t n = ctx->n_eq;
    const int cap = (tile_max + 2 + 1) & ~1;             // largest tile + the alignment entries at both ends
    const size_t per_stage = (size_t)cap * 8 + (size_t)(RT + 2) * 8 + (DOT ? (size_t)RT * 8 : 0);
    const int64_t n_tiles = (n + RT - 1) / RT;
    unsigned grid = (unsigned)std::min<int64_t>(std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * 2), nb);
    if (grid == 0) grid = 1;
    *done = true;
#define SC_CT_GO(ST)                                                                                                       \
    {                                                                                                                      \
        auto kern = k_csr32_tma<LPR, U, NGRP, RT, ST, DOT>;                                                                \
        const size_t bytes = ST * per_stage + 2 * ST * sizeof(uint64_t) + 64;                                              \
        SC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));                 \
        kern<<<grid, 32 * (CT_WARPS + 1), bytes, ctx->stream>>>(rowptr, cv, x, y, n, w, partial, nb, n_tiles, cap);        \
        SC_CHECK_LAUNCH(ctx);                                                                                              \
        return SC_OK;                                                                                                      \
    }
    if (6 * per_stage + 128 <= CT_SMEM_2CTA) SC_CT_GO(6)
    if (4 * per_stage + 128 <= CT_SMEM_2CTA) SC_CT_GO(4)
    if (3 * per_stage + 128 <= CT_SMEM_2CTA) SC_CT_GO(3)
    if (2 * per_stage + 128 <= CT_SMEM_2CTA) SC_CT_GO(2)
#undef SC_CT_GO
    *done = false;                                       // rows too long for the ring: register-pipelined kernel
    return SC_OK;
}

template <bool DOT>
int csr32_launch(sc_ctx* ctx, const sc_fsai* f, bool transposed, const double* x, double* y, const double* w, double* partial, int nb) {
    const int64_t n = ctx->n_eq;
    cudaStream_t st = ctx->stream;
    const int64_t* rowptr = transposed ? f->t_rowptr : f->rowptr;
    const int2* cv = transposed ? f->t_cv : f->cv;
    const int tile_max = transposed ? f->t_tile_max : f->tile_max;
    if (!ctx->force_no_tma && tile_max > 0) {
        bool done = false;
        if (f->lanes <= 4) SC_TRY((csr32_tma_go<4, 4, 2, 256, DOT>(ctx, tile_max, rowptr, cv, x, y, w, partial, nb, &done)));
        else SC_TRY((csr32_tma_go<8, 6, 2, 64, DOT>(ctx, tile_max, rowptr, cv, x, y, w, partial, nb, &done)));
        if (done) return SC_OK;
    }
    if (f->lanes <= 4) k_csr32_spmv<4, 4, DOT><<<nb, 256, 0, st>>>(rowptr, cv, x, y, n, w, partial);      // short rows (hexa8: ~9 entries)
    else k_csr32_spmv<8, 6, DOT><<<nb, 256, 0, st>>>(rowptr, cv, x, y, n, w, partial);                  // one pass up to FSAI_CAP
    SC_CHECK_LAUNCH(ctx);
    return SC_OK;
}

}  // namespace

void fsai_free(sc_fsai* f) {
    sc_free(&f->rowptr); sc_free(&f->cv);
    sc_free(&f->t_rowptr); sc_free(&f->t_cv);
    sc_free(&f->perm);
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
    unsigned char *cls = nullptr, *node_cls = nullptr;
    int rc = SC_OK;
    auto body = [&]() -> int {
        SC_TRY(sc_alloc(ctx, &dinv, (size_t)n));
        SC_TRY(sc_alloc(ctx, &cnt, (size_t)n + 1));
        SC_TRY(sc_alloc(ctx, &row_tau, (size_t)n));
        SC_TRY(sc_alloc(ctx, &n_fail, 3));                 // [0] rows that fell back to Jacobi, [1], [2] longest row of G^T, G
        SC_CUDA(ctx, cudaMemsetAsync(n_fail, 0, 3 * sizeof(int), st));
        SC_CUDA(ctx, cudaMemsetAsync(cnt + n, 0, sizeof(int64_t), st));
        SC_TRY(la_extract_diag(ctx, vals, dinv, true));
        const unsigned nb = nblk(n, FSAI_WARPS);
        // vertex equations before mid-side equations (fsai_before); linear elements and caller-supplied matrices: one class
        const int n_vertex = ctx->elem_type == SC_TRI6 ? 3 : ctx->elem_type == SC_QUAD8 ? 4 : ctx->elem_type == SC_TETRA10 ? 4
                             : ctx->elem_type == SC_HEXA20 ? 8 : 0;
        if (n_vertex && ctx->d_eq && ctx->d_conn && !ctx->csr_only && !ctx->fsai_no_vertex_first) {
            SC_TRY(sc_alloc(ctx, &cls, (size_t)n));
            SC_TRY(sc_alloc(ctx, &node_cls, (size_t)ctx->n_nodes));
            SC_CUDA(ctx, cudaMemsetAsync(cls, 0, (size_t)n, st));
            SC_CUDA(ctx, cudaMemsetAsync(node_cls, 0, (size_t)ctx->n_nodes, st));
            k_mark_midside<<<nblk(ctx->n_elem * (ctx->nne - n_vertex), 256), 256, 0, st>>>(ctx->d_conn, ctx->n_elem, ctx->nne, n_vertex, node_cls);
            SC_CHECK_LAUNCH(ctx);
            k_cls_of_eq<<<nblk(ctx->n_nodes * ctx->dim, 256), 256, 0, st>>>(ctx->d_eq, node_cls, ctx->n_nodes, ctx->dim, cls);
            SC_CHECK_LAUNCH(ctx);
        }
        k_fsai_count<<<nb, 32 * FSAI_WARPS, 0, st>>>(ctx->d_rowptr, ctx->d_col, vals, dinv, cls, ctx->fsai_tau, n, cnt, row_tau, n_fail + 2);
        SC_CHECK_LAUNCH(ctx);
        SC_TRY(sc_alloc(ctx, &f->rowptr, (size_t)n + 1));
        SC_TRY(scan_counts(ctx, cnt, f->rowptr, n, &f->nnz));
        SC_TRY(sc_alloc(ctx, &f->cv, (size_t)f->nnz));
        k_fsai_fill<<<nb, 32 * FSAI_WARPS, 0, st>>>(ctx->d_rowptr, ctx->d_col, vals, dinv, cls, n, f->rowptr, row_tau, f->cv, n_fail);
        SC_CHECK_LAUNCH(ctx);
        // transpose (same number of entries)
        SC_CUDA(ctx, cudaMemsetAsync(cnt + n, 0, sizeof(int64_t), st));
        k_fsai_transpose<false><<<nb, 32 * FSAI_WARPS, 0, st>>>(ctx->d_rowptr, ctx->d_col, cls, n, f->rowptr, f->cv, cnt, nullptr, nullptr, n_fail + 1);
        SC_CHECK_LAUNCH(ctx);
        SC_TRY(sc_alloc(ctx, &f->t_rowptr, (size_t)n + 1));
        int64_t tn = 0;
        SC_TRY(scan_counts(ctx, cnt, f->t_rowptr, n, &tn));
        if (tn != f->nnz) return sc_fail(ctx, SC_ERR_STATE, "FSAI transpose found %lld of %lld entries: the CSR pattern is not "
                                         "structurally symmetric", (long long)tn, (long long)f->nnz);
        SC_TRY(sc_alloc(ctx, &f->t_cv, (size_t)f->nnz));
        k_fsai_transpose<true><<<nb, 32 * FSAI_WARPS, 0, st>>>(ctx->d_rowptr, ctx->d_col, cls, n, f->rowptr, f->cv, nullptr, f->t_rowptr,
                                                               f->t_cv, nullptr);
        SC_CHECK_LAUNCH(ctx);
        int h[3] = {0, 0, 0};
        SC_CUDA(ctx, cudaMemcpyAsync(h, n_fail, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
        SC_CUDA(ctx, cudaStreamSynchronize(st));
        f->jacobi_rows = h[0];
        f->t_max_row = h[1] > 0 ? h[1] : 1;
        f->max_row = h[2] > 0 ? h[2] : 1;
        // component-major numbering (needs the node -> equation table of a mesh; caller-supplied CSR keeps its numbering)
        if (ctx->d_eq && ctx->dim > 1 && !ctx->csr_only && !ctx->fsai_no_perm) {
            unsigned char* comp = nullptr;
            int64_t* rank = nullptr;
            SC_TRY(sc_alloc(ctx, &comp, (size_t)n));
            SC_TRY(sc_alloc(ctx, &rank, (size_t)n + 1));
            SC_TRY(sc_alloc(ctx, &f->perm, (size_t)n));
            int prc = SC_OK;
            auto pbody = [&]() -> int {
                SC_CUDA(ctx, cudaMemsetAsync(comp, 0, (size_t)n, st));
                k_comp_of_eq<<<nblk(ctx->n_nodes * ctx->dim, 256), 256, 0, st>>>(ctx->d_eq, ctx->n_nodes, ctx->dim, comp);
                SC_CHECK_LAUNCH(ctx);
                int64_t offset = 0;
                for (int c = 0; c < ctx->dim; ++c) {
                    k_comp_flag<<<nblk(n + 1, 256), 256, 0, st>>>(comp, c, n, cnt);
                    SC_CHECK_LAUNCH(ctx);
                    int64_t total = 0;
                    SC_TRY(scan_counts(ctx, cnt, rank, n, &total));
                    k_comp_assign<<<nblk(n, 256), 256, 0, st>>>(comp, c, n, rank, offset, f->perm);
                    SC_CHECK_LAUNCH(ctx);
                    offset += total;
                }
                if (offset != n) return sc_fail(ctx, SC_ERR_STATE, "FSAI renumbering covers %lld of %lld equations", (long long)offset, (long long)n);
                for (int side = 0; side < 2; ++side) {
                    int64_t*& rp = side ? f->t_rowptr : f->rowptr;
                    int2*& cvp = side ? f->t_cv : f->cv;
                    int64_t* nrp = nullptr;
                    int2* ncv = nullptr;
                    k_perm_count<<<nblk(n + 1, 256), 256, 0, st>>>(rp, f->perm, n, cnt);
                    SC_CHECK_LAUNCH(ctx);
                    SC_TRY(sc_alloc(ctx, &nrp, (size_t)n + 1));
                    int64_t total = 0;
                    int rc2 = scan_counts(ctx, cnt, nrp, n, &total);
                    if (rc2 == SC_OK) rc2 = sc_alloc(ctx, &ncv, (size_t)f->nnz);
                    if (rc2 != SC_OK) { sc_free(&nrp); sc_free(&ncv); return rc2; }
                    k_perm_fill<<<nblk(n * 8, 256), 256, 0, st>>>(rp, cvp, f->perm, n, nrp, ncv);
                    ctx->launches++;
                    if (cudaStreamSynchronize(st) != cudaSuccess) { sc_free(&nrp); sc_free(&ncv); return sc_fail(ctx, SC_ERR_CUDA, "FSAI renumbering failed"); }
                    sc_free(&rp); sc_free(&cvp);
                    rp = nrp; cvp = ncv;
                }
                return SC_OK;
            };
            prc = pbody();
            sc_free(&comp); sc_free(&rank);
            SC_TRY(prc);
        }
        return SC_OK;
    };
    rc = body();
    timer.stop();
    if (rc == SC_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = sc_fail(ctx, SC_ERR_CUDA, "FSAI set-up failed: %s", cudaGetErrorString(cudaGetLastError()));
    sc_free(&dinv); sc_free(&cnt); sc_free(&row_tau); sc_free(&n_fail); sc_free(&cls); sc_free(&node_cls);
    if (rc != SC_OK) { fsai_free(f); return rc; }
    f->seconds = timer.ms() * 1e-3;
    const double avg = n > 0 ? (double)f->nnz / (double)n : 1.0;
    f->lanes = avg <= 12.0 ? 4 : 8;
    {   // ring stage size of the TMA-fed products: the fullest tile of each factor (tiles of 256 / 64 rows, csr32_launch)
        const int rt = f->lanes <= 4 ? 256 : 64;
        int* d_max = nullptr;
        int h[2] = {0, 0};
        if (sc_alloc(ctx, &d_max, 2) == SC_OK) {
            cudaMemsetAsync(d_max, 0, 2 * sizeof(int), st);
            const unsigned nt = nblk((n + rt - 1) / rt, 256);
            k_tile_max<<<nt, 256, 0, st>>>(f->rowptr, n, rt, d_max);
            k_tile_max<<<nt, 256, 0, st>>>(f->t_rowptr, n, rt, d_max + 1);
            ctx->launches += 2;
            cudaMemcpyAsync(h, d_max, 2 * sizeof(int), cudaMemcpyDeviceToHost, st);
            if (cudaStreamSynchronize(st) != cudaSuccess) h[0] = h[1] = 0;      // 0: register-pipelined kernels
            sc_free(&d_max);
        }
        f->tile_max = h[0]; f->t_tile_max = h[1];
    }
    f->for_vals = vals;
    return SC_OK;
}

// t = G r, z = G^T t, partial[b] = sum over the rows of block b of r.z   (nb blocks).  With the component-major numbering
// the two products run on renumbered vectors: rp = P r, t = Gp rp, z = Gp^T t -- z is then LEFT in that numbering
// (entry f->perm[i] belongs to equation i; fsai_unpermute, or a consumer that reads z through f->perm).
int fsai_apply(sc_ctx* ctx, const sc_fsai* f, const double* r, double* t, double* z, double* partial, int nb, double* rp) {
    if (!f->perm) {
        SC_TRY(csr32_launch<false>(ctx, f, false, r, t, nullptr, nullptr, nb));
        SC_TRY(csr32_launch<true>(ctx, f, true, t, z, r, partial, nb));
        return SC_OK;
    }
    const int64_t n = ctx->n_eq;
    k_perm_scatter<<<nblk(n, 256), 256, 0, ctx->stream>>>(r, f->perm, n, rp);
    SC_CHECK_LAUNCH(ctx);
    SC_TRY(csr32_launch<false>(ctx, f, false, rp, t, nullptr, nullptr, nb));
    SC_TRY(csr32_launch<true>(ctx, f, true, t, z, rp, partial, nb));
    return SC_OK;
}
int fsai_unpermute(sc_ctx* ctx, const sc_fsai* f, const double* zp, double* z) {
    const int64_t n = ctx->n_eq;
    if (!f->perm) {
        if (zp != z) SC_CUDA(ctx, cudaMemcpyAsync(z, zp, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        return SC_OK;
    }
    k_perm_gather<<<nblk(n, 256), 256, 0, ctx->stream>>>(zp, f->perm, n, z);
    SC_CHECK_LAUNCH(ctx);
    return SC_OK;
}
