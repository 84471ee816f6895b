// TMA-fed CSR SpMV for the time loop (sm_100a): bulk asynchronous copies stream the matrix, warps only do the math.
//
// The time loop is bound by streaming K's values and column indices from HBM (~1 kB per dof and step).
// ncu on the register-staged kernel (linalg.cu k_spmv) showed the limit is memory-level parallelism: a warp alternates
// between waiting for its (value, column) loads and waiting for the dependent x gathers, so DRAM requests are in
// flight only about half of the time.  Here the two are decoupled:
//
//   producer warp   one elected lane issues `cp.async.bulk` (TMA, 1-D) copies of the contiguous value / column slices
//                   of a tile of TR consecutive rows into a ring of shared-memory stages; completion is signalled on an
//                   mbarrier (`complete_tx::bytes`).  The ring keeps STAGES * ~31 kB per CTA in flight with no register
//                   cost, independent of what the consumers are doing.
//   consumer warps  wait on the stage's "full" barrier, read values / columns from shared memory, gather x (L2
//                   resident), reduce each row with a fixed xor butterfly, apply the fused epilogue (central-difference
//                   update / PCG dot) and release the stage through the "empty" barrier.
//
// Persistent CTAs (2 per SM) walk the tiles round-robin, so neighbouring tiles are processed at the same time on
// different SMs and share their x lines in L2.  Deterministic: fixed lane partial order + butterfly, no atomics.
#include <algorithm>
#include "common.h"

namespace {

constexpr int TMA_CONSUMER_WARPS = 8;
constexpr int TMA_THREADS = 32 * (TMA_CONSUMER_WARPS + 1);
constexpr int TMA_STAGES = 3;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// MODE 0: y = A xa   MODE 2: central-difference step (see linalg.cu)   MODE 3: y = A xa and partial[blockIdx] = xa.y
// TR rows per tile, RW = TR / 8 rows per consumer warp.  cap = tile capacity in entries (multiple of 4).
template <int MODE, int TR>
__global__ void __launch_bounds__(TMA_THREADS, 2)
k_spmv_tma(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ va,
           const double* __restrict__ xa, double* __restrict__ y, const double* __restrict__ inv_d,
           const double* __restrict__ alpha, double* __restrict__ partial, int64_t n_rows, int64_t n_tiles, int cap,
           const double* __restrict__ xe, double* __restrict__ y2, double lag) {
    constexpr int RW = TR / TMA_CONSUMER_WARPS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [STAGES][cap] doubles | [STAGES][cap] ints | barriers
    double* s_val = reinterpret_cast<double*>(smem_raw);
    int* s_col = reinterpret_cast<int*>(s_val + (size_t)TMA_STAGES * cap);
    // per-row operands of the fused epilogue (alpha, inv_d, x[row], y[row]) travel through the same ring: as single
    // 8-byte global loads they cost 2.5 ms per step on the 50 M-dof box (32-byte DRAM requests), as bulk copies ~0.4 ms
    constexpr int NVEC = (MODE == 2) ? 4 : (MODE == 3 ? 1 : 0);
    double* s_vec = reinterpret_cast<double*>(s_col + (size_t)TMA_STAGES * cap);          // [STAGES][NVEC][TR]
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(s_vec + (size_t)TMA_STAGES * (NVEC > 0 ? NVEC : 1) * TR);
    uint64_t* bar_empty = bar_full + TMA_STAGES;
    __shared__ double red[TMA_CONSUMER_WARPS];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TMA_STAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], TMA_CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t G = gridDim.x;
    double dot_acc = 0.0;

    if (warp == TMA_CONSUMER_WARPS) {
        // ------------------------------------------------ producer: lane 0 of the last warp ----------------------------
        // (one thread, slice bounds fetched one tile ahead; see spmv_node.cu for why the copies are not issued from inside a
        // warp-wide shuffle loop)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int64_t tj = blockIdx.x;
            int64_t s0 = 0, s1 = 0;
            if (tj < n_tiles) { s0 = rowptr[tj * TR]; s1 = rowptr[min(tj * TR + TR, n_rows)]; }
            for (; tj < n_tiles; tj += G) {
                const int64_t a0 = s0, a1 = s1;
                if (tj + G < n_tiles) { s0 = rowptr[(tj + G) * TR]; s1 = rowptr[min((tj + G) * TR + TR, n_rows)]; }
                mbar_wait(&bar_empty[stage], phase ^ 1u);
                const int64_t v0 = a0 & ~(int64_t)1;                 // 16-byte aligned start in the value array
                const int64_t c0 = a0 & ~(int64_t)3;                 // ... and in the column array
                const uint32_t vb = (uint32_t)(((a1 - v0 + 1) & ~(int64_t)1) * 8);
                const uint32_t cb = (uint32_t)(((a1 - c0 + 3) & ~(int64_t)3) * 4);
                if (a1 > a0) {
                    const int64_t r0 = tj * TR;
                    const uint32_t rb = (uint32_t)(((min((int64_t)TR, n_rows - r0) + 1) & ~(int64_t)1) * 8);
                    mbar_expect_tx(&bar_full[stage], vb + cb + NVEC * rb);
                    tma_load_1d(s_val + (size_t)stage * cap, va + v0, vb, &bar_full[stage]);
                    tma_load_1d(s_col + (size_t)stage * cap, col + c0, cb, &bar_full[stage]);
                    double* sv = s_vec + (size_t)stage * (NVEC > 0 ? NVEC : 1) * TR;
                    if (MODE == 2) {
                        tma_load_1d(sv, alpha + r0, rb, &bar_full[stage]);
                        tma_load_1d(sv + TR, inv_d + r0, rb, &bar_full[stage]);
                        tma_load_1d(sv + 2 * TR, xe + r0, rb, &bar_full[stage]);
                        tma_load_1d(sv + 3 * TR, y + r0, rb, &bar_full[stage]);
                    }
                    if (MODE == 3) tma_load_1d(sv, xa + r0, rb, &bar_full[stage]);
                } else {
                    mbar_arrive(&bar_full[stage]);                   // empty tile (ghost rows): nothing to copy
                }
                if (++stage == TMA_STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------ consumer warps ---------------------------------------------
        // Everything inside a tile is addressed with 32-bit offsets relative to the staged slices (the instruction
        // count per entry, not DRAM, limited the first version of this kernel: 52 % issue utilisation at 58 % of HBM).
        constexpr int LPR = 32 / RW;                 // after the reduction row r of the warp lives in lanes [r*LPR, (r+1)*LPR)
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += G) {
            const int64_t r0 = t * TR;
            const int64_t row0 = r0 + (int64_t)warp * RW;
            // row pointers of this warp's rows (lanes 0..RW) + the tile start (lane 31); issued before the barrier wait
            int64_t rp = 0;
            if (lane <= RW && row0 + lane <= n_rows) rp = rowptr[row0 + lane];
            if (lane == 31) rp = rowptr[r0];
            const int myr = lane / LPR;
            const int64_t myrow = row0 + myr;
            const bool owner = (lane % LPR) == 0 && myrow < n_rows;
            const int64_t tile_s0 = __shfl_sync(0xffffffffu, rp, 31);
            const int dvc = (int)((tile_s0 & ~(int64_t)1) - (tile_s0 & ~(int64_t)3));     // value-slice vs column-slice origin
            const int loc = (int)(rp - (tile_s0 & ~(int64_t)1));                         // offset inside the value slice
            int rs[RW], len[RW];
            int maxlen = 0;
#pragma unroll
            for (int r = 0; r < RW; ++r) {
                rs[r] = __shfl_sync(0xffffffffu, loc, r);
                const int e = __shfl_sync(0xffffffffu, loc, r + 1);
                len[r] = (row0 + r < n_rows) ? e - rs[r] : 0;
                maxlen = max(maxlen, len[r]);
                rs[r] += lane;
            }
            const double* sv = s_val + (size_t)stage * cap;
            const int* sc = s_col + (size_t)stage * cap + dvc;
            mbar_wait(&bar_full[stage], phase);
            double e_al = 0.0, e_id = 0.0, e_x = 0.0, e_y = 0.0;
            if (NVEC > 0 && owner && maxlen > 0) {
                const double* svec = s_vec + (size_t)stage * NVEC * TR + warp * RW + myr;
                if (MODE == 2) { e_al = svec[0]; e_id = svec[TR]; e_x = svec[2 * TR]; e_y = svec[3 * TR]; }
                if (MODE == 3) e_x = svec[0];
            }

            double sum[RW];
#pragma unroll
            for (int r = 0; r < RW; ++r) sum[r] = 0.0;
            constexpr int U = 3;
            for (int jb = 0; jb < maxlen; jb += 32 * U) {
                double v[RW][U], xg[RW][U];
                int c[RW][U];
#pragma unroll
                for (int r = 0; r < RW; ++r)
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int off = jb + 32 * u;
                        const bool ok = lane + off < len[r];
                        v[r][u] = 0.0;
                        c[r][u] = 0;
                        if (ok) {
                            v[r][u] = sv[rs[r] + off];
                            c[r][u] = sc[rs[r] + off];
                        }
                    }
#pragma unroll
                for (int r = 0; r < RW; ++r)
#pragma unroll
                    for (int u = 0; u < U; ++u) xg[r][u] = __ldg(xa + c[r][u]);
                __syncwarp();      // scheduling fence: all gathers of the pass are issued before the first FMA
#pragma unroll
                for (int r = 0; r < RW; ++r)
#pragma unroll
                    for (int u = 0; u < U; ++u) sum[r] += v[r][u] * xg[r][u];
            }
            // the stage's shared memory is no longer needed by this warp
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[stage]);
            if (++stage == TMA_STAGES) { stage = 0; phase ^= 1u; }

            // multi-row butterfly: halve the number of live rows per lane at every step, fixed order => deterministic
            double mine;
            if (RW == 4) {
                const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0;
                double k0 = h16 ? sum[2 % RW] : sum[0], k1 = h16 ? sum[3 % RW] : sum[1 % RW];
                const double g0 = h16 ? sum[0] : sum[2 % RW], g1 = h16 ? sum[1 % RW] : sum[3 % RW];
                k0 += __shfl_xor_sync(0xffffffffu, g0, 16);
                k1 += __shfl_xor_sync(0xffffffffu, g1, 16);
                mine = h8 ? k1 : k0;
                const double g = h8 ? k0 : k1;
                mine += __shfl_xor_sync(0xffffffffu, g, 8);
                mine += __shfl_xor_sync(0xffffffffu, mine, 4);
                mine += __shfl_xor_sync(0xffffffffu, mine, 2);
                mine += __shfl_xor_sync(0xffffffffu, mine, 1);
            } else if (RW == 2) {
                const bool h16 = (lane & 16) != 0;
                mine = h16 ? sum[1 % RW] : sum[0];
                const double g = h16 ? sum[0] : sum[1 % RW];
                mine += __shfl_xor_sync(0xffffffffu, g, 16);
                mine += __shfl_xor_sync(0xffffffffu, mine, 8);
                mine += __shfl_xor_sync(0xffffffffu, mine, 4);
                mine += __shfl_xor_sync(0xffffffffu, mine, 2);
                mine += __shfl_xor_sync(0xffffffffu, mine, 1);
            } else {
                mine = sum[0];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
            }
            int mylen = len[0];
#pragma unroll
            for (int r = 1; r < RW; ++r) mylen = (myr == r) ? len[r] : mylen;
            if (owner) {
                if (MODE == 2) {
                    if (mylen > 0) {
                        const double un = e_id * (-mine) + e_al * e_x - (e_al - 1.0) * e_y;
                        y[myrow] = un;
                        if (y2) y2[myrow] = (1.0 + lag) * un - lag * e_x;
                    }
                } else {
                    y[myrow] = mine;
                    if (MODE == 3 && mylen > 0) dot_acc += e_x * mine;
                }
            }
        }
    }
    if (MODE == 3) {
        if (warp < TMA_CONSUMER_WARPS) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dot_acc += __shfl_down_sync(0xffffffffu, dot_acc, o);
            if (lane == 0) red[warp] = dot_acc;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < TMA_CONSUMER_WARPS; ++w) s += red[w];
            partial[blockIdx.x] = s;
        }
    }
}

template <int MODE, int TR>
int launch_tr(sc_ctx* ctx, const double* va, const double* xa, double* y, const double* inv_d, const double* alpha,
              double* partial, unsigned* nblocks_out, const double* xe, double* y2, double g) {
    const int64_t n = ctx->n_eq;
    const int64_t n_tiles = (n + TR - 1) / TR;
    int cap = TR * ctx->max_rl + 8;
    cap = (cap + 31) & ~31;                                  // stage starts stay 128-byte aligned
    const size_t bytes = (size_t)TMA_STAGES * cap * (sizeof(double) + sizeof(int)) + (size_t)TMA_STAGES * 4 * TR * sizeof(double) +
                         2 * TMA_STAGES * sizeof(uint64_t);
    auto kern = k_spmv_tma<MODE, TR>;
    SC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    unsigned grid = (unsigned)std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * 2);
    if (grid == 0) grid = 1;
    if (nblocks_out) *nblocks_out = grid;
    kern<<<grid, TMA_THREADS, bytes, ctx->stream>>>(ctx->d_rowptr, ctx->d_col, va, xa, y, inv_d, alpha, partial, n, n_tiles, cap, xe, y2, g);
    SC_CHECK_LAUNCH(ctx);
    return SC_OK;
}

}  // namespace

// false when the rows are too long for the staging ring: the caller then uses the register-staged kernel (linalg.cu)
bool la_tma_usable(sc_ctx* ctx) {
    if (ctx->force_no_tma || ctx->max_rl <= 0) return false;
    const size_t per_entry = sizeof(double) + sizeof(int);
    const size_t budget = 100 * 1024;                        // per CTA, two CTAs per SM
    return (size_t)TMA_STAGES * (8 * (size_t)ctx->max_rl + 40) * per_entry + 4096 <= budget;
}

template <int MODE>
static int launch_mode(sc_ctx* ctx, const double* va, const double* xa, double* y, const double* inv_d, const double* alpha,
                       double* partial, unsigned* nblocks_out, const double* xe = nullptr, double* y2 = nullptr, double g = 0.0) {
    const size_t per_entry = sizeof(double) + sizeof(int);
    const size_t budget = 100 * 1024;
    auto fits = [&](int tr) { return (size_t)TMA_STAGES * ((size_t)tr * ctx->max_rl + 40) * per_entry + (size_t)TMA_STAGES * 32 * tr <= budget; };
    if (fits(32)) return launch_tr<MODE, 32>(ctx, va, xa, y, inv_d, alpha, partial, nblocks_out, xe, y2, g);
    if (fits(16)) return launch_tr<MODE, 16>(ctx, va, xa, y, inv_d, alpha, partial, nblocks_out, xe, y2, g);
    return launch_tr<MODE, 8>(ctx, va, xa, y, inv_d, alpha, partial, nblocks_out, xe, y2, g);
}

int la_tma_spmv(sc_ctx* ctx, const double* vals, const double* x, double* y) {
    return launch_mode<0>(ctx, vals, x, y, nullptr, nullptr, nullptr, nullptr);
}
int la_tma_cd_step(sc_ctx* ctx, const double* K, const double* w, const double* u, double* uprev_next, const double* inv_d,
                   const double* alpha, double g, double* w_next) {
    return launch_mode<2>(ctx, K, w, uprev_next, inv_d, alpha, nullptr, nullptr, u, w_next, g);
}
int la_tma_spmv_dot(sc_ctx* ctx, const double* vals, const double* p, double* q, double* partial, unsigned* nblocks) {
    return launch_mode<3>(ctx, vals, p, q, nullptr, nullptr, partial, nblocks);
}
