// Node-blocked, TMA-fed SpMV for the time loop (sm_100a).
//
// In the reference's CSR (scatter/system_matrix.py:98-121) the rows of one node -- its 1..3 free dofs -- have identical
// column lists, because two dofs couple iff their nodes share an element.  Storing the list once per node instead of
// once per row removes 2/3 of the column-index traffic (81*4 -> 27*4 bytes per row for interior hexa8 rows) and, more
// importantly for the SM, 2/3 of the x gathers and of the per-entry instructions: a lane loads one column index,
// gathers x once and feeds up to three FMAs (one per row of the node).  The value array is untouched -- it is still the
// reference CSR value array, row after row -- only the index structure differs (pattern.cu builds `ncol` and the
// 24-byte node descriptors next to the exported CSR pattern).
//
// Data movement follows spmv_tma.cu: persistent CTAs, one producer warp issuing `cp.async.bulk` copies of a tile's
// contiguous slices (values, node columns, node descriptors, epilogue vectors) into a ring of shared-memory stages,
// mbarrier full/empty handshakes, eight consumer warps (one node each per tile).  Deterministic: fixed lane partials
// and a fixed butterfly, no atomics.
#include <algorithm>
#include <vector>
#include "common.h"

namespace {

constexpr int NB_WARPS = 8;                       // consumer warps
using NodeDesc = sc_ctx::NodeDesc;

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

// streaming loads: the value slices are read exactly once per product -- L2 evict-first keeps them from displacing the
// gathered vector (three node planes of it are live at any time) and the lines prefetched for the next tiles
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;
__device__ __forceinline__ void tma_load_1d_stream(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(L2_EVICT_FIRST)
                 : "memory");
}
// MODE 0: y = A xa   MODE 3: y = A xa, partial[blockIdx] = xa.y
// MODE 2: central-difference step  y <- inv_d (-A xa) + alpha xe - (alpha - 1) y  with xe = u(t), y = u(t-dt) on entry and
//         xa = w(t) = (1+g) u(t) - g u(t-dt) the gathered vector (lagged stiffness-proportional damping, g = c1/dt);
//         y2 != null: the next gather vector w(t+dt) = (1+g) y - g xe is written by the same epilogue (g = 0: xa = xe, y2 = null)
// NG consumer groups of NB_WARPS warps share one ring: group g takes the CTA's tiles g, g + NG, ... -- with NG = 2 (one CTA
// per SM, a ring of 4-6 stages) two tiles are being reduced while the others load, which hides the per-tile latency of a
// consumer warp (gather -> FMA -> butterfly -> store) that a two-stage ring per CTA exposes.
template <int MODE, int STAGES, int NB_NPW, int NG>
__global__ void __launch_bounds__(32 * (NG * NB_WARPS + 1), NG == 1 ? 2 : 1)
k_spmv_node(const NodeDesc* __restrict__ nd, const int32_t* __restrict__ ncol, const double* __restrict__ va,
            const double* __restrict__ xa, double* __restrict__ y, const double* __restrict__ inv_d,
            const double* __restrict__ alpha, double* __restrict__ partial, int64_t n_nodes, int64_t n_rows, int64_t n_tiles,
            int cap_v, int cap_c, const int32_t* __restrict__ dict, int n_dict, int dict_stride,
            const double* __restrict__ xe, double* __restrict__ y2, double lag, int64_t t_lo1, int64_t t_n1, int64_t t_lo2) {
    // `n_tiles` counts the tiles of this launch: positions v < t_n1 map to tile t_lo1 + v, the others to t_lo2 + (v - t_n1)
    // (whole matrix: t_lo1 = 0, t_n1 = n_tiles; the halo overlap launches the two ends and the interior separately)
    constexpr int NB_NODES = NB_WARPS * NB_NPW;      // nodes per tile; NB_NPW nodes per consumer warp (gathers in flight together)
    constexpr int NB_VT = 3 * NB_NODES + 8;          // vector slots per tile (rows + alignment), multiple of 2
    constexpr int NVEC = (MODE == 2) ? 4 : (MODE == 3 ? 1 : 0);
    constexpr int NV1 = NVEC > 0 ? NVEC : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* s_val = reinterpret_cast<double*>(smem_raw);                         // [STAGES][cap_v]
    double* s_vec = s_val + (size_t)STAGES * cap_v;                              // [STAGES][NV1][NB_VT]
    NodeDesc* s_nd = reinterpret_cast<NodeDesc*>(s_vec + (size_t)STAGES * NV1 * NB_VT);   // [STAGES][NB_NODES]
    int* s_col = reinterpret_cast<int*>(s_nd + (size_t)STAGES * NB_NODES);       // [STAGES][cap_c]
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(s_col + (size_t)STAGES * cap_c);
    uint64_t* bar_empty = bar_full + STAGES;
    int* s_dict = reinterpret_cast<int*>(bar_empty + STAGES);                     // [n_dict][dict_stride] relative column lists (node_dict.cu)
    __shared__ double red[NG * NB_WARPS];
    constexpr int NTHREADS = 32 * (NG * NB_WARPS + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], NB_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int t = threadIdx.x; t < n_dict * dict_stride; t += NTHREADS) s_dict[t] = dict[t];
    __syncthreads();

    const int64_t G = gridDim.x;
    double dot_acc = 0.0;

    if (warp == NG * NB_WARPS) {
        // ------------------------------------------------ producer: lane 0 of the last warp ----------------------------
        // One thread walks this CTA's tiles: slice bounds from the node descriptors (fetched one tile ahead, so the look-up
        // overlaps the wait for a free stage), wait for the stage, arm the barrier, issue the bulk copies.  The other 31 lanes
        // leave.  (Round 1 let every lane fetch the bounds of one of 32 upcoming tiles and broadcast them with shuffles while
        // lane 0 issued the copies: with lane 0 spinning on the "empty" barrier inside the shuffle loop the ring only
        // streamed at 5.9 TB/s; this form reaches the 7.3 TB/s of a bare TMA ring -- scripts/probes/tma_stream_probe.cu.)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int64_t vj = blockIdx.x;
            auto tile_of = [&](int64_t v) { return v < t_n1 ? t_lo1 + v : t_lo2 + (v - t_n1); };
            NodeDesc n0, n1;
            if (vj < n_tiles) { const int64_t t = tile_of(vj); n0 = nd[t * NB_NODES]; n1 = nd[t * NB_NODES + NB_NODES]; }   // the descriptor array is padded by NB_NODES entries
            for (; vj < n_tiles; vj += G) {
                const int64_t tj = tile_of(vj);
                const int64_t a_v0 = n0.val_off, a_v1 = n1.val_off, a_c0 = n0.col_off, a_c1 = n1.col_off;
                const int a_r0 = n0.row0, a_r1 = n1.row0;
                if (vj + G < n_tiles) { const int64_t t = tile_of(vj + G); n0 = nd[t * NB_NODES]; n1 = nd[t * NB_NODES + NB_NODES]; }
                mbar_wait(&bar_empty[stage], phase ^ 1u);
                const int64_t vs = a_v0 & ~(int64_t)1, cs = a_c0 & ~(int64_t)3;
                const int rs = a_r0 & ~1;
                const uint32_t vb = (uint32_t)(((a_v1 - vs + 1) & ~(int64_t)1) * 8);
                const uint32_t cb = (uint32_t)(((a_c1 - cs + 3) & ~(int64_t)3) * 4);
                const uint32_t rb = (uint32_t)(((a_r1 - rs + 1) & ~1) * 8);
                const uint32_t db = NB_NODES * (uint32_t)sizeof(NodeDesc);
                const bool has = a_v1 > a_v0;
                mbar_expect_tx(&bar_full[stage], db + (has ? vb + cb + NVEC * rb : 0u));
                tma_load_1d(s_nd + (size_t)stage * NB_NODES, nd + tj * NB_NODES, db, &bar_full[stage]);
                if (has) {
                    tma_load_1d_stream(s_val + (size_t)stage * cap_v, va + vs, vb, &bar_full[stage]);
                    if (cb) tma_load_1d(s_col + (size_t)stage * cap_c, ncol + cs, cb, &bar_full[stage]);   // empty when the dictionary covers the tile
                    double* sv = s_vec + (size_t)stage * NV1 * NB_VT;
                    if (MODE == 2) {
                        tma_load_1d(sv, alpha + rs, rb, &bar_full[stage]);
                        tma_load_1d(sv + NB_VT, inv_d + rs, rb, &bar_full[stage]);
                        tma_load_1d(sv + 2 * NB_VT, xe + rs, rb, &bar_full[stage]);
                        tma_load_1d(sv + 3 * NB_VT, y + rs, rb, &bar_full[stage]);
                    }
                    if (MODE == 3) tma_load_1d(sv, xa + rs, rb, &bar_full[stage]);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------ consumer warps: NB_NPW nodes each per tile -----------------
        // The warp's NB_NPW nodes are processed side by side: LPN = 32 / NB_NPW lanes per node, every lane belongs to one
        // node, so one instruction stream serves all of them (round 1 looped over the nodes with all 32 lanes: 3x the
        // instructions per node, per-node arrays that spilled, a butterfly per node).  Lane l of a node takes entries
        // l, l + LPN, ... of each of the node's rows; U entries per lane are in flight together.
        constexpr int LPN = 32 / NB_NPW;                 // lanes per node
        constexpr int U = NB_NPW == 1 ? 4 : 6;           // entries per lane and pass (128 / 96 entries per node and pass)
        const int grp = lane / LPN, l = lane % LPN;
        const int myr = l / (LPN / 4);                   // row of the node this lane owns after the reduction (3 = none)
        const int wl = warp % NB_WARPS;                  // warp inside its consumer group
        for (int64_t i = warp / NB_WARPS; blockIdx.x + i * G < n_tiles; i += NG) {     // i: position in this CTA's tile sequence
            const int stage = (int)(i % STAGES);
            const uint32_t phase = (uint32_t)((i / STAGES) & 1);
            // With an odd number of stages the groups alternate on a stage, and a parity wait only tells phases apart that
            // are at most one step from the barrier's current one: a group that runs ahead would take the completed phase of
            // tile i - 2 STAGES for its own while tile i - STAGES is still loading.  Waiting for the release of tile
            // i - STAGES first (by the other group; never more than one phase away) pins the "full" barrier to the phase of
            // tile i or the one after.
            if (NG > 1 && (STAGES & 1) && i >= STAGES) mbar_wait(&bar_empty[stage], phase ^ 1u);
            mbar_wait(&bar_full[stage], phase);
            const NodeDesc* snd = s_nd + (size_t)stage * NB_NODES;
            const NodeDesc d0 = snd[0];
            const NodeDesc d = snd[wl * NB_NPW + grp];
            const int nfree = d.len_nfree >> 24;
            const int L = nfree > 0 ? (d.len_nfree & 0xffff) : 0;
            int maxL = L;
            if (NB_NPW == 2) maxL = max(maxL, __shfl_xor_sync(0xffffffffu, maxL, 16));
            const double* sv = s_val + (size_t)stage * cap_v + (int)(d.val_off - (d0.val_off & ~(int64_t)1)) + l;
            // columns: the node's explicit list in the ring, or row0 + a relative list of the dictionary (bits 16..23: pattern id)
            const int pid = (d.len_nfree >> 16) & 0xff;
            const int* sc = pid ? s_dict + (pid - 1) * dict_stride + l
                                : s_col + (size_t)stage * cap_c + (int)(d.col_off - (d0.col_off & ~(int64_t)3)) + l;
            const int cadd = pid ? d.row0 : 0;
            const int L1 = nfree > 1 ? L : -1, L2 = nfree > 2 ? 2 * L : -1;     // offsets of rows 1, 2 in the value slice (-1: no such row)
            // rows without entries (ghost nodes of a domain decomposition) still get y = 0: PCG reads q on every local
            // row; the fused central-difference step leaves them alone (their u comes from the halo exchange)
            const bool owner = (l % (LPN / 4)) == 0 && myr < nfree && (L > 0 || MODE != 2);
            double e_al = 0.0, e_id = 0.0, e_x = 0.0, e_y = 0.0;
            if (NVEC > 0 && owner) {
                const double* svec = s_vec + (size_t)stage * NV1 * NB_VT + (d.row0 - (d0.row0 & ~1)) + myr;
                if (MODE == 2) { e_al = svec[0]; e_id = svec[NB_VT]; e_x = svec[2 * NB_VT]; e_y = svec[3 * NB_VT]; }
                if (MODE == 3 && L > 0) e_x = svec[0];      // a tile of ghost nodes only loads its descriptors: nothing to read, and
                                                            // 0 * (stale shared memory) would poison the dot product with a NaN
            }
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            for (int kb = 0; kb < maxL; kb += LPN * U) {
                int c[U];
                double xg[U], v0[U], v1[U], v2[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int k = kb + LPN * u;
                    const bool ok = k + l < L;
                    c[u] = 0; v0[u] = 0.0; v1[u] = 0.0; v2[u] = 0.0;
                    if (ok) {
                        c[u] = sc[k] + cadd;
                        v0[u] = sv[k];
                        if (L1 >= 0) v1[u] = sv[L1 + k];
                        if (L2 >= 0) v2[u] = sv[L2 + k];
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) xg[u] = __ldg(xa + c[u]);
                __syncwarp();      // scheduling fence: all gathers of the pass are issued before the first FMA
                // last pass: everything this warp needs from the stage sits in registers (the gathers could only be issued
                // once the column reads had returned, and shared-memory reads return in order) -- hand the stage back now, so
                // that its refill overlaps the gather latency and the arithmetic instead of following them
                if (kb + LPN * U >= maxL && lane == 0) mbar_arrive(&bar_empty[stage]);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    s0 += v0[u] * xg[u];
                    s1 += v1[u] * xg[u];
                    s2 += v2[u] * xg[u];
                }
            }
            if (maxL == 0) {                                 // nothing was read beyond the descriptors / epilogue operands
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[stage]);
            }

            // three-row butterfly inside the node's LPN lanes (a fourth, empty row keeps the 4-row pattern): two steps halve the
            // live rows per lane, the rest adds up one row per quarter; row r ends up in lanes [r LPN/4, (r+1) LPN/4) of the node
            const bool hi = (l & (LPN / 2)) != 0, hb = (l & (LPN / 4)) != 0;
            double k0 = hi ? s2 : s0, k1 = hi ? 0.0 : s1;
            const double g0 = hi ? s0 : s2, g1 = hi ? s1 : 0.0;
            k0 += __shfl_xor_sync(0xffffffffu, g0, LPN / 2);
            k1 += __shfl_xor_sync(0xffffffffu, g1, LPN / 2);
            double mine = hb ? k1 : k0;
            const double gq = hb ? k0 : k1;
            mine += __shfl_xor_sync(0xffffffffu, gq, LPN / 4);
#pragma unroll
            for (int o = LPN / 8; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
            if (owner) {
                const int64_t row = (int64_t)d.row0 + myr;
                if (MODE == 2) {
                    const double un = e_id * (-mine) + e_al * e_x - (e_al - 1.0) * e_y;
                    y[row] = un;
                    if (y2) y2[row] = (1.0 + lag) * un - lag * e_x;
                } else {
                    y[row] = mine;
                    if (MODE == 3) dot_acc += e_x * mine;
                }
            }
        }
    }
    if (MODE == 3) {
        if (warp < NG * NB_WARPS) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dot_acc += __shfl_down_sync(0xffffffffu, dot_acc, o);
            if (lane == 0) red[warp] = dot_acc;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < NG * NB_WARPS; ++w) s += red[w];
            partial[blockIdx.x] = s;
        }
    }
}

struct NodeCfg { int cap_v, cap_c, stages, npw, ng; size_t bytes, max_bytes; };
constexpr size_t NODE_SMEM_MAX = 112 * 1024;       // two CTAs per SM
constexpr size_t NODE_SMEM_MAX_1CTA = 224 * 1024;  // one CTA per SM (two consumer groups)

// ring configuration (dictionary not counted: `node_dict_room` sizes the dictionary into what is left)
bool node_cfg(sc_ctx* ctx, NodeCfg& c) {
    if (ctx->force_no_node || !ctx->d_nd || ctx->max_rl <= 0 || ctx->dim > 3) return false;
    // two nodes per consumer warp when the stages fit (short rows), one node per warp for long rows (hexa20, tetra10)
    for (int npw = 2; npw >= 1; --npw) {
        const int nodes = NB_WARPS * npw, vt = 3 * nodes + 8;
        c.cap_v = (nodes * 3 * ctx->max_rl + 2 + 15) & ~15;
        c.cap_c = (nodes * ctx->max_rl + 4 + 31) & ~31;
        const size_t per_stage = (size_t)c.cap_v * 8 + (size_t)c.cap_c * 4 + 4 * (size_t)vt * 8 + nodes * sizeof(NodeDesc);
        // preferred: one CTA per SM, two consumer groups on a ring of 3-6 stages that still leaves room for a full dictionary
        const size_t dict_full = (size_t)SC_DICT_MAX * ctx->max_rl * sizeof(int32_t);
        for (int st = 6; st >= 3 && !ctx->force_one_group; --st) {
            const size_t bytes = st * per_stage + 2 * st * sizeof(uint64_t) + 64;
            if (bytes + dict_full + 64 <= NODE_SMEM_MAX_1CTA) {
                c.stages = st; c.npw = npw; c.ng = 2; c.bytes = bytes; c.max_bytes = NODE_SMEM_MAX_1CTA;
                return true;
            }
        }
        const size_t budget = (npw == 2 ? 104 : 112) * 1024;
        for (int st = 4; st >= 2; --st) {
            const size_t bytes = st * per_stage + 2 * st * sizeof(uint64_t) + 64;
            if (bytes <= budget) { c.stages = st; c.npw = npw; c.ng = 1; c.bytes = bytes; c.max_bytes = NODE_SMEM_MAX; return true; }
        }
    }
    return false;
}

template <int MODE, int STAGES, int NPW, int NG>
int launch_node(sc_ctx* ctx, const NodeCfg& c, const double* va, const double* xa, double* y, const double* inv_d, const double* alpha,
                double* partial, unsigned* nblocks_out, const double* xe, double* y2, double g, int part) {
    constexpr int NODES = NB_WARPS * NPW;
    const int64_t all_tiles = (ctx->n_nodes + NODES - 1) / NODES;
    // part 0: every tile; 1: the tiles outside [ov_tile_lo, ov_tile_hi) (they hold every row the halo exchange sends);
    // 2: the interior tiles
    int64_t n_tiles = all_tiles, lo1 = 0, n1 = all_tiles, lo2 = 0;
    if (part == 1) { lo1 = 0; n1 = ctx->ov_tile_lo; lo2 = ctx->ov_tile_hi; n_tiles = n1 + (all_tiles - lo2); }
    if (part == 2) { lo1 = ctx->ov_tile_lo; n1 = ctx->ov_tile_hi - ctx->ov_tile_lo; n_tiles = n1; }
    if (n_tiles <= 0) return SC_OK;
    auto kern = k_spmv_node<MODE, STAGES, NPW, NG>;
    const size_t bytes = c.bytes + (size_t)ctx->n_dict * ctx->dict_stride * sizeof(int32_t);
    SC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    unsigned grid = (unsigned)std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * (NG == 1 ? 2 : 1));
    // the interior launch of the halo overlap leaves a few SMs to the NCCL send / receive kernels: a persistent CTA per SM
    // with ~200 KB of shared memory would otherwise keep them waiting until the step is over
    if (part == 2 && NG == 2 && grid > 16u) grid -= (unsigned)ctx->ov_spare_sms;
    if (grid == 0) grid = 1;
    if (nblocks_out) *nblocks_out = grid;
    kern<<<grid, 32 * (NG * NB_WARPS + 1), bytes, ctx->stream>>>(ctx->d_nd, ctx->d_ncol, va, xa, y, inv_d, alpha, partial, ctx->n_nodes,
                                                                  ctx->n_eq, n_tiles, c.cap_v, c.cap_c, ctx->d_dict, ctx->n_dict,
                                                                  ctx->dict_stride, xe, y2, g, lo1, n1, lo2);
    SC_CHECK_LAUNCH(ctx);
    return SC_OK;
}

template <int MODE>
int launch_mode(sc_ctx* ctx, const double* va, const double* xa, double* y, const double* inv_d, const double* alpha, double* partial,
                unsigned* nblocks_out, const double* xe = nullptr, double* y2 = nullptr, double g = 0.0, int part = 0) {
    NodeCfg c;
    if (!node_cfg(ctx, c)) return sc_fail(ctx, SC_ERR_UNSUPPORTED, "node-blocked SpMV not usable for this pattern");
#define SC_NODE_GO(ST, NPW, NG) return launch_node<MODE, ST, NPW, NG>(ctx, c, va, xa, y, inv_d, alpha, partial, nblocks_out, xe, y2, g, part)
    if (c.ng == 2) {
        if (c.npw == 2) { if (c.stages == 6) SC_NODE_GO(6, 2, 2); if (c.stages == 5) SC_NODE_GO(5, 2, 2); if (c.stages == 4) SC_NODE_GO(4, 2, 2); SC_NODE_GO(3, 2, 2); }
        if (c.stages == 6) SC_NODE_GO(6, 1, 2); if (c.stages == 5) SC_NODE_GO(5, 1, 2); if (c.stages == 4) SC_NODE_GO(4, 1, 2); SC_NODE_GO(3, 1, 2);
    }
    if (c.npw == 2) { if (c.stages == 4) SC_NODE_GO(4, 2, 1); if (c.stages == 3) SC_NODE_GO(3, 2, 1); SC_NODE_GO(2, 2, 1); }
    if (c.stages == 4) SC_NODE_GO(4, 1, 1); if (c.stages == 3) SC_NODE_GO(3, 1, 1); SC_NODE_GO(2, 1, 1);
#undef SC_NODE_GO
}

}  // namespace

bool la_node_usable(sc_ctx* ctx) {
    NodeCfg c;
    return node_cfg(ctx, c);
}
// shared memory (bytes) the column dictionary may take without changing the ring configuration or the two CTAs per SM
int64_t node_dict_room(sc_ctx* ctx) {
    NodeCfg c;
    const int keep = ctx->n_dict;
    ctx->n_dict = 0;
    const bool ok = node_cfg(ctx, c);
    ctx->n_dict = keep;
    if (!ok || c.bytes + 64 >= c.max_bytes) return 0;
    return (int64_t)(c.max_bytes - c.bytes - 64);
}
int la_node_spmv(sc_ctx* ctx, const double* vals, const double* x, double* y) {
    return launch_mode<0>(ctx, vals, x, y, nullptr, nullptr, nullptr, nullptr);
}
int la_node_cd_step(sc_ctx* ctx, const double* K, const double* w, const double* u, double* uprev_next, const double* inv_d,
                    const double* alpha, double g, double* w_next, int part) {
    return launch_mode<2>(ctx, K, w, uprev_next, inv_d, alpha, nullptr, nullptr, u, w_next, g, part);
}

namespace {
// tile of every row the halo exchange sends (rows -> node by bisection of the node row offsets)
__global__ void k_flag_send_tiles(const int64_t* __restrict__ send_idx, int64_t ns, const int64_t* __restrict__ node_row0,
                                  int64_t n_nodes, int nodes_per_tile, unsigned char* __restrict__ flag) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= ns) return;
    const int64_t row = send_idx[t];
    int64_t lo = 0, hi = n_nodes;                      // last node with node_row0 <= row
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (node_row0[mid] <= row) lo = mid; else hi = mid;
    }
    flag[lo / nodes_per_tile] = 1;                     // every writer stores the same byte
}
}  // namespace

// Interior of the halo overlap: the longest run of tiles [lo, hi) that holds no row of the send lists (for slab / RCB
// partitions of a sorted mesh the interface rows sit at the two ends of the numbering).  Nodes of those tiles are not
// adjacent to ghost nodes either (adjacency is symmetric: a node next to a ghost is needed by the ghost's owner), so their
// products neither produce values the exchange sends nor read values it delivers.  Returns false when the run covers less
// than half of the tiles (then the step and the exchange stay serial).
int la_node_overlap_plan(sc_ctx* ctx) {
    if (ctx->ov_planned) return SC_OK;
    ctx->ov_planned = true;
    ctx->ov_ok = false;
    NodeCfg c;
    if (ctx->world <= 1 || ctx->no_overlap || !node_cfg(ctx, c) || !ctx->d_node_row0) return SC_OK;
    const int nodes = NB_WARPS * c.npw;
    const int64_t n_tiles = (ctx->n_nodes + nodes - 1) / nodes;
    const int64_t ns = ctx->send_ptr.empty() ? 0 : ctx->send_ptr.back();
    if (n_tiles < 64 || ns == 0) return SC_OK;
    unsigned char* d_flag = nullptr;
    SC_TRY(sc_alloc(ctx, &d_flag, (size_t)n_tiles));
    std::vector<unsigned char> flag((size_t)n_tiles);
    int rc = SC_OK;
    auto body = [&]() -> int {
        SC_CUDA(ctx, cudaMemsetAsync(d_flag, 0, (size_t)n_tiles, ctx->stream));
        k_flag_send_tiles<<<(unsigned)((ns + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_send_idx, ns, ctx->d_node_row0, ctx->n_nodes, nodes, d_flag);
        SC_CHECK_LAUNCH(ctx);
        SC_CUDA(ctx, cudaMemcpyAsync(flag.data(), d_flag, (size_t)n_tiles, cudaMemcpyDeviceToHost, ctx->stream));
        SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return SC_OK;
    };
    rc = body();
    sc_free(&d_flag);
    SC_TRY(rc);
    int64_t best_lo = 0, best_hi = 0, run = -1;
    for (int64_t t = 0; t <= n_tiles; ++t) {
        const bool set = t == n_tiles || flag[(size_t)t] != 0;
        if (!set && run < 0) run = t;
        if (set && run >= 0) {
            if (t - run > best_hi - best_lo) { best_lo = run; best_hi = t; }
            run = -1;
        }
    }
    if (best_hi - best_lo < n_tiles / 2) return SC_OK;
    ctx->ov_tile_lo = best_lo; ctx->ov_tile_hi = best_hi;
    // first rows of the two cut tiles (for the range filter of the sparse load kernel)
    int64_t h[2] = {0, 0};
    const int64_t n_lo = std::min<int64_t>(best_lo * nodes, ctx->n_nodes), n_hi = std::min<int64_t>(best_hi * nodes, ctx->n_nodes);
    SC_CUDA(ctx, cudaMemcpy(&h[0], ctx->d_node_row0 + n_lo, sizeof(int64_t), cudaMemcpyDeviceToHost));
    SC_CUDA(ctx, cudaMemcpy(&h[1], ctx->d_node_row0 + n_hi, sizeof(int64_t), cudaMemcpyDeviceToHost));
    ctx->ov_row_lo = h[0]; ctx->ov_row_hi = h[1];
    ctx->ov_ok = true;
    return SC_OK;
}
int la_node_spmv_dot(sc_ctx* ctx, const double* vals, const double* p, double* q, double* partial, unsigned* nblocks) {
    return launch_mode<3>(ctx, vals, p, q, nullptr, nullptr, partial, nblocks);
}
// bytes one fused central-difference launch has to move with this format: values, node column lists, node descriptors,
// and five vector passes (alpha, inv_d, u gathered once, u_prev read, u_next written); with stiffness-proportional
// damping (g != 0) two more: the gather vector w is a separate read and w_next is written
int64_t la_node_step_bytes(sc_ctx* ctx, bool lagged) {
    return ctx->nnz * 8 + ctx->ncol_total * 4 + ctx->n_nodes * (int64_t)sizeof(NodeDesc) + ctx->n_eq * (lagged ? 56 : 40);
}
