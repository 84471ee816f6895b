// Element integration + deterministic assembly in one kernel (no element-matrix scratch, no float atomics).
//
// Replaces GenerateMatrix.generate_stiffness_and_mass (scatter/system_matrix.py:35-121):
//   per element   VolumeElement/SurfaceElement.generate + compute_stiffness + compute_mass
//                 (scatter/discretisation.py:83-222, :290-417), D from material_models.py:5-43
//   scatter       k_dict[i,k] += Ke[j,l]; mass_dict[i,k] += Me[j,l] in element order (system_matrix.py:98-103)
//
// Formulation ("row gather"): every node owns DIM consecutive matrix rows and a row only receives contributions from
// the elements touching its node.  A block of consecutive nodes therefore produces its CSR rows alone: it walks the
// elements of each node in ascending element id -- the order in which the reference adds them to a slot -- evaluates
// only the DIM x (NNE*DIM) row block of Ke that belongs to the node, and sums per slot in that order.  Nobody else
// writes those rows, so the result is reproducible bit for bit.  Three kernels implement it: `k_elem_records` +
// `k_assemble_tma` (default: Jacobians once per element into HBM records; persistent CTAs fed by TMA bulk copies of
// packed block descriptors and records; row block in registers, one, two or five lanes per (node, element) pair),
// `k_assemble_blk` (round 1, same arithmetic and bits without any scratch: Jacobian set-up shared by the pairs of a
// block) and `k_assemble` (one warp per node, shared-memory staging; very high node valences).
//
// Isotropic elasticity lets the row block be formed without B or D:
//   K[(a,i),(b,j)] = sum_g w_g detJ_g ( lam dNa_i dNb_j + mu dNa_j dNb_i + delta_ij mu dNa.dNb )
//   M[(a,i),(b,j)] = delta_ij rho sum_g w_g detJ_g Na Nb          (consistent mass; zeros kept in the pattern)
// which equals B^T D B with the reference's Voigt ordering; detJ is used signed (discretisation.py:126).
#include <algorithm>
#include <cstdlib>
#include <cub/device/device_scan.cuh>
#include "common.h"
#include "tma.h"

#ifndef SC_BLK_UG
#define SC_BLK_UG 2
#endif

namespace {

struct AsmParams {
    const double* xyz; const int32_t* conn; const int32_t* eq;
    const double *E, *nu, *rho;
    const int64_t* n2e_ptr; const int32_t* n2e;
    const int64_t* nbr_ptr; const int32_t* nbr; const uint16_t* nbr_off; const uint8_t* nbr_free;
    const int32_t* node_rl; const int64_t* node_row0; const int64_t* rowptr;
    const double *tabN, *tabdN, *tabw;
    const uint8_t *pair_pos, *pair_al;
    double *K, *M, *Ml;
    int64_t n_nodes;
    int max_rl, max_nbr;
};

template <int DIM>
__device__ __forceinline__ void invert(const double* J, double* inv, double& det);

template <>
__device__ __forceinline__ void invert<2>(const double* J, double* inv, double& det) {
    det = J[0] * J[3] - J[1] * J[2];
    double r = 1.0 / det;
    inv[0] = J[3] * r; inv[1] = -J[1] * r; inv[2] = -J[2] * r; inv[3] = J[0] * r;
}
template <>
__device__ __forceinline__ void invert<3>(const double* J, double* inv, double& det) {
    double a = J[0], b = J[1], c = J[2], d = J[3], e = J[4], f = J[5], g = J[6], h = J[7], i = J[8];
    double c0 = e * i - f * h, c1 = f * g - d * i, c2 = d * h - e * g;
    det = a * c0 + b * c1 + c * c2;
    double r = 1.0 / det;
    inv[0] = c0 * r; inv[1] = (c * h - b * i) * r; inv[2] = (b * f - c * e) * r;
    inv[3] = c1 * r; inv[4] = (a * i - c * g) * r; inv[5] = (c * d - a * f) * r;
    inv[6] = c2 * r; inv[7] = (b * g - a * h) * r; inv[8] = (a * e - b * d) * r;
}

template <int NNE, int DIM, int NGP>
__global__ void k_assemble(AsmParams p, int warps) {
    constexpr int DD = DIM * DIM;
    constexpr int JS = DD + 1;                 // inverse Jacobian + detJ*w per Gauss point
    extern __shared__ double smem[];
    double* sN = smem;                          // [NGP*NNE]
    double* sdN = sN + NGP * NNE;               // [NGP*NNE*DIM]
    double* sw = sdN + NGP * NNE * DIM;         // [NGP]
    const int per_warp = NNE * DIM + NGP * JS + NGP * NNE * DIM + p.max_nbr + (NNE + 1) / 2;
    double* wbase = sw + NGP;
    double* Kst = wbase + (size_t)warps * per_warp;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* xs = wbase + (size_t)warp * per_warp;       // [NNE*DIM]
    double* sJ = xs + NNE * DIM;                         // [NGP*JS]
    double* dNg = sJ + NGP * JS;                         // [NGP*NNE*DIM]
    double* mnode = dNg + NGP * NNE * DIM;               // [max_nbr]
    int* cb = reinterpret_cast<int*>(mnode + p.max_nbr); // [NNE]

    for (int t = threadIdx.x; t < NGP * NNE; t += blockDim.x) sN[t] = p.tabN[t];
    for (int t = threadIdx.x; t < NGP * NNE * DIM; t += blockDim.x) sdN[t] = p.tabdN[t];
    for (int t = threadIdx.x; t < NGP; t += blockDim.x) sw[t] = p.tabw[t];

    const int64_t a0 = (int64_t)blockIdx.x * warps;
    const int64_t a1 = min(a0 + warps, p.n_nodes);
    const int64_t R0 = p.node_row0[a0], R1 = p.node_row0[a1];
    const int64_t base = p.rowptr[R0];
    const int span = (int)(p.rowptr[R1] - base);
    for (int t = threadIdx.x; t < span; t += blockDim.x) Kst[t] = 0.0;
    __syncthreads();

    const int64_t a = a0 + warp;
    const bool live = (a < p.n_nodes) && (p.node_rl[a] > 0);
    int64_t nb0 = 0;
    int nn_a = 0;
    int rowoff[DIM];
    if (live) {
        nb0 = p.nbr_ptr[a];
        nn_a = (int)(p.nbr_ptr[a + 1] - nb0);
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            int r = p.eq[a * DIM + i];
            rowoff[i] = r >= 0 ? (int)(p.rowptr[r] - base) : -1;
        }
        for (int t = lane; t < nn_a; t += 32) mnode[t] = 0.0;
        __syncwarp();

        for (int64_t k = p.n2e_ptr[a]; k < p.n2e_ptr[a + 1]; ++k) {
            const int e = p.n2e[k];
            if (lane < NNE) {
                int c = p.conn[(int64_t)e * NNE + lane];
                cb[lane] = c;
#pragma unroll
                for (int d = 0; d < DIM; ++d) xs[lane * DIM + d] = p.xyz[(int64_t)c * 3 + d];
            }
            __syncwarp();
            int al = 0;
#pragma unroll
            for (int b = 0; b < NNE; ++b) if (cb[b] == (int)a) al = b;
            const double E = p.E[e], nu = p.nu[e], rho = p.rho[e];
            const double lam = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
            const double mu = E / (2.0 * (1.0 + nu));

            // Jacobian J[g][d][k] = sum_b dN[g][b][d] x[b][k]   (discretisation.py:113-116)
            for (int it = lane; it < NGP * DD; it += 32) {
                const int g = it / DD, r = it % DD, d = r / DIM, kk = r % DIM;
                double s = 0.0;
#pragma unroll
                for (int b = 0; b < NNE; ++b) s += sdN[(g * NNE + b) * DIM + d] * xs[b * DIM + kk];
                sJ[g * JS + r] = s;
            }
            __syncwarp();
            for (int g = lane; g < NGP; g += 32) {
                double J[DD], inv[DD], det;
#pragma unroll
                for (int r = 0; r < DD; ++r) J[r] = sJ[g * JS + r];
                invert<DIM>(J, inv, det);
#pragma unroll
                for (int r = 0; r < DD; ++r) sJ[g * JS + r] = inv[r];
                sJ[g * JS + DD] = det * sw[g];
            }
            __syncwarp();
            // global derivatives dNg[g][b][k] = sum_d dN[g][b][d] Jinv[k][d]    (dN . inv(J^T), discretisation.py:128)
            for (int it = lane; it < NGP * NNE * DIM; it += 32) {
                const int g = it / (NNE * DIM), kk = it % DIM;
                const int gb = it / DIM;   // g*NNE + b
                double s = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; ++d) s += sdN[gb * DIM + d] * sJ[g * JS + kk * DIM + d];
                dNg[it] = s;
            }
            __syncwarp();
            // row block of node `al`: item = (b, i) -> DIM entries (j)
            for (int it = lane; it < NNE * DIM; it += 32) {
                const int b = it / DIM, i = it % DIM;
                double acc[DIM];
#pragma unroll
                for (int j = 0; j < DIM; ++j) acc[j] = 0.0;
                double macc = 0.0;
#pragma unroll 4
                for (int g = 0; g < NGP; ++g) {
                    const double wj = sJ[g * JS + DD];
                    const double* ga = dNg + (g * NNE + al) * DIM;
                    const double* gbv = dNg + (g * NNE + b) * DIM;
                    double s = 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; ++d) s += ga[d] * gbv[d];
                    const double lga = lam * ga[i], mgb = mu * gbv[i];
#pragma unroll
                    for (int j = 0; j < DIM; ++j) {
                        double t = lga * gbv[j] + mgb * ga[j];
                        if (j == i) t += mu * s;
                        acc[j] += wj * t;
                    }
                    if (i == 0) macc += wj * sN[g * NNE + al] * sN[g * NNE + b];
                }
                // slot lookup: position of node cb[b] in the (ascending) neighbour list of a
                const int nodeb = cb[b];
                int lo = 0, hi = nn_a;
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (p.nbr[nb0 + mid] < nodeb) lo = mid + 1; else hi = mid;
                }
                if (rowoff[i] >= 0) {
                    int o = rowoff[i] + p.nbr_off[nb0 + lo];
#pragma unroll
                    for (int j = 0; j < DIM; ++j)
                        if (p.eq[(int64_t)nodeb * DIM + j] >= 0) Kst[o++] += acc[j];
                }
                if (i == 0) mnode[lo] += rho * macc;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (p.K)
        for (int t = threadIdx.x; t < span; t += blockDim.x) p.K[base + t] = Kst[t];
    if (live) {
        if (p.M) {
            for (int q = lane; q < nn_a; q += 32) {
                const int nodeb = p.nbr[nb0 + q];
                const double m = mnode[q];
                const int off = p.nbr_off[nb0 + q];
#pragma unroll
                for (int i = 0; i < DIM; ++i) {
                    if (rowoff[i] < 0) continue;
                    int64_t o = base + rowoff[i] + off;
#pragma unroll
                    for (int j = 0; j < DIM; ++j)
                        if (p.eq[(int64_t)nodeb * DIM + j] >= 0) p.M[o++] = (i == j) ? m : 0.0;
                }
            }
        }
        if (p.Ml) {
            // row sum of the consistent mass: columns (b,i) that exist, neighbour order, fixed-shape reduction
#pragma unroll
            for (int i = 0; i < DIM; ++i) {
                if (rowoff[i] < 0) continue;
                double s = 0.0;
                for (int q = lane; q < nn_a; q += 32)
                    if (p.eq[(int64_t)p.nbr[nb0 + q] * DIM + i] >= 0) s += mnode[q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
                if (lane == 0) p.Ml[p.eq[a * DIM + i]] = s;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Register-resident variant for element types whose DIM x (NNE*DIM) row block fits in registers (<= 72 doubles:
// tri3, tri6, quad4, quad8, tetra4, hexa8).  One *lane* owns one (node, element) pair: it evaluates the Jacobians and
// the complete row block of the node in that element in registers (integration tables come from constant memory with
// warp-uniform addresses), so the FP64 pipe, not shared memory, is the limit (ncu of the warp-per-node kernel above:
// 90 % LSU/shared-memory utilisation, 28 % FP64).  The pairs of one node are then added to the shared-memory image of
// the node's CSR rows in ascending element order (one barrier-separated round per rank), which keeps the summation
// order of the reference (system_matrix.py:98-103) and needs no atomics.
__constant__ double c_tabN[SC_MAX_GP * SC_MAX_NNE];
__constant__ double c_tabdN[SC_MAX_GP * SC_MAX_NNE * 3];

// Block-level element sharing (round 1; now the fallback that needs no scratch memory).  Repeating the Jacobian set-up in every lane of every
// (node, element) pair costs 16 evaluations per hexa8 element and Gauss point inside one block alone (the previous
// generation of this kernel did).  Here a block first lists the *distinct* elements of its pairs (consecutive nodes share most of theirs),
// evaluates J^-1 and detJ*w once per (element, Gauss point) -- a few threads per element, coordinates in registers --
// and parks the 10 numbers in shared memory.  The pair lanes then only accumulate the gradient products
//     G_ab[i][j] = sum_g detJ w  dNa_i dNb_j          (9 FMA per node block and Gauss point, tables from __constant__)
// and apply the material law once at the end:  K_ab[i][j] = lam G[i][j] + mu G[j][i] + delta_ij mu tr G.
// All pairs park their row blocks in a shared-memory staging area and one thread per (node, neighbour) block adds the
// staged contributions in ascending element id (the reference's accumulation order), so repeated runs are bit-identical.
template <int NNE, int DIM, int NGP, int TPB, int LPP, int MINB>
__global__ void __launch_bounds__(TPB, MINB) k_assemble_blk(AsmParams p, int npb) {
    constexpr int DD = DIM * DIM, ND = NNE * DIM;
    constexpr int NBB = NNE / LPP;                       // node blocks per lane
    constexpr int PPB = TPB / LPP;                       // pairs per block
    constexpr int SST = DIM * ND + 1;                    // stride of one pair's row block in the staging area (odd)
    constexpr int GC = NGP <= 9 ? NGP : 9;               // Gauss points per chunk (27 = 3 x 9)
    constexpr int IST = DD + 1;                          // J^-1 and detJ*w
    constexpr int UST = GC * IST + 1;                    // per-element stride (odd: no bank conflicts across elements)
    constexpr int UGB = SC_BLK_UG;                       // unroll factor of the Gauss-point loop of the pair lanes
    constexpr int HBITS = 8, HSIZE = 1 << HBITS;         // hash table of the block's elements (load <= 1/2)
    static_assert(NNE % LPP == 0 && NGP % GC == 0 && (TPB / 32) % LPP == 0 && PPB * 2 <= HSIZE, "unsupported split");
    extern __shared__ double smem[];
    double* sinv = smem;                                 // [PPB][UST]  set-up of the distinct elements (phase 1)
    double* stage = smem;                                // [PPB][SST]  row block of every pair        (phase 2)
    double* stage_m = stage + (size_t)PPB * SST;         // [PPB][NNE]
    constexpr size_t SZ_STAGE = (size_t)PPB * SST + (size_t)PPB * NNE, SZ_INV = (size_t)PPB * UST;
    constexpr size_t REGION_A = SZ_STAGE > SZ_INV ? SZ_STAGE : SZ_INV;
    double* sdN = smem + REGION_A;                       // [NGP*NNE*DIM] table copies for lane-dependent rows
    double* sN = sdN + NGP * NNE * DIM;                  // [NGP*NNE]
    double* sw = sN + NGP * NNE;                         // [NGP]
    double* s_mitem = sw + NGP;                          // [npb*max_nbr]
    long long* s_rowbase = reinterpret_cast<long long*>(s_mitem + (size_t)npb * p.max_nbr);   // [npb*DIM]
    int* s_ptr = reinterpret_cast<int*>(s_rowbase + (size_t)npb * DIM);   // [npb+1]
    int* s_nptr = s_ptr + npb + 1;                       // [npb+1]
    int* s_rl = s_nptr + npb + 1;                        // [npb] row length of the block's nodes (0: ghost / fully fixed)
    int* s_conn = s_rl + npb;                            // [PPB][NNE] connectivity of the distinct elements
    int* s_wcnt = s_conn + PPB * NNE;                    // [1] number of distinct elements
    int* s_hkey = s_wcnt + 1;                            // [HSIZE] element id or -1
    int* s_hval = s_hkey + HSIZE;                        // [HSIZE] index in the distinct list
    double* s_mat = reinterpret_cast<double*>(s_hval + HSIZE + ((npb + 1) & 1));   // [PPB][3] lambda, mu, rho (8-byte aligned)
    unsigned char* s_inv = reinterpret_cast<unsigned char*>(s_mat + PPB * 3);      // [PPB][max_nbr]

    // pair lanes: warp w serves pairs (w / LPP) * 32 + lane and node blocks [half * NBB, half * NBB + NBB), half = w % LPP
    // (warp-uniform, so the table rows of the node blocks come from constant memory)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = (warp / LPP) * 32 + lane, half = warp % LPP;
    const int64_t a0 = (int64_t)blockIdx.x * npb;
    const int64_t a1 = min(a0 + npb, p.n_nodes);
    const int nbn = (int)(a1 - a0);
    const int64_t P0 = p.n2e_ptr[a0];
    const int64_t nbr0 = p.nbr_ptr[a0];
    const int npairs = (int)(p.n2e_ptr[a1] - P0);        // <= PPB by construction of npb
    // everything a pair needs from global memory is requested before the first barrier
    int e_k = -1, al = 0;
    int cn[NNE];
    unsigned char pos[NNE];
    double lam = 0.0, mu = 0.0, rho = 0.0;
    if (k < npairs) {
        e_k = p.n2e[P0 + k];
        al = p.pair_al[P0 + k];
        if (half == 0) {
#pragma unroll
            for (int b = 0; b < NNE; ++b) pos[b] = p.pair_pos[(P0 + k) * NNE + b];
#pragma unroll
            for (int b = 0; b < NNE; ++b) cn[b] = p.conn[(int64_t)e_k * NNE + b];
            const double E = p.E[e_k], nu = p.nu[e_k];
            rho = p.rho[e_k];
            lam = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
            mu = E / (2.0 * (1.0 + nu));
        }
    }
    for (int t = tid; t <= nbn; t += TPB) {
        s_ptr[t] = (int)(p.n2e_ptr[a0 + t] - P0);
        s_nptr[t] = (int)(p.nbr_ptr[a0 + t] - nbr0);
        if (t < nbn) s_rl[t] = p.node_rl[a0 + t];
    }
    for (int t = tid; t < NGP * NNE * DIM; t += TPB) sdN[t] = p.tabdN[t];
    for (int t = tid; t < NGP * NNE; t += TPB) sN[t] = p.tabN[t];
    for (int t = tid; t < NGP; t += TPB) sw[t] = p.tabw[t];
    for (int t = tid; t < (PPB * p.max_nbr + 3) / 4; t += TPB) reinterpret_cast<unsigned*>(s_inv)[t] = 0xffffffffu;
    for (int t = tid; t < HSIZE; t += TPB) s_hkey[t] = -1;
    if (tid == 0) s_wcnt[0] = 0;
    // first slot of the block's rows: a two-level load chain that is only needed in phase 2 -> stored after the barrier
    long long rb_reg[DIM];                               // nbn <= PPB <= TPB, so DIM strided passes cover nbn * DIM rows
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
        const int t = tid + c * TPB;
        rb_reg[c] = -1;
        if (t < nbn * DIM) {
            const int rr = p.eq[a0 * DIM + t];
            if (rr >= 0 && p.node_rl[a0 + t / DIM] > 0) rb_reg[c] = (long long)p.rowptr[rr];
        }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < DIM; ++c)
        if (tid + c * TPB < nbn * DIM) s_rowbase[tid + c * TPB] = rb_reg[c];
    const int n_items = s_nptr[nbn];

    // ---- distinct elements of the block -----------------------------------------------------------------------------
    if (e_k >= 0) {
        bool any_empty = false;                          // ghost nodes of a domain decomposition (and fully fixed nodes) own no rows
        for (int t = 0; t < nbn; ++t) any_empty |= s_rl[t] <= 0;
        if (any_empty) {
            int lo = 0, hi = nbn;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_ptr[mid] <= k) lo = mid; else hi = mid;
            }
            if (s_rl[lo] <= 0) e_k = -1;
        }
    }
    const bool valid = e_k >= 0;
    // open-addressing table keyed by element id; which pair wins a slot only decides where the element's set-up is
    // parked, never a summation order, so the integer atomics do not affect the result
    unsigned h = ((unsigned)e_k * 2654435761u) >> (32 - HBITS);
    if (valid && half == 0) {
        for (;;) {
            const int old = atomicCAS(&s_hkey[h], -1, e_k);
            if (old == -1) {                             // first pair of this element in the block
                const int u = atomicAdd(&s_wcnt[0], 1);
#pragma unroll
                for (int b = 0; b < NNE; ++b) s_conn[u * NNE + b] = cn[b];
                s_mat[u * 3 + 0] = lam; s_mat[u * 3 + 1] = mu; s_mat[u * 3 + 2] = rho;
                s_hval[h] = u;
                break;
            }
            if (old == e_k) break;
            h = (h + 1) & (HSIZE - 1);
        }
#pragma unroll
        for (int b = 0; b < NNE; ++b) s_inv[k * p.max_nbr + pos[b]] = (unsigned char)b;
    }
    __syncthreads();
    int ui = 0;
    if (valid) {
        while (s_hkey[h] != e_k) h = (h + 1) & (HSIZE - 1);
        ui = s_hval[h];
    }
    const int U = s_wcnt[0];

    // ---- phase 1 ----------------------------------------------------------------------------------------------------
    double acc[DIM][NBB * DIM];
    double mab[NBB];
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int c = 0; c < NBB * DIM; ++c) acc[i][c] = 0.0;
#pragma unroll
    for (int b = 0; b < NBB; ++b) mab[b] = 0.0;
    // set-up threads: tpe threads per distinct element, each takes the Gauss points part, part + tpe, ... of a chunk
    const int tpe = U > 0 ? min(GC, TPB / U) : 1;
    const int su = tid / tpe, part = tid % tpe;
    const bool setup = su < U;
    double xe[ND];
    if (setup) {
#pragma unroll
        for (int b = 0; b < NNE; ++b) {
            const int c = s_conn[su * NNE + b];
#pragma unroll
            for (int d = 0; d < DIM; ++d) xe[b * DIM + d] = p.xyz[(int64_t)c * 3 + d];
        }
    }
#pragma unroll 1
    for (int g0 = 0; g0 < NGP; g0 += GC) {
        if (setup) {
            for (int gl = part; gl < GC; gl += tpe) {
                const int g = g0 + gl;
                double J[DD], inv[DD], det;
#pragma unroll
                for (int r = 0; r < DD; ++r) J[r] = 0.0;
#pragma unroll
                for (int b = 0; b < NNE; ++b)
#pragma unroll
                    for (int d = 0; d < DIM; ++d) {
                        const double dn = sdN[(g * NNE + b) * DIM + d];
#pragma unroll
                        for (int kk = 0; kk < DIM; ++kk) J[d * DIM + kk] += dn * xe[b * DIM + kk];
                    }
                invert<DIM>(J, inv, det);
                double* o = sinv + (size_t)su * UST + gl * IST;
#pragma unroll
                for (int r = 0; r < DD; ++r) o[r] = inv[r];
                o[DD] = det * sw[g];
            }
        }
        __syncthreads();
        if (valid) {
#pragma unroll UGB
            for (int gl = 0; gl < GC; ++gl) {
                const int g = g0 + gl;
                const double* si = sinv + (size_t)ui * UST + gl * IST;
                double inv[DD];
#pragma unroll
                for (int r = 0; r < DD; ++r) inv[r] = si[r];
                const double wj = si[DD];
                double wga[DIM];
#pragma unroll
                for (int kk = 0; kk < DIM; ++kk) {
                    double s = 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; ++d) s += sdN[(g * NNE + al) * DIM + d] * inv[kk * DIM + d];
                    wga[kk] = wj * s;
                }
                const double wna = wj * sN[g * NNE + al];
#pragma unroll
                for (int bb = 0; bb < NBB; ++bb) {
                    const double* dnb = c_tabdN + (g * NNE + half * NBB + bb) * DIM;
                    double gb[DIM];
#pragma unroll
                    for (int kk = 0; kk < DIM; ++kk) {
                        double s = 0.0;
#pragma unroll
                        for (int d = 0; d < DIM; ++d) s += dnb[d] * inv[kk * DIM + d];
                        gb[kk] = s;
                    }
#pragma unroll
                    for (int i = 0; i < DIM; ++i)
#pragma unroll
                        for (int j = 0; j < DIM; ++j) acc[i][bb * DIM + j] += wga[i] * gb[j];
                    mab[bb] += wna * c_tabN[g * NNE + half * NBB + bb];
                }
            }
        }
        __syncthreads();
    }
    // material law on the accumulated gradient products; region A becomes the staging area (everybody passed the
    // barrier that ends the last chunk)
    if (valid) {
        lam = s_mat[ui * 3 + 0]; mu = s_mat[ui * 3 + 1]; rho = s_mat[ui * 3 + 2];
#pragma unroll
        for (int bb = 0; bb < NBB; ++bb) {
            double tr = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; ++d) tr += acc[d][bb * DIM + d];
            tr *= mu;
#pragma unroll
            for (int i = 0; i < DIM; ++i)
#pragma unroll
                for (int j = 0; j < DIM; ++j) {
                    double t = lam * acc[i][bb * DIM + j] + mu * acc[j][bb * DIM + i];
                    if (i == j) t += tr;
                    stage[(size_t)k * SST + i * ND + (half * NBB + bb) * DIM + j] = t;
                }
            stage_m[k * NNE + half * NBB + bb] = rho * mab[bb];
        }
    }
    __syncthreads();

    // ---- phase 2: one thread per (node, neighbour) item = one DIM x DIM block of the global matrix; it adds the staged
    //      contributions of the node's elements in ascending element id (the reference's summation order) -------------
    for (int q = tid; q < n_items; q += TPB) {
        int lo = 0, hi = nbn;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_nptr[mid] <= q) lo = mid; else hi = mid;
        }
        const int n = lo, pidx = q - s_nptr[n];
        const int off = p.nbr_off[nbr0 + q];             // requested before the summation loop hides their latency
        const int fmask = p.nbr_free[nbr0 + q];
        double blk[DIM][DIM];
#pragma unroll
        for (int i = 0; i < DIM; ++i)
#pragma unroll
            for (int j = 0; j < DIM; ++j) blk[i][j] = 0.0;
        double m = 0.0;
        const int pr1 = s_ptr[n + 1];
        for (int pr0 = s_ptr[n]; pr0 < pr1; pr0 += 8) {
            int bs[8];                                   // slot lookups of eight pairs issued together
#pragma unroll
            for (int c = 0; c < 8; ++c) bs[c] = (pr0 + c < pr1) ? s_inv[(pr0 + c) * p.max_nbr + pidx] : 0xff;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int b = bs[c];
                if (b == 0xff) continue;
                const double* sp = stage + (size_t)(pr0 + c) * SST + b * DIM;
#pragma unroll
                for (int i = 0; i < DIM; ++i)
#pragma unroll
                    for (int j = 0; j < DIM; ++j) blk[i][j] += sp[i * ND + j];
                m += stage_m[(pr0 + c) * NNE + b];
            }
        }
        s_mitem[q] = m;
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            const long long rb = s_rowbase[n * DIM + i];
            if (rb < 0) continue;
            int64_t o = rb + off;
#pragma unroll
            for (int j = 0; j < DIM; ++j) {
                if (!(fmask & (1 << j))) continue;
                if (p.K) p.K[o] = blk[i][j];
                if (p.M) p.M[o] = (i == j) ? m : 0.0;
                ++o;
            }
        }
    }
    if (p.Ml) {
        __syncthreads();
        for (int t = tid; t < nbn * DIM; t += TPB) {
            const int n = t / DIM, i = t % DIM;
            if (s_rowbase[t] < 0) continue;
            double s = 0.0;
            for (int q = s_nptr[n]; q < s_nptr[n + 1]; ++q)
                if (p.nbr_free[nbr0 + q] & (1 << i)) s += s_mitem[q];
            p.Ml[p.eq[(a0 + n) * DIM + i]] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Record-fed generations (round 2; default: k_assemble_tma below).  ncu of k_assemble_blk: 35 % of the samples in the index
// prologue and the element de-duplication, every element's Jacobians evaluated in the ~4.5 blocks that see it, six
// barrier-separated phases.  Now
//   1. k_elem_records evaluates every element ONCE: J^-1 and detJ*w per Gauss point plus lambda, mu, rho -> one record of
//      REC doubles per element in HBM (hexa8, 2x2x2 points: 688 B, written coalesced through shared memory);
//   2. the distinct elements of every node block and each pair's index into that list are structure, not values: they
//      are listed once with the pattern (k_blk_desc: ascending element id, no atomics) and packed with every other index
//      of the block into one contiguous descriptor (k_blk_pack);
//   3. the row-gather kernel fetches descriptor and records with bulk copies (TMA, cp.async.bulk) straight into shared
//      memory, completion counted on mbarriers; the pair lanes run all Gauss points without a barrier, reading each
//      point's ten numbers with 16-byte shared loads (the record stride is = 2 mod 4 doubles, so the eight lanes of a
//      quarter warp never share a bank).
// The arithmetic per pair and the summation order are those of k_assemble_blk: both kernels give the same bits.
__host__ __device__ constexpr int rec_point_stride(int dim) { return (dim * dim + 1 + 1) & ~1; }            // J^-1, detJ*w (+ pad): even
__host__ __device__ constexpr int rec_stride(int dim, int ngp) {                                            // + lambda, mu, rho; = 2 (mod 4)
    int r = ngp * rec_point_stride(dim) + 3;
    while (r % 4 != 2) ++r;
    return r;
}

template <int NNE, int DIM, int NGP>
__global__ void __launch_bounds__(128)
k_elem_records(const double* __restrict__ xyz, const int32_t* __restrict__ conn, const double* __restrict__ E,
               const double* __restrict__ nu, const double* __restrict__ rho, const double* __restrict__ tabdN,
               const double* __restrict__ tabw, int64_t n_elem, double* __restrict__ out) {
    constexpr int DD = DIM * DIM, ND = NNE * DIM;
    constexpr int ISTP = rec_point_stride(DIM), REC = rec_stride(DIM, NGP);
    constexpr int EPB = 128 / NGP;                       // elements per group, one thread per (element, Gauss point)
    constexpr bool STAGE_XE = NGP >= 4;                  // the threads of an element fetch its coordinates together
    __shared__ double s_rec[EPB * REC];
    constexpr int SDS = ND | 1;                          // odd row stride of the table copy: the Gauss-point threads of a warp
    __shared__ double sdN[NGP * SDS];                    // read different rows -- with the natural stride (24 doubles for hexa8) they
    __shared__ double s_xe[STAGE_XE ? EPB * ND : 1];     // met in two banks, and the 72 reads per thread bound the kernel (4-way conflicts)
    const int tid = threadIdx.x;
    const int el = tid / NGP, g = tid % NGP;
    for (int t = tid; t < NGP * ND; t += 128) sdN[(t / ND) * SDS + t % ND] = tabdN[t];
    for (int t = tid; t < EPB * REC; t += 128) s_rec[t] = 0.0;          // the pad entries stay zero
    const double wg = el < EPB ? tabw[g] : 0.0;
    // grid-stride over groups of EPB elements: a million 128-thread blocks that live for a microsecond each were bound by
    // the block launch rate (5.4 ms for 11 GB of records, 27 % DRAM throughput, 17 long-scoreboard stalls per issue)
    const int64_t n_groups = (n_elem + EPB - 1) / EPB;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t e0 = grp * EPB, e = e0 + el;
        const bool live = el < EPB && e < n_elem;
        // every global operand of the group is requested before the barrier; each node is fetched once per element and
        // shared through shared memory (not once per Gauss-point thread)
        double Ee = 0.0, ne = 0.0, re = 0.0;
        if (live && g == 0) { Ee = E[e]; ne = nu[e]; re = rho[e]; }
        if (STAGE_XE && live) {
            for (int b = g; b < NNE; b += NGP) {
                const int c = conn[e * NNE + b];
#pragma unroll
                for (int d = 0; d < DIM; ++d) s_xe[(el * NNE + b) * DIM + d] = xyz[(int64_t)c * 3 + d];
            }
        }
        __syncthreads();                                 // coordinates staged; the previous group's records are copied out
        if (live) {
            double xe[ND];
            if (STAGE_XE) {
#pragma unroll
                for (int t = 0; t < ND; ++t) xe[t] = s_xe[el * ND + t];
            } else {
#pragma unroll
                for (int b = 0; b < NNE; ++b) {
                    const int c = conn[e * NNE + b];
#pragma unroll
                    for (int d = 0; d < DIM; ++d) xe[b * DIM + d] = xyz[(int64_t)c * 3 + d];
                }
            }
            double J[DD], inv[DD], det;
#pragma unroll
            for (int r = 0; r < DD; ++r) J[r] = 0.0;
#pragma unroll
            for (int b = 0; b < NNE; ++b)
#pragma unroll
                for (int d = 0; d < DIM; ++d) {
                    const double dn = sdN[g * SDS + b * DIM + d];
#pragma unroll
                    for (int kk = 0; kk < DIM; ++kk) J[d * DIM + kk] += dn * xe[b * DIM + kk];
                }
            invert<DIM>(J, inv, det);
            double* o = s_rec + el * REC + g * ISTP;
#pragma unroll
            for (int r = 0; r < DD; ++r) o[r] = inv[r];
            o[DD] = det * wg;
            if (g == 0) {
                double* m = s_rec + el * REC + NGP * ISTP;
                m[0] = Ee * ne / ((1.0 + ne) * (1.0 - 2.0 * ne));
                m[1] = Ee / (2.0 * (1.0 + ne));
                m[2] = re;
            }
        }
        __syncthreads();
        const int64_t left = n_elem - e0;
        const int cnt = (int)(left < EPB ? left : EPB) * REC;
        double* dst = out + e0 * REC;
        for (int t = tid; t < cnt; t += 128) dst[t] = s_rec[t];
    }
}

// Node blocks of the record-fed kernel: as many consecutive nodes as fit into the ppb pair lanes of a CTA (greedy, restarted
// at every chunk of BLK_CHUNK nodes so that the chunks pack in parallel).  A fixed count of ppb / max_valence nodes leaves
// lanes idle wherever valences differ (hexa20: vertex nodes 8, mid-side nodes 4 elements; unstructured meshes).  Node and
// item caps bound the per-block arrays (shared memory of the kernel): see asm_build_block_desc.
// FILL = false: blocks per chunk; FILL = true: first node of every block (blocks of chunk c start at chunk_ptr[c]).
constexpr int BLK_CHUNK = 128;
template <bool FILL>
__global__ void k_blk_chunks(const int64_t* __restrict__ n2e_ptr, const int64_t* __restrict__ nbr_ptr, int64_t n_nodes, int ppb,
                             int node_cap, int item_cap, int64_t* __restrict__ cnt, const int64_t* __restrict__ chunk_ptr, int64_t* __restrict__ blk_first,
                             int* __restrict__ maxima) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t a0 = c * BLK_CHUNK;
    if (a0 >= n_nodes) return;
    const int64_t a1 = min(a0 + BLK_CHUNK, n_nodes);
    int64_t nb = 0, out = FILL ? chunk_ptr[c] : 0;
    int pairs = 0, nodes = 0, items = 0, max_nodes = 0, max_items = 0;
    for (int64_t a = a0; a < a1; ++a) {
        const int v = (int)(n2e_ptr[a + 1] - n2e_ptr[a]);
        const int ni = (int)(nbr_ptr[a + 1] - nbr_ptr[a]);
        if (nodes > 0 && (pairs + v > ppb || nodes >= node_cap || items + ni > item_cap)) {   // close the block
            max_nodes = max(max_nodes, nodes); max_items = max(max_items, items);
            pairs = 0; nodes = 0; items = 0;
        }
        if (nodes == 0) {
            if (FILL) blk_first[out + nb] = a;
            ++nb;
        }
        pairs += v; ++nodes; items += ni;
    }
    max_nodes = max(max_nodes, nodes); max_items = max(max_items, items);
    if (!FILL) { cnt[c] = nb; atomicMax(&maxima[0], max_nodes); atomicMax(&maxima[1], max_items); }
}

// distinct elements of the pairs of every node block, ascending; one thread per pair
__global__ void k_blk_desc(const int64_t* __restrict__ n2e_ptr, const int32_t* __restrict__ n2e, const int32_t* __restrict__ node_rl,
                           const int64_t* __restrict__ blk_first, int ppb, int32_t* __restrict__ blk_elem, int32_t* __restrict__ blk_U,
                           uint8_t* __restrict__ pair_ui, int* __restrict__ umax) {
    extern __shared__ int s_desc[];
    int* s_e = s_desc;                                   // [ppb] element of the pair (-1: none / node without rows)
    int* s_f = s_desc + ppb;                             // [ppb] 1: first pair of its element in the block
    const int k = threadIdx.x;
    const int64_t a0 = blk_first[blockIdx.x], a1 = blk_first[blockIdx.x + 1];
    const int64_t P0 = n2e_ptr[a0];
    const int npairs = (int)(n2e_ptr[a1] - P0);
    int e = -1;
    if (k < npairs) {
        e = n2e[P0 + k];
        int64_t a = a0;
        while (a + 1 < a1 && n2e_ptr[a + 1] - P0 <= k) ++a;
        if (node_rl[a] <= 0) e = -1;
    }
    s_e[k] = e;
    __syncthreads();
    bool first = e >= 0;
    if (e >= 0)
        for (int j = 0; j < k; ++j)
            if (s_e[j] == e) { first = false; break; }
    s_f[k] = first ? 1 : 0;
    __syncthreads();
    // record slot of an element = its rank among the block's distinct elements (ascending id): consecutive elements sit in
    // consecutive slots and are fetched as one bulk copy.  (Measured and dropped: colouring the slots so that the eight
    // elements of a node fall into different shared-memory banks -- their ranks x, x+1, x+9, x+10, ... collide pairwise
    // on the 16-byte record reads, 2.1 wavefronts instead of 1 -- saved 1.3 ms per assembly of the 255^3 box and cost
    // 31 ms in this kernel, once per pattern.)
    if (k < npairs) {
        int u = 255;
        if (e >= 0) {
            u = 0;
            for (int j = 0; j < ppb; ++j) u += (s_f[j] && s_e[j] < e) ? 1 : 0;
            if (first) blk_elem[(int64_t)blockIdx.x * ppb + u] = e;
        }
        pair_ui[P0 + k] = (uint8_t)u;
    }
    if (k == 0) {
        int U = 0;
        for (int j = 0; j < ppb; ++j) U += s_f[j];
        blk_U[blockIdx.x] = U;
        atomicMax(umax, U);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent, fully TMA-fed kernel.  A first record-fed kernel with one CTA per node block (50.9 ms at 255^3, since removed;
// profiles/r2_k_assemble_rec_128cube.txt) spent 25 % of its samples in the FP64 phases, 17 % in the index prologue (two
// dependent levels of global loads), 16 % waiting for the records, 12 % at barriers -- every block paid its whole load
// latency, and only four blocks per SM overlapped.  Here every index a
// block needs is packed at pattern time into ONE contiguous descriptor (k_blk_pack: pointers, row bases, element list,
// pair -> element / local node / slot map, item offsets and masks), and a persistent CTA runs its blocks through a
// pipeline: the descriptor of block i+2 and the records of block i+1 are in flight (bulk copies counted on mbarriers)
// while block i integrates, stages and gathers.  The compute warps never wait for global memory; their only global
// accesses are the stores of K / M.  Same arithmetic and summation order as the two kernels above (same bits).
struct DescLayout {                                      // byte offsets inside one block descriptor
    int stride;                                          // multiple of 16
    int o_ptr, o_nptr, o_rl, o_eq, o_runs, o_rowbase, o_off, o_ui, o_al, o_pos, o_fmask;
};
inline DescLayout make_desc_layout(int npb, int imax, int ppb, int nne, int dim, int umax) {   // npb / imax: most nodes / items of a block
    DescLayout L;
    int o = 16;                                          // header: npairs, nbn, n_items, U | runs << 16
    L.o_ptr = o; o += 4 * (npb + 1);
    L.o_nptr = o; o += 4 * (npb + 1);
    L.o_rl = o; o += 4 * npb;
    L.o_eq = o; o += 4 * npb * dim;
    L.o_runs = o; o += 8 * umax;                         // runs of consecutive element ids: first element, first slot | count << 16
    o = (o + 7) & ~7;
    L.o_rowbase = o; o += 8 * npb * dim;
    L.o_off = o; o += 2 * imax;
    L.o_ui = o; o += ppb;
    L.o_al = o; o += ppb;
    L.o_pos = o; o += ppb * nne;
    L.o_fmask = o; o += imax;
    L.stride = (o + 15) & ~15;
    return L;
}

__global__ void __launch_bounds__(128)
k_blk_pack(const int64_t* __restrict__ n2e_ptr, const int64_t* __restrict__ nbr_ptr, const int32_t* __restrict__ node_rl,
           const int32_t* __restrict__ eq, const int64_t* __restrict__ rowptr, const uint16_t* __restrict__ nbr_off,
           const uint8_t* __restrict__ nbr_free, const uint8_t* __restrict__ pair_al, const uint8_t* __restrict__ pair_pos,
           const int32_t* __restrict__ blk_elem, const int32_t* __restrict__ blk_U, const uint8_t* __restrict__ pair_ui,
           const int64_t* __restrict__ blk_first, int npb, int imax, int ppb, int nne, int dim, int umax, DescLayout L,
           unsigned char* __restrict__ out) {
    unsigned char* D = out + (size_t)blockIdx.x * L.stride;
    const int tid = threadIdx.x;
    const int64_t a0 = blk_first[blockIdx.x], a1 = blk_first[blockIdx.x + 1];
    const int nbn = (int)(a1 - a0);
    const int64_t P0 = n2e_ptr[a0], nbr0 = nbr_ptr[a0];
    const int npairs = (int)(n2e_ptr[a1] - P0), n_items = (int)(nbr_ptr[a1] - nbr0);
    const int U = blk_U[blockIdx.x];
    if (tid == 0) {                                      // consecutive element ids are consecutive records: one bulk copy per run
        int* runs = reinterpret_cast<int*>(D + L.o_runs);
        const int32_t* el = blk_elem + (int64_t)blockIdx.x * ppb;
        int nr = 0;
        for (int t = 0; t < U;) {
            int c = 1;
            while (t + c < U && el[t + c] == el[t] + c) ++c;
            runs[2 * nr] = el[t]; runs[2 * nr + 1] = t | (c << 16);
            ++nr; t += c;
        }
        int* h = reinterpret_cast<int*>(D);
        h[0] = npairs; h[1] = nbn; h[2] = n_items; h[3] = U | (nr << 16);
    }
    for (int t = tid; t <= npb; t += 128) {
        const int64_t a = min(a0 + t, a1);
        reinterpret_cast<int*>(D + L.o_ptr)[t] = (int)(n2e_ptr[a] - P0);
        reinterpret_cast<int*>(D + L.o_nptr)[t] = (int)(nbr_ptr[a] - nbr0);
        if (t < npb) reinterpret_cast<int*>(D + L.o_rl)[t] = t < nbn ? node_rl[a0 + t] : 0;
    }
    for (int t = tid; t < npb * dim; t += 128) {
        int e = -1;
        long long rb = -1;
        if (t < nbn * dim) {
            e = eq[a0 * dim + t];
            if (e >= 0 && node_rl[a0 + t / dim] > 0) rb = (long long)rowptr[e];
        }
        reinterpret_cast<int*>(D + L.o_eq)[t] = e;
        reinterpret_cast<long long*>(D + L.o_rowbase)[t] = rb;
    }
    for (int t = tid; t < ppb; t += 128) {
        D[L.o_ui + t] = t < npairs ? pair_ui[P0 + t] : (unsigned char)255;
        D[L.o_al + t] = t < npairs ? pair_al[P0 + t] : (unsigned char)0;
    }
    for (int t = tid; t < ppb * nne; t += 128) D[L.o_pos + t] = t < npairs * nne ? pair_pos[P0 * nne + t] : (unsigned char)0;
    for (int t = tid; t < imax; t += 128) {
        reinterpret_cast<uint16_t*>(D + L.o_off)[t] = t < n_items ? nbr_off[nbr0 + t] : (uint16_t)0;
        D[L.o_fmask + t] = t < n_items ? nbr_free[nbr0 + t] : (unsigned char)0;
    }
}

struct TmaSmem { size_t srec, stage, sdN, sN, mitem, bars, desc, inv, total; };
// staging row of one pair: its DIM x (NNE*DIM) row block followed by the NNE mass entries; odd stride, so the pair lanes
// of a half warp (consecutive pairs) and the (node, neighbour) lanes of the gather spread over the banks
__host__ __device__ constexpr int tma_stage_stride(int nne, int dim) { return (dim * nne * dim + nne) | 1; }
__host__ __device__ inline TmaSmem tma_smem_layout(int rec, int umax, int ppb, int sst, int nne, int dim, int ngp, int imax, int max_nbr, int desc_stride) {
    TmaSmem m;
    size_t o = 0;
    m.srec = o; o += (size_t)umax * rec * 8;
    m.stage = o; o += (size_t)ppb * sst * 8;
    m.sdN = o; o += (size_t)ngp * nne * dim * 8;
    m.sN = o; o += (size_t)ngp * nne * 8;
    m.mitem = o; o += (size_t)imax * 8;
    m.bars = o; o += 4 * 8;
    o = (o + 15) & ~(size_t)15;
    m.desc = o; o += 2 * (size_t)desc_stride;
    m.inv = o; o += (size_t)ppb * ((max_nbr + 3) & ~3);
    m.total = (o + 15) & ~(size_t)15;
    return m;
}

template <int NNE, int DIM, int NGP, int TPB, int LPP, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
k_assemble_tma(AsmParams p, const double* __restrict__ rec, const unsigned char* __restrict__ desc, DescLayout L, int64_t n_blocks,
               int imax, int umax) {
    constexpr int DD = DIM * DIM, ND = NNE * DIM;
    constexpr int NBB = NNE / LPP;
    constexpr int PPB = TPB / LPP;
    constexpr int SST = tma_stage_stride(NNE, DIM), SMO = DIM * ND;   // row stride and offset of the mass entries in a row
    constexpr int ISTP = rec_point_stride(DIM), REC = rec_stride(DIM, NGP);
    constexpr int UGB = SC_BLK_UG;
    static_assert(NNE % LPP == 0 && (TPB / 32) % LPP == 0 && PPB <= 255, "unsupported split");
    extern __shared__ __align__(128) unsigned char smem_b[];
    const TmaSmem sm = tma_smem_layout(REC, umax, PPB, SST, NNE, DIM, NGP, imax, p.max_nbr, L.stride);
    double* srec = reinterpret_cast<double*>(smem_b + sm.srec);
    double* stage = reinterpret_cast<double*>(smem_b + sm.stage);
    double* sdN = reinterpret_cast<double*>(smem_b + sm.sdN);
    double* sN = reinterpret_cast<double*>(smem_b + sm.sN);
    double* s_mitem = reinterpret_cast<double*>(smem_b + sm.mitem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + sm.bars);      // [0], [1]: descriptor slots, [2]: records
    unsigned char* sdesc = smem_b + sm.desc;
    unsigned char* s_inv = smem_b + sm.inv;
    const int irow = (p.max_nbr + 3) & ~3;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = (warp / LPP) * 32 + lane, half = warp % LPP;
    const int64_t G = gridDim.x;
    int64_t b = blockIdx.x;
    if (b >= n_blocks) return;
    if (tid == 0) {
        tma_mbar_init(&bars[0], 1); tma_mbar_init(&bars[1], 1); tma_mbar_init(&bars[2], 1);
        tma_mbar_fence_init();
    }
    for (int t = tid; t < NGP * NNE * DIM; t += TPB) sdN[t] = p.tabdN[t];
    for (int t = tid; t < NGP * NNE; t += TPB) sN[t] = p.tabN[t];
    __syncthreads();
    // records of the block whose descriptor sits in `D` (warp 0, after the descriptor arrived)
    auto issue_records = [&](const unsigned char* D) {
        const int hu = reinterpret_cast<const int*>(D)[3];
        const int U = hu & 0xffff, nr = hu >> 16;
        if (lane == 0) tma_mbar_expect_tx(&bars[2], (uint32_t)(U * REC * sizeof(double)));
        __syncwarp();
        const int* runs = reinterpret_cast<const int*>(D + L.o_runs);
        for (int r = lane; r < nr; r += 32) {
            const int e0 = runs[2 * r], sc = runs[2 * r + 1];
            tma_bulk_load(srec + (size_t)(sc & 0xffff) * REC, rec + (int64_t)e0 * REC, (uint32_t)((sc >> 16) * REC * sizeof(double)), &bars[2]);
        }
    };
    if (warp == 0) {
        if (lane == 0) {
            tma_mbar_expect_tx(&bars[0], (uint32_t)L.stride);
            tma_bulk_load(sdesc, desc + (size_t)b * L.stride, (uint32_t)L.stride, &bars[0]);
            if (b + G < n_blocks) {
                tma_mbar_expect_tx(&bars[1], (uint32_t)L.stride);
                tma_bulk_load(sdesc + L.stride, desc + (size_t)(b + G) * L.stride, (uint32_t)L.stride, &bars[1]);
            }
        }
        __syncwarp();
        tma_mbar_wait(&bars[0], 0);
        issue_records(sdesc);
    }

    for (int it = 0; b < n_blocks; b += G, ++it) {
        const int slot = it & 1;
        const unsigned char* D = sdesc + (size_t)slot * L.stride;
        tma_mbar_wait(&bars[slot], (uint32_t)((it >> 1) & 1));
        const int nbn = reinterpret_cast<const int*>(D)[1], n_items = reinterpret_cast<const int*>(D)[2];
        const int* s_ptr = reinterpret_cast<const int*>(D + L.o_ptr);
        const int* s_nptr = reinterpret_cast<const int*>(D + L.o_nptr);
        const int* s_rl = reinterpret_cast<const int*>(D + L.o_rl);
        const long long* s_rowbase = reinterpret_cast<const long long*>(D + L.o_rowbase);
        const int ui = D[L.o_ui + k], al = D[L.o_al + k];
        const bool valid = ui != 255;
        if (half == 0) {                                 // this pair's row of the slot -> local node map (only its owner writes it)
            unsigned* row = reinterpret_cast<unsigned*>(s_inv + (size_t)k * irow);
            for (int t = 0; t < irow / 4; ++t) row[t] = 0xffffffffu;
            if (valid) {
#pragma unroll
                for (int bb = 0; bb < NNE; ++bb) s_inv[(size_t)k * irow + D[L.o_pos + k * NNE + bb]] = (unsigned char)bb;
            }
        }
        // ---- phase 1: gradient products of every pair over all Gauss points ----------------------------------------
        double acc[DIM][NBB * DIM];
        double mab[NBB];
#pragma unroll
        for (int i = 0; i < DIM; ++i)
#pragma unroll
            for (int c = 0; c < NBB * DIM; ++c) acc[i][c] = 0.0;
#pragma unroll
        for (int bb = 0; bb < NBB; ++bb) mab[bb] = 0.0;
        tma_mbar_wait(&bars[2], (uint32_t)(it & 1));
        double lam = 0.0, mu = 0.0, rho = 0.0;
        if (valid) {
            const double* sr = srec + (size_t)ui * REC;
#pragma unroll UGB
            for (int g = 0; g < NGP; ++g) {
                const double2* si2 = reinterpret_cast<const double2*>(sr + g * ISTP);
                double rv[ISTP];
#pragma unroll
                for (int r = 0; r < ISTP / 2; ++r) { const double2 v = si2[r]; rv[2 * r] = v.x; rv[2 * r + 1] = v.y; }
                const double wj = rv[DD];
                double wga[DIM];
#pragma unroll
                for (int kk = 0; kk < DIM; ++kk) {
                    double s = 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; ++d) s += sdN[(g * NNE + al) * DIM + d] * rv[kk * DIM + d];
                    wga[kk] = wj * s;
                }
                const double wna = wj * sN[g * NNE + al];
#pragma unroll
                for (int bb = 0; bb < NBB; ++bb) {
                    const double* dnb = c_tabdN + (g * NNE + half * NBB + bb) * DIM;
                    double gb[DIM];
#pragma unroll
                    for (int kk = 0; kk < DIM; ++kk) {
                        double s = 0.0;
#pragma unroll
                        for (int d = 0; d < DIM; ++d) s += dnb[d] * rv[kk * DIM + d];
                        gb[kk] = s;
                    }
#pragma unroll
                    for (int i = 0; i < DIM; ++i)
#pragma unroll
                        for (int j = 0; j < DIM; ++j) acc[i][bb * DIM + j] += wga[i] * gb[j];
                    mab[bb] += wna * c_tabN[g * NNE + half * NBB + bb];
                }
            }
            lam = sr[NGP * ISTP + 0]; mu = sr[NGP * ISTP + 1]; rho = sr[NGP * ISTP + 2];
        }
        __syncthreads();                                 // A: the records are dead -> fetch the next block's
        if (warp == 0 && b + G < n_blocks) {
            tma_mbar_wait(&bars[slot ^ 1], (uint32_t)(((it + 1) >> 1) & 1));
            issue_records(sdesc + (size_t)(slot ^ 1) * L.stride);
        }
        if (valid) {
#pragma unroll
            for (int bb = 0; bb < NBB; ++bb) {
                double tr = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; ++d) tr += acc[d][bb * DIM + d];
                tr *= mu;
#pragma unroll
                for (int i = 0; i < DIM; ++i)
#pragma unroll
                    for (int j = 0; j < DIM; ++j) {
                        double t = lam * acc[i][bb * DIM + j] + mu * acc[j][bb * DIM + i];
                        if (i == j) t += tr;
                        stage[(size_t)k * SST + i * ND + (half * NBB + bb) * DIM + j] = t;
                    }
                stage[(size_t)k * SST + SMO + half * NBB + bb] = rho * mab[bb];
            }
        }
        __syncthreads();                                 // B
        // ---- phase 2: ordered gather, one thread per (node, neighbour) item -----------------------------------------
        const uint16_t* s_off = reinterpret_cast<const uint16_t*>(D + L.o_off);
        const unsigned char* s_fm = D + L.o_fmask;
        for (int q = tid; q < n_items; q += TPB) {
            int lo = 0, hi = nbn;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_nptr[mid] <= q) lo = mid; else hi = mid;
            }
            const int n = lo, pidx = q - s_nptr[n];
            const int off = s_off[q];
            const int fmask = s_fm[q];
            double blk[DIM][DIM];
#pragma unroll
            for (int i = 0; i < DIM; ++i)
#pragma unroll
                for (int j = 0; j < DIM; ++j) blk[i][j] = 0.0;
            double m = 0.0;
            const int pr1 = s_ptr[n + 1];
            if (s_rl[n] > 0)
            for (int pr0 = s_ptr[n]; pr0 < pr1; pr0 += 8) {
                int bs[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) bs[c] = (pr0 + c < pr1) ? s_inv[(size_t)(pr0 + c) * irow + pidx] : 0xff;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int bl = bs[c];
                    if (bl == 0xff) continue;
                    const double* sp = stage + (size_t)(pr0 + c) * SST + bl * DIM;
#pragma unroll
                    for (int i = 0; i < DIM; ++i)
#pragma unroll
                        for (int j = 0; j < DIM; ++j) blk[i][j] += sp[i * ND + j];
                    m += stage[(size_t)(pr0 + c) * SST + SMO + bl];
                }
            }
            s_mitem[q] = m;
#pragma unroll
            for (int i = 0; i < DIM; ++i) {
                const long long rb = s_rowbase[n * DIM + i];
                if (rb < 0) continue;
                int64_t o = rb + off;
#pragma unroll
                for (int j = 0; j < DIM; ++j) {
                    if (!(fmask & (1 << j))) continue;
                    if (p.K) p.K[o] = blk[i][j];
                    if (p.M) p.M[o] = (i == j) ? m : 0.0;
                    ++o;
                }
            }
        }
        __syncthreads();                                 // C: staging area and slot map are free again
        if (warp == 0) {
            // lumped mass = ordered row sums of the items' mass blocks (s_mitem is not written again before barrier B of the
            // next block); then the descriptor slot is free: fetch the one of the block after the next
            if (p.Ml) {
                const int* s_eq = reinterpret_cast<const int*>(D + L.o_eq);
                for (int t = lane; t < nbn * DIM; t += 32) {
                    const int n = t / DIM, i = t % DIM;
                    if (s_rowbase[t] < 0) continue;
                    double s = 0.0;
                    for (int q = s_nptr[n]; q < s_nptr[n + 1]; ++q)
                        if (s_fm[q] & (1 << i)) s += s_mitem[q];
                    p.Ml[s_eq[t]] = s;
                }
            }
            __syncwarp();
            if (lane == 0 && b + 2 * G < n_blocks) {
                tma_mbar_expect_tx(&bars[slot], (uint32_t)L.stride);
                tma_bulk_load(sdesc + (size_t)slot * L.stride, desc + (size_t)(b + 2 * G) * L.stride, (uint32_t)L.stride, &bars[slot]);
            }
        }
    }
}

template <int NNE, int DIM, int NGP, int TPB, int LPP, int MINB>
int launch_tma_cfg(sc_ctx* ctx, const AsmParams& p, const ShapeTable& t, bool* handled) {
    constexpr int PPB = TPB / LPP;
    constexpr int SST = tma_stage_stride(NNE, DIM);
    constexpr int REC = rec_stride(DIM, NGP);
    *handled = false;
    if (ctx->no_asm_records || !ctx->d_blk_desc || !p.pair_pos || ctx->blk_ppb != PPB || ctx->max_valence <= 0 ||
        ctx->max_valence > PPB || p.max_nbr > 255)
        return SC_OK;
    const DescLayout L = make_desc_layout(ctx->blk_npb, ctx->blk_imax, PPB, NNE, DIM, ctx->blk_umax);
    if (L.stride != ctx->blk_desc_stride) return SC_OK;
    const TmaSmem sm = tma_smem_layout(REC, ctx->blk_umax, PPB, SST, NNE, DIM, NGP, ctx->blk_imax, p.max_nbr, L.stride);
    if (sm.total > 112 * 1024) return SC_OK;
    const size_t rec_doubles = (size_t)ctx->n_elem * REC;
    if (ctx->asm_rec_cap < rec_doubles) {
        sc_free(&ctx->d_asm_rec); ctx->asm_rec_cap = 0;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || free_b < rec_doubles * sizeof(double) + (size_t(2) << 30)) return SC_OK;
        if (cudaMalloc((void**)&ctx->d_asm_rec, rec_doubles * sizeof(double) + 64) != cudaSuccess) { cudaGetLastError(); ctx->d_asm_rec = nullptr; return SC_OK; }
        ctx->asm_rec_cap = rec_doubles;
    }
    auto kern = k_assemble_tma<NNE, DIM, NGP, TPB, LPP, MINB>;
    SC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm.total));
    int occ = 0;
    SC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TPB, sm.total));
    if (occ < 1) return SC_OK;
    SC_CUDA(ctx, cudaMemcpyToSymbolAsync(c_tabN, t.N.data(), t.N.size() * sizeof(double), 0, cudaMemcpyHostToDevice, ctx->stream));
    SC_CUDA(ctx, cudaMemcpyToSymbolAsync(c_tabdN, t.dN.data(), t.dN.size() * sizeof(double), 0, cudaMemcpyHostToDevice, ctx->stream));
    constexpr int EPB = 128 / NGP;
    k_elem_records<NNE, DIM, NGP><<<(unsigned)std::min<int64_t>((ctx->n_elem + EPB - 1) / EPB, (int64_t)ctx->sm_count * 16), 128, 0, ctx->stream>>>(
        p.xyz, p.conn, p.E, p.nu, p.rho, p.tabdN, p.tabw, ctx->n_elem, ctx->d_asm_rec);
    SC_CHECK_LAUNCH(ctx);
    const int64_t n_blocks = ctx->blk_count;
    const unsigned grid = (unsigned)std::min<int64_t>(n_blocks, (int64_t)ctx->sm_count * occ);
    kern<<<grid, TPB, sm.total, ctx->stream>>>(p, ctx->d_asm_rec, ctx->d_blk_desc, L, n_blocks, ctx->blk_imax, ctx->blk_umax);
    SC_CHECK_LAUNCH(ctx);
    *handled = true;
    return SC_OK;
}

template <int NNE, int DIM, int NGP, int TPB, int LPP, int MINB>
int launch_blk_cfg(sc_ctx* ctx, const AsmParams& p, const ShapeTable& t, bool* handled) {
    constexpr int PPB = TPB / LPP;
    constexpr int ND = NNE * DIM, SST = DIM * ND + 1;
    constexpr int GC = NGP <= 9 ? NGP : 9;
    constexpr int UST = GC * (DIM * DIM + 1) + 1;
    *handled = false;
    if (!p.pair_pos || ctx->max_valence <= 0 || ctx->max_valence > PPB || p.max_nbr > 255) return SC_OK;
    const int npb = std::max(1, PPB / ctx->max_valence);
    const size_t region_a = std::max((size_t)PPB * SST + (size_t)PPB * NNE, (size_t)PPB * UST);
    const size_t bytes = (region_a + (size_t)NGP * NNE * DIM + (size_t)NGP * NNE + NGP + (size_t)npb * p.max_nbr + (size_t)npb * DIM + (size_t)PPB * 3) * sizeof(double) +
                         (3 * (size_t)(npb + 1) + (size_t)PPB * NNE + 2 + 512) * sizeof(int) + (size_t)PPB * p.max_nbr + 16;
    if (bytes > 110 * 1024) return SC_OK;
    SC_CUDA(ctx, cudaMemcpyToSymbolAsync(c_tabN, t.N.data(), t.N.size() * sizeof(double), 0, cudaMemcpyHostToDevice, ctx->stream));
    SC_CUDA(ctx, cudaMemcpyToSymbolAsync(c_tabdN, t.dN.data(), t.dN.size() * sizeof(double), 0, cudaMemcpyHostToDevice, ctx->stream));
    const unsigned grid = (unsigned)((p.n_nodes + npb - 1) / npb);
    auto kern = k_assemble_blk<NNE, DIM, NGP, TPB, LPP, MINB>;
    SC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    kern<<<grid, TPB, bytes, ctx->stream>>>(p, npb);
    SC_CHECK_LAUNCH(ctx);
    *handled = true;
    return SC_OK;
}

template <int NNE, int DIM, int NGP>
int launch_blk(sc_ctx* ctx, const AsmParams& p, const ShapeTable& t, bool* handled) {
    // measured on B200 (hexa8, 128^3 elements): 128 threads x 2 lanes per pair x 4 CTAs/SM 7.4 ms; 256 x 4 x 3 9.5 ms;
    // 128 x 4 x 6 9.4 ms (both spill at 80 registers and do 14 % more flops); 256 x 2 x 2 8.3 ms; coordinates re-read
    // per set-up task instead of held in registers 7.9 ms
    if constexpr (DIM * NNE * DIM > 72 && NNE % 5 == 0) {
        // tetra10 / hexa20: five lanes per pair (2 / 4 node blocks each), 160-thread blocks of 32 pairs
        SC_TRY((launch_tma_cfg<NNE, DIM, NGP, 160, 5, 2>(ctx, p, t, handled)));
        if (*handled) return SC_OK;
        return launch_blk_cfg<NNE, DIM, NGP, 160, 5, 3>(ctx, p, t, handled);
    } else if constexpr (DIM * NNE * DIM > 36 && NNE % 2 == 0) {
        SC_TRY((launch_tma_cfg<NNE, DIM, NGP, 128, 2, 3>(ctx, p, t, handled)));
        if (*handled) return SC_OK;
        return launch_blk_cfg<NNE, DIM, NGP, 128, 2, 4>(ctx, p, t, handled);
    } else {
        SC_TRY((launch_tma_cfg<NNE, DIM, NGP, 128, 1, 2>(ctx, p, t, handled)));
        if (*handled) return SC_OK;
        return launch_blk_cfg<NNE, DIM, NGP, 128, 1, 2>(ctx, p, t, handled);
    }
}

template <int NNE, int DIM, int NGP>
int launch(sc_ctx* ctx, const AsmParams& p) {
    constexpr int JS = DIM * DIM + 1;
    const size_t tables = (size_t)NGP * NNE + (size_t)NGP * NNE * DIM + NGP;
    const size_t per_warp = (size_t)NNE * DIM + (size_t)NGP * JS + (size_t)NGP * NNE * DIM + p.max_nbr + (NNE + 1) / 2;
    int warps = 8;
    size_t bytes = 0;
    for (; warps >= 1; warps >>= 1) {
        bytes = (tables + warps * per_warp + (size_t)warps * DIM * p.max_rl) * sizeof(double);
        if (bytes <= 200 * 1024) break;
    }
    if (warps < 1) return sc_fail(ctx, SC_ERR_UNSUPPORTED, "assembly staging does not fit in shared memory (max row length %d)", p.max_rl);
    auto kern = k_assemble<NNE, DIM, NGP>;
    SC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    const unsigned grid = (unsigned)((p.n_nodes + warps - 1) / warps);
    kern<<<grid, warps * 32, bytes, ctx->stream>>>(p, warps);
    SC_CHECK_LAUNCH(ctx);
    return SC_OK;
}

}  // namespace

void asm_release_scratch(sc_ctx* ctx) {
    sc_free(&ctx->d_asm_rec);
    ctx->asm_rec_cap = 0;
}

int asm_build_block_desc(sc_ctx* ctx) {
    sc_free(&ctx->d_blk_elem); sc_free(&ctx->d_blk_U); sc_free(&ctx->d_pair_ui); sc_free(&ctx->d_blk_desc);
    ctx->blk_npb = ctx->blk_imax = ctx->blk_ppb = ctx->blk_umax = ctx->blk_desc_stride = 0;
    ctx->blk_count = 0;
    if (!ctx->d_pair_pos || ctx->max_valence <= 0 || ctx->n_nodes <= 0 || ctx->max_nbr > 255) return SC_OK;
    int tpb = 0, lpp = 1;
    asm_blk_shape(ctx->nne, ctx->dim, &tpb, &lpp);
    const int ppb = tpb / lpp;
    if (ctx->max_valence > ppb) return SC_OK;
    cudaStream_t st = ctx->stream;
    const int64_t n_chunks = (ctx->n_nodes + BLK_CHUNK - 1) / BLK_CHUNK;
    const int64_t n_pairs = ctx->n_elem * ctx->nne;
    int64_t *d_cnt = nullptr, *d_cptr = nullptr, *d_first = nullptr;
    int* d_max = nullptr;
    int rc = SC_OK;
    auto body = [&]() -> int {
        // 1. node blocks: greedy packing by pair count, chunk by chunk
        SC_TRY(sc_alloc(ctx, &d_cnt, (size_t)n_chunks + 1));
        SC_TRY(sc_alloc(ctx, &d_cptr, (size_t)n_chunks + 1));
        SC_TRY(sc_alloc(ctx, &d_max, 3));
        SC_CUDA(ctx, cudaMemsetAsync(d_max, 0, 3 * sizeof(int), st));
        SC_CUDA(ctx, cudaMemsetAsync(d_cnt + n_chunks, 0, sizeof(int64_t), st));
        const unsigned cg = (unsigned)((n_chunks + 127) / 128);
        // caps relative to the uniform packing ppb / max_valence: twice the nodes; the same number of (node, neighbour) items
        // for linear elements (valences only differ at boundaries; the hexa8 kernel sits 1.3 kB below the shared memory of
        // three CTAs per SM), a third more for quadratic ones, whose mid-side nodes have half the valence of the vertices
        const int npb0 = std::max(1, ppb / ctx->max_valence);
        const bool quadratic = ctx->elem_type == SC_TRI6 || ctx->elem_type == SC_QUAD8 || ctx->elem_type == SC_TETRA10 || ctx->elem_type == SC_HEXA20;
        const int node_cap = std::min(2 * npb0, 64), item_cap = std::max(npb0 * ctx->max_nbr * (quadratic ? 4 : 3) / 3, ctx->max_nbr);
        k_blk_chunks<false><<<cg, 128, 0, st>>>(ctx->d_n2e_ptr, ctx->d_nbr_ptr, ctx->n_nodes, ppb, node_cap, item_cap, d_cnt, nullptr, nullptr, d_max);
        SC_CHECK_LAUNCH(ctx);
        size_t tmp_bytes = 0;
        SC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt, d_cptr, n_chunks + 1, st));
        void* tmp = nullptr;
        SC_CUDA(ctx, cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
        const cudaError_t se = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_cnt, d_cptr, n_chunks + 1, st);
        ctx->launches += 2;
        int64_t n_blocks = 0;
        cudaError_t ce = cudaMemcpyAsync(&n_blocks, d_cptr + n_chunks, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
        int h_max[3] = {0, 0, 0};
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_max, d_max, 2 * sizeof(int), cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        cudaFree(tmp);
        if (se != cudaSuccess || ce != cudaSuccess)
            return sc_fail(ctx, SC_ERR_CUDA, "node blocks of the assembly failed: %s", cudaGetErrorString(se != cudaSuccess ? se : ce));
        SC_TRY(sc_alloc(ctx, &d_first, (size_t)n_blocks + 1));
        k_blk_chunks<true><<<cg, 128, 0, st>>>(ctx->d_n2e_ptr, ctx->d_nbr_ptr, ctx->n_nodes, ppb, node_cap, item_cap, nullptr, d_cptr, d_first, nullptr);
        SC_CHECK_LAUNCH(ctx);
        SC_CUDA(ctx, cudaMemcpyAsync(d_first + n_blocks, &ctx->n_nodes, sizeof(int64_t), cudaMemcpyHostToDevice, st));
        // 2. distinct elements per block, each pair's index into that list
        SC_TRY(sc_alloc(ctx, &ctx->d_blk_elem, (size_t)n_blocks * ppb));
        SC_TRY(sc_alloc(ctx, &ctx->d_blk_U, (size_t)n_blocks));
        SC_TRY(sc_alloc(ctx, &ctx->d_pair_ui, (size_t)n_pairs));
        SC_CUDA(ctx, cudaMemsetAsync(ctx->d_blk_elem, 0xff, (size_t)n_blocks * ppb * sizeof(int32_t), st));
        SC_CUDA(ctx, cudaMemsetAsync(ctx->d_pair_ui, 0xff, (size_t)n_pairs, st));
        k_blk_desc<<<(unsigned)n_blocks, ppb, 2 * ppb * sizeof(int), st>>>(ctx->d_n2e_ptr, ctx->d_n2e, ctx->d_node_rl, d_first, ppb, ctx->d_blk_elem,
                                                                          ctx->d_blk_U, ctx->d_pair_ui, d_max + 2);
        SC_CHECK_LAUNCH(ctx);
        int umax = 0;
        SC_CUDA(ctx, cudaMemcpyAsync(&umax, d_max + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
        SC_CUDA(ctx, cudaStreamSynchronize(st));
        const int npb = std::max(h_max[0], 1), imax = std::max(h_max[1], 1);
        umax = std::max(umax, 1);
        // 3. packed per-block descriptors of the persistent kernel
        const DescLayout L = make_desc_layout(npb, imax, ppb, ctx->nne, ctx->dim, umax);
        size_t free_b = 0, total_b = 0;
        const size_t desc_bytes = (size_t)n_blocks * L.stride;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || free_b < desc_bytes + (size_t(4) << 30)) return SC_OK;
        SC_TRY(sc_alloc(ctx, &ctx->d_blk_desc, desc_bytes));
        SC_CUDA(ctx, cudaMemsetAsync(ctx->d_blk_desc, 0, desc_bytes, st));
        k_blk_pack<<<(unsigned)n_blocks, 128, 0, st>>>(ctx->d_n2e_ptr, ctx->d_nbr_ptr, ctx->d_node_rl, ctx->d_eq, ctx->d_rowptr, ctx->d_nbr_off,
                                                        ctx->d_nbr_free, ctx->d_pair_al, ctx->d_pair_pos, ctx->d_blk_elem, ctx->d_blk_U,
                                                        ctx->d_pair_ui, d_first, npb, imax, ppb, ctx->nne, ctx->dim, umax, L, ctx->d_blk_desc);
        SC_CHECK_LAUNCH(ctx);
        SC_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->blk_npb = npb; ctx->blk_imax = imax; ctx->blk_ppb = ppb; ctx->blk_umax = umax;
        ctx->blk_count = n_blocks; ctx->blk_desc_stride = L.stride;
        return SC_OK;
    };
    rc = body();
    sc_free(&d_cnt); sc_free(&d_cptr); sc_free(&d_first); sc_free(&d_max);
    sc_free(&ctx->d_blk_elem); sc_free(&ctx->d_blk_U); sc_free(&ctx->d_pair_ui);      // packed into the descriptors
    if (rc != SC_OK) { sc_free(&ctx->d_blk_desc); ctx->blk_desc_stride = 0; }
    return rc;
}

int sc_assemble_run(sc_ctx* ctx, int order, int flags, double* seconds) {
    ShapeTable t;
    std::string err;
    if (!sc_make_shape_table(ctx->elem_type, order, t, err)) return sc_fail(ctx, SC_ERR_ARG, "%s", err.c_str());
    double *dN = nullptr, *ddN = nullptr, *dw = nullptr;
    SC_TRY(sc_alloc(ctx, &dN, t.N.size()));
    SC_TRY(sc_alloc(ctx, &ddN, t.dN.size()));
    SC_TRY(sc_alloc(ctx, &dw, t.w.size()));
    SC_CUDA(ctx, cudaMemcpy(dN, t.N.data(), t.N.size() * sizeof(double), cudaMemcpyHostToDevice));
    SC_CUDA(ctx, cudaMemcpy(ddN, t.dN.data(), t.dN.size() * sizeof(double), cudaMemcpyHostToDevice));
    SC_CUDA(ctx, cudaMemcpy(dw, t.w.data(), t.w.size() * sizeof(double), cudaMemcpyHostToDevice));

    if (flags & SC_ASM_K) SC_TRY(sc_alloc(ctx, &ctx->d_K, (size_t)ctx->nnz));
    if (flags & SC_ASM_M_FULL) SC_TRY(sc_alloc(ctx, &ctx->d_M, (size_t)ctx->nnz));
    if (flags & SC_ASM_M_LUMPED) {
        SC_TRY(sc_alloc(ctx, &ctx->d_Ml, (size_t)ctx->n_eq));
        SC_CUDA(ctx, cudaMemsetAsync(ctx->d_Ml, 0, sizeof(double) * ctx->n_eq, ctx->stream));
    }
    AsmParams p;
    p.xyz = ctx->d_xyz; p.conn = ctx->d_conn; p.eq = ctx->d_eq;
    p.E = ctx->d_E; p.nu = ctx->d_nu; p.rho = ctx->d_rho;
    p.n2e_ptr = ctx->d_n2e_ptr; p.n2e = ctx->d_n2e;
    p.nbr_ptr = ctx->d_nbr_ptr; p.nbr = ctx->d_nbr; p.nbr_off = ctx->d_nbr_off; p.nbr_free = ctx->d_nbr_free;
    p.node_rl = ctx->d_node_rl; p.node_row0 = ctx->d_node_row0; p.rowptr = ctx->d_rowptr;
    p.tabN = dN; p.tabdN = ddN; p.tabw = dw;
    p.K = (flags & SC_ASM_K) ? ctx->d_K : nullptr;
    p.M = (flags & SC_ASM_M_FULL) ? ctx->d_M : nullptr;
    p.Ml = (flags & SC_ASM_M_LUMPED) ? ctx->d_Ml : nullptr;
    p.pair_pos = ctx->d_pair_pos; p.pair_al = ctx->d_pair_al;
    p.n_nodes = ctx->n_nodes; p.max_rl = ctx->max_rl; p.max_nbr = ctx->max_nbr;

    sc_gpu_timer timer(ctx->stream);
    timer.start();
    int rc = SC_ERR_UNSUPPORTED;
    const int key = t.nne * 10000 + t.dim * 1000 + t.ngp;
    switch (key) {
#define SC_CASE(NNE, DIM, NGP)                                                           \
    case NNE * 10000 + DIM * 1000 + NGP:                                                 \
        {                                                                                \
            bool done = false;                                                           \
            rc = SC_OK;                                                                  \
            if (!ctx->force_generic_assembly) rc = launch_blk<NNE, DIM, NGP>(ctx, p, t, &done); \
            if (rc == SC_OK && !done) rc = launch<NNE, DIM, NGP>(ctx, p);                \
        }                                                                                \
        break;
        SC_CASE(3, 2, 1) SC_CASE(3, 2, 3) SC_CASE(3, 2, 4)
        SC_CASE(6, 2, 1) SC_CASE(6, 2, 3) SC_CASE(6, 2, 4)
        SC_CASE(4, 2, 1) SC_CASE(4, 2, 4) SC_CASE(4, 2, 9)
        SC_CASE(8, 2, 1) SC_CASE(8, 2, 4) SC_CASE(8, 2, 9)
        SC_CASE(4, 3, 1) SC_CASE(4, 3, 4)
        SC_CASE(10, 3, 1) SC_CASE(10, 3, 4)
        SC_CASE(8, 3, 1) SC_CASE(8, 3, 8) SC_CASE(8, 3, 27)
        SC_CASE(20, 3, 1) SC_CASE(20, 3, 8) SC_CASE(20, 3, 27)
#undef SC_CASE
        default: rc = sc_fail(ctx, SC_ERR_UNSUPPORTED, "no assembly kernel for nne=%d dim=%d ngp=%d", t.nne, t.dim, t.ngp);
    }
    timer.stop();
    cudaError_t se = cudaStreamSynchronize(ctx->stream);
    const float ms = timer.ms();
    sc_free(&dN); sc_free(&ddN); sc_free(&dw);
    if (rc != SC_OK) return rc;
    if (se != cudaSuccess) return sc_fail(ctx, SC_ERR_CUDA, "assembly kernel failed: %s", cudaGetErrorString(se));
    if (seconds) *seconds = ms * 1e-3;
    if (flags & SC_ASM_K) ctx->have_K = true;
    if (flags & SC_ASM_M_FULL) ctx->have_M = true;
    if (flags & SC_ASM_M_LUMPED) ctx->have_Ml = true;
    return SC_OK;
}
