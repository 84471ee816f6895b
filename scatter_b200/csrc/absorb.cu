// Absorbing (Lysmer-Kuhlemeyer) boundary faces on the device -- the arithmetic of GenerateMatrix.absorbing_boundaries
// (scatter/system_matrix.py:320-376) and of compute_abs_bound (scatter/discretisation.py:419-433).
//
// The host plans (scatter_b200/system_matrix.py:absorbing_plan): which element faces absorb, their nodes in the
// reference's face-node order, the sorted equation numbers the reference pairs them with, which of those dofs are
// perpendicular to their boundary, and -- because neighbouring faces share dofs -- the grouping of the per-face entries
// by matrix position in face order.  The device computes:
//   k_abs_faces   one thread per face: unit consistent face matrix S_ab = sum_g N_a N_b detJ w on the face's 2-D element
//                 (face coordinates = node coordinates without the face's normal direction), wave speeds
//                 vp = sqrt(Ec / rho), vs = sqrt(G / rho) of the face's element, and the per-face entries
//                 C: S_ab * (p0 rho vp | p1 rho vs)_b      K: |S_ab| * (Ec | G)_b
//   k_abs_reduce  one thread per matrix position: sum of its entries in face order (the order in which the reference's
//                 `C[i1, i1] += ...` statements run), K additionally divided by `absorbing_BC_stiff`.
// Both are O(boundary faces): latency-sized kernels, no roofline to chase; they are here so that no floating-point work
// of the path runs on the host.
#include "common.h"

namespace {

constexpr int ABS_MAX_NL = 8;      // quad8 faces of hexa20 meshes
constexpr int ABS_MAX_GP = 9;      // 3 x 3 Gauss points

struct FaceTables {
    int nl, ngp;
    double N[ABS_MAX_GP * ABS_MAX_NL];
    double dN[ABS_MAX_GP * ABS_MAX_NL * 2];
    double w[ABS_MAX_GP];
};

__global__ void k_abs_faces(FaceTables tb, int64_t n_faces, const int32_t* __restrict__ face_nodes, const int32_t* __restrict__ face_elem,
                            const int32_t* __restrict__ face_dir, const uint8_t* __restrict__ perp, const double* __restrict__ xyz,
                            const double* __restrict__ E, const double* __restrict__ nu, const double* __restrict__ rho, double p0, double p1,
                            double* __restrict__ cv, double* __restrict__ kv) {
    const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    const int nl = tb.nl;
    const int d = face_dir[f];
    const int k0 = d == 0 ? 1 : 0, k1 = d == 2 ? 1 : 2;          // in-plane coordinates: all but the face's direction
    double x[ABS_MAX_NL], y[ABS_MAX_NL];
    for (int a = 0; a < nl; ++a) {
        const int64_t n = face_nodes[f * nl + a];
        x[a] = xyz[n * 3 + k0];
        y[a] = xyz[n * 3 + k1];
    }
    double S[ABS_MAX_NL * ABS_MAX_NL];
    for (int i = 0; i < nl * nl; ++i) S[i] = 0.0;
    for (int g = 0; g < tb.ngp; ++g) {
        double j00 = 0.0, j01 = 0.0, j10 = 0.0, j11 = 0.0;       // J = dN^T xy (discretisation.py:315-329)
        for (int a = 0; a < nl; ++a) {
            const double du = tb.dN[(g * nl + a) * 2], dv = tb.dN[(g * nl + a) * 2 + 1];
            j00 += du * x[a]; j01 += du * y[a];
            j10 += dv * x[a]; j11 += dv * y[a];
        }
        const double dw = (j00 * j11 - j01 * j10) * tb.w[g];
        for (int a = 0; a < nl; ++a) {
            const double na = tb.N[g * nl + a] * dw;
            for (int b = 0; b < nl; ++b) S[a * nl + b] += na * tb.N[g * nl + b];
        }
    }
    const int e = face_elem[f];
    const double Ee = E[e], ne = nu[e], re = rho[e];
    const double Ec = Ee * (1.0 - ne) / ((1.0 + ne) * (1.0 - 2.0 * ne));        // system_matrix.py:282-287
    const double G = Ee / (2.0 * (1.0 + ne));
    const double fp = p0 * re * sqrt(Ec / re), fs = p1 * re * sqrt(G / re);
    for (int a = 0; a < nl; ++a)
        for (int b = 0; b < nl; ++b) {
            const bool pb = perp[f * nl + b] != 0;
            const double s = S[a * nl + b];
            cv[(f * nl + a) * nl + b] = s * (pb ? fp : fs);
            kv[(f * nl + a) * nl + b] = fabs(s) * (pb ? Ec : G);
        }
}

__global__ void k_abs_reduce(int64_t n_unique, const int64_t* __restrict__ grp_ptr, const int64_t* __restrict__ grp_entry,
                             const double* __restrict__ cv, const double* __restrict__ kv, double stiff, double* __restrict__ csum,
                             double* __restrict__ ksum) {
    const int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (u >= n_unique) return;
    double c = 0.0, k = 0.0;
    for (int64_t q = grp_ptr[u]; q < grp_ptr[u + 1]; ++q) {      // ascending entry id = face order
        const int64_t e = grp_entry[q];
        c += cv[e];
        k += kv[e];
    }
    csum[u] = c;
    ksum[u] = k / stiff;
}

template <typename T>
int to_device(sc_ctx* ctx, T** dst, const T* src, size_t n) {
    SC_TRY(sc_alloc(ctx, dst, n));
    if (n) SC_CUDA(ctx, cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return SC_OK;
}

}  // namespace

int abs_faces_eval(sc_ctx* ctx, int face_type, int order, int64_t n_faces, const int32_t* face_nodes, const int32_t* face_elem,
                   const int32_t* face_dir, const uint8_t* perp, int64_t n_unique, const int64_t* grp_ptr, const int64_t* grp_entry,
                   double p0, double p1, double stiff, double** d_csum, double** d_ksum) {
    ShapeTable t;
    std::string err;
    if (!sc_make_shape_table(face_type, order, t, err)) return sc_fail(ctx, SC_ERR_ARG, "%s", err.c_str());
    if (t.dim != 2 || t.nne > ABS_MAX_NL || t.ngp > ABS_MAX_GP) return sc_fail(ctx, SC_ERR_UNSUPPORTED, "unsupported face element / Gauss order");
    FaceTables tb;
    tb.nl = t.nne; tb.ngp = t.ngp;
    for (size_t i = 0; i < t.N.size(); ++i) tb.N[i] = t.N[i];
    for (size_t i = 0; i < t.dN.size(); ++i) tb.dN[i] = t.dN[i];
    for (size_t i = 0; i < t.w.size(); ++i) tb.w[i] = t.w[i];
    const int nl = t.nne;
    const int64_t n_entries = n_faces * nl * nl;
    for (int64_t i = 0; i < n_faces * nl; ++i)
        if (face_nodes[i] < 0 || face_nodes[i] >= ctx->n_nodes) return sc_fail(ctx, SC_ERR_ARG, "face node %lld out of range", (long long)i);
    for (int64_t q = 0; q < grp_ptr[n_unique]; ++q)
        if (grp_entry[q] < 0 || grp_entry[q] >= n_entries) return sc_fail(ctx, SC_ERR_ARG, "entry id %lld out of range", (long long)q);

    int32_t *d_fn = nullptr, *d_fe = nullptr, *d_fd = nullptr;
    uint8_t* d_perp = nullptr;
    int64_t *d_gp = nullptr, *d_ge = nullptr;
    double *d_cv = nullptr, *d_kv = nullptr;
    auto body = [&]() -> int {
        SC_TRY(to_device(ctx, &d_fn, face_nodes, (size_t)(n_faces * nl)));
        SC_TRY(to_device(ctx, &d_fe, face_elem, (size_t)n_faces));
        SC_TRY(to_device(ctx, &d_fd, face_dir, (size_t)n_faces));
        SC_TRY(to_device(ctx, &d_perp, perp, (size_t)(n_faces * nl)));
        SC_TRY(to_device(ctx, &d_gp, grp_ptr, (size_t)(n_unique + 1)));
        SC_TRY(to_device(ctx, &d_ge, grp_entry, (size_t)grp_ptr[n_unique]));
        SC_TRY(sc_alloc(ctx, &d_cv, (size_t)n_entries));
        SC_TRY(sc_alloc(ctx, &d_kv, (size_t)n_entries));
        SC_TRY(sc_alloc(ctx, d_csum, (size_t)n_unique));
        SC_TRY(sc_alloc(ctx, d_ksum, (size_t)n_unique));
        k_abs_faces<<<(unsigned)((n_faces + 63) / 64), 64, 0, ctx->stream>>>(tb, n_faces, d_fn, d_fe, d_fd, d_perp, ctx->d_xyz, ctx->d_E, ctx->d_nu,
                                                                             ctx->d_rho, p0, p1, d_cv, d_kv);
        SC_CHECK_LAUNCH(ctx);
        k_abs_reduce<<<(unsigned)((n_unique + 127) / 128), 128, 0, ctx->stream>>>(n_unique, d_gp, d_ge, d_cv, d_kv, stiff, *d_csum, *d_ksum);
        SC_CHECK_LAUNCH(ctx);
        SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return SC_OK;
    };
    const int rc = body();
    sc_free(&d_fn); sc_free(&d_fe); sc_free(&d_fd); sc_free(&d_perp); sc_free(&d_gp); sc_free(&d_ge); sc_free(&d_cv); sc_free(&d_kv);
    return rc;
}
