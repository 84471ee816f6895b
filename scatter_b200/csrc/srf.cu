// Spectral random field evaluation (randomisation method) -- the sampler behind the reference's per-element random
// material field (`scatter/random_fields.py:59-104`: gstools `SRF(model, mean, seed)` evaluated at the element
// centroids, optional lognormal transform).  gstools is an un-vendored dependency (requirements.txt:5,
// gstools==1.7.0); its `RandMeth` generator computes, for every point x,
//
//     field(x) = mean + sqrt(var / N) * sum_{j<N} ( z1_j cos(k_j . x) + z2_j sin(k_j . x) )
//
// with N (default 1000) wave vectors k_j drawn from the spectral density of the covariance model and z1, z2 ~ N(0,1).
// The wave vectors and amplitudes are drawn on the host (`scatter_b200/random_fields.py`); this kernel does the
// O(points x modes) evaluation: FP64 sincos bound (no reuse to exploit beyond the mode table), one point per thread, the
// mode table staged through shared memory in chunks, the sum over j strictly in order so the result is reproducible
// and matches the sequential CPU restatement (oracle/fem_np.py:srf_field) to rounding of sin/cos.
#include "common.h"

namespace {

constexpr int SRF_TPB = 256;
constexpr int SRF_CHUNK = 256;     // modes per shared-memory chunk (5 doubles each = 10 kB)

__global__ void __launch_bounds__(SRF_TPB) k_srf(const double* __restrict__ pos, int64_t n, const double* __restrict__ kv,
                                                 const double* __restrict__ z1, const double* __restrict__ z2, int n_modes,
                                                 double scale, double mean, int lognormal, double* __restrict__ out) {
    __shared__ double s_k[SRF_CHUNK * 3];
    __shared__ double s_z1[SRF_CHUNK];
    __shared__ double s_z2[SRF_CHUNK];
    const int64_t i = (int64_t)blockIdx.x * SRF_TPB + threadIdx.x;
    double x = 0.0, y = 0.0, z = 0.0;
    if (i < n) { x = pos[i * 3]; y = pos[i * 3 + 1]; z = pos[i * 3 + 2]; }
    double acc = 0.0;
    for (int m0 = 0; m0 < n_modes; m0 += SRF_CHUNK) {
        const int mc = min(SRF_CHUNK, n_modes - m0);
        __syncthreads();
        for (int t = threadIdx.x; t < mc * 3; t += SRF_TPB) s_k[t] = kv[(size_t)m0 * 3 + t];
        for (int t = threadIdx.x; t < mc; t += SRF_TPB) { s_z1[t] = z1[m0 + t]; s_z2[t] = z2[m0 + t]; }
        __syncthreads();
        if (i < n) {
#pragma unroll 4
            for (int j = 0; j < mc; ++j) {
                // explicit roundings: the phase is bit-identical to the CPU restatement (no FMA contraction)
                const double ph = __dadd_rn(__dadd_rn(__dmul_rn(s_k[j * 3], x), __dmul_rn(s_k[j * 3 + 1], y)), __dmul_rn(s_k[j * 3 + 2], z));
                double sn, cs;
                sincos(ph, &sn, &cs);
                acc = __dadd_rn(acc, __dadd_rn(__dmul_rn(s_z1[j], cs), __dmul_rn(s_z2[j], sn)));
            }
        }
    }
    if (i < n) {
        const double f = __dadd_rn(mean, __dmul_rn(scale, acc));
        out[i] = lognormal ? exp(f) : f;
    }
}

}  // namespace

int srf_sample(sc_ctx* ctx, int64_t n_points, const double* pos, int n_modes, const double* k, const double* z1, const double* z2,
               double scale, double mean, int lognormal, double* out, double* seconds) {
    double *d_pos = nullptr, *d_k = nullptr, *d_z1 = nullptr, *d_z2 = nullptr, *d_out = nullptr;
    int rc = SC_OK;
    auto body = [&]() -> int {
        SC_TRY(sc_alloc(ctx, &d_pos, (size_t)n_points * 3));
        SC_TRY(sc_alloc(ctx, &d_k, (size_t)n_modes * 3));
        SC_TRY(sc_alloc(ctx, &d_z1, (size_t)n_modes));
        SC_TRY(sc_alloc(ctx, &d_z2, (size_t)n_modes));
        SC_TRY(sc_alloc(ctx, &d_out, (size_t)n_points));
        cudaStream_t st = ctx->stream;
        SC_CUDA(ctx, cudaMemcpyAsync(d_pos, pos, sizeof(double) * 3 * n_points, cudaMemcpyHostToDevice, st));
        SC_CUDA(ctx, cudaMemcpyAsync(d_k, k, sizeof(double) * 3 * n_modes, cudaMemcpyHostToDevice, st));
        SC_CUDA(ctx, cudaMemcpyAsync(d_z1, z1, sizeof(double) * n_modes, cudaMemcpyHostToDevice, st));
        SC_CUDA(ctx, cudaMemcpyAsync(d_z2, z2, sizeof(double) * n_modes, cudaMemcpyHostToDevice, st));
        sc_gpu_timer timer(st);
        timer.start();
        const unsigned grid = (unsigned)((n_points + SRF_TPB - 1) / SRF_TPB);
        k_srf<<<grid, SRF_TPB, 0, st>>>(d_pos, n_points, d_k, d_z1, d_z2, n_modes, scale, mean, lognormal, d_out);
        SC_CHECK_LAUNCH(ctx);
        timer.stop();
        SC_CUDA(ctx, cudaMemcpyAsync(out, d_out, sizeof(double) * n_points, cudaMemcpyDeviceToHost, st));
        SC_CUDA(ctx, cudaStreamSynchronize(st));
        if (seconds) *seconds = timer.ms() * 1e-3;
        return SC_OK;
    };
    rc = body();
    sc_free(&d_pos); sc_free(&d_k); sc_free(&d_z1); sc_free(&d_z2); sc_free(&d_out);
    return rc;
}
