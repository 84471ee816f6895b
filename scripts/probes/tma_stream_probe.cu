// Measurement aid (not product code): how fast can a persistent TMA ring stream HBM on this GPU, as a function of tile size,
// ring depth and CTAs per SM?  One producer lane issues cp.async.bulk copies of consecutive tiles (grid-stride), NC consumer
// warps wait for the tile and release the stage at once.  Build: nvcc -O3 -arch=sm_100a tma_stream_probe.cu -o tma_stream_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// split: number of bulk copies one tile is cut into
__global__ void k_stream(const char* __restrict__ src, int64_t n_tiles, int tile_bytes, int stages, int nc, int split, double* sink, int lookup, const int64_t* __restrict__ offs) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* bar_empty = bar_full + 16;
    unsigned char* buf = smem + 256;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], nc); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t G = gridDim.x;
    if (warp == nc) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += G) {
                const int64_t off = lookup ? offs[t] : t * (int64_t)tile_bytes;      // lookup: dependent global load per tile
                mbar_wait(&bar_empty[stage], phase ^ 1u);
                mbar_expect_tx(&bar_full[stage], tile_bytes);
                const int part = tile_bytes / split;
                for (int k = 0; k < split; ++k)
                    tma_load_1d(buf + (size_t)stage * tile_bytes + (size_t)k * part, src + off + (int64_t)k * part, part, &bar_full[stage]);
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        int stage = 0; uint32_t phase = 0;
        double acc = 0.0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += G) {
            mbar_wait(&bar_full[stage], phase);
            acc += reinterpret_cast<const double*>(buf + (size_t)stage * tile_bytes)[lane];
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[stage]);
            if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        if (acc == 1.2345e300) *sink = acc;
    }
}
__global__ void k_fill_random(double* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long z = i * 0x9E3779B97F4A7C15ull + 0x1234567ull;
        z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
        p[i] = (double)(long long)z * 1e-3;
    }
}
int main() {
    const size_t total = (size_t)32 << 30;
    char* d; double* sink;
    if (cudaMalloc(&d, total) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 8);
    cudaMemset(d, 0, total);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    // the SpMV's geometry: 31104-byte tiles (16 nodes x 243 values), 2 stages, 2 CTAs per SM; vary the start alignment and
    // whether the producer has to look the offset up first
    const int tile_bytes = 31104, stages = 2, cps = 2, nc = 8;
    const int64_t n_tiles = (total - 4096) / tile_bytes;
    int64_t* offs; cudaMalloc(&offs, n_tiles * 8);
    int64_t* h = (int64_t*)malloc(n_tiles * 8);
    const size_t sm_bytes = 256 + (size_t)stages * (tile_bytes + 128);
    cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_bytes);
    for (int64_t t = 0; t < n_tiles; ++t) h[t] = t * (int64_t)tile_bytes;
    cudaMemcpy(offs, h, n_tiles * 8, cudaMemcpyHostToDevice);
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) { k_fill_random<<<148 * 8, 256>>>((double*)d, total / 8); cudaDeviceSynchronize(); }
        // sustained: 300 back-to-back launches (~1.5 s); bandwidth of the first and the last 20
        float first = 0, last = 0;
        for (int rep = 0; rep < 300; ++rep) {
            cudaEventRecord(e0);
            k_stream<<<sms * cps, (nc + 1) * 32, sm_bytes>>>(d, n_tiles, tile_bytes, stages, nc, 1, sink, 1, offs);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep < 20) first += ms; if (rep >= 280) last += ms;
        }
        printf("%s data, 300 launches back to back: first 20 -> %6.0f GB/s, last 20 -> %6.0f GB/s\n", pass ? "random" : "zero  ",
               20.0 * n_tiles * tile_bytes / (first * 1e-3) / 1e9, 20.0 * n_tiles * tile_bytes / (last * 1e-3) / 1e9);
    }
    return 0;
}
