// Multi-GPU plumbing: one process per GPU; halo exchange over NVLink/NVSwitch either through peer memory (default: the
// sender's pack kernel stores straight into a receive window in the neighbour's HBM and raises a flag there, the
// receiver's unpack kernel waits on its own flags -- two small kernels per exchange, no staging copy, no library call)
// or through NCCL point-to-point calls (fallback, and the transport of the one-off set-up messages and all-reduces).
//
// The reference is serial; the domain decomposition is this build's own.  NCCL is bound at run time
// with dlopen (the torch-bundled libnccl.so.2 is already mapped into a process that imported torch), so the library
// has no link-time NCCL dependency and single-GPU use needs no NCCL at all.
#include <dlfcn.h>
#include <algorithm>
#include <cstring>
#include "common.h"

typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_getuid)(nccl_uid*);
typedef int (*fn_initrank)(void**, int, nccl_uid, int);
typedef int (*fn_destroy)(void*);
typedef int (*fn_sendrecv)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_group)(void);
typedef const char* (*fn_errstr)(int);

struct NcclApi {
    void* lib = nullptr;
    fn_getuid get_unique_id = nullptr;
    fn_initrank comm_init_rank = nullptr;
    fn_destroy comm_destroy = nullptr;
    fn_sendrecv send = nullptr;
    fn_sendrecv recv = nullptr;
    fn_allreduce all_reduce = nullptr;
    fn_group group_start = nullptr, group_end = nullptr;
    fn_errstr err_string = nullptr;
};

namespace {

constexpr int NCCL_DOUBLE = 8, NCCL_SUM = 0, NCCL_CHAR = 0;
NcclApi g_api;
bool g_loaded = false;

bool load_nccl(std::string& err) {
    if (g_loaded) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        g_api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_api.lib) break;
    }
    if (!g_api.lib) { err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
#define SYM(field, name, type) g_api.field = (type)dlsym(g_api.lib, name); if (!g_api.field) { err = "missing NCCL symbol " name; return false; }
    SYM(get_unique_id, "ncclGetUniqueId", fn_getuid)
    SYM(comm_init_rank, "ncclCommInitRank", fn_initrank)
    SYM(comm_destroy, "ncclCommDestroy", fn_destroy)
    SYM(send, "ncclSend", fn_sendrecv)
    SYM(recv, "ncclRecv", fn_sendrecv)
    SYM(all_reduce, "ncclAllReduce", fn_allreduce)
    SYM(group_start, "ncclGroupStart", fn_group)
    SYM(group_end, "ncclGroupEnd", fn_group)
    SYM(err_string, "ncclGetErrorString", fn_errstr)
#undef SYM
    g_loaded = true;
    return true;
}

__global__ void k_pack(const double* __restrict__ x, const int64_t* __restrict__ idx, double* __restrict__ buf, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) buf[i] = x[idx[i]];
}
__global__ void k_unpack(double* __restrict__ x, const int64_t* __restrict__ idx, const double* __restrict__ buf, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) x[idx[i]] = buf[i];
}


// ---- peer-memory halo exchange ------------------------------------------------------------------------------------
// Every rank owns a receive window [2][n_recv] (double buffered by the parity of the exchange counter) and one 64-bit
// flag per neighbour, both mapped into the neighbours' address spaces with CUDA IPC.  Exchange number e:
//   k_peer_push        every send entry i of neighbour k:  win_k[e & 1][off_k + i] = x[send_idx[i]]  (remote stores over
//                      NVLink), system-scope fence, and the last block to finish stores e into the neighbours' flags
//   k_peer_wait_unpack waits until all of the rank's own flags have reached e, then x[recv_idx[j]] = win[e & 1][j]
// Why two buffers suffice: a neighbour can only start exchange e + 2 (which overwrites buffer e & 1) after it has seen
// this rank's flag e + 1, and this rank raises that flag after -- in stream order -- its unpack of exchange e.
// All ranks run the same sequence of exchanges (the time loops are SPMD), so one counter per context is enough.
struct PeerMsg { cudaIpcMemHandle_t win, flags; int64_t off, slot, n_recv; };

__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256)
k_peer_push(const double* __restrict__ x, const int64_t* __restrict__ send_idx, const int64_t* __restrict__ send_ptr, int n_nbr,
            double* const* __restrict__ peer_win, const int64_t* __restrict__ peer_off, const int64_t* __restrict__ peer_nrecv,
            unsigned long long* const* __restrict__ peer_flag, unsigned long long epoch, unsigned int* __restrict__ counter) {
    const int64_t ns = send_ptr[n_nbr];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ns; i += (int64_t)gridDim.x * blockDim.x) {
        int k = 0;
        while (k + 1 < n_nbr && send_ptr[k + 1] <= i) ++k;
        peer_win[k][(epoch & 1ull) * peer_nrecv[k] + peer_off[k] + (i - send_ptr[k])] = x[send_idx[i]];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(counter, 1u);            // integer arrival count, not a float sum
        if (done == gridDim.x - 1) {
            __threadfence_system();
            for (int k = 0; k < n_nbr; ++k) {
                asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(peer_flag[k]), "l"(epoch) : "memory");
            }
            *counter = 0u;
        }
    }
}

__global__ void __launch_bounds__(256)
k_peer_wait_unpack(double* __restrict__ x, const int64_t* __restrict__ recv_idx, int64_t nr, const double* __restrict__ win,
                   const unsigned long long* __restrict__ flags, int n_nbr, unsigned long long epoch, int* __restrict__ timed_out) {
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int k = 0; k < n_nbr; ++k) {
            while (ld_flag(flags + k) < epoch) {
                if (clock64() - t0 > 40000000000ll) { *timed_out = 1; break; }     // ~20 s: a neighbour died; the host reports it
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    const double* w = win + (epoch & 1ull) * nr;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < nr; j += (int64_t)gridDim.x * blockDim.x) {
        double v;
        asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(w + j) : "memory");   // written by a peer: no cached copy
        x[recv_idx[j]] = v;
    }
}

#define SC_NCCL(ctx, call)                                                                                 \
    do {                                                                                                   \
        int _r = (call);                                                                                   \
        if (_r != 0) return sc_fail((ctx), SC_ERR_NCCL, "%s failed: %s", #call, g_api.err_string(_r));     \
    } while (0)

}  // namespace

int dist_unique_id(void* out) {
    std::string err;
    if (!load_nccl(err)) { sc_set_global_error(err.c_str()); return SC_ERR_NCCL; }
    nccl_uid id;
    int r = g_api.get_unique_id(&id);
    if (r != 0) { sc_set_global_error(g_api.err_string(r)); return SC_ERR_NCCL; }
    std::memcpy(out, &id, 128);
    return SC_OK;
}

int dist_init(sc_ctx* ctx, int rank, int world, const void* idbytes) {
    if (world <= 1) { ctx->rank = 0; ctx->world = 1; return SC_OK; }
    std::string err;
    if (!load_nccl(err)) return sc_fail(ctx, SC_ERR_NCCL, "%s", err.c_str());
    nccl_uid id;
    std::memcpy(&id, idbytes, 128);
    SC_CUDA(ctx, cudaSetDevice(ctx->device));
    SC_NCCL(ctx, g_api.comm_init_rank(&ctx->comm, world, id, rank));
    ctx->nccl = &g_api;
    ctx->rank = rank;
    ctx->world = world;
    return SC_OK;
}

int dist_allreduce_sum(sc_ctx* ctx, double* d_vals, int n, cudaStream_t s);

// one-off: windows, flags, IPC handles exchanged with the neighbours over NCCL, peer tables on the device
int dist_peer_setup(sc_ctx* ctx) {
    ctx->peer_tried = true;
    ctx->peer_ok = false;
    if (ctx->world <= 1 || ctx->n_nbr_ranks == 0 || ctx->no_peer_halo) return SC_OK;
    const int nn = ctx->n_nbr_ranks;
    const int64_t nr = ctx->recv_ptr.back();
    cudaStream_t st = ctx->stream;
    bool ok = true;
    SC_TRY(sc_alloc(ctx, &ctx->d_peer_win, (size_t)std::max<int64_t>(2 * nr, 1)));
    SC_TRY(sc_alloc(ctx, &ctx->d_peer_flags, (size_t)nn + 2));           // [nn] flags, then the arrival counter and the time-out mark
    SC_CUDA(ctx, cudaMemset(ctx->d_peer_flags, 0, ((size_t)nn + 2) * sizeof(unsigned long long)));
    std::vector<PeerMsg> out(nn), in(nn);
    cudaIpcMemHandle_t hw, hf;
    if (cudaIpcGetMemHandle(&hw, ctx->d_peer_win) != cudaSuccess || cudaIpcGetMemHandle(&hf, ctx->d_peer_flags) != cudaSuccess) {
        cudaGetLastError();
        ok = false;
        std::memset(&hw, 0, sizeof(hw)); std::memset(&hf, 0, sizeof(hf));
    }
    for (int k = 0; k < nn; ++k) { out[k].win = hw; out[k].flags = hf; out[k].off = ctx->recv_ptr[k]; out[k].slot = k; out[k].n_recv = nr; }
    char *d_out = nullptr, *d_in = nullptr;
    SC_TRY(sc_alloc(ctx, &d_out, sizeof(PeerMsg) * nn));
    SC_TRY(sc_alloc(ctx, &d_in, sizeof(PeerMsg) * nn));
    int rc = SC_OK;
    auto body = [&]() -> int {
        SC_CUDA(ctx, cudaMemcpyAsync(d_out, out.data(), sizeof(PeerMsg) * nn, cudaMemcpyHostToDevice, st));
        SC_NCCL(ctx, g_api.group_start());
        for (int k = 0; k < nn; ++k) {
            SC_NCCL(ctx, g_api.send(d_out + sizeof(PeerMsg) * k, sizeof(PeerMsg), NCCL_CHAR, ctx->nbr_rank[k], ctx->comm, st));
            SC_NCCL(ctx, g_api.recv(d_in + sizeof(PeerMsg) * k, sizeof(PeerMsg), NCCL_CHAR, ctx->nbr_rank[k], ctx->comm, st));
        }
        SC_NCCL(ctx, g_api.group_end());
        SC_CUDA(ctx, cudaMemcpyAsync(in.data(), d_in, sizeof(PeerMsg) * nn, cudaMemcpyDeviceToHost, st));
        SC_CUDA(ctx, cudaStreamSynchronize(st));
        return SC_OK;
    };
    rc = body();
    sc_free(&d_out); sc_free(&d_in);
    SC_TRY(rc);
    std::vector<double*> pw(nn, nullptr);
    std::vector<unsigned long long*> pf(nn, nullptr);
    std::vector<int64_t> poff(nn), pnr(nn);
    ctx->peer_maps.clear();
    for (int k = 0; k < nn && ok; ++k) {
        void *w = nullptr, *f = nullptr;
        if (cudaIpcOpenMemHandle(&w, in[k].win, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        ctx->peer_maps.push_back(w);
        if (cudaIpcOpenMemHandle(&f, in[k].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        ctx->peer_maps.push_back(f);
        pw[k] = static_cast<double*>(w);
        pf[k] = static_cast<unsigned long long*>(f) + in[k].slot;
        poff[k] = in[k].off; pnr[k] = in[k].n_recv;
    }
    // every rank must take the same path
    double* d_ok = nullptr;
    SC_TRY(sc_alloc(ctx, &d_ok, 1));
    const double mine = ok ? 0.0 : 1.0;
    double sum = 0.0;
    rc = [&]() -> int {
        SC_CUDA(ctx, cudaMemcpyAsync(d_ok, &mine, sizeof(double), cudaMemcpyHostToDevice, st));
        SC_TRY(dist_allreduce_sum(ctx, d_ok, 1, st));
        SC_CUDA(ctx, cudaMemcpyAsync(&sum, d_ok, sizeof(double), cudaMemcpyDeviceToHost, st));
        SC_CUDA(ctx, cudaStreamSynchronize(st));
        return SC_OK;
    }();
    sc_free(&d_ok);
    SC_TRY(rc);
    if (sum != 0.0) return SC_OK;                                        // somebody could not map a window: NCCL path for all
    SC_TRY(sc_alloc(ctx, &ctx->d_peer_win_ptr, (size_t)nn));
    SC_TRY(sc_alloc(ctx, &ctx->d_peer_flag_ptr, (size_t)nn));
    SC_TRY(sc_alloc(ctx, &ctx->d_peer_off, (size_t)2 * nn));
    SC_TRY(sc_alloc(ctx, &ctx->d_send_ptr, (size_t)nn + 1));
    SC_CUDA(ctx, cudaMemcpy(ctx->d_peer_win_ptr, pw.data(), sizeof(double*) * nn, cudaMemcpyHostToDevice));
    SC_CUDA(ctx, cudaMemcpy(ctx->d_peer_flag_ptr, pf.data(), sizeof(unsigned long long*) * nn, cudaMemcpyHostToDevice));
    SC_CUDA(ctx, cudaMemcpy(ctx->d_peer_off, poff.data(), sizeof(int64_t) * nn, cudaMemcpyHostToDevice));
    SC_CUDA(ctx, cudaMemcpy(ctx->d_peer_off + nn, pnr.data(), sizeof(int64_t) * nn, cudaMemcpyHostToDevice));
    SC_CUDA(ctx, cudaMemcpy(ctx->d_send_ptr, ctx->send_ptr.data(), sizeof(int64_t) * (nn + 1), cudaMemcpyHostToDevice));
    ctx->peer_epoch = 0;
    ctx->peer_ok = true;
    return SC_OK;
}

int dist_peer_check(sc_ctx* ctx) {
    if (!ctx->peer_ok) return SC_OK;
    int mark = 0;
    SC_CUDA(ctx, cudaMemcpy(&mark, reinterpret_cast<int*>(ctx->d_peer_flags + ctx->n_nbr_ranks + 1), sizeof(int), cudaMemcpyDeviceToHost));
    if (mark != 0) return sc_fail(ctx, SC_ERR_NCCL, "halo exchange timed out waiting for a neighbouring rank (exchange %llu)", ctx->peer_epoch);
    return SC_OK;
}

void dist_peer_release(sc_ctx* ctx) {
    for (void* m : ctx->peer_maps) cudaIpcCloseMemHandle(m);
    ctx->peer_maps.clear();
    sc_free(&ctx->d_peer_win); sc_free(&ctx->d_peer_flags); sc_free(&ctx->d_peer_win_ptr); sc_free(&ctx->d_peer_flag_ptr);
    sc_free(&ctx->d_peer_off); sc_free(&ctx->d_send_ptr);
    ctx->peer_ok = false; ctx->peer_tried = false;
}

void dist_destroy(sc_ctx* ctx) {
    dist_peer_release(ctx);
    if (ctx->comm && g_loaded) g_api.comm_destroy(ctx->comm);
    ctx->comm = nullptr;
}

int dist_halo(sc_ctx* ctx, double* d_x, cudaStream_t s) {
    if (ctx->world <= 1 || ctx->n_nbr_ranks == 0) return SC_OK;
    const int64_t ns = ctx->send_ptr.back(), nr = ctx->recv_ptr.back();
    if (!ctx->peer_tried) SC_TRY(dist_peer_setup(ctx));
    if (ctx->peer_ok) {
        const unsigned long long e = ++ctx->peer_epoch;
        const int nn = ctx->n_nbr_ranks;
        unsigned int* counter = reinterpret_cast<unsigned int*>(ctx->d_peer_flags + nn);
        int* timed_out = reinterpret_cast<int*>(ctx->d_peer_flags + nn + 1);
        const unsigned gs = (unsigned)std::min<int64_t>(std::max<int64_t>((ns + 255) / 256, 1), 2 * (int64_t)ctx->sm_count);
        k_peer_push<<<gs, 256, 0, s>>>(d_x, ctx->d_send_idx, ctx->d_send_ptr, nn, ctx->d_peer_win_ptr, ctx->d_peer_off, ctx->d_peer_off + nn,
                                       ctx->d_peer_flag_ptr, e, counter);
        SC_CHECK_LAUNCH(ctx);
        const unsigned gr = (unsigned)std::min<int64_t>(std::max<int64_t>((nr + 255) / 256, 1), 2 * (int64_t)ctx->sm_count);
        k_peer_wait_unpack<<<gr, 256, 0, s>>>(d_x, ctx->d_recv_idx, nr, ctx->d_peer_win, ctx->d_peer_flags, nn, e, timed_out);
        SC_CHECK_LAUNCH(ctx);
        return SC_OK;
    }
    if (ns > 0) {
        k_pack<<<(unsigned)((ns + 255) / 256), 256, 0, s>>>(d_x, ctx->d_send_idx, ctx->d_send_buf, ns);
        SC_CHECK_LAUNCH(ctx);
    }
    SC_NCCL(ctx, g_api.group_start());
    for (int k = 0; k < ctx->n_nbr_ranks; ++k) {
        const int64_t s0 = ctx->send_ptr[k], s1 = ctx->send_ptr[k + 1];
        const int64_t r0 = ctx->recv_ptr[k], r1 = ctx->recv_ptr[k + 1];
        if (s1 > s0) SC_NCCL(ctx, g_api.send(ctx->d_send_buf + s0, (size_t)(s1 - s0), NCCL_DOUBLE, ctx->nbr_rank[k], ctx->comm, s));
        if (r1 > r0) SC_NCCL(ctx, g_api.recv(ctx->d_recv_buf + r0, (size_t)(r1 - r0), NCCL_DOUBLE, ctx->nbr_rank[k], ctx->comm, s));
    }
    SC_NCCL(ctx, g_api.group_end());
    if (nr > 0) {
        k_unpack<<<(unsigned)((nr + 255) / 256), 256, 0, s>>>(d_x, ctx->d_recv_idx, ctx->d_recv_buf, nr);
        SC_CHECK_LAUNCH(ctx);
    }
    return SC_OK;
}

int dist_allreduce_sum(sc_ctx* ctx, double* d_vals, int n, cudaStream_t s) {
    if (ctx->world <= 1) return SC_OK;
    SC_NCCL(ctx, g_api.all_reduce(d_vals, d_vals, (size_t)n, NCCL_DOUBLE, NCCL_SUM, ctx->comm, s));
    return SC_OK;
}
