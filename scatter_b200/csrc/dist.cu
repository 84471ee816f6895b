// Multi-GPU plumbing: one process per GPU, NCCL point-to-point halo exchange over NVLink/NVSwitch.
//
// The reference is serial; the domain decomposition is this build's own.  NCCL is bound at run time
// with dlopen (the torch-bundled libnccl.so.2 is already mapped into a process that imported torch), so the library
// has no link-time NCCL dependency and single-GPU use needs no NCCL at all.
#include <dlfcn.h>
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

constexpr int NCCL_DOUBLE = 8, NCCL_SUM = 0;
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

void dist_destroy(sc_ctx* ctx) {
    if (ctx->comm && g_loaded) g_api.comm_destroy(ctx->comm);
    ctx->comm = nullptr;
}

int dist_halo(sc_ctx* ctx, double* d_x, cudaStream_t s) {
    if (ctx->world <= 1 || ctx->n_nbr_ranks == 0) return SC_OK;
    const int64_t ns = ctx->send_ptr.back(), nr = ctx->recv_ptr.back();
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
