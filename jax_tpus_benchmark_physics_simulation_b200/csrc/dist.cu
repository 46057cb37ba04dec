// dist.cu — multi-GPU plumbing (new; the reference script is single-device).
//
// One process per GPU.  NCCL (over NVLink 5 / NVSwitch) is used for what happens ONCE per call:
// bootstrap (exchange of CUDA-IPC handles), replication of the final state, and the single
// all-reduce of the energy trace.  The per-step position exchange is NOT an NCCL call: the
// all-pairs kernel's integrate epilogue stores each new position straight into every peer's
// next-position buffer through the IPC-mapped peer pointers set up here, and a per-step arrival
// word per rank replaces the collective's synchronisation (allpairs.cu).
//
// libnccl is dlopen()ed so that single-GPU users need no NCCL at all; inside a torch process the
// soname resolves to the copy torch already loaded (one NCCL per process).
#include "ljmd_internal.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <cstring>

namespace ljmd {

namespace {

// types and enum values come from the NCCL header; the functions are resolved at run time
struct Nccl {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId)    GetUniqueId = nullptr;
    decltype(&ncclCommInitRank)   CommInitRank = nullptr;
    decltype(&ncclCommDestroy)    CommDestroy = nullptr;
    decltype(&ncclAllGather)      AllGather = nullptr;
    decltype(&ncclAllReduce)      AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

Nccl g_nccl;     // published only when every symbol has been resolved

int load_nccl() {
    if (g_nccl.lib) return 0;
    Nccl n;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (n.lib) break;
    }
    if (!n.lib) { set_error("cannot dlopen libnccl.so.2: %s", dlerror()); return LJMD_E_NCCL; }
#define LJ_SYM(field, name)                                                        \
    *(void**)(&n.field) = dlsym(n.lib, name);                                      \
    if (!n.field) { set_error("libnccl lacks %s", name); dlclose(n.lib); return LJMD_E_NCCL; }
    LJ_SYM(GetUniqueId, "ncclGetUniqueId")
    LJ_SYM(CommInitRank, "ncclCommInitRank")
    LJ_SYM(CommDestroy, "ncclCommDestroy")
    LJ_SYM(AllGather, "ncclAllGather")
    LJ_SYM(AllReduce, "ncclAllReduce")
    LJ_SYM(GetErrorString, "ncclGetErrorString")
#undef LJ_SYM
    g_nccl = n;
    return 0;
}

#define LJ_NCCL(expr)                                                              \
    do {                                                                           \
        int _r = (expr);                                                           \
        if (_r != 0) {                                                             \
            set_error("%s failed: %s", #expr, g_nccl.GetErrorString((ncclResult_t)_r)); \
            return LJMD_E_NCCL;                                                    \
        }                                                                          \
    } while (0)

}  // namespace

struct Dist {
    ncclComm_t comm = nullptr;
    float* scratch = nullptr;       // one float: payload of the stream barrier
    void* peer_mapped[LJMD_MAX_RANKS] = {};
    int n_mapped = 0;
};

int dist_init(ljmd_handle* h, const void* nccl_unique_id) {
    int r = load_nccl();
    if (r) return r;
    Dist* d = new Dist();
    h->dist = d;
    static_assert(sizeof(ncclUniqueId) == 128, "ljmd.h promises a 128-byte unique id");
    ncclUniqueId id;
    memcpy(&id, nccl_unique_id, sizeof(id));
    LJ_NCCL(g_nccl.CommInitRank(&d->comm, h->nranks, id, h->rank));
    LJ_CUDA(cudaMalloc(&d->scratch, sizeof(float)));
    LJ_CUDA(cudaMemset(d->scratch, 0, sizeof(float)));
    return 0;
}

// Cross-rank barrier IN STREAM ORDER (a one-float all-reduce): whatever a rank enqueues after it
// starts only when every rank has reached this point of its stream.  The persistent kernels spin
// on words their peers write, with a bail-out timer: the ranks have to enter them together even
// if their host threads are seconds apart.
int dist_barrier(ljmd_handle* h) {
    Dist* d = h->dist;
    if (!d) return 0;
    LJ_NCCL(g_nccl.AllReduce(d->scratch, d->scratch, 1, ncclFloat32, ncclSum, d->comm, h->stream));
    return 0;
}

void dist_destroy(ljmd_handle* h) {
    Dist* d = h->dist;
    if (!d) return;
    for (int q = 0; q < LJMD_MAX_RANKS; ++q)
        if (d->peer_mapped[q]) cudaIpcCloseMemHandle(d->peer_mapped[q]);
    if (d->scratch) cudaFree(d->scratch);
    if (d->comm) g_nccl.CommDestroy(d->comm);
    delete d;
    h->dist = nullptr;
}

// Exchange the CUDA-IPC handle of `local_base` (the base of one cudaMalloc allocation) with all
// ranks through NCCL and map every peer's allocation into this process.
int dist_share(ljmd_handle* h, void* local_base, void** peer_bases) {
    Dist* d = h->dist;
    if (!d) { set_error("dist_share without a communicator"); return LJMD_E_STATE; }
    cudaIpcMemHandle_t mine;
    LJ_CUDA(cudaIpcGetMemHandle(&mine, local_base));
    const size_t hs = sizeof(cudaIpcMemHandle_t);
    cudaIpcMemHandle_t all[LJMD_MAX_RANKS];
    char* dev = nullptr;
    LJ_CUDA(cudaMalloc(&dev, hs * h->nranks));
    auto exchange = [&]() -> int {          // (dev is released on every path below)
        LJ_CUDA(cudaMemcpy(dev + hs * h->rank, &mine, hs, cudaMemcpyHostToDevice));
        LJ_NCCL(g_nccl.AllGather(dev + hs * h->rank, dev, hs, ncclInt8, d->comm, h->stream));
        LJ_CUDA(cudaStreamSynchronize(h->stream));
        LJ_CUDA(cudaMemcpy(all, dev, hs * h->nranks, cudaMemcpyDeviceToHost));
        return 0;
    };
    const int r = exchange();
    cudaFree(dev);
    if (r) return r;
    for (int q = 0; q < h->nranks; ++q) {
        if (q == h->rank) { peer_bases[q] = local_base; continue; }
        void* pq = nullptr;
        // (mappings opened so far are recorded in d->peer_mapped and closed by dist_destroy)
        LJ_CUDA(cudaIpcOpenMemHandle(&pq, all[q], cudaIpcMemLazyEnablePeerAccess));
        peer_bases[q] = pq;
        d->peer_mapped[q] = pq;
    }
    return 0;
}

int dist_allgather(ljmd_handle* h, void* buf, size_t bytes_per_rank) {
    Dist* d = h->dist;
    char* b = reinterpret_cast<char*>(buf);
    LJ_NCCL(g_nccl.AllGather(b + bytes_per_rank * h->rank, b, bytes_per_rank, ncclInt8, d->comm, h->stream));
    return 0;
}

int dist_allgather_from(ljmd_handle* h, const void* send, void* recv, size_t bytes_per_rank) {
    Dist* d = h->dist;
    LJ_NCCL(g_nccl.AllGather(send, recv, bytes_per_rank, ncclInt8, d->comm, h->stream));
    return 0;
}

int dist_allreduce_f32(ljmd_handle* h, float* buf, size_t n) {
    Dist* d = h->dist;
    LJ_NCCL(g_nccl.AllReduce(buf, buf, n, ncclFloat32, ncclSum, d->comm, h->stream));
    return 0;
}

}  // namespace ljmd

using namespace ljmd;

extern "C" {

int ljmd_get_unique_id(void* id128) {
    if (!id128) { set_error("null argument"); return LJMD_E_INVALID; }
    int r = load_nccl();
    if (r) return r;
    ncclUniqueId id;
    LJ_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int ljmd_create_dist(ljmd_t** out, const ljmd_params* p, const void* nccl_unique_id, int32_t rank,
                     int32_t nranks) {
    if (nranks == 1) return create_common(out, p, 0, 1, nullptr);
    return create_common(out, p, rank, nranks, nccl_unique_id);
}

}  // extern "C"
