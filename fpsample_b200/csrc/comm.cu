// comm.cu -- the ONE exchange step of the multi-GPU path (kernel K6): the index arrays of every shard are
// gathered to rank 0 over NCCL (NVLink 5 / NVSwitch).  Clouds are independent, so nothing else ever crosses GPUs.
//
//   * one process per GPU (torchrun): fps_b200_comm_unique_id on rank 0, the 128 bytes travel through the launcher's
//     rendezvous (fpsample_b200/dist.py: a TCP exchange on MASTER_ADDR), fps_b200_comm_init on every rank (ncclCommInitRank);
//   * one process, several GPUs: fps_b200_comm_init_local (ncclCommInitAll), one communicator and one stream per device.
// The gather itself: every rank narrows its [nb][k] uint64 indices to uint32 on its device (indices < 2^32 is checked by the
// C ABI), ranks 1.. ncclSend them, rank 0 ncclRecv's every shard at its offset inside one ncclGroup, widens on the device
// and copies [n_clouds][k] uint64 to the host once.
//
// NCCL is bound at run time (dlopen "libnccl.so.2": the copy a host process such as PyTorch already loaded, else the system
// one), so the library itself has no link-time dependency on it and loads on hosts without NCCL -- the comm entry points then
// fail loudly with FPS_ERR_NCCL.
#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/fps_b200.h"
#include "engine.h"

namespace fps {

// ---- the slice of the NCCL 2.x ABI this file uses (nccl.h: ncclUniqueId = 128 opaque bytes, ncclUint32 = 3) -------------
typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclUint32 = 3;

struct NcclApi {
    void *so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    bool ok = false;
};
static NcclApi g_nccl;
static std::once_flag g_nccl_once;
static char g_nccl_path[1024] = "";   // fps_b200_nccl_library
static bool g_nccl_bound = false;

static const NcclApi &nccl() {
    std::call_once(g_nccl_once, [] {
        // Which libnccl: (1) the one this process already has (a host framework such as PyTorch); (2) the one the host named
        // (fps_b200_nccl_library / FPS_B200_NCCL_LIB -- the python package points at the pip-installed nvidia-nccl wheel, the
        // copy PyTorch will load LATER if it is imported after us: two different libnccl.so.2 in one process do not work, the
        // second user would be handed the first one's symbols); (3) the system's.
        void *so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);
        const char *named = g_nccl_path[0] ? g_nccl_path : getenv("FPS_B200_NCCL_LIB");
        if (!so && named && named[0]) so = dlopen(named, RTLD_NOW | RTLD_GLOBAL);
        if (!so) so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!so) so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!so) return;
        g_nccl.so = so;
        g_nccl_bound = true;
#define BIND(field, sym) g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(so, sym))
        BIND(GetUniqueId, "ncclGetUniqueId");
        BIND(CommInitRank, "ncclCommInitRank");
        BIND(CommInitAll, "ncclCommInitAll");
        BIND(CommDestroy, "ncclCommDestroy");
        BIND(Send, "ncclSend");
        BIND(Recv, "ncclRecv");
        BIND(GroupStart, "ncclGroupStart");
        BIND(GroupEnd, "ncclGroupEnd");
        BIND(GetErrorString, "ncclGetErrorString");
        BIND(GetVersion, "ncclGetVersion");
#undef BIND
        g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommInitAll && g_nccl.CommDestroy && g_nccl.Send &&
                    g_nccl.Recv && g_nccl.GroupStart && g_nccl.GroupEnd && g_nccl.GetErrorString;
    });
    return g_nccl;
}

void comm_set_err(const char *fmt, ...);   // capi.cu: the thread-local error text behind fps_b200_last_error

// ---- one endpoint per device of this process ------------------------------------------------------------------------------
struct Endpoint {
    int dev = -1, rank = -1;
    ncclComm_t comm = nullptr;
    cudaStream_t st = nullptr;
    void *wire = nullptr;      // uint32 [.][k]: this rank's shard (ranks 1..) or every shard (rank 0)
    size_t wire_cap = 0;
    void *stage = nullptr;     // device copy of host-resident local indices / rank 0's widened result
    size_t stage_cap = 0;
};
static std::mutex g_comm_mu;
static std::vector<Endpoint> g_eps;   // one entry (multi-process) or one per local device (single process)
static int g_nranks = 0;

static int ensure(void **p, size_t *cap, size_t bytes) {
    if (bytes <= *cap) return FPS_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(p, bytes + 256);
    if (e != cudaSuccess) {
        comm_set_err("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return FPS_ERR_CUDA + (int)e;
    }
    *cap = bytes + 256;
    return FPS_OK;
}

#define NCK(call)                                                                                   \
    do {                                                                                            \
        ncclResult_t r__ = (call);                                                                  \
        if (r__ != 0) {                                                                             \
            comm_set_err("%s failed: %s (%s:%d)", #call, nccl().GetErrorString(r__), __FILE__, __LINE__); \
            return FPS_ERR_NCCL;                                                                    \
        }                                                                                           \
    } while (0)
#define CCK(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            comm_set_err("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return FPS_ERR_CUDA + (int)e__;                                                         \
        }                                                                                           \
    } while (0)

__global__ void narrow_kernel(const u64 *__restrict__ src, u32 *__restrict__ dst, size_t count) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) dst[i] = (u32)src[i];
}
__global__ void widen_kernel(const u32 *__restrict__ src, u64 *__restrict__ dst, size_t count) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) dst[i] = (u64)src[i];
}
static unsigned blocks_for(size_t count) {
    size_t b = (count + 255) / 256;
    return (unsigned)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

static void shard_of(size_t n_clouds, int world, int rank, size_t *b0, size_t *nb) {   // the rule of run_batch (capi.cu)
    const size_t base = n_clouds / (size_t)world, rem = n_clouds % (size_t)world;
    *nb = base + ((size_t)rank < rem ? 1 : 0);
    *b0 = (size_t)rank * base + ((size_t)rank < rem ? (size_t)rank : rem);
}

struct DevRestore {
    int prev = -1;
    DevRestore() {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1, cudaGetLastError();
    }
    ~DevRestore() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Gather over the endpoints of THIS process.  locals[i] = endpoint i's [nb_i][k] uint64 indices, host or device pointer
// (device: on that endpoint's device; `after[i]`, if not null, is a stream whose queued work produces them).
int comm_gather(const u64 *const *locals, const cudaStream_t *after, size_t k, size_t n_clouds, u64 *out_rank0) {
    std::lock_guard<std::mutex> lk(g_comm_mu);
    const NcclApi &N = nccl();
    if (g_eps.empty() || !N.ok) {
        comm_set_err("no communicator: call fps_b200_comm_init / fps_b200_comm_init_local first (NCCL %s)", N.ok ? "loaded" : "not found");
        return FPS_ERR_NCCL;
    }
    DevRestore restore;
    const int world = g_nranks;
    std::vector<cudaEvent_t> evs;
    int rc = FPS_OK;
    // 1. every endpoint: its shard as uint32 in its wire buffer (rank 0: at its offset of the full buffer)
    for (size_t i = 0; i < g_eps.size() && rc == FPS_OK; ++i) {
        Endpoint &ep = g_eps[i];
        size_t b0, nb;
        shard_of(n_clouds, world, ep.rank, &b0, &nb);
        CCK(cudaSetDevice(ep.dev));
        const size_t mine = nb * k, all = n_clouds * k;
        if ((rc = ensure(&ep.wire, &ep.wire_cap, (ep.rank == 0 ? all : mine) * sizeof(u32)))) break;
        if (after && after[i]) {   // order behind the sampler's stream without a host synchronisation
            cudaEvent_t ev;
            CCK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            evs.push_back(ev);
            CCK(cudaEventRecord(ev, after[i]));
            CCK(cudaStreamWaitEvent(ep.st, ev, 0));
        }
        const u64 *src = locals[i];
        cudaPointerAttributes at;
        const bool on_dev = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeDevice;
        if (!on_dev) {
            cudaGetLastError();
            if ((rc = ensure(&ep.stage, &ep.stage_cap, (ep.rank == 0 ? all : mine) * sizeof(u64)))) break;
            CCK(cudaMemcpyAsync(ep.stage, src, mine * sizeof(u64), cudaMemcpyHostToDevice, ep.st));
            src = static_cast<const u64 *>(ep.stage);
        }
        if (mine) narrow_kernel<<<blocks_for(mine), 256, 0, ep.st>>>(src, static_cast<u32 *>(ep.wire) + (ep.rank == 0 ? b0 * k : 0), mine);
        count_launch();
    }
    // 2. one NCCL group: ranks 1.. send, rank 0 receives every shard at its offset
    if (rc == FPS_OK && world > 1) {
        NCK(N.GroupStart());
        for (Endpoint &ep : g_eps) {
            if (ep.rank == 0) {
                for (int r = 1; r < world; ++r) {
                    size_t b0, nb;
                    shard_of(n_clouds, world, r, &b0, &nb);
                    if (nb) NCK(N.Recv(static_cast<u32 *>(ep.wire) + b0 * k, nb * k, kNcclUint32, r, ep.comm, ep.st));
                }
            } else {
                size_t b0, nb;
                shard_of(n_clouds, world, ep.rank, &b0, &nb);
                if (nb) NCK(N.Send(ep.wire, nb * k, kNcclUint32, 0, ep.comm, ep.st));
            }
        }
        NCK(N.GroupEnd());
    }
    // 3. rank 0 (if it lives in this process): widen on the device, one copy to the host
    for (Endpoint &ep : g_eps) {
        if (rc != FPS_OK) break;
        CCK(cudaSetDevice(ep.dev));
        if (ep.rank == 0) {
            if (!out_rank0) {
                comm_set_err("rank 0 needs an output buffer");
                rc = FPS_ERR_ARG;
                break;
            }
            const size_t all = n_clouds * k;
            if ((rc = ensure(&ep.stage, &ep.stage_cap, all * sizeof(u64)))) break;
            widen_kernel<<<blocks_for(all), 256, 0, ep.st>>>(static_cast<const u32 *>(ep.wire), static_cast<u64 *>(ep.stage), all);
            count_launch();
            CCK(cudaMemcpyAsync(out_rank0, ep.stage, all * sizeof(u64), cudaMemcpyDeviceToHost, ep.st));
        }
    }
    for (Endpoint &ep : g_eps) {
        cudaSetDevice(ep.dev);
        cudaError_t e = cudaStreamSynchronize(ep.st);
        if (e != cudaSuccess && rc == FPS_OK) {
            comm_set_err("gather failed: %s", cudaGetErrorString(e));
            rc = FPS_ERR_CUDA + (int)e;
        }
    }
    for (cudaEvent_t ev : evs) cudaEventDestroy(ev);
    return rc;
}

int comm_world() { return g_nranks; }
int comm_local_endpoints() { return (int)g_eps.size(); }
int comm_endpoint_device(int i) { return g_eps[(size_t)i].dev; }
int comm_endpoint_rank(int i) { return g_eps[(size_t)i].rank; }

static void destroy_locked() {
    for (Endpoint &ep : g_eps) {
        cudaSetDevice(ep.dev);
        if (ep.comm && nccl().ok) nccl().CommDestroy(ep.comm);
        if (ep.st) cudaStreamDestroy(ep.st);
        if (ep.wire) cudaFree(ep.wire);
        if (ep.stage) cudaFree(ep.stage);
    }
    g_eps.clear();
    g_nranks = 0;
}

}  // namespace fps

using namespace fps;

extern "C" {

int fps_b200_comm_unique_id(void *id128) {
    const NcclApi &N = nccl();
    if (!N.ok || !id128) {
        comm_set_err(N.ok ? "null id buffer" : "NCCL not found (dlopen libnccl.so.2)");
        return N.ok ? FPS_ERR_ARG : FPS_ERR_NCCL;
    }
    ncclUniqueId id;
    NCK(N.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return FPS_OK;
}

int fps_b200_comm_init(const void *id128, int n_ranks, int rank) {
    const NcclApi &N = nccl();
    if (!N.ok) {
        comm_set_err("NCCL not found (dlopen libnccl.so.2)");
        return FPS_ERR_NCCL;
    }
    if (!id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) {
        comm_set_err("bad argument: need an id, n_ranks >= 1 and 0 <= rank < n_ranks");
        return FPS_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(g_comm_mu);
    destroy_locked();
    Endpoint ep;
    CCK(cudaGetDevice(&ep.dev));
    ep.rank = rank;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    NCK(N.CommInitRank(&ep.comm, n_ranks, id, rank));
    CCK(cudaStreamCreateWithFlags(&ep.st, cudaStreamNonBlocking));
    g_eps.push_back(ep);
    g_nranks = n_ranks;
    return FPS_OK;
}

int fps_b200_comm_init_local(const int *devices, int n_devices) {
    const NcclApi &N = nccl();
    if (!N.ok) {
        comm_set_err("NCCL not found (dlopen libnccl.so.2)");
        return FPS_ERR_NCCL;
    }
    if (!devices || n_devices < 1) {
        comm_set_err("bad argument: need a device list");
        return FPS_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(g_comm_mu);
    destroy_locked();
    DevRestore restore;
    std::vector<ncclComm_t> comms((size_t)n_devices, nullptr);
    NCK(N.CommInitAll(comms.data(), n_devices, devices));
    for (int i = 0; i < n_devices; ++i) {
        Endpoint ep;
        ep.dev = devices[i];
        ep.rank = i;
        ep.comm = comms[(size_t)i];
        CCK(cudaSetDevice(ep.dev));
        CCK(cudaStreamCreateWithFlags(&ep.st, cudaStreamNonBlocking));
        g_eps.push_back(ep);
    }
    g_nranks = n_devices;
    return FPS_OK;
}

void fps_b200_comm_destroy(void) {
    std::lock_guard<std::mutex> lk(g_comm_mu);
    DevRestore restore;
    destroy_locked();
}

int fps_b200_comm_ranks(void) { return g_nranks; }

int fps_b200_nccl_library(const char *path) {
    if (!path || strlen(path) >= sizeof g_nccl_path) {
        comm_set_err("bad argument: need a path of fewer than %zu characters", sizeof g_nccl_path);
        return FPS_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(g_comm_mu);
    if (g_nccl_bound) {
        comm_set_err("NCCL is already bound; name the library before the first comm call");
        return FPS_ERR_NCCL;
    }
    snprintf(g_nccl_path, sizeof g_nccl_path, "%s", path);
    return FPS_OK;
}

int fps_b200_nccl_version(void) {
    const NcclApi &N = nccl();
    int v = 0;
    if (N.ok && N.GetVersion) N.GetVersion(&v);
    return v;
}

int fps_b200_gather_indices(const uint64_t *local, size_t nb, size_t k, size_t n_clouds, uint64_t *out_rank0) {
    if (g_eps.size() != 1) {
        comm_set_err("fps_b200_gather_indices is the one-process-per-GPU entry: call fps_b200_comm_init first");
        return FPS_ERR_NCCL;
    }
    size_t b0, want;
    shard_of(n_clouds, g_nranks, g_eps[0].rank, &b0, &want);
    if (!local || want != nb || k == 0) {
        comm_set_err("rank %d holds %zu clouds, its shard of %zu is %zu", g_eps[0].rank, nb, n_clouds, want);
        return FPS_ERR_ARG;
    }
    const u64 *loc[1] = {reinterpret_cast<const u64 *>(local)};
    return comm_gather(loc, nullptr, k, n_clouds, reinterpret_cast<u64 *>(out_rank0));
}

}  // extern "C"
