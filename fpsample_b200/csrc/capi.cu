// capi.cu -- the C-ABI layer (include/fps_b200.h): validation, device contexts, host<->device staging,
// batch sharding over the devices of one box.  No torch / python types; no CPU fallback.
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fps_b200.h"
#include "engine.h"

namespace fps {
cudaError_t kb_debug_counters(unsigned long long *out16);
// comm.cu
int comm_gather(const u64 *const *locals, const cudaStream_t *after, size_t k, size_t n_clouds, u64 *out_rank0);
int comm_world();
int comm_local_endpoints();
int comm_endpoint_device(int i);
int comm_endpoint_rank(int i);

static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static thread_local char tl_err[512] = "";
static thread_local char tl_plan[512] = "";

static void set_err(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_err, sizeof(tl_err), fmt, ap);
    va_end(ap);
}
void comm_set_err(const char *fmt, ...) {   // comm.cu reports through the same thread-local text
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_err, sizeof(tl_err), fmt, ap);
    va_end(ap);
}
static void set_plan(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_plan, sizeof(tl_plan), fmt, ap);
    va_end(ap);
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            set_err("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return FPS_ERR_CUDA + (int)e__;                                                   \
        }                                                                                     \
    } while (0)

// ---- tuning knobs: the environment is read ONCE, the call path reads this struct ------------------------------------------
static Tuning g_tuning;
static std::once_flag g_tuning_once;
struct KnobDesc {
    const char *name;
    long Tuning::*lf;
    int Tuning::*f;
};
static const KnobDesc kKnobs[] = {
    {"GRID", nullptr, &Tuning::grid},           {"GROUP", nullptr, &Tuning::group},
    {"GRIDBUILD", nullptr, &Tuning::gridbuild}, {"VANILLA_KD", nullptr, &Tuning::vanilla_kd},
    {"PIPE", nullptr, &Tuning::pipe},           {"ZEROCOPY", nullptr, &Tuning::zerocopy},
    {"GRID_ECAP", nullptr, &Tuning::grid_ecap}, {"WARP", nullptr, &Tuning::warp},
    {"WARP_TMEM", nullptr, &Tuning::warp_tmem}, {"WARP_LAZY", nullptr, &Tuning::warp_lazy},
    {"WARP_HYBRID", nullptr, &Tuning::warp_hybrid}, {"WARP_GLOBAL_MINB", &Tuning::warp_global_minb, nullptr},
    {"KDSMALL", nullptr, &Tuning::kdsmall},     {"STREAM_WARPS", nullptr, &Tuning::stream_warps},
    {"STREAM_SPLIT", nullptr, &Tuning::stream_split},
    {"PSUM", nullptr, &Tuning::psum},           {"STAGE", nullptr, &Tuning::stage},
    {"COUNT", nullptr, &Tuning::count},         {"PREFETCH", nullptr, &Tuning::prefetch},
};
static bool set_knob(const char *name, long v) {
    for (const KnobDesc &k : kKnobs)
        if (!strcmp(k.name, name)) {
            if (k.lf) g_tuning.*(k.lf) = v;
            else g_tuning.*(k.f) = (int)v;
            return true;
        }
    return false;
}
const Tuning &tuning() {
    std::call_once(g_tuning_once, [] {
        for (const KnobDesc &k : kKnobs) {
            const std::string env = std::string("FPS_B200_") + k.name;
            if (const char *e = getenv(env.c_str())) set_knob(k.name, atol(e));
        }
    });
    return g_tuning;
}

// ---- the caller's current device is restored on every exit path -----------------------------------------------------
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            cudaGetLastError();
            prev = -1;
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- optional phase timing of the *_dev entries (bench.py: the dominant kernel's own duration) -----------------------
// When enabled, the calling thread's next *_dev call records CUDA events on ITS stream around the build and the
// sampling launches; fps_b200_last_phase_ms synchronises on them.  Off by default: no events, no overhead.
static std::atomic<int> g_phase_timing{0};
struct PhaseTimer {
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    bool armed = false;
    void mark(int i, cudaStream_t st) {
        if (!g_phase_timing.load(std::memory_order_relaxed)) return;
        if (!ev[i]) cudaEventCreate(&ev[i]);
        cudaEventRecord(ev[i], st);
        if (i == 2) armed = true;
    }
};
static thread_local PhaseTimer tl_phase;

// ---- devices ---------------------------------------------------------------------------------------------
struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return FPS_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            set_err("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return FPS_ERR_CUDA + (int)e;
        }
        cap = want;
        return FPS_OK;
    }
};

struct Lane {  // one in-flight chunk: its own stream and buffers
    cudaStream_t st = nullptr;
    Buf in, out, ws, starts;
};

// ---- pageable host inputs: a few persistent host threads copy slices into page-locked slots and enqueue each slot's
// transfer themselves ---------------------------------------------------------------------------------------------------
// cudaMemcpyAsync from pageable memory goes through the driver's single bounce buffer at ~11 GB/s (measured on the B200
// boxes: what a numpy user of the drop-in API gets).  Four threads memcpy at ~4 x one core's rate into their own pinned
// slots, and the copy of the next slice overlaps the PCIe transfer of the previous one.  One pool per device context,
// created at first use, never destroyed (its threads sleep on a condition variable between calls).
struct StagePool {
    static constexpr int T = 4, SLOTS = 2;
    static constexpr size_t SLOT = (size_t)4 << 20;
    struct Task {
        const char *src;
        char *dst;
        size_t bytes;
    };
    struct Worker {
        std::thread th;
        void *slot[SLOTS] = {};
        cudaEvent_t ev[SLOTS] = {};
        cudaEvent_t done = nullptr;   // everything this worker enqueued so far (recorded by wait_all's caller)
        cudaStream_t st = nullptr;
        int next = 0;
    };
    int dev;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<Task> q;
    int pending = 0;
    cudaError_t err = cudaSuccess;
    Worker w[T];

    explicit StagePool(int device) : dev(device) {   // (called with `device` current)
        cudaError_t e = cudaSuccess;
        for (int t = 0; t < T && e == cudaSuccess; ++t) {
            e = cudaStreamCreateWithFlags(&w[t].st, cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w[t].done, cudaEventDisableTiming);
            for (int i = 0; i < SLOTS && e == cudaSuccess; ++i) {
                e = cudaMallocHost(&w[t].slot[i], SLOT);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w[t].ev[i], cudaEventDisableTiming);
            }
        }
        err = e;
        for (int t = 0; t < T; ++t) w[t].th = std::thread([this, t] { run(t); });
        for (int t = 0; t < T; ++t) w[t].th.detach();
    }
    void run(int t) {
        Worker &me = w[t];
        cudaError_t e = cudaSetDevice(dev);
        for (;;) {
            Task task;
            {
                std::unique_lock<std::mutex> lk(mu);
                if (e != cudaSuccess && err == cudaSuccess) err = e;
                cv_work.wait(lk, [this] { return !q.empty(); });
                task = q.front();
                q.pop_front();
            }
            if (e == cudaSuccess) {
                const int i = me.next;
                me.next = (i + 1) % SLOTS;
                e = cudaEventSynchronize(me.ev[i]);   // the slot's previous transfer has left it
                if (e == cudaSuccess) {
                    memcpy(me.slot[i], task.src, task.bytes);
                    e = cudaMemcpyAsync(task.dst, me.slot[i], task.bytes, cudaMemcpyHostToDevice, me.st);
                }
                if (e == cudaSuccess) e = cudaEventRecord(me.ev[i], me.st);
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                if (e != cudaSuccess && err == cudaSuccess) err = e;
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    // every slice of [src, src + bytes) is copied and its transfer enqueued when this returns; `after` then waits for them
    cudaError_t upload(void *dst, const void *src, size_t bytes, cudaStream_t after) {
        {
            // slices: all workers share a small piece (a quarter each, not below 256 KB), big ones go slot by slot
            size_t step = (bytes + T - 1) / T;
            step = (step + 4095) & ~(size_t)4095;
            if (step < ((size_t)256 << 10)) step = (size_t)256 << 10;
            if (step > SLOT) step = SLOT;
            std::lock_guard<std::mutex> lk(mu);
            for (size_t off = 0; off < bytes; off += step) {
                q.push_back(Task{static_cast<const char *>(src) + off, static_cast<char *>(dst) + off, bytes - off < step ? bytes - off : step});
                ++pending;
            }
        }
        cv_work.notify_all();
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [this] { return pending == 0; });
            if (err != cudaSuccess) return err;
        }
        for (int t = 0; t < T; ++t) {
            cudaError_t e = cudaEventRecord(w[t].done, w[t].st);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(after, w[t].done, 0);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
};

struct DevCtx {
    int dev = -1, n_sms = 0;
    StagePool *stage = nullptr;
    std::mutex mu;
    Lane lane[2];
    Buf gout;                 // indices kept on the device for the NCCL gather (sharded entries)
    cudaEvent_t ev[8] = {};   // upload-chunk-landed events of the pipelined kd-line path
    cudaEvent_t ev_in = nullptr;   // 'device-resident input is ready' (recorded on the producer's stream)
    bool ready = false;
};

static std::mutex g_mu;
static std::vector<int> g_devs;       // usable device ordinals
static std::vector<DevCtx *> g_ctx;   // indexed by ordinal
static bool g_scanned = false;

static void scan_devices() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_scanned) return;
    g_scanned = true;
    int nd = 0;
    if (cudaGetDeviceCount(&nd) != cudaSuccess) {
        cudaGetLastError();
        nd = 0;
    }
    g_ctx.assign(nd, nullptr);
    for (int d = 0; d < nd; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) {
            g_devs.push_back(d);
            g_ctx[d] = new DevCtx();
            g_ctx[d]->dev = d;
            cudaDeviceGetAttribute(&g_ctx[d]->n_sms, cudaDevAttrMultiProcessorCount, d);
        }
    }
}

static DevCtx *get_ctx(int dev) {
    scan_devices();
    if (dev < 0 || dev >= (int)g_ctx.size() || !g_ctx[dev]) return nullptr;
    return g_ctx[dev];
}

static int n_sms_current(int *dev_out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    DevCtx *cx = get_ctx(dev);
    if (dev_out) *dev_out = dev;
    return cx ? cx->n_sms : 0;
}

// ---- validation shared by every entry ------------------------------------------------------------------------
static int check_common(const void *pts, size_t B, size_t n, size_t dim, size_t k, const void *out) {
    if (!pts || !out || B == 0 || n == 0 || dim == 0 || k == 0 || k > n) {
        set_err("bad argument: need points/out non-null, B,n,dim >= 1 and 1 <= k <= n (B=%zu n=%zu dim=%zu k=%zu)", B, n,
                dim, k);
        return FPS_ERR_ARG;
    }
    if (n >= 0xfffffff0ull || k >= 0xfffffff0ull || B >= 0xfffffff0ull || dim > 4096 ||
        n * dim >= 0xfffffff0ull) {
        set_err("shape too large for 32-bit indexing (n=%zu dim=%zu k=%zu B=%zu)", n, dim, k, B);
        return FPS_ERR_UNSUPPORTED;
    }
    return FPS_OK;
}

static int check_kdline(size_t n, size_t dim, size_t h) {
    if (h == 0) {
        set_err("height must be >= 1");
        return FPS_ERR_ARG;
    }
    // the reference's python front-end asserts 2**h <= n (src/fpsample/__init__.py:197); the slot-indexed
    // build keeps 2^h bucket slots, so absurd heights are refused instead of allocating 2^h slots
    if (h > 24 || (((size_t)1 << h) > 2 * n && h > 6)) {
        set_err("height %zu is not supported for n=%zu (need 2^h <= 2n or h <= 6)", h, n);
        return FPS_ERR_UNSUPPORTED;
    }
    (void)dim;
    return FPS_OK;
}

// ---- enqueue on the current device ------------------------------------------------------------------------------
struct WsLayout {
    bool cluster, kd;   // kd: big clouds go through a kd permutation + the batched-pick grid sampler (exact, pruned)
    size_t kd_h, kd_total;
    VanillaPlan vp;
    VanillaGridPlan gp;
    size_t off_scratch, off_slots, off_counters, total;
};
static bool vanilla_kd_layout(size_t B, size_t n, size_t dim, size_t n_starts, int n_sms, size_t *h, size_t *total);

static void vanilla_layout(size_t B, size_t n, size_t dim, int n_sms, WsLayout *L, size_t n_starts = 1) {
    L->kd = vanilla_kd_layout(B, n, dim, n_starts, n_sms, &L->kd_h, &L->kd_total);
    L->cluster = plan_vanilla_cluster(n, dim, B, n_sms, &L->vp);
    L->off_scratch = L->off_slots = L->off_counters = 0;
    L->total = 256;
    if (!L->cluster) {
        plan_vanilla_grid(n, dim, B, n_sms, &L->gp);
        size_t o = 0;
        L->off_counters = o;
        o += ((size_t)L->gp.groups * 32 * 4 + 255) & ~(size_t)255;
        L->off_slots = o;
        o += ((size_t)L->gp.groups * 2 * L->gp.G * 8 + 255) & ~(size_t)255;
        L->off_scratch = o;
        o += (L->gp.scratch_floats * 4 + 255) & ~(size_t)255;
        L->total = o + 256;
    }
    if (L->kd && L->kd_total > L->total) L->total = L->kd_total;
}

static int enqueue_kdline(const float *d_pts, size_t B, size_t n, size_t dim, size_t k, const u64 *d_starts, size_t h,
                          u64 *d_out, u32 *perm_out, u32 *leaf_lo_out, float *leaf_box_out, void *ws,
                          size_t ws_bytes, int n_sms, cudaStream_t st, const float *van_pts = nullptr, size_t van_nstarts = 1);

static int enqueue_vanilla(const float *d_pts, size_t B, size_t n, size_t dim, size_t k, const u64 *d_starts,
                           size_t n_starts, u64 *d_out, void *ws, size_t ws_bytes, int n_sms, cudaStream_t st) {
    WsLayout L;
    vanilla_layout(B, n, dim, n_sms, &L, d_starts ? n_starts : 1);
    if (L.kd)
        return enqueue_kdline(d_pts, B, n, dim, k, d_starts, L.kd_h, d_out, nullptr, nullptr, nullptr, ws, ws_bytes, n_sms, st,
                              d_pts, n_starts);
    if (L.cluster) {
        VanillaArgs a;
        a.pts = d_pts;
        a.starts = d_starts;
        a.out = d_out;
        a.n = (u32)n;
        a.dim = (u32)dim;
        a.k = (u32)k;
        a.n_starts = (u32)(d_starts ? n_starts : 0);
        a.slice = L.vp.slice;
        set_plan("vanilla_cluster_kernel<DIM=%d,PPT=%d> clouds=%zu cluster=%u slice=%u smem=%zu", L.vp.dimp, L.vp.ppt, B,
                 L.vp.C, L.vp.slice, L.vp.smem);
        tl_phase.mark(0, st);
        tl_phase.mark(1, st);
        CK(launch_vanilla_cluster(L.vp, a, (u32)B, st));
        tl_phase.mark(2, st);
        count_launch();
        return FPS_OK;
    }
    if (!ws || ws_bytes < L.total || (reinterpret_cast<uintptr_t>(ws) & 255)) {
        set_err("workspace too small or misaligned: need %zu bytes, 256-byte aligned (got %zu)", L.total, ws_bytes);
        return FPS_ERR_WORKSPACE;
    }
    unsigned char *w = static_cast<unsigned char *>(ws);
    VanillaGridArgs g;
    g.pts = d_pts;
    g.starts = d_starts;
    g.out = d_out;
    g.counters = reinterpret_cast<u32 *>(w + L.off_counters);
    g.slots = reinterpret_cast<u64 *>(w + L.off_slots);
    g.scratch = reinterpret_cast<float *>(w + L.off_scratch);
    g.B = (u32)B;
    g.n = (u32)n;
    g.dim = (u32)dim;
    g.k = (u32)k;
    g.n_starts = (u32)(d_starts ? n_starts : 0);
    set_plan("vanilla_grid_kernel clouds=%zu groups=%u G=%u slice=%u smem=%zu", B, L.gp.groups, L.gp.G, L.gp.slice,
             L.gp.smem);
    tl_phase.mark(0, st);
    tl_phase.mark(1, st);
    CK(launch_vanilla_grid(L.gp, g, st));
    tl_phase.mark(2, st);
    count_launch();
    return FPS_OK;
}

// ---- the kd-line route table -------------------------------------------------------------------------------------------
// Which kernel builds the per-cloud regions and which one samples them.  The thresholds were measured on one B200 (the
// script behind each is named); every decision can be overridden through fps_b200_set_tuning / FPS_B200_<KNOB>.
//
//   sampler       | takes                                                             | builder
//   --------------+-------------------------------------------------------------------+--------------------------------------
//   WarpOnChip    | cloud fits a shared-memory / TMEM slot, 2^h <= 128                | Small (kdsmall_kernel)
//   Grid (groups) | n >= kGroupMinPoints and 1..8 CTAs hold a cloud, unless WarpStream | Small if the coordinates fit one SM, else
//   Grid (whole)  | 1-4 clouds of >= 131 072 points that fit the SMs' shared memory   |   GridWide for few / huge / L2-bound
//   WarpStream    | cloud not on chip, batch >= kStreamMinCloudsPerSm clouds per SM   |   batches, else PerCloud (kdline_kernel)
//   AsyncCluster  | what is left of the clouds beyond one SM                          |
//   FusedCta      | everything else (one CTA builds and samples, any h)               | Fused
namespace route {
constexpr size_t kGroupMinPoints = 8192;        // groups of CTAs with batched picks beat one warp per cloud from here (scripts/cmp_group.py)
constexpr u32 kStreamGroupCtas = 8;             // a cloud that would tie up >= 8 CTAs of a group ...
constexpr size_t kStreamMinCloudsPerSm = 2;     //   ... streams from HBM once the batch stacks 2 clouds per SM (scripts/cmp_cfg5.py:
                                                //   16-CTA groups sample 4.7 k 100 k-point clouds/s at any batch size, the streaming kernel overtakes at ~300)
constexpr size_t kStreamMinCloudsPerSmMid = 4;  //   3..7 CTAs per cloud: from 4 clouds per SM
constexpr size_t kGridBuildMinPoints = 262144;  // one grid-wide launch per phase per level from here, whatever the batch (scripts/cmp_build5.py)
}  // namespace route

enum class Sampler { BuildOnly, WarpOnChip, WarpStream, Grid, AsyncCluster, FusedCta };
enum class Builder { Small, GridWide, PerCloud, Fused };

struct KdLayout {
    KdlinePlan pl;
    AsyncPlan ap;
    WarpPlan wp;
    StreamPlan stp;
    GridPlan gp;
    KdSmallPlan sp;
    Sampler sampler;
    Builder builder;
    size_t region_off, region_stride, aux_off, counter_off, pub_off, qv_off, total;
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static cudaError_t kd_layout(size_t B, size_t n, size_t dim, size_t h, int n_sms, bool build_only, KdLayout *L, bool ids = false) {
    cudaError_t e = plan_kdline(n, dim, h, B, n_sms, &L->pl);
    if (e != cudaSuccess) return e;
    const Tuning &tu = tuning();
    L->region_off = L->region_stride = L->aux_off = L->counter_off = L->pub_off = L->qv_off = 0;
    L->total = L->pl.ws_bytes;
    L->sampler = Sampler::FusedCta;
    L->builder = Builder::Fused;
    const bool cloud_in_one_sm = (L->pl.in_smem & 1) != 0;

    // ---- the sampler ----------------------------------------------------------------------------------------------
    bool have_small = false;
    if (build_only) {
        L->sampler = Sampler::BuildOnly;
        have_small = plan_kdsmall(n, dim, h, B, n_sms, &L->sp);   // the build-only entry runs the kernel the samplers are fed by
        if (!have_small) return cudaSuccess;                      // ... or the fused kernel's build phase
    } else {
        const bool force_grid = tu.grid == 1 || tu.group == 1;    // tests force the grid sampler onto clouds the planner would keep on one SM
        bool prefer_group = false;
        if (!force_grid && n >= route::kGroupMinPoints) {
            GridPlan tmp;
            const size_t stream_from = route::kStreamMinCloudsPerSm * (size_t)n_sms;
            prefer_group = plan_kdline_grid(n, dim, h, B, n_sms, &tmp, ids) && tmp.flat &&
                           (ids || tmp.gc <= 2 ||
                            B < (tmp.gc >= route::kStreamGroupCtas ? stream_from : route::kStreamMinCloudsPerSmMid * (size_t)n_sms));
        }
        // not on chip: the points stay in global memory, a team of warps per cloud -- worth it only when the batch keeps every SM
        // busy with clouds (throughput from clouds in flight, HBM-bound); small batches go to the cluster / grid kernels
        const size_t stream_minB = tu.warp_global_minb >= 0 ? (size_t)tu.warp_global_minb : route::kStreamMinCloudsPerSm * (size_t)n_sms;
        const bool warp_ok = !force_grid && !prefer_group && !ids && tu.warp != 0;
        if (warp_ok && plan_kdline_warp(n, dim, h, B, n_sms, &L->wp))
            L->sampler = Sampler::WarpOnChip;
        else if (warp_ok && B >= stream_minB && plan_kdline_stream(n, dim, h, B, n_sms, &L->stp))
            L->sampler = Sampler::WarpStream;
        else if ((force_grid || prefer_group || ids || !cloud_in_one_sm) && plan_kdline_grid(n, dim, h, B, n_sms, &L->gp, ids))
            L->sampler = Sampler::Grid;
        else if (ids)
            return cudaErrorNotSupported;
        else if (!cloud_in_one_sm && plan_kdline_async(n, dim, h, B, n_sms, &L->ap))
            L->sampler = Sampler::AsyncCluster;
        else
            return cudaSuccess;   // FusedCta
    }

    // ---- the builder that fills the regions ---------------------------------------------------------------------------
    bool gridbuild;
    if (L->sampler == Sampler::BuildOnly) {
        gridbuild = false;
    } else if (L->sampler == Sampler::WarpOnChip || L->sampler == Sampler::WarpStream) {
        have_small = plan_kdsmall(n, dim, h, B, n_sms, &L->sp);
        // clouds that stay in global memory: the grid-wide per-level launches build a batch faster than one CTA per cloud
        // working out of L2 (scripts/cmp_build5.py: 512 x 100 k points 10.9 -> 6.4 ms)
        gridbuild = !have_small && (tu.gridbuild >= 0 ? tu.gridbuild != 0 : L->sampler == Sampler::WarpStream);
    } else {
        // few clouds: one CTA per cloud would idle most SMs -> grid-wide; a cloud whose coordinates fit one SM's shared memory is
        // built by one CTA (0.1 ms per wave of 148 clouds against 49 grid-wide launches, 0.77 ms at BASELINE.json cfg 3)
        have_small = tu.gridbuild < 0 && plan_kdsmall(n, dim, h, B, n_sms, &L->sp);
        gridbuild = !have_small && (tu.gridbuild >= 0 ? tu.gridbuild != 0
                                                      : (B * 2 <= (size_t)n_sms || n >= route::kGridBuildMinPoints || !cloud_in_one_sm));
    }
    L->builder = have_small ? Builder::Small : gridbuild ? Builder::GridWide : Builder::PerCloud;

    // ---- workspace: [builder scratch][regions][256 B counter][grid-wide build aux][grid sampler exchange][reversed input] ----
    const size_t head = (have_small && L->sp.ws_bytes > L->pl.ws_bytes) ? L->sp.ws_bytes : L->pl.ws_bytes;
    L->region_off = align256(head);
    L->region_stride = kd_region_bytes(n, dim, h);
    L->counter_off = L->region_off + B * L->region_stride;
    L->total = L->counter_off + 256;
    if (gridbuild) {
        L->aux_off = align256(L->total);
        L->total = L->aux_off + B * kd_gridbuild_aux_bytes(n, dim, h);
    }
    if (L->sampler == Sampler::Grid) {
        L->pub_off = align256(L->total);
        L->total = L->pub_off + kd_grid_pub_bytes(L->gp);
        if (ids) {   // the reversed SoA copy of the input
            L->qv_off = align256(L->total);
            L->total = L->qv_off + B * dim * ((n + 31) & ~(size_t)31) * sizeof(float);
        }
    }
    return cudaSuccess;
}

// vanilla FPS on big clouds: the same exact recurrence, pruned.  A kd permutation (height chosen here: leaves of ~128-256
// points) groups the points into slices; ties are decided by the original index, so the result is fps_sampling's own.
static bool vanilla_kd_layout(size_t B, size_t n, size_t dim, size_t n_starts, int n_sms, size_t *h, size_t *total) {
    const int want = tuning().vanilla_kd;
    if (want == 0 || dim > FPS_B200_MAX_KDLINE_DIM || n_starts > 256 || n_starts == 0) return false;
    if (want < 0 && n < 16384) return false;   // smaller clouds: brute force in registers (vanilla_cluster_kernel) wins
    if (n < 2048) return false;
    size_t hh = 1;
    while (hh < 12 && (n >> hh) > 192) ++hh;
    if (n >= 262144 && hh > 9) hh = 9;   // one huge cloud on the whole grid: slices are cut inside leaves, few leaves
    KdLayout L;
    if (kd_layout(B, n, dim, hh, n_sms, false, &L, true) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    *h = hh;
    *total = L.total;
    return true;
}

static int enqueue_kdline(const float *d_pts, size_t B, size_t n, size_t dim, size_t k, const u64 *d_starts, size_t h,
                          u64 *d_out, u32 *perm_out, u32 *leaf_lo_out, float *leaf_box_out, void *ws,
                          size_t ws_bytes, int n_sms, cudaStream_t st, const float *van_pts, size_t van_nstarts) {
    KdLayout L;
    CK(kd_layout(B, n, dim, h, n_sms, d_out == nullptr, &L, van_pts != nullptr));
    const KdlinePlan &pl = L.pl;
    if (!ws || ws_bytes < L.total || (reinterpret_cast<uintptr_t>(ws) & 255)) {
        set_err("workspace too small or misaligned: need %zu bytes, 256-byte aligned (got %zu)", L.total, ws_bytes);
        return FPS_ERR_WORKSPACE;
    }
    unsigned char *w = static_cast<unsigned char *>(ws);
    KdlineArgs a;
    memset(&a, 0, sizeof(a));
    a.pts = d_pts;
    a.starts = d_starts;
    a.out = d_out;
    a.perm_out = perm_out;
    a.leaf_lo_out = leaf_lo_out;
    a.leaf_box_out = leaf_box_out;
    a.B = (u32)B;
    a.n = (u32)n;
    a.dim = (u32)dim;
    a.k = (u32)k;
    a.h = (u32)h;
    if (L.builder != Builder::Fused) {
        a.region = w + L.region_off;
        a.region_stride = L.region_stride;
    }
    // the build that fills the per-cloud regions, by plan
    auto build_regions = [&]() -> int {
        if (L.builder == Builder::Small)
            CK(launch_kdsmall(L.sp, d_pts, a.region, a.region_stride, static_cast<u32 *>(ws), (u32)B, (u32)n, (u32)dim, (u32)h, st));
        else if (L.builder == Builder::GridWide)
            CK(launch_kd_gridbuild(d_pts, a.region, a.region_stride, w + L.aux_off, (u32)B, (u32)n, (u32)dim, (u32)h, st));
        else
            CK(launch_kdline(pl, a, w, st));
        return FPS_OK;
    };
    char bdesc[128];
    if (L.builder == Builder::Small)
        snprintf(bdesc, sizeof bdesc, "kdsmall_kernel<DIM=%d,T=%d>(build in shared memory, %u CTA%s per SM, smem=%zu)", L.sp.dimp,
                 L.sp.big ? 1024 : 256, L.sp.occ, L.sp.occ > 1 ? "s" : "", L.sp.smem);
    else
        snprintf(bdesc, sizeof bdesc, "%s", L.builder == Builder::GridWide ? "gb_* grid-wide build (7 launches per level)" : "kdline_kernel(build, 1 CTA per cloud)");
    switch (L.sampler) {
    case Sampler::BuildOnly:
        if (L.builder == Builder::Small) {
            set_plan("kdsmall_kernel<DIM=%d>(build only, %u CTAs per SM, smem=%zu) + kdsmall_export_kernel", L.sp.dimp, L.sp.occ, L.sp.smem);
            tl_phase.mark(0, st);
            CK(launch_kdsmall(L.sp, d_pts, a.region, a.region_stride, static_cast<u32 *>(ws), (u32)B, (u32)n, (u32)dim, (u32)h, st));
            CK(launch_kdsmall_export(a.region, a.region_stride, (u32)B, (u32)n, (u32)dim, (u32)h, perm_out, leaf_lo_out, leaf_box_out, st));
            tl_phase.mark(1, st);
            tl_phase.mark(2, st);
            return FPS_OK;
        }
        break;   // the fused kernel's build phase, below
    case Sampler::WarpOnChip:
        set_plan("%s + kdline_warp%s_kernel<DIM=%d,BPL=%u> %s R=%u clouds=%zu grid=%u "
                 "warps/CTA=%u (tmem %u + smem %u) smem=%zu store/cloud=%u",
                 bdesc, L.wp.hybrid ? "(hybrid smem+tmem)" : "", L.wp.dimp, L.wp.bpl, L.wp.lazy ? "lazy" : "eager", L.wp.rs, B, L.wp.grid, L.wp.n_tmem_warps + L.wp.n_smem_warps, L.wp.n_tmem_warps,
                 L.wp.n_smem_warps, L.wp.smem, L.wp.slot_bytes);
        tl_phase.mark(0, st);
        if (int rcb = build_regions()) return rcb;
        tl_phase.mark(1, st);
        CK(launch_kdline_warp(L.wp, a.region, a.region_stride, d_starts, d_out, reinterpret_cast<u32 *>(w + L.counter_off), (u32)B, (u32)n,
                              (u32)dim, (u32)k, (u32)h, st));
        tl_phase.mark(2, st);
        return FPS_OK;
    case Sampler::WarpStream:
        set_plan("%s + kdline_stream_kernel<DIM=%d> (points in HBM, a team of warps per cloud): %s; clouds=%zu", bdesc, L.stp.dimp,
                 L.stp.desc, B);
        tl_phase.mark(0, st);
        if (int rcb = build_regions()) return rcb;
        tl_phase.mark(1, st);
        CK(launch_kdline_stream(L.stp, a.region, a.region_stride, d_starts, d_out, reinterpret_cast<u32 *>(w + L.counter_off), (u32)B, (u32)n,
                                (u32)dim, (u32)k, (u32)h, tuning().count == 1, st));
        tl_phase.mark(2, st);
        return FPS_OK;
    case Sampler::Grid:
        tl_phase.mark(0, st);
        set_plan("%s%s + kdline_grid_kernel<DIM=%d,%s> clouds=%zu grid=%u (%u CTAs per cloud, %u clouds in flight) threads=1024 "
                 "points/thread=%u candidates/round<=%u smem=%zu region/cloud=%zu",
                 van_pts ? "vanilla FPS via kd permutation (ties by original index): " : "",
                 bdesc, L.gp.dimp,
                 L.gp.flat ? "flat" : "merged", B, L.gp.G, L.gp.gc, L.gp.groups, L.gp.ppt, L.gp.ecap, L.gp.smem, L.region_stride);
        if (int rcb = build_regions()) return rcb;
        tl_phase.mark(1, st);
        a.starts = van_pts ? nullptr : d_starts;   // (the build kernels never read it)
        CK(launch_kdline_grid(L.gp, a.region, a.region_stride, d_starts, d_out, w + L.pub_off,
                              (u32)B, (u32)n, (u32)dim, (u32)k, (u32)h, st, van_pts,
                              van_pts ? reinterpret_cast<float *>(w + L.qv_off) : nullptr,
                              (u32)van_nstarts));
        tl_phase.mark(2, st);
        return FPS_OK;
    case Sampler::AsyncCluster:
        tl_phase.mark(0, st);
        set_plan("%s + kdline_async_kernel<DIM=%d> clouds=%zu clusters=%u "
                 "cluster=%u threads=%u smem=%zu R=%u region/cloud=%zu",
                 bdesc, L.ap.dimp, B, L.ap.clusters, L.ap.C, L.ap.threads, L.ap.smem, L.ap.R,
                 L.region_stride);
        if (int rcb = build_regions()) return rcb;
        tl_phase.mark(1, st);
        CK(launch_kdline_async(L.ap, a.region, a.region_stride, d_starts, d_out, (u32)B, (u32)n, (u32)dim, (u32)k,
                               (u32)h, st));
        tl_phase.mark(2, st);
        return FPS_OK;
    case Sampler::FusedCta:
        break;
    }
    set_plan("kdline_kernel<DIM=%d> clouds=%zu grid=%u threads=%u smem=%zu placement=%s ws/cta=%zu", pl.dimp, B, pl.grid,
             pl.threads, pl.smem, pl.in_smem == 3 ? "smem" : (pl.in_smem == 2 ? "meta-smem,data-L2" : "L2"),
             pl.ws_stride);
    tl_phase.mark(0, st);
    tl_phase.mark(1, st);
    a.region = nullptr;
    a.region_stride = 0;
    CK(launch_kdline(pl, a, w, st));
    tl_phase.mark(2, st);
    return FPS_OK;
}

// ---- full kd tree (bucket_fps_kdtree_sampling): GPU build of the full permutation, vanilla kernels over the permuted rows ----
struct KtLayout {
    WsLayout v;
    size_t region_off, region_stride, rows_off, vws_off, total;
};
static void kdtree_layout(size_t B, size_t n, size_t dim, int n_sms, KtLayout *L) {
    vanilla_layout(B, n, dim, n_sms, &L->v);
    L->region_off = 0;
    L->region_stride = kdtree_region_bytes(n, dim);
    L->rows_off = L->region_off + B * L->region_stride;
    L->vws_off = (L->rows_off + B * n * dim * sizeof(float) + 255) & ~(size_t)255;
    L->total = L->vws_off + L->v.total;
}

static int enqueue_kdtree(const float *d_pts, size_t B, size_t n, size_t dim, size_t k, const u64 *d_starts, u64 *d_out,
                          void *ws, size_t ws_bytes, int n_sms, cudaStream_t st) {
    KtLayout L;
    kdtree_layout(B, n, dim, n_sms, &L);
    if (!ws || ws_bytes < L.total || (reinterpret_cast<uintptr_t>(ws) & 255)) {
        set_err("workspace too small or misaligned: need %zu bytes, 256-byte aligned (got %zu)", L.total, ws_bytes);
        return FPS_ERR_WORKSPACE;
    }
    unsigned char *w = static_cast<unsigned char *>(ws);
    float *rows = reinterpret_cast<float *>(w + L.rows_off);
    tl_phase.mark(0, st);
    CK(launch_kdtree_build(d_pts, w + L.region_off, L.region_stride, rows, (u32)B, (u32)n, (u32)dim, n_sms, st));
    // positions: exact FPS over the permuted rows from POSITION start, ties to the highest position (KDNode.h:41-46)
    PhaseTimer keep = tl_phase;   // the vanilla enqueue marks its own phases: keep ours
    int rc = enqueue_vanilla(rows, B, n, dim, k, d_starts, 1, d_out, w + L.vws_off, L.v.total, n_sms, st);
    if (rc) return rc;
    std::string vplan = tl_plan;
    CK(launch_kdtree_map(d_out, w + L.region_off, L.region_stride, (u32)B, (u32)n, (u32)k, (u32)dim, st));
    tl_phase = keep;
    tl_phase.mark(1, st);
    tl_phase.mark(2, st);
    set_plan("kdtree_build_kernel(full kd permutation, 1 CTA per cloud) + %s + kdtree_map_kernel", vplan.c_str());
    return FPS_OK;
}

// ---- host-pointer shard on one device ------------------------------------------------------------------------------
struct ShardJob {
    int algo;
    const float *pts;
    size_t B, n, dim, k, h;
    const size_t *start;  // per cloud [B][n_starts] or nullptr
    size_t n_starts;
    size_t *out;
    bool keep_dev = false;   // leave the indices in cx->gout ([B][k] uint64 on the device) instead of copying them to `out`
};

// the stream that produced a device-resident input (set by fps_b200_set_producer_stream for the calling thread's next call);
// default: the legacy default stream, which orders behind every blocking stream of the caller
static thread_local cudaStream_t tl_producer = cudaStreamLegacy;

// Host -> device copy of one piece of the input, ordered before whatever is enqueued on `st` afterwards.  Page-locked sources
// go straight to the copy engine; big pageable ones through the staging pool above.
static int upload(DevCtx *cx, void *dst, const void *src, size_t bytes, cudaStream_t st) {
    bool pageable = false;
    if (bytes >= ((size_t)1 << 20) && tuning().stage != 0) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, src) == cudaSuccess) pageable = at.type == cudaMemoryTypeUnregistered;
        else cudaGetLastError();
    }
    if (!pageable) {
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return FPS_OK;
    }
    if (!cx->stage) cx->stage = new StagePool(cx->dev);
    CK(cx->stage->upload(dst, src, bytes, st));
    return FPS_OK;
}

static int shard_enqueue(DevCtx *cx, const ShardJob &j, cudaStream_t producer) {
    const int dev = cx->dev;
    const Tuning &tu = tuning();
    const size_t in_per = j.n * j.dim * sizeof(float), out_per = j.k * sizeof(u64);
    // input already in this device's memory (a pointer from cudaMalloc, torch, cupy ...): no upload, the kernels read it
    bool dev_in = false;
    {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, j.pts) == cudaSuccess && at.type == cudaMemoryTypeDevice) {
            if (at.device != dev) {
                set_err("points live on device %d but the shard runs on device %d", at.device, dev);
                return FPS_ERR_ARG;
            }
            dev_in = true;
            // whatever produced the buffer is ordered before our lanes by an event on the producer's stream: no device-wide
            // synchronisation, the caller's other streams keep running
            CK(cudaEventRecord(cx->ev_in, producer));
            for (auto &ln : cx->lane) CK(cudaStreamWaitEvent(ln.st, cx->ev_in, 0));
        } else {
            cudaGetLastError();
        }
    }
    // chunks overlap the upload of one with the kernels of the other (two lanes); each chunk must still be a batch the
    // samplers like: >= 1024 clouds (the one-warp-per-cloud kernels want every SM stacked) and >= 32 MB of input
    size_t nch = 1;
    {
        const size_t by_bytes = (j.B * in_per) / (32u << 20), by_clouds = j.B / 1024;
        nch = by_bytes < by_clouds ? by_bytes : by_clouds;
        if (nch < 1) nch = 1;
        if (nch > 8) nch = 8;
        if (dev_in) nch = 1;
    }
    size_t chunk = (j.B + nch - 1) / nch;
    if (nch > 1 && j.algo == FPS_ALGO_KDLINE) {
        // the streaming sampler works in waves of 8 two-warp teams per SM (kdline_stream.cu: plan_kdline_stream): chunks of
        // whole waves, so that only the last chunk of the batch has a partial wave (which then runs on 4-warp teams)
        KdLayout L;
        if (kd_layout(chunk, j.n, j.dim, j.h, cx->n_sms, false, &L) == cudaSuccess && L.sampler == Sampler::WarpStream) {
            const size_t wave = (size_t)8 * cx->n_sms;
            chunk = (chunk + wave - 1) / wave * wave;
        }
    }
    static_assert(sizeof(size_t) == sizeof(u64), "size_t must be 64-bit");
    int rc = FPS_OK;
    // One batch of small clouds (the on-chip sampler wants all of them in one launch): the upload is cut into pieces and
    // every piece is BUILT (kdsmall_kernel, into its clouds' regions) while the next one is still crossing PCIe; the
    // sampler then runs once over the whole batch.  Copy engine on lane 0's stream, kernels on lane 1's.
    if (!dev_in && j.algo == FPS_ALGO_KDLINE && nch == 1 && j.B >= 64 && j.B * in_per >= ((size_t)4 << 20) && tu.pipe != 0) {
        KdLayout L;
        CK(kd_layout(j.B, j.n, j.dim, j.h, cx->n_sms, false, &L));
        if ((L.sampler == Sampler::WarpOnChip || L.sampler == Sampler::WarpStream) &&
            (L.builder == Builder::Small || L.builder == Builder::GridWide)) {
            Lane &cp = cx->lane[0], &ex = cx->lane[1];
            if ((rc = cp.in.ensure(j.B * in_per)) || (rc = cp.out.ensure(j.B * out_per)) || (rc = cp.ws.ensure(L.total))) return rc;
            u64 *d_starts = nullptr;
            if (j.start) {
                if ((rc = cp.starts.ensure(j.B * sizeof(u64)))) return rc;
                d_starts = static_cast<u64 *>(cp.starts.p);
                CK(cudaMemcpyAsync(d_starts, j.start, j.B * sizeof(u64), cudaMemcpyHostToDevice, ex.st));
            }
            unsigned char *ws = static_cast<unsigned char *>(cp.ws.p);
            unsigned char *region = ws + L.region_off;
            const float *d_in = static_cast<const float *>(cp.in.p);
            size_t np = (j.B * in_per) / ((size_t)6 << 20);   // pieces of >= 6 MB
            if (np < 2) np = 2;
            if (np > 8) np = 8;
            const size_t per = (j.B + np - 1) / np;
            size_t c = 0;
            for (size_t b0 = 0; b0 < j.B; b0 += per, ++c) {
                const size_t nb = (j.B - b0 < per) ? j.B - b0 : per;
                if ((rc = upload(cx, const_cast<float *>(d_in) + b0 * j.n * j.dim, j.pts + b0 * j.n * j.dim, nb * in_per, cp.st))) return rc;
                CK(cudaEventRecord(cx->ev[c], cp.st));
                CK(cudaStreamWaitEvent(ex.st, cx->ev[c], 0));
                if (L.builder == Builder::Small) {
                    KdSmallPlan sp = L.sp;
                    const size_t gmax = (size_t)sp.occ * (size_t)cx->n_sms;
                    sp.grid = (u32)(nb < gmax ? nb : gmax);
                    CK(launch_kdsmall(sp, d_in + b0 * j.n * j.dim, region + b0 * L.region_stride, L.region_stride,
                                      reinterpret_cast<u32 *>(ws), (u32)nb, (u32)j.n, (u32)j.dim, (u32)j.h, ex.st));
                } else {   // clouds that stay in HBM (one shard of BASELINE cfg 5): the grid-wide build, piece by piece
                    CK(launch_kd_gridbuild(d_in + b0 * j.n * j.dim, region + b0 * L.region_stride, L.region_stride, ws + L.aux_off,
                                           (u32)nb, (u32)j.n, (u32)j.dim, (u32)j.h, ex.st));
                }
            }
            // page-locked output (what the python module hands out): the sampler writes its 32-pick blocks straight into
            // host memory over PCIe while it runs, no device-to-host copy afterwards
            u64 *d_out = static_cast<u64 *>(j.keep_dev ? cx->gout.p : cp.out.p);
            bool zero_copy = false;
            if (tu.zerocopy != 0 && !j.keep_dev) {
                cudaPointerAttributes at;
                if (cudaPointerGetAttributes(&at, j.out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
                    d_out = static_cast<u64 *>(at.devicePointer);
                    zero_copy = true;
                } else {
                    cudaGetLastError();
                }
            }
            if (L.sampler == Sampler::WarpOnChip)
                CK(launch_kdline_warp(L.wp, region, L.region_stride, d_starts, d_out,
                                      reinterpret_cast<u32 *>(ws + L.counter_off), (u32)j.B, (u32)j.n, (u32)j.dim, (u32)j.k, (u32)j.h,
                                      ex.st));
            else
                CK(launch_kdline_stream(L.stp, region, L.region_stride, d_starts, d_out,
                                        reinterpret_cast<u32 *>(ws + L.counter_off), (u32)j.B, (u32)j.n, (u32)j.dim, (u32)j.k, (u32)j.h,
                                        false, ex.st));
            if (!zero_copy && !j.keep_dev) CK(cudaMemcpyAsync(j.out, cp.out.p, j.B * out_per, cudaMemcpyDeviceToHost, ex.st));
            char bd[48];
            if (L.builder == Builder::Small) snprintf(bd, sizeof bd, "kdsmall_kernel<DIM=%d>", L.sp.dimp);
            else snprintf(bd, sizeof bd, "gb_* grid-wide build");
            if (L.sampler == Sampler::WarpOnChip)
                set_plan("pipelined upload (%zu pieces) + %s per piece + kdline_warp_kernel clouds=%zu%s", c, bd, j.B,
                         zero_copy ? " + indices written straight to page-locked host memory" : "");
            else
                set_plan("pipelined upload (%zu pieces) + %s per piece + kdline_stream_kernel<DIM=%d>: %s; clouds=%zu%s", c, bd, L.stp.dimp,
                         L.stp.desc, j.B, zero_copy ? " + indices written straight to page-locked host memory" : "");
            return FPS_OK;
        }
    }
    // the partial chunk goes FIRST: nothing hides the first upload, so it should be the small one (and for the streaming sampler
    // the partial wave then runs on its wide teams while the next full chunk is crossing PCIe)
    const size_t nchunks = (j.B + chunk - 1) / chunk, first = j.B - (nchunks - 1) * chunk;
    for (size_t c = 0, b0 = 0, nb = first; b0 < j.B; ++c, b0 += nb, nb = chunk) {
        Lane &ln = cx->lane[c & 1];
        size_t ws_need;
        if (j.algo == FPS_ALGO_NPDU) {
            ws_need = npdu_workspace_bytes(nb, j.n);
        } else if (j.algo == FPS_ALGO_NPDU_KNN) {
            ws_need = npdu_knn_workspace_bytes(nb, j.n);
        } else if (j.algo == FPS_ALGO_VANILLA) {
            WsLayout L;
            vanilla_layout(nb, j.n, j.dim, cx->n_sms, &L);
            ws_need = L.total;
        } else if (j.algo == FPS_ALGO_KDTREE) {
            KtLayout L;
            kdtree_layout(nb, j.n, j.dim, cx->n_sms, &L);
            ws_need = L.total;
        } else {
            KdLayout L;
            CK(kd_layout(nb, j.n, j.dim, j.h, cx->n_sms, false, &L));
            ws_need = L.total;
        }
        if ((!dev_in && (rc = ln.in.ensure(nb * in_per))) || (rc = ln.out.ensure(nb * out_per)) || (rc = ln.ws.ensure(ws_need))) return rc;
        const float *d_in = dev_in ? j.pts + b0 * j.n * j.dim : static_cast<const float *>(ln.in.p);
        u64 *d_res = j.keep_dev ? static_cast<u64 *>(cx->gout.p) + b0 * j.k : static_cast<u64 *>(ln.out.p);
        u64 *d_starts = nullptr;
        if (j.start) {
            if ((rc = ln.starts.ensure(nb * j.n_starts * sizeof(u64)))) return rc;
            d_starts = static_cast<u64 *>(ln.starts.p);
            CK(cudaMemcpyAsync(d_starts, j.start + b0 * j.n_starts, nb * j.n_starts * sizeof(u64),
                               cudaMemcpyHostToDevice, ln.st));
        }
        if (!dev_in && (rc = upload(cx, ln.in.p, j.pts + b0 * j.n * j.dim, nb * in_per, ln.st))) return rc;
        if (j.algo == FPS_ALGO_NPDU || j.algo == FPS_ALGO_NPDU_KNN) {
            cudaError_t e = j.algo == FPS_ALGO_NPDU
                                ? launch_npdu(d_in, nb, j.n, j.dim, j.k, j.h /* window */, d_starts, d_res, ln.ws.p, cx->n_sms, ln.st)
                                : launch_npdu_knn(d_in, nb, j.n, j.dim, j.k, j.h /* neighbours */, d_starts, d_res, ln.ws.p, cx->n_sms, ln.st);
            if (e == cudaErrorNotSupported) {
                cudaGetLastError();
                set_err("fps_npdu: dim > 64 or more than 6.5 M points are not supported");
                rc = FPS_ERR_UNSUPPORTED;
            } else if (e != cudaSuccess) {
                set_err("npdu launch failed: %s", cudaGetErrorString(e));
                rc = FPS_ERR_CUDA + (int)e;
            } else {
                if (j.algo == FPS_ALGO_NPDU)
                    set_plan("npdu_kernel clouds=%zu (one CTA per cloud, 256-point segment maxima in shared memory) window=%zu", nb, j.h);
                else
                    set_plan("npdu_knn_kernel clouds=%zu (one CTA per cloud, radix select of the k-th nearest distance) neighbours=%zu", nb, j.h);
            }
        } else if (j.algo == FPS_ALGO_VANILLA)
            rc = enqueue_vanilla(d_in, nb, j.n, j.dim, j.k, d_starts, j.n_starts, d_res, ln.ws.p, ln.ws.cap, cx->n_sms, ln.st);
        else if (j.algo == FPS_ALGO_KDTREE)
            rc = enqueue_kdtree(d_in, nb, j.n, j.dim, j.k, d_starts, d_res, ln.ws.p, ln.ws.cap, cx->n_sms, ln.st);
        else
            rc = enqueue_kdline(d_in, nb, j.n, j.dim, j.k, d_starts, j.h, d_res, nullptr, nullptr, nullptr, ln.ws.p, ln.ws.cap,
                                cx->n_sms, ln.st);
        if (rc) return rc;
        if (!j.keep_dev) CK(cudaMemcpyAsync(j.out + b0 * j.k, ln.out.p, nb * out_per, cudaMemcpyDeviceToHost, ln.st));
    }
    return FPS_OK;
}

// One shard on one device.  The caller's current device is restored on every exit path, and BOTH lanes are drained before
// returning -- also after an error: copies into the caller's `out` (or zero-copy kernel writes into it) may still be queued,
// and the lane buffers are reused by the next call.
static int run_shard(int dev, const ShardJob &j) {
    DevCtx *cx = get_ctx(dev);
    if (!cx) {
        set_err("device %d is not a usable sm_100 device", dev);
        return FPS_ERR_NO_DEVICE;
    }
    const cudaStream_t producer = tl_producer;
    tl_producer = cudaStreamLegacy;   // one call only
    std::lock_guard<std::mutex> lk(cx->mu);
    DeviceGuard guard;
    CK(cudaSetDevice(dev));
    if (!cx->ready) {
        for (auto &ln : cx->lane) CK(cudaStreamCreateWithFlags(&ln.st, cudaStreamNonBlocking));
        for (auto &e : cx->ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&cx->ev_in, cudaEventDisableTiming));
        cx->ready = true;
    }
    int rc = j.keep_dev ? cx->gout.ensure(j.B * j.k * sizeof(u64)) : FPS_OK;
    if (rc == FPS_OK) rc = shard_enqueue(cx, j, producer);
    for (auto &ln : cx->lane) {
        cudaError_t e = cudaStreamSynchronize(ln.st);
        if (e != cudaSuccess && rc == FPS_OK) {
            set_err("kernel execution failed: %s", cudaGetErrorString(e));
            rc = FPS_ERR_CUDA + (int)e;
        }
    }
    return rc;
}

static int run_batch(ShardJob j, const int *devices, int n_devices) {
    scan_devices();
    std::vector<int> devs;
    if (devices && n_devices > 0) {
        devs.assign(devices, devices + n_devices);
    } else if (j.B == 1) {
        // a single cloud runs where the caller is (one process per GPU: torch.cuda.set_device(local) and no device list)
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess) cudaGetLastError();
        if (cur >= 0 && get_ctx(cur)) devs.push_back(cur);
        else devs = g_devs;
    } else {
        devs = g_devs;
    }
    if (devs.empty()) {
        set_err("no usable CUDA device (need compute capability 10.x); there is no CPU fallback");
        return FPS_ERR_NO_DEVICE;
    }
    for (int d : devs)
        if (!get_ctx(d)) {
            set_err("device %d is not a usable sm_100 device", d);
            return FPS_ERR_NO_DEVICE;
        }
    {   // a device-resident batch is sampled where it lives
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, j.pts) == cudaSuccess && at.type == cudaMemoryTypeDevice) {
            bool listed = !(devices && n_devices > 0);
            for (int d : devs) listed = listed || d == at.device;
            if (!listed || !get_ctx(at.device)) {
                set_err("points live on device %d, which is not among the requested / usable devices", at.device);
                return FPS_ERR_ARG;
            }
            return run_shard(at.device, j);
        }
        cudaGetLastError();
    }
    size_t G = devs.size();
    if (G > j.B) G = j.B;
    if (G == 1) return run_shard(devs[0], j);
    // contiguous shards, remainder to the low ranks; no inter-device traffic
    std::vector<int> rcs(G, FPS_OK);
    std::vector<std::string> errs(G);
    std::vector<std::thread> th;
    size_t base = j.B / G, rem = j.B % G, b0 = 0;
    for (size_t g = 0; g < G; ++g) {
        size_t nb = base + (g < rem ? 1 : 0);
        ShardJob s = j;
        s.B = nb;
        s.pts = j.pts + b0 * j.n * j.dim;
        s.out = j.out + b0 * j.k;
        s.start = j.start ? j.start + b0 * j.n_starts : nullptr;
        int dev = devs[g];
        th.emplace_back([&, g, s, dev]() {
            rcs[g] = run_shard(dev, s);
            if (rcs[g]) errs[g] = tl_err;
        });
        b0 += nb;
    }
    for (auto &t : th) t.join();
    for (size_t g = 0; g < G; ++g)
        if (rcs[g]) {
            set_err("device %d: %s", devs[g], errs[g].c_str());
            return rcs[g];
        }
    return FPS_OK;
}

// ---- multi-GPU with the NCCL gather ------------------------------------------------------------------------
// Every endpoint of this process (one under torchrun, one per device after fps_b200_comm_init_local) samples its contiguous
// shard on its own GPU, the indices STAY on the device, and comm.cu gathers them to rank 0 as uint32 over NVLink.
static void shard_range(size_t n_clouds, int world, int rank, size_t *b0, size_t *nb) {
    const size_t base = n_clouds / (size_t)world, rem = n_clouds % (size_t)world;
    *nb = base + ((size_t)rank < rem ? 1 : 0);
    *b0 = (size_t)rank * base + ((size_t)rank < rem ? (size_t)rank : rem);
}

static int run_sharded(ShardJob j, size_t n_clouds, u64 *out_rank0) {
    const int E = comm_local_endpoints(), world = comm_world();
    if (E < 1) {
        set_err("no communicator: call fps_b200_comm_init (one process per GPU) or fps_b200_comm_init_local first");
        return FPS_ERR_NCCL;
    }
    scan_devices();
    // one endpoint: `j.pts` is this rank's shard.  Several: this process holds the whole batch and cuts it itself.
    std::vector<int> rcs((size_t)E, FPS_OK);
    std::vector<std::string> errs((size_t)E);
    std::vector<const u64 *> locals((size_t)E, nullptr);
    std::vector<std::thread> th;
    for (int i = 0; i < E; ++i) {
        size_t b0, nb;
        shard_range(n_clouds, world, comm_endpoint_rank(i), &b0, &nb);
        ShardJob s = j;
        s.B = nb;
        s.keep_dev = true;
        s.out = nullptr;
        if (E > 1) {
            s.pts = j.pts + b0 * j.n * j.dim;
            s.start = j.start ? j.start + b0 * j.n_starts : nullptr;
        }
        const int dev = comm_endpoint_device(i);
        DevCtx *cx = get_ctx(dev);
        if (!cx) {
            set_err("device %d is not a usable sm_100 device", dev);
            return FPS_ERR_NO_DEVICE;
        }
        auto work = [&, i, s, dev, cx]() {
            if (s.B) rcs[(size_t)i] = run_shard(dev, s);
            if (rcs[(size_t)i]) errs[(size_t)i] = tl_err;
            locals[(size_t)i] = static_cast<const u64 *>(cx->gout.p);
        };
        if (E == 1) work();
        else th.emplace_back(work);
    }
    for (auto &t : th) t.join();
    for (int i = 0; i < E; ++i)
        if (rcs[(size_t)i]) {
            set_err("device %d: %s", comm_endpoint_device(i), errs[(size_t)i].c_str());
            return rcs[(size_t)i];
        }
    return comm_gather(locals.data(), nullptr, j.k, n_clouds, out_rank0);
}

}  // namespace fps

using namespace fps;

// ======================================================================================================
//  exported C ABI
// ======================================================================================================
extern "C" {

int fps_b200_device_count(void) {
    scan_devices();
    return (int)g_devs.size();
}
const char *fps_b200_version(void) { return "fpsample-b200 0.1.0 (sm_100a)"; }
const char *fps_b200_last_error(void) { return tl_err; }
const char *fps_b200_last_plan(void) { return tl_plan; }
uint64_t fps_b200_kernel_launches(void) { return g_launches.load(); }

int fps_b200_debug_counters(int which, uint64_t *out16) {
    if (!out16) return FPS_ERR_ARG;
    CK(cudaDeviceSynchronize());
    switch (which) {
        case FPS_DBG_WARP: CK(warp_debug_counters(reinterpret_cast<u64 *>(out16))); break;
        case FPS_DBG_BUILD: CK(kb_debug_counters(reinterpret_cast<unsigned long long *>(out16))); break;
        case FPS_DBG_GRID: CK(grid_debug_counters(reinterpret_cast<u64 *>(out16))); break;
        case FPS_DBG_ASYNC: CK(async_debug_counters(reinterpret_cast<u64 *>(out16))); break;
        case FPS_DBG_STREAM: CK(stream_debug_counters(reinterpret_cast<u64 *>(out16))); break;
        default: set_err("unknown counter set %d", which); return FPS_ERR_ARG;
    }
    return FPS_OK;
}

int fps_b200_set_tuning(const char *name, long value) {
    (void)tuning();   // the environment is read first, so an explicit setting wins
    if (!name || !set_knob(name, value)) {
        set_err("unknown tuning knob '%s'", name ? name : "(null)");
        return FPS_ERR_ARG;
    }
    return FPS_OK;
}

int fps_b200_describe_stream_plan(size_t n_clouds, size_t n, size_t dim, size_t height, int n_sms, char *buf, size_t buf_len) {
    if (!buf || buf_len == 0 || n_sms <= 0) {
        set_err("bad argument: need a buffer and a positive SM count");
        return FPS_ERR_ARG;
    }
    StreamPlan pl;
    if (!plan_kdline_stream(n, dim, height, n_clouds, n_sms, &pl)) {
        set_err("the streaming sampler does not take %zu clouds of %zu x %zu points at height %zu", n_clouds, n, dim, height);
        return FPS_ERR_UNSUPPORTED;
    }
    snprintf(buf, buf_len, "%s", pl.desc);
    return FPS_OK;
}

void fps_b200_set_producer_stream(void *stream) { tl_producer = stream ? static_cast<cudaStream_t>(stream) : cudaStreamLegacy; }

void fps_b200_phase_timing(int enable) { g_phase_timing.store(enable ? 1 : 0); }

int fps_b200_last_phase_ms(float *build_ms, float *sample_ms) {
    if (!tl_phase.armed) {
        set_err("no timed *_dev call on this thread (enable fps_b200_phase_timing first)");
        return FPS_ERR_ARG;
    }
    CK(cudaEventSynchronize(tl_phase.ev[2]));
    float b = 0.f, s = 0.f;
    CK(cudaEventElapsedTime(&b, tl_phase.ev[0], tl_phase.ev[1]));
    CK(cudaEventElapsedTime(&s, tl_phase.ev[1], tl_phase.ev[2]));
    if (build_ms) *build_ms = b;
    if (sample_ms) *sample_ms = s;
    return FPS_OK;
}

void *fps_b200_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void fps_b200_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int fps_b200_vanilla(const float *points, size_t n, size_t dim, size_t k, const size_t *starts, size_t n_starts,
                     size_t *out) {
    int rc = check_common(points, 1, n, dim, k, out);
    if (rc) return rc;
    if (!starts || n_starts == 0 || n_starts > k) {
        set_err("need 1 <= n_starts <= k start indices (n_starts=%zu k=%zu)", n_starts, k);
        return FPS_ERR_ARG;
    }
    for (size_t i = 0; i < n_starts; ++i)
        if (starts[i] >= n) {
            set_err("start index %zu out of range (n=%zu)", starts[i], n);
            return FPS_ERR_START;
        }
    ShardJob j{FPS_ALGO_VANILLA, points, 1, n, dim, k, 0, starts, n_starts, out};
    return run_batch(j, nullptr, 1);
}

int bucket_fps_kdline(const float *raw_data, size_t n_points, size_t dim, size_t n_samples, size_t start_idx,
                      size_t height, size_t *out) {
    // same order and codes as src/wrapper.hpp:121-127
    if (dim == 0 || dim > FPS_B200_MAX_KDLINE_DIM) {
        set_err("only 1 to %d dimensions are supported (dim=%zu)", FPS_B200_MAX_KDLINE_DIM, dim);
        return FPS_ERR_DIM;
    }
    if (start_idx >= n_points) {
        set_err("start_idx %zu out of range (n=%zu)", start_idx, n_points);
        return FPS_ERR_START;
    }
    int rc = check_common(raw_data, 1, n_points, dim, n_samples, out);
    if (rc) return rc;
    if ((rc = check_kdline(n_points, dim, height))) return rc;
    ShardJob j{FPS_ALGO_KDLINE, raw_data, 1, n_points, dim, n_samples, height, &start_idx, 1, out};
    return run_batch(j, nullptr, 1);
}

int bucket_fps_kdtree(const float *raw_data, size_t n_points, size_t dim, size_t n_samples, size_t start_idx, size_t *out) {
    // same order and codes as src/wrapper.hpp:105-111
    if (dim == 0 || dim > FPS_B200_MAX_KDLINE_DIM) {
        set_err("only 1 to %d dimensions are supported (dim=%zu)", FPS_B200_MAX_KDLINE_DIM, dim);
        return FPS_ERR_DIM;
    }
    if (start_idx >= n_points) {
        set_err("start_idx %zu out of range (n=%zu)", start_idx, n_points);
        return FPS_ERR_START;
    }
    int rc = check_common(raw_data, 1, n_points, dim, n_samples, out);
    if (rc) return rc;
    ShardJob j{FPS_ALGO_KDTREE, raw_data, 1, n_points, dim, n_samples, 0, &start_idx, 1, out};
    return run_batch(j, nullptr, 1);
}

int fps_b200_kdtree_batch(const float *points, size_t B, size_t n, size_t dim, size_t k, const size_t *start, size_t *out,
                          const int *devices, int n_devices) {
    if (dim == 0 || dim > FPS_B200_MAX_KDLINE_DIM) {
        set_err("only 1 to %d dimensions are supported (dim=%zu)", FPS_B200_MAX_KDLINE_DIM, dim);
        return FPS_ERR_DIM;
    }
    if (start)
        for (size_t b = 0; b < B; ++b)
            if (start[b] >= n) {
                set_err("start[%zu]=%zu out of range (n=%zu)", b, start[b], n);
                return FPS_ERR_START;
            }
    int rc = check_common(points, B, n, dim, k, out);
    if (rc) return rc;
    ShardJob j{FPS_ALGO_KDTREE, points, B, n, dim, k, 0, start, 1, out};
    return run_batch(j, devices, n_devices);
}

int fps_b200_kdtree_batch_dev(const float *d_points, size_t B, size_t n, size_t dim, size_t k, const uint64_t *d_start,
                              uint64_t *d_out, void *d_workspace, size_t workspace_bytes, void *stream) {
    if (dim == 0 || dim > FPS_B200_MAX_KDLINE_DIM) {
        set_err("only 1 to %d dimensions are supported (dim=%zu)", FPS_B200_MAX_KDLINE_DIM, dim);
        return FPS_ERR_DIM;
    }
    int rc = check_common(d_points, B, n, dim, k, d_out);
    if (rc) return rc;
    int n_sms = n_sms_current(nullptr);
    if (n_sms <= 0) {
        set_err("current device is not a usable sm_100 device; there is no CPU fallback");
        return FPS_ERR_NO_DEVICE;
    }
    return enqueue_kdtree(d_points, B, n, dim, k, reinterpret_cast<const u64 *>(d_start), reinterpret_cast<u64 *>(d_out),
                          d_workspace, workspace_bytes, n_sms, static_cast<cudaStream_t>(stream));
}

int fps_b200_vanilla_batch(const float *points, size_t B, size_t n, size_t dim, size_t k, const size_t *start,
                           size_t *out, const int *devices, int n_devices) {
    int rc = check_common(points, B, n, dim, k, out);
    if (rc) return rc;
    if (start)
        for (size_t b = 0; b < B; ++b)
            if (start[b] >= n) {
                set_err("start[%zu]=%zu out of range (n=%zu)", b, start[b], n);
                return FPS_ERR_START;
            }
    ShardJob j{FPS_ALGO_VANILLA, points, B, n, dim, k, 0, start, 1, out};
    return run_batch(j, devices, n_devices);
}

int fps_b200_kdline_batch(const float *points, size_t B, size_t n, size_t dim, size_t k, const size_t *start,
                          size_t height, size_t *out, const int *devices, int n_devices) {
    if (dim == 0 || dim > FPS_B200_MAX_KDLINE_DIM) {
        set_err("only 1 to %d dimensions are supported (dim=%zu)", FPS_B200_MAX_KDLINE_DIM, dim);
        return FPS_ERR_DIM;
    }
    if (start)
        for (size_t b = 0; b < B; ++b)
            if (start[b] >= n) {
                set_err("start[%zu]=%zu out of range (n=%zu)", b, start[b], n);
                return FPS_ERR_START;
            }
    int rc = check_common(points, B, n, dim, k, out);
    if (rc) return rc;
    if ((rc = check_kdline(n, dim, height))) return rc;
    ShardJob j{FPS_ALGO_KDLINE, points, B, n, dim, k, height, start, 1, out};
    return run_batch(j, devices, n_devices);
}

int fps_b200_kdline_batch_sharded(const float *points, size_t n_clouds, size_t n, size_t dim, size_t k, const size_t *start,
                                  size_t height, size_t *out_rank0) {
    if (dim == 0 || dim > FPS_B200_MAX_KDLINE_DIM) {
        set_err("only 1 to %d dimensions are supported (dim=%zu)", FPS_B200_MAX_KDLINE_DIM, dim);
        return FPS_ERR_DIM;
    }
    static size_t dummy_out;
    int rc = check_common(points, n_clouds, n, dim, k, &dummy_out);
    if (rc) return rc;
    if ((rc = check_kdline(n, dim, height))) return rc;
    ShardJob j{FPS_ALGO_KDLINE, points, 0, n, dim, k, height, start, 1, nullptr};
    return run_sharded(j, n_clouds, reinterpret_cast<u64 *>(out_rank0));
}

int fps_b200_vanilla_batch_sharded(const float *points, size_t n_clouds, size_t n, size_t dim, size_t k, const size_t *start,
                                   size_t *out_rank0) {
    static size_t dummy_out;
    int rc = check_common(points, n_clouds, n, dim, k, &dummy_out);
    if (rc) return rc;
    ShardJob j{FPS_ALGO_VANILLA, points, 0, n, dim, k, 0, start, 1, nullptr};
    return run_sharded(j, n_clouds, reinterpret_cast<u64 *>(out_rank0));
}

int fps_b200_npdu(const float *points, size_t n, size_t dim, size_t n_samples, size_t window, size_t start_idx, size_t *out) {
    int rc = check_common(points, 1, n, dim, n_samples, out);
    if (rc) return rc;
    if (start_idx >= n) {
        set_err("start_idx %zu out of range (n=%zu)", start_idx, n);
        return FPS_ERR_START;
    }
    ShardJob j{FPS_ALGO_NPDU, points, 1, n, dim, n_samples, window, &start_idx, 1, out};
    return run_batch(j, nullptr, 1);
}

int fps_b200_npdu_kdtree(const float *points, size_t n, size_t dim, size_t n_samples, size_t k, size_t start_idx, size_t *out) {
    int rc = check_common(points, 1, n, dim, n_samples, out);
    if (rc) return rc;
    if (start_idx >= n) {
        set_err("start_idx %zu out of range (n=%zu)", start_idx, n);
        return FPS_ERR_START;
    }
    ShardJob j{FPS_ALGO_NPDU_KNN, points, 1, n, dim, n_samples, k, &start_idx, 1, out};
    return run_batch(j, nullptr, 1);
}

int fps_b200_npdu_kdtree_batch(const float *points, size_t B, size_t n, size_t dim, size_t n_samples, size_t k,
                               const size_t *start, size_t *out, const int *devices, int n_devices) {
    int rc = check_common(points, B, n, dim, n_samples, out);
    if (rc) return rc;
    if (start)
        for (size_t b = 0; b < B; ++b)
            if (start[b] >= n) {
                set_err("start[%zu]=%zu out of range (n=%zu)", b, start[b], n);
                return FPS_ERR_START;
            }
    ShardJob j{FPS_ALGO_NPDU_KNN, points, B, n, dim, n_samples, k, start, 1, out};
    return run_batch(j, devices, n_devices);
}

int fps_b200_npdu_batch(const float *points, size_t B, size_t n, size_t dim, size_t n_samples, size_t window,
                        const size_t *start, size_t *out, const int *devices, int n_devices) {
    int rc = check_common(points, B, n, dim, n_samples, out);
    if (rc) return rc;
    if (start)
        for (size_t b = 0; b < B; ++b)
            if (start[b] >= n) {
                set_err("start[%zu]=%zu out of range (n=%zu)", b, start[b], n);
                return FPS_ERR_START;
            }
    ShardJob j{FPS_ALGO_NPDU, points, B, n, dim, n_samples, window, start, 1, out};
    return run_batch(j, devices, n_devices);
}

size_t fps_b200_workspace_bytes(int algo, size_t B, size_t n, size_t dim, size_t k, size_t height) {
    (void)k;
    int n_sms = n_sms_current(nullptr);
    if (n_sms <= 0 || B == 0 || n == 0 || dim == 0) return 0;
    if (algo == FPS_ALGO_VANILLA) {
        WsLayout L;
        vanilla_layout(B, n, dim, n_sms, &L);
        return L.total;
    }
    if (algo == FPS_ALGO_KDTREE) {
        if (dim > FPS_B200_MAX_KDLINE_DIM) return 0;
        KtLayout L;
        kdtree_layout(B, n, dim, n_sms, &L);
        return L.total;
    }
    if (dim > FPS_B200_MAX_KDLINE_DIM || check_kdline(n, dim, height)) return 0;
    KdLayout L;
    if (kd_layout(B, n, dim, height, n_sms, false, &L) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    size_t total = L.total;
    KdLayout Lb;   // the build-only entry (fps_b200_kdline_build_dev) shares this query
    if (kd_layout(B, n, dim, height, n_sms, true, &Lb) == cudaSuccess && Lb.total > total) total = Lb.total;
    cudaGetLastError();
    return total;
}

int fps_b200_vanilla_batch_dev(const float *d_points, size_t B, size_t n, size_t dim, size_t k,
                               const uint64_t *d_start, uint64_t *d_out, void *d_workspace, size_t workspace_bytes,
                               void *stream) {
    int rc = check_common(d_points, B, n, dim, k, d_out);
    if (rc) return rc;
    int n_sms = n_sms_current(nullptr);
    if (n_sms <= 0) {
        set_err("current device is not a usable sm_100 device; there is no CPU fallback");
        return FPS_ERR_NO_DEVICE;
    }
    return enqueue_vanilla(d_points, B, n, dim, k, reinterpret_cast<const u64 *>(d_start), 1,
                           reinterpret_cast<u64 *>(d_out), d_workspace, workspace_bytes, n_sms,
                           static_cast<cudaStream_t>(stream));
}

int fps_b200_kdline_batch_dev(const float *d_points, size_t B, size_t n, size_t dim, size_t k, const uint64_t *d_start,
                              size_t height, uint64_t *d_out, void *d_workspace, size_t workspace_bytes,
                              void *stream) {
    if (dim == 0 || dim > FPS_B200_MAX_KDLINE_DIM) {
        set_err("only 1 to %d dimensions are supported (dim=%zu)", FPS_B200_MAX_KDLINE_DIM, dim);
        return FPS_ERR_DIM;
    }
    int rc = check_common(d_points, B, n, dim, k, d_out);
    if (rc) return rc;
    if ((rc = check_kdline(n, dim, height))) return rc;
    int n_sms = n_sms_current(nullptr);
    if (n_sms <= 0) {
        set_err("current device is not a usable sm_100 device; there is no CPU fallback");
        return FPS_ERR_NO_DEVICE;
    }
    return enqueue_kdline(d_points, B, n, dim, k, reinterpret_cast<const u64 *>(d_start), height,
                          reinterpret_cast<u64 *>(d_out), nullptr, nullptr, nullptr, d_workspace, workspace_bytes, n_sms,
                          static_cast<cudaStream_t>(stream));
}

int fps_b200_kdline_build_dev(const float *d_points, size_t B, size_t n, size_t dim, size_t height, uint32_t *d_perm,
                              uint32_t *d_leaf_lo, float *d_leaf_box, void *d_workspace, size_t workspace_bytes,
                              void *stream) {
    if (dim == 0 || dim > FPS_B200_MAX_KDLINE_DIM) {
        set_err("only 1 to %d dimensions are supported (dim=%zu)", FPS_B200_MAX_KDLINE_DIM, dim);
        return FPS_ERR_DIM;
    }
    int rc = check_common(d_points, B, n, dim, 1, d_perm);
    if (rc) return rc;
    if ((rc = check_kdline(n, dim, height))) return rc;
    int n_sms = n_sms_current(nullptr);
    if (n_sms <= 0) {
        set_err("current device is not a usable sm_100 device; there is no CPU fallback");
        return FPS_ERR_NO_DEVICE;
    }
    return enqueue_kdline(d_points, B, n, dim, 1, nullptr, height, nullptr, d_perm, d_leaf_lo, d_leaf_box, d_workspace,
                          workspace_bytes, n_sms, static_cast<cudaStream_t>(stream));
}

int fps_b200_seqsum_dev(const float *d_values, size_t n, float *d_sum, uint32_t *d_fast_tiles, int tile, void *stream) {
    if (!d_values || !d_sum || n == 0 || (tile != 256 && tile != 512 && tile != -512)) {
        set_err("bad argument: need values/sum non-null, n >= 1, tile 256, 512 or -512 (two-phase)");
        return FPS_ERR_ARG;
    }
    if (n_sms_current(nullptr) <= 0) {
        set_err("current device is not a usable sm_100 device; there is no CPU fallback");
        return FPS_ERR_NO_DEVICE;
    }
    CK(launch_seqsum(d_values, n, d_sum, d_fast_tiles, tile < 0 ? -1 : tile / 32, static_cast<cudaStream_t>(stream)));
    return FPS_OK;
}

}  // extern "C"
