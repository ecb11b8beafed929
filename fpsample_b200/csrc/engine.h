// engine.h -- internal interface between the kernel translation units and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace fps {

typedef unsigned long long u64;
typedef unsigned int u32;

// ---- vanilla -------------------------------------------------------------------------------------------
struct VanillaArgs {
    const float *pts;   // [B][n][dim] row-major
    const u64 *starts;  // nullptr (start 0) or [B][n_starts]
    u64 *out;           // [B][k]
    u32 n, dim, k, n_starts;
    u32 slice;          // points per CTA of a cluster
    u64 negzero;        // two binary32 -0.0, set by launch_vanilla_cluster: an operand ptxas cannot see through (vanilla.cu)
};

struct VanillaPlan {
    int dimp, ppt;
    u32 C, slice;
    size_t smem;
};

struct VanillaGridArgs {
    const float *pts;
    const u64 *starts;
    u64 *out;
    float *scratch;   // per-CTA [dim+1][slice] when the slice does not fit in shared memory
    u64 *slots;       // [groups][2][G]
    u32 *counters;    // [groups][32]
    u32 B, n, dim, k, n_starts;
    u32 G, slice, use_smem;
};

struct VanillaGridPlan {
    u32 G, groups, slice;
    bool use_smem;
    size_t smem, scratch_floats;
};

bool plan_vanilla_cluster(size_t n, size_t dim, size_t B, int n_sms, VanillaPlan *pl);
cudaError_t launch_vanilla_cluster(const VanillaPlan &pl, VanillaArgs a, u32 B, cudaStream_t st);
void plan_vanilla_grid(size_t n, size_t dim, size_t B, int n_sms, VanillaGridPlan *pl);
cudaError_t launch_vanilla_grid(const VanillaGridPlan &pl, VanillaGridArgs a, cudaStream_t st);

// ---- kd-line -------------------------------------------------------------------------------------------
struct KdlineArgs {
    const float *pts;   // [B][n][dim]
    const u64 *starts;  // nullptr or [B] (POSITION in the permuted array, src/wrapper.hpp:54-55)
    u64 *out;           // [B][k] or nullptr (build only)
    unsigned char *ws;  // per-cloud workspace region base (global), ws_stride bytes each
    size_t ws_stride;
    // optional build outputs (nullptr unless the build-only entry asked for them)
    u32 *perm_out;      // [B][n]
    u32 *leaf_lo_out;   // [B][2^h+1]
    float *leaf_box_out;// [B][2^h][2][dim]
    unsigned char *region;  // build-only into per-cloud regions (see kd_region_bytes), nullptr = fused build+sample
    size_t region_stride;
    u32 B, n, dim, k, h;
    u32 in_smem;        // bit0: coordinates + scratch in shared memory, bit1: node/bucket metadata in shared memory
};

struct KdlinePlan {
    int dimp;
    u32 threads, in_smem, grid;
    size_t smem;        // dynamic shared memory bytes
    size_t ws_stride;   // global workspace bytes per resident CTA
    size_t ws_bytes;    // total workspace: 256-byte work counter + grid * ws_stride
};

cudaError_t plan_kdline(size_t n, size_t dim, size_t h, size_t B, int n_sms, KdlinePlan *pl);
cudaError_t launch_kdline(const KdlinePlan &pl, KdlineArgs a, unsigned char *ws_base, cudaStream_t st);

// ---- kd-line build for batches of small clouds, everything in one CTA's shared memory (kdsmall.cu) --------
struct KdSmallPlan {
    int dimp;
    u32 grid, occ, big;   // big: 1024 threads, one CTA per SM, index arrays in global memory
    size_t smem;
    size_t ws_bytes;      // 256-byte scheduler counter + the big variant's index arrays
};
bool plan_kdsmall(size_t n, size_t dim, size_t h, size_t B, int n_sms, KdSmallPlan *pl);
// ws: pl.ws_bytes of workspace (dynamic cloud scheduler counter, then the big variant's index arrays)
cudaError_t launch_kdsmall(const KdSmallPlan &pl, const float *pts, unsigned char *region, size_t region_stride,
                           u32 *ws, u32 B, u32 n, u32 dim, u32 h, cudaStream_t st);
cudaError_t launch_kdsmall_export(const unsigned char *region, size_t region_stride, u32 B, u32 n, u32 dim, u32 h, u32 *perm_out,
                                  u32 *leaf_lo_out, float *leaf_box_out, cudaStream_t st);

// ---- kd-line, asynchronous coordinator/worker sampling over prebuilt regions (kdline_async.cu) -----------
// per-cloud region: [q dim*npad f32][dis npad f32][perm npad u32][nlo pad32(S+1) u32][fbox S*2*dim f32]
size_t kd_region_bytes(size_t n, size_t dim, size_t h);
struct AsyncPlan {
    int dimp;
    u32 C, threads, R, clusters;
    size_t smem;
};
bool plan_kdline_async(size_t n, size_t dim, size_t h, size_t B, int n_sms, AsyncPlan *pl);
cudaError_t launch_kdline_async(const AsyncPlan &pl, unsigned char *region, size_t region_stride, const u64 *starts,
                                u64 *out, u32 B, u32 n, u32 dim, u32 k, u32 h, cudaStream_t st);

cudaError_t async_debug_counters(u64 *out16);
cudaError_t warp_debug_counters(u64 *out16);

// ---- kd-line, one warp per cloud over prebuilt regions, records in shared memory or tensor memory (kdline_warp.cu) --
struct WarpPlan {
    int dimp;
    u32 rs /* pending samples per bucket */, bpl, n_tmem_warps, n_smem_warps, slot_bytes, meta_bytes, grid, lazy, nch, hybrid;
    size_t smem;
};
bool plan_kdline_warp(size_t n, size_t dim, size_t h, size_t B, int n_sms, WarpPlan *pl);
cudaError_t launch_kdline_warp(const WarpPlan &pl, unsigned char *region, size_t region_stride, const u64 *starts, u64 *out,
                               u32 *counter, u32 B, u32 n, u32 dim, u32 k, u32 h, cudaStream_t st);

// ---- kd-line, a team of 1 / 2 / 4 warps per cloud over prebuilt regions left in global memory (kdline_stream.cu) --------
struct StreamSeg {   // one launch: `clouds` consecutive clouds of the batch on teams of `wpc` warps
    u32 wpc /* warps per cloud */, bpl /* buckets per lane */, rs /* pending samples per bucket */, team_bytes, grid, clouds;
    size_t smem;
};
struct StreamPlan {   // a batch is cut into at most three runs of clouds, narrow teams first (kdline_stream.cu: plan_kdline_stream)
    int dimp;
    u32 nseg;
    StreamSeg seg[3];
    char desc[256];   // "1184 clouds x WPC=2 (...) + 16 clouds x WPC=4 (...)"
};
bool plan_kdline_stream(size_t n, size_t dim, size_t h, size_t B, int n_sms, StreamPlan *pl);
cudaError_t launch_kdline_stream(const StreamPlan &pl, unsigned char *region, size_t region_stride, const u64 *starts, u64 *out,
                                 u32 *counter, u32 B, u32 n, u32 dim, u32 k, u32 h, bool count, cudaStream_t st);
cudaError_t stream_debug_counters(u64 *out16);   // points scanned, point-updates, passes, early passes, bucket tests, picks, clouds, distances stored

// ---- kd-line, one huge cloud on the whole GPU: points in shared memory, batched picks per grid-wide exchange (kdline_grid.cu) --
struct GridPlan {
    int dimp;
    u32 ppt, G /* CTAs launched */, ecap, gc /* CTAs per cloud */, groups /* clouds in flight */, flat;
    size_t smem;
};
size_t kd_grid_pub_bytes(const GridPlan &pl);
bool plan_kdline_grid(size_t n, size_t dim, size_t h, size_t B, int n_sms, GridPlan *pl, bool ids = false);
// pts_vanilla != nullptr: vanilla FPS over the same machinery (starts = [B][n_starts] original indices, ties to the highest
// index); qv = scratch for the reversed SoA copy of the input, B * dim * npad floats
cudaError_t launch_kdline_grid(const GridPlan &pl, unsigned char *region, size_t region_stride, const u64 *starts, u64 *out,
                               unsigned char *pub, u32 B, u32 n, u32 dim, u32 k, u32 h, cudaStream_t st,
                               const float *pts_vanilla = nullptr, float *qv = nullptr, u32 n_starts = 1);
cudaError_t grid_debug_counters(u64 *out16);

// ---- kd-line build with the whole grid per level (kdbuild.cu), into the same per-cloud regions -------------
size_t kd_gridbuild_aux_bytes(size_t n, size_t dim, size_t h);
cudaError_t launch_kd_gridbuild(const float *pts, unsigned char *region, size_t region_stride, unsigned char *aux,
                                u32 B, u32 n, u32 dim, u32 h, cudaStream_t st);

// ---- full kd tree (bucket_fps_kdtree_sampling): GPU build of the permutation, then the vanilla kernels (kdtree.cu) ----
size_t kdtree_region_bytes(size_t n, size_t dim);
cudaError_t launch_kdtree_build(const float *pts, unsigned char *region, size_t region_stride, float *rows, u32 B, u32 n,
                                u32 dim, int n_sms, cudaStream_t st);
cudaError_t launch_kdtree_map(u64 *out, const unsigned char *region, size_t region_stride, u32 B, u32 n, u32 k, u32 dim,
                              cudaStream_t st);

// ---- fps_npdu_sampling: index-window heuristic, one CTA per cloud (npdu.cu) ---------------------------------------------
size_t npdu_workspace_bytes(size_t B, size_t n);
cudaError_t launch_npdu(const float *pts, size_t B, size_t n, size_t dim, size_t k, size_t w, const u64 *starts, u64 *out,
                        void *ws, int n_sms, cudaStream_t st);

// ---- fps_npdu_kdtree_sampling: k-nearest-neighbour update, one CTA per cloud, radix select of the k-th distance (npdu.cu) ------
size_t npdu_knn_workspace_bytes(size_t B, size_t n);
cudaError_t launch_npdu_knn(const float *pts, size_t B, size_t n, size_t dim, size_t k, size_t w, const u64 *starts, u64 *out,
                            void *ws, int n_sms, cudaStream_t st);

// ---- test entry for the tile-parallel sequential sum (seqsum.cu) ------------------------------------------------------
cudaError_t launch_seqsum(const float *x, size_t n, float *out, u32 *fast_tiles, int epl, cudaStream_t st);

void count_launch();

// ---- tuning knobs ------------------------------------------------------------------------------------------------------
// Read from the environment ONCE (FPS_B200_<NAME>, first use); afterwards only fps_b200_set_tuning changes them.  -1 =
// the planner's own choice.  The call path reads this struct, never the environment.
struct Tuning {
    int grid = -1;            // GRID: 1 forces the whole-GPU / grouped grid sampler, 0 forbids it
    int group = -1;           // GROUP: 1 forces the grouped (flat) grid sampler, 0 forbids it
    int gridbuild = -1;       // GRIDBUILD: 1 forces the grid-wide build launches, 0 forbids them
    int vanilla_kd = -1;      // VANILLA_KD: 1 forces fps_sampling through the kd permutation route, 0 forbids it
    int pipe = -1;            // PIPE: 0 switches the pipelined upload + per-piece build off
    int zerocopy = -1;        // ZEROCOPY: 0 never writes indices straight into page-locked host memory
    int grid_ecap = -1;       // GRID_ECAP: candidates per round of the grid sampler (32..G_ECAP)
    int warp = -1;            // WARP: 0 forbids the one-warp-per-cloud samplers
    int warp_tmem = -1;       // WARP_TMEM: 0 keeps the on-chip sampler out of tensor memory
    int warp_lazy = -1;       // WARP_LAZY: 0 = eager flushes
    int warp_hybrid = -1;     // WARP_HYBRID: 1 = coordinates in shared memory + distances in TMEM (opt-in)
    long warp_global_minb = -1;   // WARP_GLOBAL_MINB: smallest batch the streaming sampler takes
    int kdsmall = -1;         // KDSMALL: 0 forbids the shared-memory build kernel
    int stream_warps = -1;    // STREAM_WARPS: warps per cloud of the streaming sampler (1, 2, 4)
    int stream_split = -1;    // STREAM_SPLIT: 0 = one team size per batch; 1 (default) = full waves of narrow teams + a tail of wide ones; 2 = also one-warp teams
    int prefetch = -1;        // PREFETCH: 0 = the streaming sampler does not prefetch the buckets it is about to pass over into L2
    int psum = -1;            // PSUM: 0 = the grid-wide build never prepares the sequential sum's tiles in parallel (two-phase sum)
    int stage = -1;           // STAGE: 0 = pageable host inputs use plain cudaMemcpyAsync instead of the threaded page-locked staging pool
    int count = -1;           // COUNT: 1 = the streaming sampler counts the work it executes (fps_b200_debug_counters)
};
const Tuning &tuning();

}  // namespace fps
