/*
 * fps_b200.h -- C ABI of the B200-native farthest-point-sampling engine (libfps_b200.so).
 *
 * This is the drop-in boundary for the FPS hot path of leonardodalinky/fpsample v1.0.2.  Plain
 * pointers and sizes only; no torch / numpy / pybind types.  Every entry point cites the reference
 * interface it replaces (paths are relative to the reference repository root).
 *
 * Conventions shared by all entry points
 *   - points are row-major float32 [n][dim] (one cloud) or [B][n][dim] (a batch of equally sized
 *     clouds), exactly the buffer the reference's pybind layer hands to its C++ code
 *     (src/lib.cpp:249-253, 522-527: array_t<float, c_style|forcecast>).
 *   - indices out are size_t / uint64 (src/lib.cpp:240-245, 561-562), [k] or [B][k].
 *   - return 0 on success.  1 and 2 keep the reference's meaning (src/wrapper.hpp:121-127).
 *   - host-pointer entry points are synchronous and re-entrant; device memory, streams and
 *     workspaces are private to the library; the caller's current device is left as it was.  Single-cloud
 *     entries run on the caller's current device.  `points` may also be a DEVICE pointer (the clouds are then
 *     sampled on the device they live on, without an upload, ordered behind the producer's stream by an
 *     event -- fps_b200_set_producer_stream; no device-wide synchronisation); page-locked
 *     buffers make the host path faster (pipelined upload, indices written straight into `out`).  *_dev entry points take device pointers and a CUDA
 *     stream (void* == cudaStream_t) and only enqueue work.
 *   - there is NO CPU fallback: without a usable sm_100 device the calls fail with FPS_ERR_NO_DEVICE.
 */
#ifndef FPS_B200_H
#define FPS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FPS_API __attribute__((visibility("default")))
#else
#define FPS_API
#endif

#define FPS_OK 0
#define FPS_ERR_DIM 1        /* kd-line: dim == 0 or dim > 8            (src/wrapper.hpp:121-123) */
#define FPS_ERR_START 2      /* a start index >= n                      (src/wrapper.hpp:124-127) */
#define FPS_ERR_ARG 3        /* k == 0, k > n, n == 0, h == 0, null pointer, n_starts > k ...      */
#define FPS_ERR_NO_DEVICE 4  /* no CUDA device / not an sm_100 part                                */
#define FPS_ERR_WORKSPACE 5  /* *_dev: workspace too small or misaligned                           */
#define FPS_ERR_UNSUPPORTED 6/* shape outside what the kernels cover               */
#define FPS_ERR_NCCL 7       /* NCCL missing / no communicator / a collective failed                  */
#define FPS_ERR_CUDA 100     /* 100 + cudaError_t of the failing runtime call                      */

#define FPS_B200_MAX_KDLINE_DIM 8 /* BUCKET_FPS_MAX_DIM, src/wrapper.hpp:8-11 */

/* ---- single cloud, host pointers ---------------------------------------------------------------- */

/* Vanilla FPS.  Replaces fps_sampling (src/lib.cpp:188-246) and fps_sampling_multi_start_index
 * (src/lib.cpp:111-186), which the reference runs inline in its pybind function (it has no C entry
 * for this path).  The first n_starts picks are forced to starts[0..n_starts), every pick still
 * min-updates; free picks take the HIGHEST index among maximal distances (the '>=' at lib.cpp:226). */
FPS_API int fps_b200_vanilla(const float *points, size_t n, size_t dim, size_t k, const size_t *starts,
                     size_t n_starts, size_t *out_indices);

/* QuickFPS kd-line.  SAME NAME AND SIGNATURE as the reference's C ABI (src/wrapper.hpp:118-132), so a
 * build of the reference can link this library in place of wrapper.hpp.  start_idx addresses the
 * POSITION in the array after the kd build permuted it (src/wrapper.hpp:54-55).
 *
 * One restriction the reference does not have: the build keeps 2^height bucket slots, so heights with 2^height > 2 * n_points
 * (beyond height 6) and heights above 24 return FPS_ERR_UNSUPPORTED (6).  The reference's python front-end asserts
 * 2^height <= n_points (src/fpsample/__init__.py:197), so this is only reachable through the C ABI or under `python -O`.
 * (Clamping the height instead would NOT be exact: the mean split is unbalanced, a leaf at the clamped depth can still hold
 * several points that the reference would keep partitioning.) */
FPS_API int bucket_fps_kdline(const float *raw_data, size_t n_points, size_t dim, size_t n_samples,
                      size_t start_idx, size_t height, size_t *sampled_point_indices);

/* QuickFPS full kd tree.  SAME NAME AND SIGNATURE as the reference's C ABI (src/wrapper.hpp:102-116 ->
 * kdtree_sample :29-43 -> src/_ext/KDTree.h:13-52).  start_idx addresses the POSITION in the array after the
 * (full-depth) kd build permuted it (src/wrapper.hpp:36-37); ties go to the HIGHEST position (the right child
 * wins, src/_ext/KDNode.h:41-46).  Returns 1 for dim outside [1, 8], 2 for start_idx >= n_points. */
FPS_API int bucket_fps_kdtree(const float *raw_data, size_t n_points, size_t dim, size_t n_samples,
                      size_t start_idx, size_t *sampled_point_indices);

/* ---- batches, host pointers (new; the reference has no batched entry) --------------------------- */

/* B independent clouds [B][n][dim] -> [B][k].  start: NULL (all 0) or [B].  devices: NULL/0 = all
 * visible devices; the batch is split into contiguous shards, one per device, no inter-device traffic. */
FPS_API int fps_b200_vanilla_batch(const float *points, size_t B, size_t n, size_t dim, size_t k,
                           const size_t *start, size_t *out_indices, const int *devices, int n_devices);
FPS_API int fps_b200_kdline_batch(const float *points, size_t B, size_t n, size_t dim, size_t k,
                          const size_t *start, size_t height, size_t *out_indices, const int *devices,
                          int n_devices);

FPS_API int fps_b200_kdtree_batch(const float *points, size_t B, size_t n, size_t dim, size_t k,
                          const size_t *start, size_t *out_indices, const int *devices, int n_devices);

/* FPS with the nearest-point-distance-updating heuristic over an index window (NOT exact FPS).  Replaces
 * fps_npdu_sampling (src/lib.cpp:272-340), which the reference runs inline in its pybind function: after a full
 * min-update against the start point, every pick min-updates only the points whose index lies within window / 2 of it
 * (window shifted at the array ends) and the next pick is the arg-max over all points, lowest index among equals.
 * Same indices as the reference for the same `window` (the python front-end's default is n / n_samples * 16). */
FPS_API int fps_b200_npdu(const float *points, size_t n, size_t dim, size_t n_samples, size_t window, size_t start_idx,
                  size_t *out_indices);
FPS_API int fps_b200_npdu_batch(const float *points, size_t B, size_t n, size_t dim, size_t n_samples, size_t window,
                        const size_t *start, size_t *out_indices, const int *devices, int n_devices);

/* FPS with the nearest-point-distance-updating heuristic over the k NEAREST points (NOT exact FPS).  Replaces
 * fps_npdu_kdtree_sampling_py (src/lib.cpp:369-465; its neighbour search is nanoflann's, src/nanoflann.hpp): after a full
 * min-update against the start point, every pick min-updates its k nearest points (binary32 distances in the reference's
 * arithmetic, lib.cpp:33-41; k capped at n, :407) and the next pick is the arg-max over all points, lowest index among equals
 * (:438-442).  Same indices as the reference whenever the k-th nearest distance of a pick is not shared by more candidates
 * than there are places left (exact ties there are taken in index order here, in nanoflann's traversal order there). */
FPS_API int fps_b200_npdu_kdtree(const float *points, size_t n, size_t dim, size_t n_samples, size_t k, size_t start_idx,
                         size_t *out_indices);
FPS_API int fps_b200_npdu_kdtree_batch(const float *points, size_t B, size_t n, size_t dim, size_t n_samples, size_t k,
                               const size_t *start, size_t *out_indices, const int *devices, int n_devices);

/* ---- multi-GPU: shards on every GPU, indices gathered to rank 0 over NCCL (new) ----------------------
 * Clouds are independent: a batch of n_clouds is cut into contiguous shards (remainder to the low ranks, the rule of the
 * *_batch entries), every GPU samples its shard with NO inter-GPU traffic, and the only exchange is the gather of the index
 * arrays to rank 0 (uint32 on the wire, grouped ncclSend / ncclRecv over NVLink, widened on rank 0's device, one copy to the
 * host).  NCCL is bound at run time (dlopen libnccl.so.2); without it these entries return FPS_ERR_NCCL.
 *   one process per GPU : rank 0 calls fps_b200_comm_unique_id, the 128 bytes reach every rank through the launcher's
 *                         rendezvous, every rank calls fps_b200_comm_init on its (current) device.
 *   one process, G GPUs : fps_b200_comm_init_local(devices, G)   (ncclCommInitAll, one communicator + stream per device). */
#define FPS_COMM_ID_BYTES 128
FPS_API int fps_b200_comm_unique_id(void *id128);
FPS_API int fps_b200_comm_init(const void *id128, int n_ranks, int rank);
FPS_API int fps_b200_comm_init_local(const int *devices, int n_devices);
FPS_API void fps_b200_comm_destroy(void);
FPS_API int fps_b200_comm_ranks(void);    /* 0 = no communicator */
FPS_API int fps_b200_nccl_version(void);  /* e.g. 22809; 0 = NCCL not found */
/* Optional, before the first comm call: the libnccl to bind when the process has not loaded one yet.  Search order: a
 * libnccl.so.2 already in the process (a host framework's) -> this path (or the environment's FPS_B200_NCCL_LIB) ->
 * the system's libnccl.so.2.  A process must not end up with two different libnccl.so.2: a host that will load its own
 * copy LATER (PyTorch imported after the first comm call) names that copy here; the python package does so by itself
 * (the pip-installed nvidia-nccl wheel).  FPS_ERR_NCCL if NCCL is already bound. */
FPS_API int fps_b200_nccl_library(const char *path);
/* Collective.  One process per GPU: `points` / `start` are THIS RANK's shard ([nb][n][dim], nb = its share of n_clouds; host or
 * device memory).  One process with several endpoints: the whole batch [n_clouds][n][dim] in host memory.  The indices of
 * all n_clouds clouds arrive in out_rank0 ([n_clouds][k], host memory) where rank 0 lives; other ranks pass NULL. */
FPS_API int fps_b200_kdline_batch_sharded(const float *points, size_t n_clouds, size_t n, size_t dim, size_t k,
                                  const size_t *start, size_t height, size_t *out_rank0);
FPS_API int fps_b200_vanilla_batch_sharded(const float *points, size_t n_clouds, size_t n, size_t dim, size_t k,
                                   const size_t *start, size_t *out_rank0);
/* Collective: gather index arrays that were sampled separately (local = this rank's [nb][k] uint64, host or device). */
FPS_API int fps_b200_gather_indices(const uint64_t *local, size_t nb, size_t k, size_t n_clouds, uint64_t *out_rank0);

/* ---- batches, device pointers (inputs already resident in HBM) ---------------------------------- */

#define FPS_ALGO_VANILLA 0
#define FPS_ALGO_KDLINE 1
#define FPS_ALGO_KDTREE 2
#define FPS_ALGO_NPDU 3     /* host-pointer entries only */
#define FPS_ALGO_NPDU_KNN 4 /* host-pointer entries only */

/* bytes of scratch the *_dev calls need on the current device for this shape (256-byte aligned base) */
FPS_API size_t fps_b200_workspace_bytes(int algo, size_t B, size_t n, size_t dim, size_t k, size_t height);

/* d_points [B][n][dim] float32, d_start NULL or [B] uint64, d_out [B][k] uint64, all on the current
 * device.  Work is enqueued on `stream`; nothing is synchronised. */
FPS_API int fps_b200_vanilla_batch_dev(const float *d_points, size_t B, size_t n, size_t dim, size_t k,
                               const uint64_t *d_start, uint64_t *d_out, void *d_workspace,
                               size_t workspace_bytes, void *stream);
FPS_API int fps_b200_kdline_batch_dev(const float *d_points, size_t B, size_t n, size_t dim, size_t k,
                              const uint64_t *d_start, size_t height, uint64_t *d_out,
                              void *d_workspace, size_t workspace_bytes, void *stream);
FPS_API int fps_b200_kdtree_batch_dev(const float *d_points, size_t B, size_t n, size_t dim, size_t k,
                              const uint64_t *d_start, uint64_t *d_out, void *d_workspace,
                              size_t workspace_bytes, void *stream);

/* kd-line build only (testing / inspection): d_perm [B][n] uint32 (position -> original id),
 * d_leaf_lo [B][2^h + 1] uint32 (slot s covers positions [lo[s], lo[s+1]); empty slots allowed),
 * d_leaf_box [B][2^h][2][dim] float32 (lows then highs; undefined for empty slots). */
FPS_API int fps_b200_kdline_build_dev(const float *d_points, size_t B, size_t n, size_t dim, size_t height,
                              uint32_t *d_perm, uint32_t *d_leaf_lo, float *d_leaf_box,
                              void *d_workspace, size_t workspace_bytes, void *stream);

/* The kd build's split value (testing / inspection): d_sum[0] = the STRICTLY SEQUENTIAL binary32 sum of d_values[0..n)
 * (`float s = 0; for (...) s += x;`, src/_ext/KDTreeBase.h:151-158), evaluated by one warp through the tile-parallel code
 * of the build kernels (csrc/seqsum.cuh; tile = 256 or 512 elements; tile = -512: the two-phase form of the grid-wide build,
 * where 512-element tiles are prepared in parallel under a guessed binade and one warp walks their records).  d_fast_tiles
 * (may be NULL) receives how many tiles took the integer path instead of the dependent add chain. */
FPS_API int fps_b200_seqsum_dev(const float *d_values, size_t n, float *d_sum, uint32_t *d_fast_tiles, int tile,
                        void *stream);

/* ---- utilities ---------------------------------------------------------------------------------- */

FPS_API int fps_b200_device_count(void);            /* usable (sm_100) devices; 0 if none                     */
FPS_API const char *fps_b200_version(void);
FPS_API const char *fps_b200_last_error(void);      /* thread-local description of the last failure           */
FPS_API uint64_t fps_b200_kernel_launches(void);    /* kernels launched by this library in this process       */
FPS_API const char *fps_b200_last_plan(void);       /* thread-local: which kernel/shape the last call picked  */
/* diagnostics: phase / executed-work counters of the last launch of one sampler family (16 words, meaning per family in
 * fpsample_b200/capi.py); the kernels only count when the library is built with -DFPS_COUNTERS=1 (bench.py's W_exec). */
#define FPS_DBG_ASYNC 0
#define FPS_DBG_WARP 1
#define FPS_DBG_BUILD 2
#define FPS_DBG_GRID 3
#define FPS_DBG_STREAM 4 /* executed work of the streaming sampler (tuning knob COUNT=1): points scanned, point-updates,
                          bucket passes, early passes (full pending list), bucket tests, picks, clouds, distances written back */
FPS_API int fps_b200_debug_counters(int which, uint64_t *out16);
/* Planner overrides (tests, experiments).  The library reads FPS_B200_<NAME> from the environment ONCE, at first use; after
 * that only this call changes a knob.  value -1 = the planner's own choice.  Names: GRID, GROUP, GRIDBUILD, VANILLA_KD, PIPE,
 * ZEROCOPY, GRID_ECAP, WARP, WARP_TMEM, WARP_LAZY, WARP_HYBRID, WARP_GLOBAL_MINB, KDSMALL, STREAM_WARPS, STREAM_SPLIT, COUNT, PREFETCH, PSUM, STAGE. */
FPS_API int fps_b200_set_tuning(const char *name, long value);
/* Device-resident inputs of the host-pointer entries: the calling thread's NEXT call waits (on the device, through an event)
 * for the work queued on `stream` -- the stream that produced the points -- instead of the legacy default stream.  Nothing is
 * synchronised device-wide. */
FPS_API void fps_b200_set_producer_stream(void *stream);
/* Phase timing of the *_dev entries (measurement only, off by default): when enabled, the calling thread's next
 * *_dev call records CUDA events on its stream around the kd build launches and the sampling launch;
 * fps_b200_last_phase_ms waits for them and returns both durations (build = 0 for the vanilla entry). */
FPS_API void fps_b200_phase_timing(int enable);
FPS_API int fps_b200_last_phase_ms(float *build_ms, float *sample_ms);
/* Latency floor of a sampler's synchronisation structure (measurement only, csrc/floors.cu): `rounds` empty rounds of the
 * exchange the sampler performs per round, timed with CUDA events on the current device -> nanoseconds per round.
 *   FPS_FLOOR_WARP: one warp's arg-max collectives.  FPS_FLOOR_CLUSTER: `ctas` = cluster size (512 threads per CTA).
 *   FPS_FLOOR_GRID: `ctas` CTAs of 1024 threads, `words` = (CTAs per group << 16) | stamped 16-byte words per CTA
 *   (group 0 = the whole grid). */
#define FPS_FLOOR_WARP 0
#define FPS_FLOOR_CLUSTER 1
#define FPS_FLOOR_GRID 2
FPS_API int fps_b200_sync_floor(int kind, int ctas, int words, int rounds, float *ns_per_round);
/* How the streaming sampler (big batches of clouds that stay in HBM: BASELINE.json configs[4]) would cut a batch of
 * `n_clouds` clouds on a GPU of `n_sms` SMs into launches: full waves of narrow teams of warps + a partial wave of wide
 * ones, e.g. "3552 clouds x WPC=2 (...) + 544 clouds x WPC=4 (...)".  Pure host arithmetic (no device needed); honours
 * the tuning knobs STREAM_WARPS / STREAM_SPLIT.  FPS_ERR_UNSUPPORTED if the shape is not one the sampler takes. */
FPS_API int fps_b200_describe_stream_plan(size_t n_clouds, size_t n, size_t dim, size_t height, int n_sms, char *buf, size_t buf_len);
FPS_API void *fps_b200_host_alloc(size_t bytes);    /* page-locked host memory for the host-pointer entries   */
FPS_API void fps_b200_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* FPS_B200_H */
