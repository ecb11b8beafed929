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
 *     workspaces are private to the library.  `points` may also be a DEVICE pointer (the clouds are then
 *     sampled on the device they live on, without an upload, after a device synchronise); page-locked
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
 * POSITION in the array after the kd build permuted it (src/wrapper.hpp:54-55). */
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

/* ---- batches, device pointers (inputs already resident in HBM) ---------------------------------- */

#define FPS_ALGO_VANILLA 0
#define FPS_ALGO_KDLINE 1
#define FPS_ALGO_KDTREE 2
#define FPS_ALGO_NPDU 3 /* host-pointer entries only */

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
 * of the build kernels (csrc/seqsum.cuh; tile = 256 or 512 elements).  d_fast_tiles (may be NULL) receives how many tiles
 * took the integer prefix-scan path instead of the dependent add chain. */
FPS_API int fps_b200_seqsum_dev(const float *d_values, size_t n, float *d_sum, uint32_t *d_fast_tiles, int tile,
                        void *stream);

/* ---- utilities ---------------------------------------------------------------------------------- */

FPS_API int fps_b200_device_count(void);            /* usable (sm_100) devices; 0 if none                     */
FPS_API const char *fps_b200_version(void);
FPS_API const char *fps_b200_last_error(void);      /* thread-local description of the last failure           */
FPS_API uint64_t fps_b200_kernel_launches(void);    /* kernels launched by this library in this process       */
FPS_API const char *fps_b200_last_plan(void);       /* thread-local: which kernel/shape the last call picked  */
FPS_API int fps_b200_debug_counters(uint64_t *out16); /* phase counters of the last kd-line cluster launch (diagnostics) */
/* Phase timing of the *_dev entries (measurement only, off by default): when enabled, the calling thread's next
 * *_dev call records CUDA events on its stream around the kd build launches and the sampling launch;
 * fps_b200_last_phase_ms waits for them and returns both durations (build = 0 for the vanilla entry). */
FPS_API void fps_b200_phase_timing(int enable);
FPS_API int fps_b200_last_phase_ms(float *build_ms, float *sample_ms);
FPS_API void *fps_b200_host_alloc(size_t bytes);    /* page-locked host memory for the host-pointer entries   */
FPS_API void fps_b200_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* FPS_B200_H */
