// pymodule.cpp -- pybind11 module `_fpsample`: the host-side mirror of the reference's binding layer for
// the two hot entry points (src/lib.cpp:249-270 `_fps_sampling`, :522-579 `_bucket_fps_kdline_sampling`),
// same positional signatures, same exception types, numpy arrays in, uint64 index arrays out.  All
// compute goes through the C ABI in include/fps_b200.h; there is no CPU path in this file.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <map>
#include <mutex>
#include <optional>
#include <string>
#include <vector>

#include "../../include/fps_b200.h"

namespace py = pybind11;
using farray = py::array_t<float, py::array::c_style | py::array::forcecast>;
using iarray = py::array_t<size_t, py::array::c_style | py::array::forcecast>;

static void raise_rc(const char *fn, int rc) {
    // the reference surfaces a non-zero C return code as RuntimeError (src/lib.cpp:574-576)
    std::string msg = std::string(fn) + " failed with error code " + std::to_string(rc);
    const char *why = fps_b200_last_error();
    if (why && *why) msg += std::string(": ") + why;
    throw std::runtime_error(msg);
}

// src/lib.cpp:249-270 + check_py_input (:52-109)
static py::array_t<size_t> fps_sampling_py(farray points, size_t n_samples, py::object start_idx_obj) {
    std::vector<size_t> starts;
    if (py::isinstance<py::int_>(start_idx_obj)) {
        starts.push_back(start_idx_obj.cast<size_t>());
    } else if (py::isinstance<py::array_t<size_t>>(start_idx_obj)) {
        auto arr = start_idx_obj.cast<py::array_t<size_t>>();
        if (arr.ndim() != 1) throw py::type_error("start_idx must be int or 1D numpy array of size_t");
        auto v = arr.unchecked<1>();
        for (py::ssize_t i = 0; i < v.shape(0); ++i) starts.push_back(v(i));
        if (starts.empty()) throw py::value_error("start_idx array must not be empty");
    } else {
        throw py::type_error("start_idx must be int or 1D numpy array of size_t");
    }
    if (points.ndim() != 2)
        throw py::value_error("points must be a 2D array, but got shape " + std::to_string(points.ndim()));
    const size_t P = (size_t)points.shape(0), C = (size_t)points.shape(1);
    if (C == 0) throw py::value_error("points must have at least one column");
    if (n_samples > P)
        throw py::value_error("n_samples must be less than the number of points: n_samples=" +
                              std::to_string(n_samples) + ", P=" + std::to_string(P));
    if (starts.size() > n_samples && !(starts.size() == 1 && n_samples == 0))
        throw py::value_error("The number of start indices must be less than or equal to n_samples: " +
                              std::to_string(starts.size()) + ", n_samples=" + std::to_string(n_samples));
    for (size_t s : starts)
        if (s >= P)
            throw py::value_error("start_idx must be less than the number of points: start_idx=" + std::to_string(s) +
                                  ", P=" + std::to_string(P));
    py::array_t<size_t> out(n_samples);
    if (n_samples == 0) return out;
    int rc;
    {
        const float *src = points.data();
        size_t *dst = out.mutable_data();
        py::gil_scoped_release rel;
        rc = fps_b200_vanilla(src, P, C, n_samples, starts.data(), starts.size(), dst);
    }
    if (rc != 0) raise_rc("fps_b200_vanilla", rc);
    return out;
}

// src/lib.cpp:522-579
static py::array_t<size_t> kdline_py(farray points, size_t n_samples, size_t height, py::object start_idx_obj) {
    size_t start = 0;
    if (py::isinstance<py::int_>(start_idx_obj)) {
        start = start_idx_obj.cast<size_t>();
    } else if (py::isinstance<py::array_t<size_t>>(start_idx_obj)) {
        PyErr_SetString(PyExc_NotImplementedError, "Array of start indices not implemented yet");
        throw py::error_already_set();
    } else {
        throw py::type_error("start_idx must be int or 1D numpy array of size_t");
    }
    if (points.ndim() != 2) throw py::value_error("points must be a 2D float32 array");
    const size_t P = (size_t)points.shape(0), C = (size_t)points.shape(1);
    if (start >= P) throw py::value_error("start_idx out of range");
    if (n_samples == 0 || n_samples > P) throw py::value_error("n_samples must be in [1, num_points]");
    if (height == 0) throw py::value_error("height must be >= 1");
    py::array_t<size_t> out(n_samples);
    int rc;
    {
        const float *src = points.data();
        size_t *dst = out.mutable_data();
        py::gil_scoped_release rel;
        rc = bucket_fps_kdline(src, P, C, n_samples, start, height, dst);
    }
    if (rc != 0) raise_rc("bucket_fps_kdline", rc);
    return out;
}

// src/lib.cpp:467-520
static py::array_t<size_t> kdtree_py(farray points, size_t n_samples, py::object start_idx_obj) {
    size_t start = 0;
    if (py::isinstance<py::int_>(start_idx_obj)) {
        start = start_idx_obj.cast<size_t>();
    } else if (py::isinstance<py::array_t<size_t>>(start_idx_obj)) {
        PyErr_SetString(PyExc_NotImplementedError, "Array of start indices not implemented yet");
        throw py::error_already_set();
    } else {
        throw py::type_error("start_idx must be int or 1D numpy array of size_t");
    }
    if (points.ndim() != 2) throw py::value_error("points must be a 2D float32 array");
    const size_t P = (size_t)points.shape(0), C = (size_t)points.shape(1);
    if (C == 0) throw py::value_error("points must have at least one column");
    if (start >= P) throw py::value_error("start_idx out of range");
    if (n_samples == 0 || n_samples > P) throw py::value_error("n_samples must be in [1, num_points]");
    py::array_t<size_t> out(n_samples);
    int rc;
    {
        const float *src = points.data();
        size_t *dst = out.mutable_data();
        py::gil_scoped_release rel;
        rc = bucket_fps_kdtree(src, P, C, n_samples, start, dst);
    }
    if (rc != 0) raise_rc("bucket_fps_kdtree", rc);
    return out;
}

// src/lib.cpp:342-366 (fps_npdu_sampling_py) + check_py_input (:52-109) -> fps_npdu_sampling (:272-340)
static py::array_t<size_t> npdu_py(farray points, size_t n_samples, size_t k, py::object start_idx_obj) {
    size_t start = 0;
    if (py::isinstance<py::int_>(start_idx_obj)) {
        start = start_idx_obj.cast<size_t>();
    } else if (py::isinstance<py::array_t<size_t>>(start_idx_obj)) {
        PyErr_SetString(PyExc_NotImplementedError, "Array of start indices not implemented yet");
        throw py::error_already_set();
    } else {
        throw py::type_error("start_idx must be int or 1D numpy array of size_t");
    }
    if (points.ndim() != 2)
        throw py::value_error("points must be a 2D array, but got shape " + std::to_string(points.ndim()));
    const size_t P = (size_t)points.shape(0), C = (size_t)points.shape(1);
    if (C == 0) throw py::value_error("points must have at least one column");
    if (n_samples > P)
        throw py::value_error("n_samples must be less than the number of points: n_samples=" + std::to_string(n_samples) +
                              ", P=" + std::to_string(P));
    if (start >= P)
        throw py::value_error("start_idx must be less than the number of points: start_idx=" + std::to_string(start) +
                              ", P=" + std::to_string(P));
    py::array_t<size_t> out(n_samples);
    if (n_samples == 0) return out;
    int rc;
    {
        const float *src = points.data();
        size_t *dst = out.mutable_data();
        py::gil_scoped_release rel;
        rc = fps_b200_npdu(src, P, C, n_samples, k, start, dst);
    }
    if (rc != 0) raise_rc("fps_b200_npdu", rc);
    return out;
}

// src/lib.cpp:369-465 (fps_npdu_kdtree_sampling_py) + check_py_input (:52-109)
static py::array_t<size_t> npdu_kdtree_py(farray points, size_t n_samples, size_t k, py::object start_idx_obj) {
    size_t start = 0;
    bool is_array = false;
    if (py::isinstance<py::int_>(start_idx_obj)) start = start_idx_obj.cast<size_t>();
    else if (py::isinstance<py::array_t<size_t>>(start_idx_obj)) is_array = true;
    else throw py::type_error("start_idx must be int or 1D numpy array of size_t");
    if (points.ndim() != 2)
        throw py::value_error("points must be a 2D array, but got shape " + std::to_string(points.ndim()));
    const size_t P = (size_t)points.shape(0), C = (size_t)points.shape(1);
    if (C == 0) throw py::value_error("points must have at least one column");
    if (n_samples > P)
        throw py::value_error("n_samples must be less than the number of points: n_samples=" + std::to_string(n_samples) +
                              ", P=" + std::to_string(P));
    if (is_array) {   // the reference validates first and refuses afterwards (src/lib.cpp:385-390)
        PyErr_SetString(PyExc_NotImplementedError, "Array of start indices not implemented yet");
        throw py::error_already_set();
    }
    if (start >= P)
        throw py::value_error("start_idx must be less than the number of points: start_idx=" + std::to_string(start) +
                              ", P=" + std::to_string(P));
    py::array_t<size_t> out(n_samples);
    if (n_samples == 0) return out;
    int rc;
    {
        const float *src = points.data();
        size_t *dst = out.mutable_data();
        py::gil_scoped_release rel;
        rc = fps_b200_npdu_kdtree(src, P, C, n_samples, k, start, dst);
    }
    if (rc != 0) raise_rc("fps_b200_npdu_kdtree", rc);
    return out;
}

// ---- batched entries (new) -----------------------------------------------------------------------------
// Large index arrays are handed out over page-locked memory from the library (fps_b200_host_alloc): the device-to-host
// copy then runs at PCIe speed straight into the array the caller gets, with no pageable bounce buffer and no first-touch
// page faults.  A buffer goes back to a small free list when its array (and every view of it) is collected, so a
// result never aliases a later one while it is alive.
namespace {
struct PinnedPool {
    std::mutex mu;
    std::multimap<size_t, void *> free_;
    size_t held = 0;
    static constexpr size_t kMaxHeld = (size_t)1 << 30;
    void *get(size_t cap) {
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = free_.find(cap);
            if (it != free_.end()) {
                void *p = it->second;
                free_.erase(it);
                held -= cap;
                return p;
            }
        }
        return fps_b200_host_alloc(cap);
    }
    void put(void *p, size_t cap) {
        {
            std::lock_guard<std::mutex> lk(mu);
            if (held + cap <= kMaxHeld) {
                free_.emplace(cap, p);
                held += cap;
                return;
            }
        }
        fps_b200_host_free(p);
    }
};
PinnedPool g_pool;
struct PinnedBlock {
    void *p;
    size_t cap;
};
}  // namespace

static py::array_t<size_t> index_array(size_t B, size_t k) {
    const size_t bytes = B * k * sizeof(size_t);
    if (bytes >= ((size_t)1 << 20)) {
        const size_t cap = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        if (void *p = g_pool.get(cap)) {
            auto *blk = new PinnedBlock{p, cap};
            py::capsule owner(blk, [](void *q) {
                auto *b = static_cast<PinnedBlock *>(q);
                g_pool.put(b->p, b->cap);
                delete b;
            });
            return py::array_t<size_t>({(py::ssize_t)B, (py::ssize_t)k}, static_cast<size_t *>(p), owner);
        }
    }
    return py::array_t<size_t>({(py::ssize_t)B, (py::ssize_t)k});
}

static std::vector<size_t> batch_starts(py::object start_obj, size_t B, size_t P) {
    std::vector<size_t> st;
    if (start_obj.is_none()) return st;
    if (py::isinstance<py::int_>(start_obj)) {
        st.assign(B, start_obj.cast<size_t>());
    } else {
        auto arr = iarray::ensure(start_obj);
        if (!arr || arr.ndim() != 1 || (size_t)arr.shape(0) != B)
            throw py::type_error("start_idx must be None, int or a 1D integer array of length B");
        st.assign(arr.data(), arr.data() + B);
    }
    for (size_t s : st)
        if (s >= P) throw py::value_error("start_idx out of range");
    return st;
}

static py::array_t<size_t> batch_core(int algo, const float *src, size_t B, size_t P, size_t C, size_t n_samples, size_t height,
                                      py::object start_obj, py::object devices_obj) {
    if (B == 0) throw py::value_error("batch must hold at least one cloud");
    if (C == 0) throw py::value_error("points must have at least one column");
    if (n_samples == 0 || n_samples > P) throw py::value_error("n_samples must be in [1, num_points]");
    if (algo == FPS_ALGO_KDLINE && height == 0) throw py::value_error("height must be >= 1");
    std::vector<size_t> st = batch_starts(start_obj, B, P);
    std::vector<int> devs;
    if (!devices_obj.is_none()) devs = devices_obj.cast<std::vector<int>>();
    py::array_t<size_t> out = index_array(B, n_samples);
    int rc;
    {
        size_t *dst = out.mutable_data();
        const size_t *sp = st.empty() ? nullptr : st.data();
        const int *dp = devs.empty() ? nullptr : devs.data();
        py::gil_scoped_release rel;
        if (algo == FPS_ALGO_VANILLA)
            rc = fps_b200_vanilla_batch(src, B, P, C, n_samples, sp, dst, dp, (int)devs.size());
        else if (algo == FPS_ALGO_KDTREE)
            rc = fps_b200_kdtree_batch(src, B, P, C, n_samples, sp, dst, dp, (int)devs.size());
        else
            rc = fps_b200_kdline_batch(src, B, P, C, n_samples, sp, height, dst, dp, (int)devs.size());
    }
    if (rc != 0)
        raise_rc(algo == FPS_ALGO_VANILLA ? "fps_b200_vanilla_batch" : algo == FPS_ALGO_KDTREE ? "fps_b200_kdtree_batch" : "fps_b200_kdline_batch", rc);
    return out;
}

static py::array_t<size_t> batch_py(int algo, farray points, size_t n_samples, size_t height, py::object start_obj,
                                    py::object devices_obj) {
    if (points.ndim() != 3) throw py::value_error("points must be a 3D array [B, N, D]");
    return batch_core(algo, points.data(), (size_t)points.shape(0), (size_t)points.shape(1), (size_t)points.shape(2), n_samples,
                      height, start_obj, devices_obj);
}

// the same over a raw address: a C-contiguous float32 [B, N, D] buffer that already lives in GPU memory (torch / cupy /
// numba arrays through __cuda_array_interface__); the library samples it where it is, no upload
static py::array_t<size_t> batch_ptr_py(int algo, size_t address, size_t B, size_t P, size_t C, size_t n_samples, size_t height,
                                        py::object start_obj, py::object devices_obj) {
    if (address == 0) throw py::value_error("null device pointer");
    return batch_core(algo, reinterpret_cast<const float *>(address), B, P, C, n_samples, height, start_obj, devices_obj);
}

PYBIND11_MODULE(_fpsample, m, py::mod_gil_not_used()) {
    m.doc() = "B200-native farthest point sampling: drop-in for fpsample._fpsample's FPS hot path";
    m.def("_fps_sampling", &fps_sampling_py,
          "Vanilla FPS. points: N x C float32; n_samples; start_idx: int or 1D uint64 array. Returns uint64[n_samples].");
    m.def("_bucket_fps_kdline_sampling", &kdline_py,
          "QuickFPS kd-line. points: N x C float32 (C <= 8); n_samples; height; start_idx: int. Returns uint64[n_samples].");
    m.def("_bucket_fps_kdtree_sampling", &kdtree_py,
          "QuickFPS full kd tree. points: N x C float32 (C <= 8); n_samples; start_idx: int. Returns uint64[n_samples].");
    m.def("_fps_npdu_sampling", &npdu_py,
          "FPS with the NPDU index-window heuristic. points: N x C float32; n_samples; k (window); start_idx: int. Returns uint64[n_samples].");
    m.def("_fps_npdu_kdtree_sampling", &npdu_kdtree_py,
          "FPS with the NPDU heuristic over the k nearest points. points: N x C float32; n_samples; k; start_idx: int. Returns uint64[n_samples].");
    m.def("_bucket_fps_kdtree_sampling_batch",
          [](farray p, size_t k, py::object s, py::object d) { return batch_py(FPS_ALGO_KDTREE, p, k, 0, s, d); },
          "Batched QuickFPS full kd tree. points: B x N x C; start_idx: None|int|int[B]; devices: None|list[int].");
    m.def("_fps_sampling_batch",
          [](farray p, size_t k, py::object s, py::object d) { return batch_py(FPS_ALGO_VANILLA, p, k, 0, s, d); },
          "Batched vanilla FPS. points: B x N x C; start_idx: None|int|int[B]; devices: None|list[int].");
    m.def("_bucket_fps_kdline_sampling_batch",
          [](farray p, size_t k, size_t h, py::object s, py::object d) { return batch_py(FPS_ALGO_KDLINE, p, k, h, s, d); },
          "Batched QuickFPS kd-line. points: B x N x C; height; start_idx: None|int|int[B]; devices: None|list[int].");
    m.def("_batch_ptr", &batch_ptr_py,
          "Batched entry over a raw float32 [B, N, D] address (device memory): algo 0 vanilla / 1 kd-line / 2 kd tree.");
    m.def("_set_producer_stream", [](size_t stream) { fps_b200_set_producer_stream(reinterpret_cast<void *>(stream)); },
          "The calling thread's next call with a device-resident input waits (on the device) for this cudaStream_t.");
    m.def("_device_count", []() { return fps_b200_device_count(); });
    m.def("_last_plan", []() { return std::string(fps_b200_last_plan()); });
    m.def("_kernel_launches", []() { return fps_b200_kernel_launches(); });
    m.attr("__version__") = "1.0.2+b200.0.1.0";
}
