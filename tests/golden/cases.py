"""The golden-vector case list shared by make_golden.py (generation, from the compiled reference) and the
tests (oracle vs golden on CPU, CUDA path vs golden on the GPU).  Inputs are regenerated from seeds
(fpsample_b200.synth); only the reference's OUTPUT indices are stored (tests/golden/golden.npz), plus a
sha256 of every input so a drifting generator is detected instead of silently "passing".

case = (id, input spec, call, params);   input spec = (generator, *args);   call in {"vanilla","kdline","kdtree"}
"""
from __future__ import annotations

import hashlib

import numpy as np

from fpsample_b200 import synth


def np_rand42(n, d):  # the reference bench's own input (bench/test_bench.py:19-21), float64
    np.random.seed(42)
    return np.random.rand(n, d)


def duplicates(n, d):  # every point identical: pure tie-rule probe
    return np.full((n, d), 0.25, dtype=np.float32)


GENERATORS = {
    "rand42": np_rand42,
    "uniform": synth.uniform,
    "lidar": synth.lidar,
    "grid": synth.grid_ties,
    "dup": duplicates,
}


def make_input(spec):
    return GENERATORS[spec[0]](*spec[1:])


def input_sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float32).tobytes()).hexdigest()[:16]


CASES = [
    # --- SURVEY.md Appendix C known answers (G0..G6) -------------------------------------------------------
    ("G0_vanilla", ("rand42", 4096, 3), "vanilla", dict(k=1024, start=0)),
    ("G0_kd_h3", ("rand42", 4096, 3), "kdline", dict(k=1024, h=3, start=0)),
    ("G0_kd_h5", ("rand42", 4096, 3), "kdline", dict(k=1024, h=5, start=0)),
    ("G0_kd_h7", ("rand42", 4096, 3), "kdline", dict(k=1024, h=7, start=0)),
    ("G0_kd_h5_s1", ("rand42", 4096, 3), "kdline", dict(k=64, h=5, start=1)),
    ("G0_kd_h5_s4095", ("rand42", 4096, 3), "kdline", dict(k=64, h=5, start=4095)),
    ("G1_vanilla", ("uniform", 1, 4096, 3), "vanilla", dict(k=1024, start=0)),          # BASELINE cfg 1
    ("G1_kd_h5", ("uniform", 1, 4096, 3), "kdline", dict(k=1024, h=5, start=0)),
    ("G2_vanilla", ("uniform", 2, 16384, 3), "vanilla", dict(k=4096, start=0)),
    ("G2_kd_h7", ("uniform", 2, 16384, 3), "kdline", dict(k=4096, h=7, start=0)),
    ("G3_vanilla", ("uniform", 3, 100000, 3), "vanilla", dict(k=8192, start=0)),
    ("G3_kd_h7", ("uniform", 3, 100000, 3), "kdline", dict(k=8192, h=7, start=0)),
    ("G4_vanilla", ("uniform", 4, 100000, 6), "vanilla", dict(k=8192, start=0)),
    ("G4_kd_h7", ("uniform", 4, 100000, 6), "kdline", dict(k=8192, h=7, start=0)),
    ("G5_kd_h9", ("uniform", 5, 2**20, 3), "kdline", dict(k=65536, h=9, start=0)),      # BASELINE cfg 4
    ("G6_kd_h9", ("lidar", 6, 2**20), "kdline", dict(k=65536, h=9, start=0)),           # BASELINE cfg 4 (lidar)
    # --- BASELINE cfg 2 / 3 / 5: the first clouds of each batch -----------------------------------------------
    *[(f"cfg2_b{b}_kd", ("uniform", 1000 + b, 4096, 3), "kdline", dict(k=1024, h=5, start=0)) for b in range(4)],
    *[(f"cfg2_b{b}_vanilla", ("uniform", 1000 + b, 4096, 3), "vanilla", dict(k=1024, start=0)) for b in range(2)],
    *[(f"cfg3_b{b}_kd", ("uniform", 2000 + b, 16384, 3), "kdline", dict(k=4096, h=7, start=0)) for b in range(2)],
    ("cfg3_b0_vanilla", ("uniform", 2000, 16384, 3), "vanilla", dict(k=4096, start=0)),
    ("cfg5_b0_d3_kd", ("uniform", 3000, 100000, 3), "kdline", dict(k=8192, h=7, start=0)),
    ("cfg5_b0_d6_kd", ("uniform", 3000, 100000, 6), "kdline", dict(k=8192, h=7, start=0)),
    ("cfg5_b1_d3_vanilla", ("uniform", 3001, 100000, 3), "vanilla", dict(k=8192, start=0)),
    # --- tie rules, duplicates, forced starts, odd dims, degenerate splits ------------------------------------
    ("dup_vanilla", ("dup", 10, 3), "vanilla", dict(k=5, start=2)),
    ("dup_kd", ("dup", 10, 3), "kdline", dict(k=5, h=2, start=2)),
    *[(f"grid_d{d}_vanilla", ("grid", 10 + d, 3000, d), "vanilla", dict(k=500, start=[5, 1, 9]))
      for d in (1, 2, 3, 6)],
    *[(f"grid_d{d}_kd_h6", ("grid", 10 + d, 3000, d), "kdline", dict(k=500, h=6, start=d)) for d in (1, 2, 3, 6)],
    ("grid_kd_exhaust", ("grid", 99, 64, 2, 3), "kdline", dict(k=64, h=6, start=0)),   # h exhausts the points
    ("grid_vanilla_all", ("grid", 98, 200, 3, 4), "vanilla", dict(k=200, start=7)),    # k == n, repeats legal
    ("multi_start", ("uniform", 77, 4096, 3), "vanilla", dict(k=256, start=[1, 2, 50, 4000])),
    ("lidar_small_vanilla", ("lidar", 21, 20000), "vanilla", dict(k=2048, start=3)),
    ("lidar_small_kd_h7", ("lidar", 21, 20000), "kdline", dict(k=2048, h=7, start=3)),
    ("lidar_100k_kd_h7", ("lidar", 22, 100000), "kdline", dict(k=8192, h=7, start=0)),
    *[(f"dim{d}_vanilla", ("uniform", 40 + d, 2000, d), "vanilla", dict(k=300, start=d)) for d in (1, 2, 4, 5, 7, 8, 12)],
    *[(f"dim{d}_kd_h4", ("uniform", 40 + d, 2000, d), "kdline", dict(k=300, h=4, start=d)) for d in (1, 2, 4, 5, 7, 8)],
    ("n_odd_vanilla", ("uniform", 60, 4099, 3), "vanilla", dict(k=1000, start=4098)),
    ("n_odd_kd_h5", ("uniform", 60, 4099, 3), "kdline", dict(k=1000, h=5, start=4098)),
    ("k1_vanilla", ("uniform", 61, 100, 3), "vanilla", dict(k=1, start=42)),
    ("k1_kd", ("uniform", 61, 100, 3), "kdline", dict(k=1, h=3, start=42)),
    # --- bucket_fps_kdtree_sampling (full kd tree) ----------------------------------
    ("G0_kdtree", ("rand42", 4096, 3), "kdtree", dict(k=1024, start=0)),
    ("G1_kdtree", ("uniform", 1, 4096, 3), "kdtree", dict(k=1024, start=0)),
    ("G2_kdtree", ("uniform", 2, 16384, 3), "kdtree", dict(k=4096, start=5)),
    ("G3_kdtree", ("uniform", 3, 100000, 3), "kdtree", dict(k=2048, start=0)),
    ("dup_kdtree", ("dup", 10, 3), "kdtree", dict(k=5, start=2)),
    *[(f"grid_d{d}_kdtree", ("grid", 10 + d, 3000, d), "kdtree", dict(k=500, start=d)) for d in (1, 2, 3, 6)],
    ("grid_kdtree_all", ("grid", 98, 200, 3, 4), "kdtree", dict(k=200, start=7)),      # k == n, repeats legal
    ("lidar_small_kdtree", ("lidar", 21, 20000), "kdtree", dict(k=2048, start=3)),
    *[(f"dim{d}_kdtree", ("uniform", 40 + d, 2000, d), "kdtree", dict(k=300, start=d)) for d in (1, 2, 4, 5, 7, 8)],
    ("n_odd_kdtree", ("uniform", 60, 4099, 3), "kdtree", dict(k=1000, start=4098)),
    ("k1_kdtree", ("uniform", 61, 100, 3), "kdtree", dict(k=1, start=42)),
    # --- fps_npdu_sampling (index-window heuristic, src/lib.cpp:272-340); w = window ----
    ("G0_npdu", ("rand42", 4096, 3), "npdu", dict(k=1024, w=64, start=0)),
    ("cfg3_npdu", ("uniform", 2000, 16384, 3), "npdu", dict(k=4096, w=64, start=5)),
    ("grid_d2_npdu", ("grid", 12, 3000, 2), "npdu", dict(k=500, w=10, start=7)),
    ("grid_d1_npdu", ("grid", 11, 777, 1), "npdu", dict(k=300, w=33, start=776)),
    ("wide_npdu", ("uniform", 70, 4099, 3), "npdu", dict(k=1000, w=2000, start=0)),
    ("full_npdu", ("uniform", 71, 5000, 3), "npdu", dict(k=5000, w=4999, start=3)),
    ("w1_npdu", ("uniform", 72, 2000, 6), "npdu", dict(k=300, w=1, start=5)),
    ("w0_npdu", ("uniform", 73, 500, 3), "npdu", dict(k=100, w=0, start=499)),
    ("lidar_npdu", ("lidar", 21, 20000), "npdu", dict(k=2048, w=156, start=11)),
    ("dim12_npdu", ("uniform", 52, 2000, 12), "npdu", dict(k=300, w=100, start=12)),
    ("dup_npdu", ("dup", 10, 3), "npdu", dict(k=5, w=4, start=2)),
    # --- SURVEY.md section 8(f) row 4, second half: fps_npdu_kdtree_sampling (k nearest neighbours, src/lib.cpp:369-465); w = k.
    #     Generic clouds only: exact ties AT the k-th nearest distance are resolved by nanoflann's traversal order there ----
    ("G0_npdukd", ("rand42", 4096, 3), "npdukd", dict(k=1024, w=64, start=0)),
    ("cfg3_npdukd", ("uniform", 2000, 16384, 3), "npdukd", dict(k=2048, w=128, start=5)),
    ("wide_npdukd", ("uniform", 70, 4099, 3), "npdukd", dict(k=1000, w=2000, start=0)),
    ("full_npdukd", ("uniform", 71, 3000, 3), "npdukd", dict(k=3000, w=3000, start=3)),
    ("over_npdukd", ("uniform", 74, 1500, 2), "npdukd", dict(k=200, w=10**6, start=1499)),
    ("w1_npdukd", ("uniform", 72, 2000, 6), "npdukd", dict(k=300, w=1, start=5)),
    ("lidar_npdukd", ("lidar", 21, 20000), "npdukd", dict(k=1024, w=312, start=11)),
    ("dim12_npdukd", ("uniform", 52, 2000, 12), "npdukd", dict(k=300, w=100, start=12)),
    ("d1_npdukd", ("uniform", 75, 5000, 1), "npdukd", dict(k=400, w=40, start=0)),
]

CASE_BY_ID = {c[0]: c for c in CASES}
# cases whose oracle replay is slow on one core (vanilla on 100k points): CPU suite runs them once, GPU suite always
SLOW_ON_CPU = {"G3_vanilla", "G4_vanilla", "cfg5_b1_d3_vanilla"}
