#!/usr/bin/env python
"""bench.py -- the measurement contract of this repo (one JSON line on stdout, rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg1|cfg3|cfg4|cfg4l|cfg5d3|cfg5d6]
                  [--algo kdline|vanilla] [--impl b200|reference] [--no-extras] [--no-cfg5]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N

A "step" is one pass of the FPS hot path over one batch of synthetic clouds (BASELINE.json configs).

Headline line (default cfg2 = `bucket_fps_kdline_sampling` 4096x3 -> 1024, h=5, batch of 1024 clouds PER GPU -- weak
scaling: clouds are independent, every rank samples its own shard, no data-path collective):
  value        : clouds/s, inputs already resident in HBM, device pointers through the C ABI (*_batch_dev), timed with
                 CUDA events on the launching stream, per step, L2 flushed between steps; max over ranks.
  e2e          : the same metric through the public python API (fpsample_b200.*_batch) with HOST page-locked buffers:
                 H2D of the clouds and D2H of the indices inside the timed region, every step.
  e2e_pageable : the same call with a plain (pageable) numpy array, what a drop-in user of the reference holds.
  h2d_floor    : the host->device copy alone (no kernels), all ranks at once -- what e2e can at best reach.
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md section "Measurement".

extra.cfg5d3 / extra.cfg5d6 (every N): BASELINE.json configs[4] as STRONG scaling -- ONE batch of 4096 clouds x 100 000
points cut into contiguous shards over the N GPUs; e2e there includes the NCCL gather of all indices to rank 0
(fps_b200_kdline_batch_sharded, csrc/comm.cu).  Its roofline is HBM with the bytes the sampler really moved
(executed-work counters of the kernel).
extra.<cfgX> (N = 1): the other configs, device-resident, single clouds with their latency floor.

torch is used for device buffers, events, streams and the launcher's barrier / max-reduction only.  The oracle
(oracle/) is executed here ONLY in the cpu_baseline leg, the `--impl reference` arm and the post-run sanity checks.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B, n, d, k, h, generator, base seed, BASELINE.json config it is)   B: per GPU (cfg5: the whole job)
    "cfg1": (1, 4096, 3, 1024, 5, "uniform", 1, "configs[0] fps_sampling 4096x3->1024 single cloud"),
    "cfg2": (1024, 4096, 3, 1024, 5, "uniform", 1000, "configs[1] 4096x3->1024, h=5, batch of 1024 clouds"),
    "cfg3": (64, 16384, 3, 4096, 7, "uniform", 2000, "configs[2] PointNet++ SA batch B=64 x 16384x3->4096, h=7"),
    "cfg4": (1, 2**20, 3, 65536, 9, "uniform", 5, "configs[3] single cloud 1,048,576x3->65536, h=9 (uniform)"),
    "cfg4l": (1, 2**20, 3, 65536, 9, "lidar", 6, "configs[3] single LiDAR-like cloud 1,048,576x3->65536, h=9"),
    "cfg5d3": (4096, 100000, 3, 8192, 7, "uniform", 3000, "configs[4] batch of 4096 clouds x 100k x3 -> 8192"),
    "cfg5d6": (4096, 100000, 6, 8192, 7, "uniform", 3000, "configs[4] batch of 4096 clouds x 100k x6 -> 8192"),
}
SM_COUNT, LANES = 148, 128

_SYNTH = None


def synth():
    """fpsample_b200/synth.py loaded by PATH: the reference arm must not import the package (its __init__ loads the CUDA
    extension), so that `--impl reference` maps nothing but oracle/_ref."""
    global _SYNTH
    if _SYNTH is None:
        spec = importlib.util.spec_from_file_location("fps_bench_synth", os.path.join(ROOT, "fpsample_b200", "synth.py"))
        _SYNTH = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_SYNTH)
    return _SYNTH


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), float(j.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def make_cloud(gen, seed, n, d):
    return synth().lidar(seed, n) if gen == "lidar" else synth().uniform(seed, n, d)


def config_of(workload, algo):
    """identical in both arms (the driver compares them key by key)"""
    B, n, d, k, h, gen, seed, desc = WORKLOADS[workload]
    return {"workload": f"{workload}: {desc}; entry={'fps_sampling' if algo == 'vanilla' else 'bucket_fps_kdline_sampling'}, "
                        f"start_idx=0, clouds seeded {seed}+i",
            "clouds_per_gpu": B, "n": n, "d": d, "k": k, "h": h if algo == "kdline" else None, "algo": algo,
            "l2": "b200 arm: inputs (%.0f MB/GPU) < L2, so a 256 MiB write flushes L2 between timed steps, outside the per-step "
                  "CUDA-event pairs" % (B * n * d * 4 / 1e6)}


# ---- CPU arm: the reference's own implementation on the host cores, one cloud per core ----------------------
_REF = None
_CLOUDS = []   # generated in the parent BEFORE the pool forks: workers inherit them, timers see compute only


def _cpu_init():
    global _REF
    from oracle import oracle as O
    r = O.load_reference()
    _REF = ("reference", r) if r is not None else ("port", O)


def _cpu_one(task):
    algo, i, k, h = task
    pc = _CLOUDS[i]
    kind, m = _REF
    t = time.perf_counter()
    if kind == "reference":
        out = m.fps_sampling(pc, k, 0) if algo == "vanilla" else m.bucket_fps_kdline_sampling(pc, k, h, 0)
    else:
        out = m.fps_vanilla(pc, k, 0) if algo == "vanilla" else m.kdline(pc, k, h, 0)
    return time.perf_counter() - t, int(out[-1])


class CpuArm:
    """multiprocessing pool over all host cores (the reference holds the GIL: processes, not threads).
    A bounded sample of the workload's clouds is generated up front; run() = pool.map over it, wall-clock."""

    def __init__(self, algo, wl, budget_s):
        import multiprocessing as mp
        global _CLOUDS
        self.algo, self.wl = algo, wl
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        B, n, d, k, h, gen, seed, _ = WORKLOADS[wl]
        self.workers = max(1, min(self.cores, B))
        _cpu_init()
        self.kind = _REF[0]
        _CLOUDS = [make_cloud(gen, seed, n, d)]
        self.t1 = min(_cpu_one((algo, 0, k, h))[0] for _ in range(2 if n * k < 1e9 else 1))   # probe, one core
        waves = max(1, int(budget_s / max(self.t1 * 1.5, 1e-4)))
        self.n_clouds = max(1, min(B, self.workers * waves, max(self.workers, int(2e9 // (n * d * 4)))))
        _CLOUDS = [make_cloud(gen, seed + i, n, d) for i in range(self.n_clouds)]
        self.reps = max(1, min(50, int(budget_s * self.workers / max(self.t1 * 1.5 * self.n_clouds, 1e-4))))
        self.pool = mp.get_context("fork").Pool(self.workers, initializer=_cpu_init)
        self.pool.map(_cpu_one, [(algo, i % self.n_clouds, k, h) for i in range(self.workers)], chunksize=1)  # warm

    def run(self):
        """one pass over the sample -> (wall seconds, clouds done, mean per-cloud compute seconds)"""
        B, n, d, k, h, gen, seed, _ = WORKLOADS[self.wl]
        tasks = [(self.algo, i, k, h) for i in range(self.n_clouds)]
        cs = max(1, self.n_clouds // (self.workers * 8))
        t = time.perf_counter()
        res = self.pool.map(_cpu_one, tasks, chunksize=cs)
        wall = time.perf_counter() - t
        return wall, self.n_clouds, statistics.mean(r[0] for r in res)

    def describe(self, per):
        return (f"{self.n_clouds} clouds of the workload per pass, one cloud per core on {self.workers} of {self.cores} "
                f"host cores, mean {per * 1e3:.3f} ms/cloud/core under load ({self.t1 * 1e3:.3f} ms alone), "
                f"{'compiled unmodified reference' if self.kind == 'reference' else 'C oracle port'}, {self.algo}")

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_measure(arm: CpuArm):
    walls, pers = [], []
    for _ in range(arm.reps):
        w, done, per = arm.run()
        walls.append(w)
        pers.append(per)
    thr = arm.n_clouds * len(walls) / sum(walls)
    return dict(value=thr, unit="clouds/s", cores=arm.workers, kind=arm.kind,
                sample=f"{len(walls)} passes; " + arm.describe(statistics.mean(pers)),
                ms_per_cloud_per_core=statistics.mean(pers) * 1e3)


# ---- clocks sampler ------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev, self.rows, self.p = dev, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.t.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---- GPU arm ---------------------------------------------------------------------------------------------------
def algorithmic_work(algo, wl, sample=8):
    """per-cloud algorithmic work: point-updates and bucket tests of the REFERENCE
    algorithm (oracle counters on a seeded sample of the workload's clouds); vanilla: n*(k-1) exactly."""
    B, n, d, k, h, gen, seed, _ = WORKLOADS[wl]
    if algo == "vanilla":
        return float(n) * (k - 1), 0.0
    from oracle import oracle as O
    pu = bt = 0
    m = min(sample, B)
    for i in range(m):
        _, st = O.kdline(make_cloud(gen, seed + i, n, d), k, h, 0, return_stats=True)
        pu += st["point_updates"]
        bt += st["bucket_tests"]
    return pu / m, bt / m


class Dist:
    """the launcher's plumbing: barrier + max over ranks (torch.distributed, NCCL), nothing on the data path"""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.dev = None

    def init(self):
        import torch
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            import torch.distributed as dist
            import datetime
            dist.init_process_group("nccl", device_id=self.dev, timeout=datetime.timedelta(seconds=300))
            self.dist = dist

    def barrier(self):
        import torch
        torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
            torch.cuda.synchronize()

    def maxr(self, x):
        if not self.dist:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.dist:
            self.dist.barrier()
            self.dist.destroy_process_group()


def pin_to_gpu_cores(D):
    if D.world == 1 or os.environ.get("FPS_BENCH_NO_AFFINITY"):
        return None
    # one process per GPU: run on (and first-touch the pinned buffers from) the cores next to this rank's GPU
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(D.local)
        words = pynvml.nvmlDeviceGetCpuAffinity(hnd, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cores next to GPU {D.local}"
    except Exception as e:   # affinity is an optimisation of the host side only
        return f"unavailable ({type(e).__name__})"
    return None


def h2d_floor(D, nbytes, reps=5):
    """GB/s per GPU of the host->device copy alone from page-locked memory, every rank copying at once (max time over ranks)"""
    import torch
    pin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dev = torch.empty(nbytes, dtype=torch.uint8, device=D.dev)
    dev.copy_(pin, non_blocking=True)
    D.barrier()
    best = 1e9
    for _ in range(reps):
        D.barrier()
        t0 = time.perf_counter()
        dev.copy_(pin, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, D.maxr(time.perf_counter() - t0))
    return nbytes / best / 1e9


def gpu_arm(args, D):
    import torch
    import fpsample_b200 as fps
    from fpsample_b200 import capi

    rank, world, local, dev = D.rank, D.world, D.local, D.dev
    B, n, d, k, h, gen, seed, desc = WORKLOADS[args.workload]
    algo = args.algo

    # this rank's shard: B clouds, globally distinct seeds (weak scaling)
    host = capi.pinned_empty((B, n, d), np.float32)
    for b in range(B):
        host[b] = make_cloud(gen, seed + rank * B + b, n, d)
    pageable = np.array(host)   # what a numpy user holds
    dpts = torch.from_numpy(host).to(dev)
    dout = torch.empty((B, k), dtype=torch.int64, device=dev)
    a = capi.ALGO_VANILLA if algo == "vanilla" else capi.ALGO_KDLINE
    wsb = capi.workspace_bytes(a, B, n, d, k, h)
    ws = torch.empty(wsb + 512, dtype=torch.uint8, device=dev)
    wp = (ws.data_ptr() + 255) & ~255
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.Stream(device=dev)

    def launch():
        if algo == "vanilla":
            capi.vanilla_batch_dev(dpts.data_ptr(), B, n, d, k, 0, dout.data_ptr(), wp, wsb, stream.cuda_stream)
        else:
            capi.kdline_batch_dev(dpts.data_ptr(), B, n, d, k, 0, h, dout.data_ptr(), wp, wsb, stream.cuda_stream)

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            launch()
    D.barrier()
    plan = capi.last_plan()

    clocks = Clocks(local)
    clocks.start()
    l0 = capi.kernel_launches()
    evs, phases = [], []
    D.barrier()
    capi.phase_timing(True)      # CUDA events on OUR stream around the build and the sampling launch of every step
    with torch.cuda.stream(stream):
        for _ in range(args.steps):
            flush.fill_(1)                                  # L2 flush, outside the event pair
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            launch()
            e1.record(stream)
            evs.append((e0, e1))
            phases.append(capi.last_phase_ms())             # waits for this step; the next step starts cold again
    capi.phase_timing(False)
    D.barrier()
    launches = capi.kernel_launches() - l0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    dev_ms = D.maxr(sum(step_ms))                             # K steps, max over ranks
    ms_per_step = dev_ms / args.steps
    value = world * B * args.steps / (dev_ms * 1e-3)
    dev_idx = dout.cpu().numpy().astype(np.uint64)

    # ---- e2e: public API, host buffers, H2D + D2H inside the timed region ------------------------------------
    def e2e_leg(src):
        api = (lambda: fps.fps_sampling_batch(src, k, 0, devices=[local])) if algo == "vanilla" else \
              (lambda: fps.bucket_fps_kdline_sampling_batch(src, k, h, 0, devices=[local]))
        for _ in range(max(1, min(args.warmup, 3))):
            idx = api()
        D.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            idx = api()
        torch.cuda.synchronize()
        s = D.maxr(time.perf_counter() - t0)
        D.barrier()
        return s, idx, capi.last_plan()

    e2e_s, e2e_idx, e2e_plan = e2e_leg(host)
    pg_s, pg_idx, pg_plan = e2e_leg(pageable)
    clk = clocks.stop()
    e2e_value = world * B * args.steps / e2e_s
    same = bool(np.array_equal(e2e_idx, dev_idx)) and bool(np.array_equal(pg_idx, dev_idx))
    floor_gbs = h2d_floor(D, B * n * d * 4)

    # ---- sanity: a few clouds of this rank against the oracle (every rank checks its own shard) ----------------------
    from oracle import oracle as O
    checked = 0
    for b in sorted({0, B // 2, B - 1}):
        pc = host[b]
        if algo == "kdline":
            ok = np.array_equal(dev_idx[b], O.kdline(pc, k, h, 0))
        else:
            ok = O.certify_vanilla(pc, dev_idx[b])[0] if n * k > 2e8 else np.array_equal(dev_idx[b], O.fps_vanilla(pc, k, 0))
        if not ok:
            raise SystemExit(f"bench.py: rank {rank} cloud {b} differs from the oracle -- the number would be invalid")
        checked += 1

    # ---- roofline ----------------------------------------------------------------------------------------------
    if rank != 0:
        return None
    hbm_gbs, sm_mhz, how = peaks()
    fp32_peak = SM_COUNT * LANES * sm_mhz * 1e6 / 1e12        # T lane-op/s, no FMA allowed on this path
    pu, bt = algorithmic_work(algo, args.workload)
    ops_cloud = pu * (3 * d + 1) + bt * (8 * d)
    build_ms = statistics.mean(p[0] for p in phases)
    sample_ms = statistics.mean(p[1] for p in phases)
    t_launch = sample_ms * 1e-3                                # the dominant kernel's own duration (CUDA events)
    names = [x.split("(")[0].split("<")[0].strip() for x in plan.split(" + ")]
    kernel = next((x for x in names if "kdline_" in x and x != "kdline_kernel"), names[0]).split(" ")[0]
    fp32_ach = ops_cloud * B / t_launch / 1e12
    bytes_cloud = n * d * 4 + k * 8
    bf_ops = float(n) * (k - 1) * (3 * d + 1) * B
    traffic = None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tj):
        traffic = json.load(open(tj)).get(f"{args.workload}:{algo}", {}).get("dram_bytes_per_launch")
    in_bytes = B * n * d * 4
    return {
        "metric": "clouds/sec (BxN->K)", "value": value, "unit": "clouds/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(args.workload, algo),
        "run": {"plan": plan, "parallelism": f"dp{world} (independent clouds, one contiguous shard of {B} clouds per GPU, no data-path collective)",
                "host_affinity": args.numa, "e2e_plan": e2e_plan, "e2e_pageable_plan": pg_plan},
        "e2e": {"value": e2e_value, "unit": "clouds/s", "h2d_bytes_per_step": in_bytes,
                "d2h_bytes_per_step": B * k * 8, "ms_per_step": e2e_s / args.steps * 1e3,
                "api": "fpsample_b200.%s(page-locked host ndarray) -> host ndarray" % ("fps_sampling_batch" if algo == "vanilla" else "bucket_fps_kdline_sampling_batch"),
                "matches_device_path": same},
        "e2e_pageable": {"value": world * B * args.steps / pg_s, "unit": "clouds/s", "ms_per_step": pg_s / args.steps * 1e3,
                         "api": "the same call on a plain (pageable) numpy array, as a drop-in user of the reference would make it"},
        "h2d_floor": {"gb_per_s_per_gpu": floor_gbs, "ms_per_step": in_bytes / floor_gbs / 1e6, "ranks_copying_at_once": world,
                      "e2e_over_floor": (in_bytes / floor_gbs / 1e6) / (e2e_s / args.steps * 1e3),
                      "note": "page-locked host->device copy of one step's input alone (no kernels), all ranks at once, max over ranks; "
                              "e2e_over_floor = that time / the e2e step"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "phases_ms": {"kd_build": build_ms, "sampling": sample_ms, "step": statistics.mean(step_ms),
                      "how": "CUDA events recorded by the library on the launching stream around its own launches (fps_b200_phase_timing)"},
        "roofline": {"bound": "fp32", "achieved": fp32_ach, "peak": fp32_peak, "unit": "Tlaneop/s",
                     "frac": fp32_ach / fp32_peak, "traffic": traffic,
                     "kernel": kernel, "kernel_ms": sample_ms,
                     "note": f"governing roofline per SURVEY.md 8(d): FP32 pipe without FMA = 148 SM x 128 lanes x {sm_mhz:.0f} MHz ({how}); "
                             f"algorithmic work = reference algorithm's {pu:.0f} point-updates x {3 * d + 1} + {bt:.0f} bucket tests x {8 * d} lane-ops per cloud",
                     "brute_force_equiv_frac": bf_ops / t_launch / 1e12 / fp32_peak},
        "roofline_hbm": {"bound": "hbm", "achieved": bytes_cloud * B / t_launch / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                         "frac": bytes_cloud * B / t_launch / 1e9 / hbm_gbs, "traffic": traffic, "kernel": kernel,
                         "note": f"compulsory bytes only: {bytes_cloud:.0f} B per cloud (coords in, uint64 indices out); clouds stay on chip for all k rounds; peak {how}"},
        "parity_checked_clouds": checked,
    }


# ---- BASELINE.json configs[4]: one batch of 4096 big clouds, sharded over the GPUs of the job (strong scaling) ------------
def cfg5(wl, D, steps=2):
    import torch
    from fpsample_b200 import capi, dist as FD
    from oracle import oracle as O
    Btot, n, d, k, h, gen, seed, desc = WORKLOADS[wl]
    b0, B = FD.shard_range(Btot, D.world, D.rank)
    host = capi.pinned_empty((B, n, d), np.float32)
    for b in range(B):
        host[b] = make_cloud(gen, seed + b0 + b, n, d)
    dpts = torch.from_numpy(host).to(D.dev)
    dout = torch.empty((B, k), dtype=torch.int64, device=D.dev)
    wsb = capi.workspace_bytes(capi.ALGO_KDLINE, B, n, d, k, h)
    ws = torch.empty(wsb + 512, dtype=torch.uint8, device=D.dev)
    wp = (ws.data_ptr() + 255) & ~255
    st = torch.cuda.current_stream().cuda_stream
    run = lambda: capi.kdline_batch_dev(dpts.data_ptr(), B, n, d, k, 0, h, dout.data_ptr(), wp, wsb, st)
    run()                                                      # warm-up (inputs are 1.2 - 9.8 GB: far beyond L2, no flush needed)
    D.barrier()
    plan = capi.last_plan()
    capi.phase_timing(True)
    ev, ph = [], []
    l0 = capi.kernel_launches()
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        ev.append((e0, e1))
        ph.append(capi.last_phase_ms())
    capi.phase_timing(False)
    launches = capi.kernel_launches() - l0
    D.barrier()
    dev_ms = D.maxr(sum(a.elapsed_time(b) for a, b in ev)) / steps
    build_ms, sample_ms = D.maxr(statistics.mean(p[0] for p in ph)), D.maxr(statistics.mean(p[1] for p in ph))
    # executed work of the sampler (one extra, untimed launch with the kernel's counters on)
    capi.set_tuning("COUNT", 1)
    run()
    cnt = capi.debug_counters(capi.DBG_STREAM)
    capi.set_tuning("COUNT", -1)
    dev_idx = dout.cpu().numpy().astype(np.uint64)
    del ws, dpts, dout
    torch.cuda.empty_cache()
    # e2e: host shard in, every index of the job on rank 0 (H2D + sampling + NCCL gather + one D2H inside the timed region)
    if capi.comm_ranks() != D.world:
        FD.init_comm(D.rank, D.world)
    api = lambda: FD.bucket_fps_kdline_sampling_sharded(host, Btot, k, h, 0)
    allidx = api()
    allidx = api()   # (two warm-ups: the result of the previous call is still alive when the next one starts, so rank 0's pool
    #                   of page-locked result buffers needs two of them before it stops allocating)
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        allidx = api()
    e2e_s = D.maxr(time.perf_counter() - t0) / steps
    D.barrier()
    ok = True
    checked = 0
    if D.rank == 0:
        ok = bool(np.array_equal(allidx[:B], dev_idx))
        for b in sorted({0, Btot // 2 + 1, Btot - 1}):         # first shard, a middle shard, the last shard
            pc = make_cloud(gen, seed + b, n, d)
            if not np.array_equal(allidx[b], O.kdline(pc, k, h, 0)):
                raise SystemExit(f"bench.py: {wl} cloud {b} differs from the oracle -- the number would be invalid")
            checked += 1
    if D.rank != 0:
        return None
    hbm_gbs, sm_mhz, how = peaks()
    pts, pu, passes, early, tests, picks, clouds, stored = [int(x) for x in cnt[:8]]
    byts = pts * 4 * (d + 1) + stored * 4                      # executed bytes: D coordinates + distance read per point, 4 per distance written back
    traffic = None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tj):
        traffic = json.load(open(tj)).get(f"{wl}:kdline:{B}", {}).get("dram_bytes_per_launch")
    fp32_peak = SM_COUNT * LANES * sm_mhz * 1e6 / 1e12
    return {"what": desc + f"; one batch of {Btot} clouds cut into {D.world} contiguous shard(s) of {B}",
            "scaling": "strong", "n_gpus": D.world, "clouds_total": Btot, "clouds_per_gpu": B, "n": n, "d": d, "k": k, "h": h,
            "value": Btot / (dev_ms * 1e-3), "unit": "clouds/s", "ms_per_step": dev_ms, "steps": steps,
            "phases_ms": {"kd_build": build_ms, "sampling": sample_ms},
            "e2e": {"value": Btot / e2e_s, "unit": "clouds/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": Btot * n * d * 4,
                    "d2h_bytes_per_step": Btot * k * 8, "nccl_gather_bytes": (Btot - B) * k * 4,
                    "api": "fpsample_b200.dist.bucket_fps_kdline_sampling_sharded(host shard) -> all indices on rank 0 "
                           "(fps_b200_kdline_batch_sharded: H2D, sampling, grouped ncclSend/ncclRecv of uint32 indices, one D2H)",
                    "nccl_version": capi.nccl_version(), "first_shard_matches_device_path": ok},
            "roofline": {"bound": "hbm", "achieved": byts / (sample_ms * 1e-3) / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                         "frac": byts / (sample_ms * 1e-3) / 1e9 / hbm_gbs, "traffic": traffic, "kernel": "kdline_stream_kernel",
                         "kernel_ms": sample_ms,
                         "executed_per_pick": {"points_scanned": pts / max(picks, 1), "point_updates": pu / max(picks, 1),
                                               "distances_stored": stored / max(picks, 1),
                                               "bucket_passes": passes / max(picks, 1), "early_passes": early / max(picks, 1),
                                               "bucket_tests": tests / max(picks, 1)},
                         "fp32_frac": (pu * (3 * d + 1) + tests * 8 * d) / (sample_ms * 1e-3) / 1e12 / fp32_peak,
                         "note": f"W_exec from the kernel's own counters on this rank's shard (SURVEY.md 8(d)): every point of a bucket pass "
                                 f"reads 4(D+1) = {4 * (d + 1)} bytes (D coordinates + its distance) and a distance is written back only if it "
                                 f"changed (counted); the reference algorithm's 4(D+2) per point-update would be {pu * 4 * (d + 2) / 1e9:.0f} GB. "
                                 f"`traffic` = ncu dram bytes of the same launch shape (profiles/traffic.json); peak {how}"},
            "plan": plan, "gpu_launches": int(launches), "parity_checked_clouds": checked}


def latency_floor(capi, plan, ms, k, clouds):
    """roofline with bound = "latency" for the pick-latency-bound configurations: the cost of the sampler's own
    synchronisation structure with no work in it (csrc/floors.cu) against what a round / pick takes"""
    try:
        if "vanilla_cluster_kernel" in plan and clouds == 1:
            C = int(plan.split("cluster=")[1].split()[0])
            ns = capi.sync_floor(capi.FLOOR_CLUSTER, ctas=C)
            return {"bound": "latency", "unit": "ns/pick", "peak": ns, "achieved": ms * 1e6 / (k - 1), "frac": ns * (k - 1) / (ms * 1e6),
                    "structure": f"empty round of vanilla_cluster_kernel's exchange, cluster of {C} CTAs x 512 threads"}
        if "kdline_grid_kernel" in plan:
            G = int(plan.split("grid=")[1].split()[0])
            gc = int(plan.split(" CTAs per cloud")[0].split("(")[-1])
            flat = ",flat>" in plan
            words = 32 * 2 if flat else 5           # flat: each of 32 warps publishes 4 keys = two 16-byte words; merged: ten 8-byte words per CTA
            ns = capi.sync_floor(capi.FLOOR_GRID, ctas=G, words=words, group=gc)
            dbg = capi.debug_counters(capi.DBG_GRID)
            rounds, picks = int(dbg[0]), int(dbg[1])
            passes = max(1, (clouds * gc + G - 1) // G)
            return {"bound": "latency", "unit": "ns/round", "peak": ns, "achieved": ms * 1e6 / max(rounds, 1),
                    "frac": ns * rounds / (ms * 1e6), "rounds": rounds, "picks_per_round": picks / max(rounds, 1),
                    "structure": f"empty round of kdline_grid_kernel's exchange: groups of {gc} CTAs x 1024 threads, {words} stamped 16-byte words "
                                 f"per CTA through L2, {G} CTAs in flight ({passes} pass(es) over the clouds; rounds counted by group 0)"}
        if "kdline_warp" in plan and clouds == 1:
            ns = capi.sync_floor(capi.FLOOR_WARP)
            return {"bound": "latency", "unit": "ns/pick", "peak": ns, "achieved": ms * 1e6 / (k - 1), "frac": ns * (k - 1) / (ms * 1e6),
                    "structure": "one warp's arg-max collectives per pick (2 redux, ballot, 3 shuffles, 1 shared load), kdline_warp_kernel"}
    except Exception as e:   # a floor never invalidates a timing
        return {"bound": "latency", "error": repr(e)}
    return None


def extras(args):
    """the other BASELINE.json configs next to the headline line, device-resident: ms for the single 1M -> 64K cloud
    (uniform and lidar-like), the PointNet++ batch (cfg 3), the single 4096-point cloud through both entries, the
    vanilla entry on the headline batch; each with the latency floor of its sampler's synchronisation structure."""
    import torch
    from fpsample_b200 import capi
    out = {}
    hbm_gbs, sm_mhz, how = peaks()
    fp32_peak = SM_COUNT * LANES * sm_mhz * 1e6 / 1e12
    for wl, algo, reps in (("cfg4", "kdline", 3), ("cfg4l", "kdline", 3), ("cfg3", "kdline", 5), ("cfg1", "vanilla", 20), ("cfg1", "kdline", 20),
                           ("cfg2", "vanilla", 5)):
        B, n, d, k, h, gen, seed, desc = WORKLOADS[wl]
        pc = np.stack([make_cloud(gen, seed + b, n, d) for b in range(B)])
        dp = torch.from_numpy(pc).cuda()
        do = torch.empty((B, k), dtype=torch.int64, device="cuda")
        a = capi.ALGO_VANILLA if algo == "vanilla" else capi.ALGO_KDLINE
        wsb = capi.workspace_bytes(a, B, n, d, k, h)
        ws = torch.empty(wsb + 512, dtype=torch.uint8, device="cuda")
        wp = (ws.data_ptr() + 255) & ~255
        st = torch.cuda.current_stream()
        fn = (lambda: capi.vanilla_batch_dev(dp.data_ptr(), B, n, d, k, 0, do.data_ptr(), wp, wsb, st.cuda_stream)) if algo == "vanilla" \
            else (lambda: capi.kdline_batch_dev(dp.data_ptr(), B, n, d, k, 0, h, do.data_ptr(), wp, wsb, st.cuda_stream))
        fn()
        torch.cuda.synchronize()
        ts, ph = [], []
        capi.phase_timing(True)
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            ph.append(capi.last_phase_ms())
        capi.phase_timing(False)
        plan = capi.last_plan()
        best = min(range(reps), key=lambda i: ts[i])
        ent = {"ms": ts[best], "ms_mean": statistics.mean(ts), "ns_per_pick": ts[best] * 1e6 / max(k - 1, 1),
               "phases_ms": {"kd_build": ph[best][0], "sampling": ph[best][1]},
               "clouds": B, "clouds_per_s": B / (ts[best] * 1e-3), "what": desc, "plan": plan}
        lf = latency_floor(capi, plan, ph[best][1], k, B)
        if lf:
            ent["roofline"] = lf
        if algo == "vanilla":   # brute force: every point-update is executed, the FP32 pipe is the other bound
            ent["roofline_fp32"] = {"bound": "fp32", "unit": "Tlaneop/s", "peak": fp32_peak,
                                    "achieved": float(n) * (k - 1) * (3 * d + 1) * B / (ph[best][1] * 1e-3) / 1e12,
                                    "frac": float(n) * (k - 1) * (3 * d + 1) * B / (ph[best][1] * 1e-3) / 1e12 / fp32_peak}
        out[f"{wl}_{algo}"] = ent
        del ws, dp, do
    return out


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return None
    B, n, d, k, h, gen, seed, desc = WORKLOADS[args.workload]
    arm = CpuArm(args.algo, args.workload, max(2.0, args.cpu_budget / max(args.steps, 1)))
    for _ in range(args.warmup):
        arm.run()
    walls, pers = [], []
    for _ in range(args.steps):                                    # a step = one pass over the bounded sample
        w, done, per = arm.run()
        walls.append(w)
        pers.append(per)
    value = arm.n_clouds * args.steps / sum(walls)
    cb = dict(value=value, unit="clouds/s", cores=arm.workers, kind=arm.kind,
              sample=f"{args.steps} steps; " + arm.describe(statistics.mean(pers)))
    arm.close()
    return {"impl": "reference", "metric": "clouds/sec (BxN->K)", "value": value, "unit": "clouds/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(walls) / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args.workload, args.algo),
            "run": {"note": "the reference's own CPU implementation on this box's host cores (all of them, one cloud per core); "
                            "wall-clock over the pool, inputs pre-generated; no GPU, no library of this repo is loaded"},
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--algo", default="kdline", choices=["kdline", "vanilla"])
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU-baseline work per core")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        line = reference_arm(args)
        if line is not None:
            print(json.dumps(line), flush=True)
        return
    D = Dist()
    cpu = None
    if D.rank == 0 and D.world == 1 and not args.no_cpu:   # before CUDA init: the pool forks
        arm = CpuArm(args.algo, args.workload, args.cpu_budget)
        cpu = cpu_measure(arm)
        arm.close()
    from fpsample_b200 import capi
    if capi.device_count() < 1:
        raise SystemExit("bench.py: no sm_100 device visible; there is no CPU fallback")
    D.init()
    args.numa = pin_to_gpu_cores(D)
    if args.workload.startswith("cfg5"):
        line = None
        ent = cfg5(args.workload, D, steps=max(1, min(args.steps, 3)))
        if ent is not None:
            line = {"metric": "clouds/sec (BxN->K)", "value": ent["value"], "unit": "clouds/s", "n_gpus": D.world, "steps": ent["steps"],
                    "warmup": 1, "ms_per_step": ent["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": "f32", "data": "synthetic", "config": config_of(args.workload, "kdline"), "e2e": ent["e2e"],
                    "roofline": ent["roofline"], "phases_ms": ent["phases_ms"], "run": {"plan": ent["plan"]}, "gpu_launches": ent["gpu_launches"],
                    "cpu_baseline": cpu}
    else:
        line = gpu_arm(args, D)
        if line is not None:
            line["cpu_baseline"] = cpu
        extra = {}
        if not args.no_cfg5:
            for wl in ("cfg5d3", "cfg5d6"):
                try:
                    ent = cfg5(wl, D)
                    if ent is not None:
                        extra[wl] = ent
                except SystemExit:
                    raise
                except Exception as e:   # extras never invalidate the main line
                    extra[wl] = {"error": repr(e)}
                    D.barrier()
        if line is not None and D.world == 1 and not args.no_extras:
            try:
                extra.update(extras(args))
            except Exception as e:
                extra["error"] = repr(e)
        if line is not None:
            line["extra"] = extra
    from fpsample_b200 import capi as _c
    if _c.comm_ranks():
        _c.comm_destroy()
    D.close()
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
