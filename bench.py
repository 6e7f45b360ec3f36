#!/usr/bin/env python
"""bench.py — trajectories scored per second of the DWA + social-force scoring path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C1] [--impl b200|reference] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one control tick: the whole (v, w) grid of one planning scene per GPU rolled out, scored and
reduced to the arg-min command (reference src/sfw_planner.cpp:345-417 and everything under it).

HEADLINE (every N): BASELINE.json configs[1] (256x256 samples, 64 steps, 20 pedestrians, 400x400 costmap), one
scene per GPU, winners exchanged by the scorer's own epilogue — weak scaling over the scene-batch axis.  Every rank
scores the SAME scene content (scene 0), so per-GPU work is identical and the 1 -> N curve shows what the multi-GPU
machinery costs; kernel time depends on scene content by up to ~15 %, which the "heterogeneous" leg (scene
1000 + rank per rank, the round-1 headline) reports next to it with the per-rank kernel times.

  value   device-resident inputs, CUDA-event time of the K steps on the launching stream, max over ranks
  e2e     the same tick through ``sfw_score_batch`` with HOST buffers (pack + H2D + kernel + D2H of the cost vector
          and the winner inside the timed region) — the number to compare with the reference arm

EXTRA LEGS in the same JSON line (``--no-extras`` drops them), each at the size BASELINE.json states:
  c3  configs[3]: 4096 independent scenes (64x64 samples, 32 steps, 10 pedestrians) block-partitioned over the N
      ranks (4096 / N scenes per rank), every rank ends up with all 4096 winners through the fused exchange
      (verified against an NCCL all-gather) — strong scaling; device-resident and e2e
  c4  configs[4]: ONE scene, 1024x1024 samples, linvel rows split into N slabs, slab winners merged on the device
      by the exchange's wait kernel (verified against the full-grid winner) — strong scaling
  c2  configs[2] (N = 1 only): 128x128 samples, 128 steps, 500 pedestrians, full grid, + its CPU baseline
  c0  configs[0] (N = 1 only): the reference's own tick (21x21 samples, 20 steps, 5 pedestrians), device + end to end,
      + the reference on one core over the whole grid

``cpu_baseline`` / ``--impl reference``: the reference's own sources (oracle/_ref, compiled unmodified) on the
box's host cores on a bounded sub-grid of the same scene — all cores, and ONE core ("as shipped": the reference has
no threads).  ``roofline``: algorithmic bytes / kernel time against the measured HBM peak — the path
is FP32/MUFU-issue bound, so that fraction is tiny by construction; ``issue`` reports the binding bound next to it.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "trajectories_scored_per_sec"
UNIT = "traj/s"
DTYPE = "f32 social forces / f64 rollout+accumulation / u8 costmap"

# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's own sources (oracle/_ref) or the C restatement, on the host cores
# --------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(workload_name, scene_index, use_ref):
    from social_force_window_planner_b200 import scenes as S
    wl = S.WORKLOADS[workload_name]
    _CPU["wl"] = wl
    _CPU["scene"] = S.make_scene(wl, scene_index)
    _CPU["params"] = wl.params()
    _CPU["use_ref"] = use_ref


def _cpu_rows(args):
    import oracle_lib as ol
    lin_rows, ang = args
    if _CPU["use_ref"]:
        return ol.ref_score(_CPU["params"], _CPU["scene"], lin_rows, ang, want_best=False)[0]
    return ol.oracle_score(_CPU["params"], _CPU["scene"], lin_rows, ang)[0]


class CpuArm:
    """Process pool (fork, created before CUDA is touched) that scores a sub-grid of one scene."""

    def __init__(self, workload_name, scene_index=0, cores=None):
        import multiprocessing as mp
        import oracle_lib as ol
        self.kind = "reference" if ol.have_ref() else "port"
        if self.kind == "port":
            ol.oracle()  # builds oracle/libsfw_oracle.so if needed
        avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
        self.cores = min(avail, cores) if cores else avail
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_init,
                                                initargs=(workload_name, scene_index, self.kind == "reference"))

    def score(self, lin_rows, ang):
        """costs[len(lin_rows) * len(ang)] (float64), rows farmed over the pool."""
        chunks = [c for c in np.array_split(np.arange(len(lin_rows)), min(len(lin_rows), self.cores * 2))
                  if len(c)]
        parts = self.pool.map(_cpu_rows, [(np.ascontiguousarray(lin_rows[c]), ang) for c in chunks])
        return np.concatenate(parts)

    def close(self):
        self.pool.close()
        self.pool.join()


def subgrid(wl, n_rows, n_cols):
    """Evenly strided sub-grid of the workload's sample arrays (row / column indices)."""
    ri = np.unique(np.linspace(0, wl.n_v - 1, min(n_rows, wl.n_v)).round().astype(int))
    ci = np.unique(np.linspace(0, wl.n_w - 1, min(n_cols, wl.n_w)).round().astype(int))
    return ri, ci


def cpu_rate(workload_name, n_rows, n_cols, cores=None, want_costs=False):
    """traj/s of the CPU arm on a strided sub-grid of scene 0 (one timed pass after a warm-up)."""
    from social_force_window_planner_b200 import scenes as S
    wl = S.WORKLOADS[workload_name]
    lin, ang = wl.sample_arrays()
    arm = CpuArm(workload_name, 0, cores=cores)
    ri, ci = subgrid(wl, n_rows, n_cols)
    lin_s, ang_s = lin[ri], np.ascontiguousarray(ang[ci])
    arm.score(lin_s[:1], ang_s[:min(len(ang_s), 2 * arm.cores)])  # warm the pool
    t0 = time.perf_counter()
    costs = arm.score(lin_s, ang_s)
    dt = time.perf_counter() - t0
    arm.close()
    out = {"value": len(costs) / dt, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
           "sample": f"{len(lin_s)}x{len(ang_s)} strided sub-grid of the {wl.n_v}x{wl.n_w} samples of scene 0, "
                     f"one pass ({dt:.1f} s)"}
    if want_costs:
        out.update(_costs=costs, _ri=ri, _ci=ci)
    return out


# --------------------------------------------------------------------------------------------------
# clocks sampler (NVML): SM clock + throttle reasons during the timed region
# --------------------------------------------------------------------------------------------------
class Clocks:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv:
            self._stop.clear()
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
            self._t = None

    def summary(self):
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None),
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
def work_model(wl):
    """Algorithmic work per trajectory: pair and obstacle force evaluations."""
    n = wl.n_peds + 1
    pair = wl.steps * (n * (n - 1) + (n - 1))
    obst = wl.steps * n * wl.n_obstacles
    return pair, obst


def workload_config(wl, n_gpus, l2):
    return {"workload": f"{wl.name}: {wl.n_v}x{wl.n_w} (v,w) samples, {wl.steps} steps, {wl.n_peds} pedestrians, "
                        f"{wl.map_w}x{wl.map_h} costmap, {wl.n_obstacles} obstacle points, 16-gon footprint",
            "scenes_per_gpu": 1, "scenes_total": n_gpus, "trajectories_per_step": wl.samples * n_gpus,
            "parallelism": f"scene-batch sharding x{n_gpus} + winner exchange" if n_gpus > 1 else "single GPU",
            "seeds": "scene 0 (seed 1000) on every rank: identical per-GPU work (heterogeneous leg: 1000 + rank)",
            "l2": l2}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation on a bounded sub-grid per step."""
    if rank != 0:
        return
    from social_force_window_planner_b200 import scenes as S
    wl = S.WORKLOADS[args.workload]
    one = cpu_rate(args.workload, 2, args.ref_cols, cores=1)
    arm = CpuArm(args.workload, 0)
    lin, ang = wl.sample_arrays()
    ri, ci = subgrid(wl, args.ref_rows, args.ref_cols)
    lin_s, ang_s = lin[ri], np.ascontiguousarray(ang[ci])
    n = len(lin_s) * len(ang_s)
    for _ in range(args.warmup):
        arm.score(lin_s, ang_s)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arm.score(lin_s, ang_s)
    dt = time.perf_counter() - t0
    arm.close()
    value = n * args.steps / dt
    sample = f"{len(lin_s)}x{len(ang_s)} strided sub-grid of the {wl.n_v}x{wl.n_w} samples of scene 0 per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(wl, args.gpus, "n/a (CPU)"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": sample},
        "cpu_baseline_1core": {**one, "note": "as shipped: the reference has no threads (src/sfw_planner.cpp:345-417 is serial)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
class Rig:
    """Per-process plumbing: device, stream, process group, L2 flush buffer, timing helpers."""

    def __init__(self, rank, local_rank, world):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = rank, world
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.local_rank = local_rank
        if world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def quiesce(self):
        """This rank's GPU work done, then every rank's: nobody stores into a peer's gather buffer afterwards."""
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def reduce(self, values, op="max"):
        """Element-wise max / min over ranks of a list of floats."""
        if self.world == 1:
            return list(values)
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.MIN)
        return [float(v) for v in t.cpu()]

    def gather_floats(self, value):
        if self.world == 1:
            return [float(value)]
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self.dev)
        out = self.torch.empty(self.world, dtype=self.torch.float64, device=self.dev)
        self.dist.all_gather_into_tensor(out, t)
        return [float(v) for v in out.cpu()]

    def scorer(self, max_scenes):
        """A scoring context on this rank's stream, connected to its peers' gather buffers when N > 1."""
        from social_force_window_planner_b200.scorer import Scorer
        sc = Scorer(self.local_rank, self.stream.cuda_stream)
        if self.world > 1:
            handles = [None] * self.world
            self.dist.all_gather_object(handles, sc.exchange_export(max_scenes))
            sc.exchange_connect(self.rank, self.world, handles)
            self.dist.barrier()
        return sc

    def nccl_gather_best(self, sc, n_local):
        """NCCL all-gather of this rank's device SfwBest[n_local] -> BEST_DTYPE[world * n_local] (the check the
        fused exchange is verified against)."""
        from social_force_window_planner_b200._abi import BEST_DTYPE
        torch = self.torch
        nb = BEST_DTYPE.itemsize * n_local
        mine = _wrap_device_bytes(torch, sc._lib.sfw_device_best(sc._ctx), nb, self.dev)
        out = torch.empty(self.world * nb, dtype=torch.uint8, device=self.dev)
        with torch.cuda.stream(self.stream):
            self.dist.all_gather_into_tensor(out, mine)
        self.torch.cuda.synchronize(self.dev)
        return out.cpu().numpy().view(BEST_DTYPE)

    def timed_ticks(self, steps, warmup, tick, after=None, clocks=None):
        """W warm-up + K timed ticks on the scorer's stream.  ``tick()`` enqueues the scoring launches, ``after()``
        what follows them inside the tick (exchange wait / merge).  L2 is flushed between ticks, outside the event
        pairs.  Returns (per-tick ms incl. after, per-tick ms of tick() alone, wall seconds)."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                tick()
                if after:
                    after()
            ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
            self.barrier()
            if clocks:
                clocks.start()
            t0 = time.perf_counter()
            for k in range(steps):
                self.flush_buf.fill_(k & 0xFF)
                ev[k][0].record(self.stream)
                tick()
                ev[k][1].record(self.stream)
                if after:
                    after()
                ev[k][2].record(self.stream)
            self.barrier()
            wall = time.perf_counter() - t0
            if clocks:
                clocks.stop()
        return ([e[0].elapsed_time(e[2]) for e in ev], [e[0].elapsed_time(e[1]) for e in ev], wall)


def _wrap_device_bytes(torch, ptr, nbytes, dev):
    """Zero-copy uint8 tensor over device memory owned by the scorer context."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3,
                                  "strides": None}
    return torch.as_tensor(h, device=dev)


def _skew(rig, kern_ms):
    per_rank = rig.gather_floats(statistics.mean(kern_ms))
    return {"min": min(per_rank), "max": max(per_rank), "all": [round(v, 4) for v in per_rank]}


# --------------------------------------------------------------------------------------------------
def leg_one_scene_per_rank(rig, wl, params, lin, ang, steps, warmup, scene_index, clocks=None, e2e=True):
    """One scene of the workload per rank (weak scaling over the scene-batch axis)."""
    from social_force_window_planner_b200 import scenes as S
    from social_force_window_planner_b200._abi import SceneArray
    scene = S.make_scene(wl, scene_index)
    sc = rig.scorer(1)
    fused = rig.world > 1
    with rig.torch.cuda.stream(rig.stream):
        sc.upload(params, [scene], lin, ang)
        sc.sync()
    launches0 = sc.kernel_launches
    step_ms, kern_ms, wall = rig.timed_ticks(steps, warmup, sc.run, sc.exchange_sync if fused else None, clocks)
    launches = (sc.kernel_launches - launches0) * steps // (steps + warmup)
    total_ms = rig.reduce([sum(step_ms)])[0]
    out = {"total_ms": total_ms, "kern_ms": kern_ms, "wall": wall, "launches": int(launches),
           "exchange_ms": statistics.mean(step_ms) - statistics.mean(kern_ms),
           "per_rank_kernel_ms": _skew(rig, kern_ms), "kernel": sc.last_kernel,
           "shared_steps": sc.shared_prefix_steps, "obst_skip": sc.obstacle_skip_fraction,
           "algo_bytes": sc.algorithmic_bytes}
    with rig.torch.cuda.stream(rig.stream):
        costs_dev, best_dev = sc.download()
    out.update(costs=costs_dev, best=best_dev)
    if fused:  # every rank must hold every rank's winner: fused exchange vs an NCCL all-gather of the same records
        got = sc.exchange_fetch().reshape(-1)
        ref = rig.nccl_gather_best(sc, 1)
        ok = bool(np.array_equal(got, ref)) and got[rig.rank] == best_dev[0]
        out["exchange_verified"] = bool(rig.reduce([1.0 if ok else 0.0], "min")[0] == 1.0)
    if e2e:
        scene_host = SceneArray([scene])  # the caller's SfwScene structs over its host buffers (built once, like a
        #                                   C++ caller's); every call below still packs + copies them to the device
        with rig.torch.cuda.stream(rig.stream):
            bufs = None
            for _ in range(3):
                bufs = sc.score(params, scene_host, lin, ang, want_costs=True, out=bufs)
            rig.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):  # the caller's output buffers are reused from tick to tick, like its inputs
                costs_e2e, best_e2e = bufs = sc.score(params, scene_host, lin, ang, want_costs=True, out=bufs)
                if fused:
                    sc.exchange_sync()
                    sc.sync()
            rig.barrier()
            e2e_s = rig.reduce([time.perf_counter() - t0])[0]
        assert np.array_equal(costs_e2e, costs_dev) and best_e2e[0] == best_dev[0], "e2e and resident arms disagree"
        out.update(e2e_s=e2e_s, h2d=sc.h2d_bytes, d2h=sc.d2h_bytes)
    rig.quiesce()
    sc.close()
    return out


def leg_c3(rig, steps, warmup):
    """BASELINE configs[3]: 4096 independent scenes block-partitioned over the ranks."""
    from social_force_window_planner_b200 import scenes as S, sharding
    from social_force_window_planner_b200._abi import SceneArray
    wl = S.WORKLOADS["C3"]
    params = wl.params()
    lin, ang = wl.sample_arrays()
    total = wl.n_scenes
    b, e = sharding.block_partition(total, rig.world, rig.rank)
    scenes = SceneArray(S.make_scenes(wl, e - b, first=b))
    counts = [sharding.block_partition(total, rig.world, r) for r in range(rig.world)]
    counts = [c[1] - c[0] for c in counts]
    sc = rig.scorer(max(counts))
    fused = rig.world > 1
    if fused:
        sc.exchange_expect(counts)
    with rig.torch.cuda.stream(rig.stream):
        sc.upload(params, scenes, lin, ang)
        sc.sync()
    step_ms, kern_ms, _ = rig.timed_ticks(steps, warmup, sc.run, sc.exchange_sync if fused else None)
    total_ms = rig.reduce([sum(step_ms)])[0]
    with rig.torch.cuda.stream(rig.stream):
        costs_dev, best_dev = sc.download()
    verified = None
    if fused:
        got = sc.exchange_fetch().reshape(-1)
        ok = len(got) == total and bool(np.array_equal(got[b:e], best_dev))
        if len(set(counts)) == 1:
            ok = ok and bool(np.array_equal(got, rig.nccl_gather_best(sc, e - b)))
        verified = bool(rig.reduce([1.0 if ok else 0.0], "min")[0] == 1.0)
    with rig.torch.cuda.stream(rig.stream):
        bufs = None
        for _ in range(2):
            bufs = sc.score(params, scenes, lin, ang, want_costs=True, out=bufs)
        rig.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            costs_e2e, best_e2e = bufs = sc.score(params, scenes, lin, ang, want_costs=True, out=bufs)
            if fused:
                sc.exchange_sync()
                sc.sync()
        rig.barrier()
        e2e_s = rig.reduce([time.perf_counter() - t0])[0]
    assert np.array_equal(costs_e2e, costs_dev) and np.array_equal(best_e2e, best_dev)
    traj = total * wl.samples
    h2d, d2h = rig.reduce([sc.h2d_bytes, sc.d2h_bytes])
    out = {"workload": f"C3: {total} independent scenes x {wl.n_v}x{wl.n_w} samples, {wl.steps} steps, {wl.n_peds} "
                       f"pedestrians, {wl.map_w}x{wl.map_h} costmap each (seeds 1000 + scene index)",
           "scaling": "strong", "scenes_total": total, "scenes_per_rank": counts, "steps": steps,
           "value": traj * steps / (total_ms * 1e-3), "unit": UNIT, "ms_per_step": total_ms / steps,
           "per_rank_kernel_ms": _skew(rig, kern_ms),
           "exchange_ms": statistics.mean(step_ms) - statistics.mean(kern_ms), "exchange_verified": verified,
           "kernel": sc.last_kernel, "valid_winners_local": int(best_dev["valid"].sum()),
           "e2e": {"value": traj * steps / e2e_s, "unit": UNIT, "ms_per_step": e2e_s / steps * 1e3,
                   "h2d_bytes_per_step_per_rank": int(h2d), "d2h_bytes_per_step_per_rank": int(d2h),
                   "vs_device_time": (e2e_s / steps * 1e3) / (total_ms / steps),
                   "api": "sfw_score_batch (C ABI, host buffers, cost vectors + winners back)"}}
    rig.quiesce()
    sc.close()
    return out


def leg_c4(rig, steps, warmup):
    """BASELINE configs[4]: one scene, linvel rows split into one slab per rank, winners merged on the device."""
    from social_force_window_planner_b200 import scenes as S, sharding
    wl = S.WORKLOADS["C4"]
    params = wl.params()
    lin, ang = wl.sample_arrays()
    scene = S.make_scene(wl, 0)
    sc = rig.scorer(1)
    sc.set_policy(sc.POLICY_THROUGHPUT)  # a slab keeps the kernel family of its full grid either way; be explicit
    fused = rig.world > 1
    b, e = sharding.block_partition(wl.n_v, rig.world, rig.rank)
    with rig.torch.cuda.stream(rig.stream):
        sc.upload(params, [scene], lin, ang)
        if fused:
            sc.set_row_slab(b, e)
        sc.sync()
    step_ms, kern_ms, _ = rig.timed_ticks(steps, warmup, sc.run,
                                          (lambda: sc.exchange_merge(sync=False)) if fused else None)
    total_ms = rig.reduce([sum(step_ms)])[0]
    kernel = sc.last_kernel
    equal = None
    with rig.torch.cuda.stream(rig.stream):
        if fused:
            merged = sc.exchange_merge()[0]
            sc.set_row_slab(0, wl.n_v)  # the full grid on every rank: what one GPU alone picks
            sc.run()
            _, full = sc.download(want_costs=False)
            ok = merged == full[0]
            equal = bool(rig.reduce([1.0 if ok else 0.0], "min")[0] == 1.0)
            winner = merged
        else:
            _, full = sc.download(want_costs=False)
            winner = full[0]
    out = {"workload": f"C4: one scene, {wl.n_v}x{wl.n_w} samples, {wl.steps} steps, {wl.n_peds} pedestrians, "
                       f"{wl.map_w}x{wl.map_h} costmap", "scaling": "strong", "steps": steps,
           "partition": f"linvel rows in {rig.world} contiguous slabs ({e - b} rows on rank {rig.rank})",
           "value": wl.samples * steps / (total_ms * 1e-3), "unit": UNIT, "ms_per_step": total_ms / steps,
           "per_rank_kernel_ms": _skew(rig, kern_ms),
           "merge_ms": statistics.mean(step_ms) - statistics.mean(kern_ms),
           "merged_winner_equals_full_grid": equal, "kernel": kernel,
           "winner": {"valid": int(winner["valid"]), "index": int(winner["index"]), "v": float(winner["v"]),
                      "w": float(winner["w"])}}
    rig.quiesce()
    sc.close()
    return out


def leg_c2(rig, cpu):
    """BASELINE configs[2] on one GPU: the full 128x128 grid with 500 pedestrians."""
    from social_force_window_planner_b200 import scenes as S
    wl = S.WORKLOADS["C2"]
    params = wl.params()
    lin, ang = wl.sample_arrays()
    sc = rig.scorer(1)
    with rig.torch.cuda.stream(rig.stream):
        sc.upload(params, [S.make_scene(wl, 0)], lin, ang)
        sc.sync()
    step_ms, kern_ms, _ = rig.timed_ticks(3, 1, sc.run)
    with rig.torch.cuda.stream(rig.stream):
        costs, best = sc.download()
    out = {"workload": f"C2: {wl.n_v}x{wl.n_w} samples, {wl.steps} steps, {wl.n_peds} pedestrians, "
                       f"{wl.map_w}x{wl.map_h} costmap (full grid, 3 ticks after 1 warm-up)",
           "value": wl.samples / (statistics.mean(step_ms) * 1e-3), "unit": UNIT,
           "ms_per_step": statistics.mean(step_ms), "kernel": sc.last_kernel,
           "valid_fraction": float((costs[0] >= 0).mean()),
           "parity": "tests/test_gpu_parity.py::test_full_size_c2_rows_vs_oracle (1024 trajectories of this grid)"}
    if cpu is not None:
        cc, ri, ci = cpu.pop("_costs"), cpu.pop("_ri"), cpu.pop("_ci")
        g = costs[0].reshape(wl.n_v, wl.n_w)[np.ix_(ri, ci)].reshape(-1).astype(np.float64)
        both = (cc >= 0) & (g >= 0)
        cpu["max_rel_err_vs_gpu"] = float(np.max(np.abs(g[both] - cc[both]) / np.abs(cc[both]))) if both.any() else 0.0
        cpu["validity_equal"] = bool(np.array_equal(cc >= 0, g >= 0))
        out["cpu_baseline"] = cpu
    rig.quiesce()
    sc.close()
    return out


def leg_c0(rig, cpu):
    """BASELINE configs[0] on one GPU: the reference's own CPU-sized tick (21x21 samples, 20 steps, 5 pedestrians),
    device-resident and end to end through ``sfw_score``, next to the reference on ONE core (how it ships)."""
    from social_force_window_planner_b200 import scenes as S
    from social_force_window_planner_b200._abi import SceneArray
    wl = S.WORKLOADS["C0"]
    params = wl.params()
    lin, ang = wl.sample_arrays()
    scene = S.make_scene(wl, 0)
    sc = rig.scorer(1)
    with rig.torch.cuda.stream(rig.stream):
        sc.upload(params, [scene], lin, ang)
        sc.sync()
    step_ms, kern_ms, _ = rig.timed_ticks(50, 5, sc.run)
    with rig.torch.cuda.stream(rig.stream):
        costs, best = sc.download()
        host = SceneArray([scene])
        bufs = None
        for _ in range(5):
            bufs = sc.score(params, host, lin, ang, want_costs=True, out=bufs)
        ts = []
        for _ in range(200):
            t0 = time.perf_counter()
            c2, b2 = bufs = sc.score(params, host, lin, ang, want_costs=True, out=bufs)
            ts.append(time.perf_counter() - t0)
    assert np.array_equal(c2, costs) and b2[0] == best[0]
    e2e_ms = statistics.median(ts) * 1e3
    out = {"workload": f"C0: {wl.n_v}x{wl.n_w} samples, {wl.steps} steps, {wl.n_peds} pedestrians, "
                       f"{wl.map_w}x{wl.map_h} costmap — one computeVelocityCommands tick",
           "value": wl.samples / (statistics.mean(step_ms) * 1e-3), "unit": UNIT,
           "ms_per_step": statistics.mean(step_ms), "kernel": sc.last_kernel, "block_threads": sc.block_threads,
           "e2e": {"value": wl.samples / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                   "api": "sfw_score (C ABI, host buffers; median of 200 ticks)"}}
    if cpu is not None:
        cc = cpu.pop("_costs")
        cpu.pop("_ri"), cpu.pop("_ci")
        g = costs[0].astype(np.float64)
        both = (cc >= 0) & (g >= 0)
        cpu["ms_per_tick"] = wl.samples / cpu["value"] * 1e3
        cpu["max_rel_err_vs_gpu"] = float(np.max(np.abs(g[both] - cc[both]) / np.abs(cc[both]))) if both.any() else 0.0
        cpu["validity_equal"] = bool(np.array_equal(cc >= 0, g >= 0))
        out["cpu_baseline_1core"] = cpu
    rig.quiesce()
    sc.close()
    return out


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C1")
    ap.add_argument("--ref-rows", type=int, default=16, help="reference arm: sub-grid rows per step")
    ap.add_argument("--ref-cols", type=int, default=32)
    ap.add_argument("--cpu-rows", type=int, default=96, help="cpu_baseline leg: sub-grid rows")
    ap.add_argument("--cpu-cols", type=int, default=96)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline only (no heterogeneous / c3 / c4 / c2 legs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return 0

    from social_force_window_planner_b200 import scenes as S
    wl = S.WORKLOADS[args.workload]
    lin, ang = wl.sample_arrays()
    params = wl.params()
    extras = not args.no_extras

    # ---- CPU baseline legs first (rank 0, N = 1 only): fork pools before CUDA is initialised ----------
    cpu = cpu1 = cpu_c2 = cpu_c0 = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_rate(args.workload, args.cpu_rows, args.cpu_cols, want_costs=True)
        cpu1 = cpu_rate(args.workload, 2, 48, cores=1)
        if extras:
            cpu_c2 = cpu_rate("C2", 4, 8, want_costs=True)
            cpu_c0 = cpu_rate("C0", 21, 21, cores=1, want_costs=True)  # the whole tick, on the one core it ships for

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this framework has no CPU fallback)")
    rig = Rig(rank, local_rank, world)

    # ---- headline: one scene per rank, identical content ---------------------------------------------------
    clocks = Clocks(local_rank)
    hd = leg_one_scene_per_rank(rig, wl, params, lin, ang, args.steps, args.warmup, 0, clocks=clocks)
    extra = {}
    if extras:
        k2 = max(5, min(args.steps, 10))
        if world > 1:
            het = leg_one_scene_per_rank(rig, wl, params, lin, ang, k2, 3, rank, e2e=False)
            extra["heterogeneous"] = {
                "seeds": "1000 + rank (a different scene per rank)", "steps": k2,
                "value": wl.samples * world * k2 / (het["total_ms"] * 1e-3), "unit": UNIT,
                "ms_per_step": het["total_ms"] / k2, "per_rank_kernel_ms": het["per_rank_kernel_ms"],
                "exchange_verified": het.get("exchange_verified"),
                "note": "kernel time depends on scene content (dead trajectories, obstacle-cutoff share); the tick "
                        "ends with the slowest rank's scene"}
        extra["c3"] = leg_c3(rig, k2, 2)
        extra["c4"] = leg_c4(rig, k2, 3)
        if world == 1:
            extra["c2"] = leg_c2(rig, cpu_c2)
            extra["c0"] = leg_c0(rig, cpu_c0)

    traj_per_step = wl.samples * world
    value = traj_per_step * args.steps / (hd["total_ms"] * 1e-3)
    e2e_value = traj_per_step * args.steps / hd["e2e_s"]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        k_ms = statistics.mean(hd["kern_ms"])
        achieved = hd["algo_bytes"] / (k_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        except Exception:
            pass
        pair, obst = work_model(wl)
        ck = clocks.summary()
        f_sm = (ck["sm_mhz"] or 1965.0) * 1e6
        evals_per_s = (pair + obst) * wl.samples / (k_ms * 1e-3)
        # MUFU operations the kernels actually execute per trajectory: every unordered
        # pedestrian pair and every robot-pedestrian pair once per step (2 rsqrt + 2 ex2, + 1 sqrt for the
        # social-work magnitude of the robot pairs), 2 per obstacle term (rsqrt, ex2), 2 rsqrt per
        # pedestrian update (goal direction, speed cap), one extra robot-pedestrian pass after the last step.
        # With rollout prefix sharing a sample does not execute its first `shared` steps (mean over the grid, from
        # the library); the shared paths themselves add 2 (n_v + n_w) + 4 path-prefixes, counted too.
        # Obstacle clusters a pedestrian pair skips (far-field cutoff) are not executed: the library's skip
        # fraction at the pedestrians' start positions stands in for the whole rollout (the robot never skips).
        shared_steps, obst_skip = hd["shared_steps"], hd["obst_skip"]
        shared_kmax = min(wl.steps, 2.0 * shared_steps)
        P_, M_, S_ = wl.n_peds, wl.n_obstacles, wl.steps
        per_step = 4 * (P_ * (P_ - 1) // 2 + P_) + P_ + 2 * (P_ * (1.0 - obst_skip) + 1) * M_ + 2 * P_
        path_steps = (2 * (wl.n_v + wl.n_w) + 4) * shared_kmax / wl.samples if shared_steps > 0 else 0.0
        mufu_exec = (S_ - shared_steps + path_steps) * per_step + 5 * P_
        mufu_per_s = mufu_exec * wl.samples / (k_ms * 1e-3)
        best = hd["best"][0]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": hd["total_ms"] / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "value_is": "device_resident (inputs in HBM when the timed region starts); e2e is the SURVEY 8(d) tick: "
                        "first H2D of the scene to winner + cost vector on the host",
            "config": workload_config(wl, world, "flushed between steps (256 MiB device write outside the event pair)"),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(hd["h2d"]),
                    "d2h_bytes_per_step": int(hd["d2h"]), "ms_per_step": hd["e2e_s"] / args.steps * 1e3,
                    "api": "sfw_score_batch (C ABI, host buffers)"},
            "gpu_launches": hd["launches"],
            "exchange": ("winner records stored into every rank's gather buffer over NVLink by the scorer's epilogue; "
                         "device-side bounded arrival wait (1 tiny kernel per tick); no collective on the data path")
            if world > 1 else "single GPU",
            "exchange_ms": hd["exchange_ms"] if world > 1 else 0.0,
            "exchange_verified": hd.get("exchange_verified"),
            "per_rank_kernel_ms": hd["per_rank_kernel_ms"],
            "kernel": hd["kernel"], "kernel_ms": k_ms, "wall_s_timed_region": hd["wall"],
            "prefix_sharing": {"mean_shared_steps": shared_steps, "of_steps": wl.steps,
                               "launches_per_tick": hd["launches"] // max(args.steps, 1),
                               "note": "samples whose velocity ramps are still saturated start from the record of a "
                                       "shared path (bit-identical to the unshared run); kernel_ms covers the path "
                                       "launches and the sample launch"},
            "obstacle_cutoff": {"skip_fraction_at_start": obst_skip,
                                "note": "pedestrian pairs skip obstacle clusters whose every term is below 2^-24 of the "
                                        "force factor (sfw_set_obstacle_cutoff; on-vs-off cost difference about 1e-7 "
                                        "relative, tests/test_gpu_obstacle_cutoff.py)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(hd["algo_bytes"]),
                         "note": "path is bound by register-file operand bandwidth / MUFU rate (about 1e5 flop per "
                                 "algorithmic byte); see issue"},
            "issue": {"reference_interaction_evals_per_s": evals_per_s,
                      "mufu_executed_per_s": mufu_per_s, "mufu_peak_per_s": 148 * 16 * f_sm,
                      "mufu_frac": mufu_per_s / (148 * 16 * f_sm),
                      "model": "reference work (SURVEY.md 8d): S*[N(N-1)+(N-1)] pair + S*N*M obstacle evaluations per "
                               "trajectory (all S steps, as the reference computes them); MUFU count: what the kernels "
                               "execute (each unordered pair once, 4 MUFU; obstacle term 2 MUFU; steps taken from a shared "
                               "path and obstacle clusters skipped by the far-field cutoff are not counted) against 16 MUFU/clk/SM at the sampled SM clock",
                      "binding": "from the ncu source pages of the same kernels (profiles/r2m_c1_loops.txt, DESIGN.md 4.6; not "
                                 "measured live): the cross-pair loop (42 % of the C1 warp samples) takes 235 cycles per trip "
                                 "and sub-partition against 218 of register-operand delivery (2 x 32-bit operands per lane "
                                 "and cycle, scripts/dbg/mix_bench.cu), the obstacle loop (25 %) runs at the MUFU rate, 10 % "
                                 "of the warp time waits at the final barrier (14 warps sit 4/4/3/3 on the schedulers)"},
            "clocks": ck,
            "winner": {"valid": int(best["valid"]), "index": int(best["index"]),
                       "v": float(best["v"]), "w": float(best["w"]), "cost": float(best["cost"])},
        }
        if cpu is not None:
            cc = cpu.pop("_costs")
            ri, ci = cpu.pop("_ri"), cpu.pop("_ci")
            g = hd["costs"][0].reshape(wl.n_v, wl.n_w)[np.ix_(ri, ci)].reshape(-1).astype(np.float64)
            both = (cc >= 0) & (g >= 0)
            rel = float(np.max(np.abs(g[both] - cc[both]) / np.maximum(np.abs(cc[both]), 1e-12))) if both.any() else 0.0
            line["cpu_baseline"] = cpu
            line["cpu_baseline_1core"] = {**cpu1, "note": "as shipped: the reference has no threads "
                                                          "(src/sfw_planner.cpp:345-417 is a serial double loop)"}
            line["parity_on_sample"] = {"n": int(len(cc)), "max_rel_err": rel,
                                        "validity_equal": bool(np.array_equal(cc >= 0, g >= 0)),
                                        "argmin_equal": bool(int(np.argmin(np.where(cc >= 0, cc, np.inf))) ==
                                                             int(np.argmin(np.where(g >= 0, g, np.inf))))}
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        rig.dist.barrier()
        rig.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
