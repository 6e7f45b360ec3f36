#!/usr/bin/env python
"""bench.py — trajectories scored per second of the DWA + social-force scoring path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C1] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one control tick: the whole (v, w) grid of one planning scene per GPU rolled out, scored and
reduced to the arg-min command (reference src/sfw_planner.cpp:345-417 and everything under it).  At N = 1
the workload is BASELINE.json configs[1] (256x256 samples, 64 steps, 20 pedestrians, 400x400 costmap); at
N > 1 every rank scores its own independent scene of that shape (seed 1000 + rank) and the winners are
all-gathered with NCCL — weak scaling over the scene batch axis, as BASELINE.json's north_star states.

One JSON line on stdout (rank 0).  ``value``: device-resident inputs, CUDA-event time of the K steps on the
launching stream, max over ranks.  ``e2e``: the same tick through ``sfw_score_batch`` with HOST buffers
(pack + H2D + kernel + D2H of cost vector and winner inside the timed region).  ``cpu_baseline`` / ``--impl
reference``: the reference's own sources (oracle/_ref) on the box's host cores on a bounded sub-grid of
the same scene.  ``roofline``: algorithmic bytes / kernel time against the measured HBM
peak — the path is FP32/MUFU-issue bound, so that fraction is tiny by construction; ``issue`` reports the
binding bound (interaction evaluations/s) next to it.
"""
from __future__ import annotations

import argparse
import json
import math
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


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's own sources (oracle/_ref) or the C restatement, on all host cores
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

    def __init__(self, workload_name, scene_index=0):
        import multiprocessing as mp
        import oracle_lib as ol
        self.kind = "reference" if ol.have_ref() else "port"
        if self.kind == "port":
            ol.oracle()  # builds oracle/libsfw_oracle.so if needed
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
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


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation on a bounded sub-grid per step."""
    if rank != 0:
        return
    from social_force_window_planner_b200 import scenes as S
    wl = S.WORKLOADS[args.workload]
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
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(wl, n_gpus, l2):
    return {"workload": f"{wl.name}: {wl.n_v}x{wl.n_w} (v,w) samples, {wl.steps} steps, {wl.n_peds} pedestrians, "
                        f"{wl.map_w}x{wl.map_h} costmap, {wl.n_obstacles} obstacle points, 16-gon footprint",
            "scenes_per_gpu": 1, "scenes_total": n_gpus, "trajectories_per_step": wl.samples * n_gpus,
            "parallelism": f"scene-batch sharding x{n_gpus} + winner exchange" if n_gpus > 1 else "single GPU",
            "seeds": "1000 + rank", "l2": l2}


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
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="N > 1: winners exchanged by the scorer's own epilogue over NVLink peer memory (fused) "
                         "or by an NCCL all-gather after the kernel (nccl)")
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

    # ---- CPU baseline leg first (rank 0, N = 1 only): fork pool before CUDA is initialised ------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(args.workload, 0)
        ri, ci = subgrid(wl, args.cpu_rows, args.cpu_cols)
        lin_s, ang_s = lin[ri], np.ascontiguousarray(ang[ci])
        arm.score(lin_s[:2], ang_s)  # warm the pool
        t0 = time.perf_counter()
        cpu_costs = arm.score(lin_s, ang_s)
        cdt = time.perf_counter() - t0
        arm.close()
        cpu = {"value": len(cpu_costs) / cdt, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
               "sample": f"{len(lin_s)}x{len(ang_s)} strided sub-grid of the {wl.n_v}x{wl.n_w} samples of scene 0, "
                         f"one pass ({cdt:.1f} s)",
               "_costs": cpu_costs, "_ri": ri, "_ci": ci}

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this framework has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from social_force_window_planner_b200.scorer import Scorer
    from social_force_window_planner_b200._abi import BEST_DTYPE

    stream = torch.cuda.Stream(device=dev)
    scorer = Scorer(local_rank, stream.cuda_stream)
    scene = S.make_scene(wl, rank)  # every rank owns one independent scene (seed 1000 + rank)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    n_best = BEST_DTYPE.itemsize

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- winner exchange: fused into the kernel epilogue (peer stores over NVLink) unless --exchange nccl ----
    exchange = "none"
    if world > 1:
        exchange = args.exchange
        if exchange == "fused":
            ok = True
            try:
                handles = [None] * world
                dist.all_gather_object(handles, scorer.exchange_export(1))
                scorer.exchange_connect(rank, world, handles)
            except Exception as e:  # e.g. cudaIpc unavailable in a restricted container
                sys.stderr.write(f"[rank {rank}] fused exchange unavailable ({e}); using NCCL all-gather\n")
                ok = False
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                if ok:
                    raise SystemExit("bench.py: fused exchange connected on some ranks only")
                exchange = "nccl"
        dist.barrier()

    def gather_winners():
        """Every rank learns every scene's winner: the records were already stored into all ranks' gather
        buffers by the scorer's epilogue (fused; only a device-side arrival wait is enqueued here), or an
        NCCL all-gather of the 32-byte records on the scorer's stream."""
        if world == 1:
            return None
        if exchange == "fused":
            scorer.exchange_sync()
            return None
        import ctypes as C
        ptr = scorer._lib.sfw_device_best(scorer._ctx)
        # wrap the device SfwBest[1] as a tensor without copying
        mine = _wrap_device_bytes(torch, ptr, n_best, dev)
        out = torch.empty(world * n_best, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, mine)
        return out

    with torch.cuda.stream(stream):
        scorer.upload(params, [scene], lin, ang)
        scorer.sync()
        # ---- device-resident arm -----------------------------------------------------------------------
        for _ in range(args.warmup):
            scorer.run()
            gather_winners()
        scorer.sync()
        clocks = Clocks(local_rank)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(args.steps)]
        launches0 = scorer.kernel_launches
        barrier()
        clocks.start()
        t_wall0 = time.perf_counter()
        for k in range(args.steps):
            flush_buf.fill_(k & 0xFF)  # L2 flush between timed iterations (outside the event pair)
            ev[k][0].record(stream)
            scorer.run()
            ev[k][1].record(stream)
            gather_winners()
            ev[k][2].record(stream)
        barrier()
        t_wall = time.perf_counter() - t_wall0
        clocks.stop()
        launches = scorer.kernel_launches - launches0
        step_ms = [e[0].elapsed_time(e[2]) for e in ev]
        kern_ms = [e[0].elapsed_time(e[1]) for e in ev]
        total_ms = float(sum(step_ms))
        if world > 1:
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        costs_dev, best_dev = scorer.download()
        kernel_name = scorer.last_kernel
        shared_steps = scorer.shared_prefix_steps
        obst_skip = scorer.obstacle_skip_fraction
        shared_kmax = min(wl.steps, 2.0 * shared_steps)  # estimate of a shared path's length (the longest ramp)
        algo_bytes = scorer.algorithmic_bytes

        # ---- end-to-end arm: host buffers in, host cost vector + winner out, every step ----------------
        from social_force_window_planner_b200._abi import SceneArray
        scene_host = SceneArray([scene])  # the caller's SfwScene structs over its host buffers (built once, like a
        #                                   C++ caller's); every call below still packs + copies them to the device
        for _ in range(3):
            scorer.score(params, scene_host, lin, ang, want_costs=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            costs_e2e, best_e2e = scorer.score(params, scene_host, lin, ang, want_costs=True)
            if world > 1:
                gather_winners()
                torch.cuda.current_stream().synchronize()
        barrier()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        h2d, d2h = scorer.h2d_bytes, scorer.d2h_bytes
        assert np.array_equal(costs_e2e, costs_dev) and best_e2e[0] == best_dev[0], "e2e and resident arms disagree"

    traj_per_step = wl.samples * world
    value = traj_per_step * args.steps / (total_ms * 1e-3)
    e2e_value = traj_per_step * args.steps / e2e_s

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        k_ms = statistics.mean(kern_ms)
        achieved = algo_bytes / (k_ms * 1e-3) / 1e9
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
        P_, M_, S_ = wl.n_peds, wl.n_obstacles, wl.steps
        per_step = 4 * (P_ * (P_ - 1) // 2 + P_) + P_ + 2 * (P_ * (1.0 - obst_skip) + 1) * M_ + 2 * P_
        path_steps = (2 * (wl.n_v + wl.n_w) + 4) * shared_kmax / wl.samples if shared_steps > 0 else 0.0
        mufu_exec = (S_ - shared_steps + path_steps) * per_step + 5 * P_
        mufu_per_s = mufu_exec * wl.samples / (k_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 social forces / f64 rollout+accumulation / u8 costmap",
            "data": "synthetic",
            "config": workload_config(wl, world, "flushed between steps (256 MiB device write outside the event pair)"),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s / args.steps * 1e3, "api": "sfw_score_batch (C ABI, host buffers)"},
            "gpu_launches": int(launches),
            "exchange": {"fused": "winner records stored into every rank's gather buffer over NVLink by the scorer's "
                                  "epilogue; device-side arrival wait (1 tiny kernel per tick)",
                         "nccl": "NCCL all-gather of the 32-byte winner records after the kernel",
                         "none": "single GPU"}[exchange],
            "kernel": kernel_name, "kernel_ms": k_ms, "wall_s_timed_region": t_wall,
            "prefix_sharing": {"mean_shared_steps": shared_steps, "of_steps": wl.steps,
                               "launches_per_tick": int(launches) // max(args.steps, 1),
                               "note": "samples whose velocity ramps are still saturated start from the record of a "
                                       "shared path (bit-identical to the unshared run); kernel_ms covers the path "
                                       "launches and the sample launch"},
            "obstacle_cutoff": {"skip_fraction_at_start": obst_skip,
                                "note": "pedestrian pairs skip obstacle clusters whose every term is below 2^-24 of the "
                                        "force factor (sfw_set_obstacle_cutoff; on-vs-off cost difference about 1e-7 "
                                        "relative, tests/test_gpu_obstacle_cutoff.py)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(algo_bytes),
                         "note": "path is FP32/MUFU-issue bound (about 1e5 flop per algorithmic byte); see issue"},
            "issue": {"reference_interaction_evals_per_s": evals_per_s,
                      "mufu_executed_per_s": mufu_per_s, "mufu_peak_per_s": 148 * 16 * f_sm,
                      "mufu_frac": mufu_per_s / (148 * 16 * f_sm),
                      "model": "reference work (SURVEY.md 8d): S*[N(N-1)+(N-1)] pair + S*N*M obstacle evaluations per "
                               "trajectory (all S steps, as the reference computes them); MUFU count: what the kernels "
                               "execute (each unordered pair once, 4 MUFU; obstacle term 2 MUFU; steps taken from a shared "
                               "path and obstacle clusters skipped by the far-field cutoff are not counted) against 16 MUFU/clk/SM at the sampled SM clock"},
            "clocks": ck,
            "winner": {"valid": int(best_dev[0]["valid"]), "index": int(best_dev[0]["index"]),
                       "v": float(best_dev[0]["v"]), "w": float(best_dev[0]["w"]), "cost": float(best_dev[0]["cost"])},
        }
        if cpu is not None:
            cc = cpu.pop("_costs")
            ri, ci = cpu.pop("_ri"), cpu.pop("_ci")
            g = costs_dev[0].reshape(wl.n_v, wl.n_w)[np.ix_(ri, ci)].reshape(-1).astype(np.float64)
            both = (cc >= 0) & (g >= 0)
            rel = float(np.max(np.abs(g[both] - cc[both]) / np.maximum(np.abs(cc[both]), 1e-12))) if both.any() else 0.0
            line["cpu_baseline"] = cpu
            line["parity_on_sample"] = {"n": int(len(cc)), "max_rel_err": rel,
                                        "validity_equal": bool(np.array_equal(cc >= 0, g >= 0)),
                                        "argmin_equal": bool(int(np.argmin(np.where(cc >= 0, cc, np.inf))) ==
                                                             int(np.argmin(np.where(g >= 0, g, np.inf))))}
        print(json.dumps(line), flush=True)
    scorer.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _wrap_device_bytes(torch, ptr, nbytes, dev):
    """Zero-copy uint8 tensor over device memory owned by the scorer context."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3,
                                  "strides": None}
    return torch.as_tensor(h, device=dev)


if __name__ == "__main__":
    sys.exit(main())
