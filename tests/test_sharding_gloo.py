"""CPU, world_size 2 over gloo: the host-side sharding logic of the N > 1 path (scene-batch partition,
row-slab partition, winner all-gather and tie-break merge).  The per-rank scorer is stood in for by the
oracle here (no GPU); the -m gpu tests run the same functions with the CUDA scorer."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_cases as G
import oracle_lib as ol
from social_force_window_planner_b200 import sharding
from social_force_window_planner_b200._abi import BEST_DTYPE


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _best_record(b):
    r = np.zeros((), dtype=BEST_DTYPE)
    r["valid"], r["index"], r["cost"], r["v"], r["w"] = b.valid, b.index, b.cost, b.v, b.w
    return r


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # --- scene-batch sharding: 4 scenes, 2 per rank ---------------------------------------------
        names = ["c0_seed0", "c0_seed1", "c0_hazards_40steps", "c0_seed3"]
        b, e = sharding.block_partition(len(names), world, rank)
        mine = []
        for n in names[b:e]:
            p, sc, lin, ang = G.CASES[n]()
            mine.append(_best_record(ol.oracle_score(p, sc, lin, ang)[1]))
        allb = sharding.all_gather_best(np.array(mine, dtype=BEST_DTYPE)).reshape(-1)
        # --- row-slab sharding of one scene ----------------------------------------------------------
        p, sc, lin, ang = G.CASES["c0_hazards_seed5"]()
        rb, re = sharding.block_partition(len(lin), world, rank)
        costs, _, _ = ol.oracle_score(p, sc, lin[rb:re], ang)
        full = np.full(len(lin) * len(ang), -2.0)
        full[rb * len(ang):re * len(ang)] = costs
        import ctypes as C
        from social_force_window_planner_b200._abi import SfwBest
        sb = SfwBest()
        # arg-min of the slab with GLOBAL indices: rows outside the slab are "skipped"
        lin_m = lin.copy()
        dp = C.POINTER(C.c_double)
        f32 = np.ascontiguousarray(full.astype(np.float32).astype(np.float64))  # the scorer reports float costs
        ol.oracle().sfw_oracle_argmin(f32.ctypes.data_as(dp), lin_m.ctypes.data_as(dp), len(lin),
                                      np.ascontiguousarray(ang).ctypes.data_as(dp), len(ang), C.byref(sb))
        merged = sharding.merge_winners(sharding.all_gather_best(np.array([_best_record(sb)], dtype=BEST_DTYPE)).reshape(-1))
        if rank == 0:
            q.put((allb.tolist(), merged.tolist()))
    finally:
        dist.destroy_process_group()


def test_block_partition():
    for n in (0, 1, 7, 8, 4096):
        for w in (1, 2, 3, 8):
            parts = [sharding.block_partition(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def test_merge_winners_tie_breaks():
    """Reference order (sfw_planner.cpp:394-414): cost up, linvel down, |w| up, later index wins."""
    def rec(valid, idx, cost, v, w):
        r = np.zeros((), dtype=BEST_DTYPE)
        r["valid"], r["index"], r["cost"], r["v"], r["w"] = valid, idx, cost, v, w
        return r
    m = sharding.merge_winners(np.array([rec(1, 3, 5.0, 0.2, 0.1), rec(1, 9, 4.0, 0.1, 0.3)], dtype=BEST_DTYPE))
    assert int(m["index"]) == 9
    m = sharding.merge_winners(np.array([rec(1, 3, 4.0, 0.2, 0.1), rec(1, 9, 4.0, 0.1, 0.0)], dtype=BEST_DTYPE))
    assert int(m["index"]) == 3  # equal cost: higher linvel
    m = sharding.merge_winners(np.array([rec(1, 3, 4.0, 0.2, -0.3), rec(1, 9, 4.0, 0.2, 0.1)], dtype=BEST_DTYPE))
    assert int(m["index"]) == 9  # then lower |w|
    m = sharding.merge_winners(np.array([rec(1, 3, 4.0, 0.2, -0.1), rec(1, 9, 4.0, 0.2, 0.1)], dtype=BEST_DTYPE))
    assert int(m["index"]) == 9  # then the later sample
    m = sharding.merge_winners(np.array([rec(0, 0, 0.0, 0, 0), rec(0, 0, 0.0, 0, 0)], dtype=BEST_DTYPE))
    assert int(m["valid"]) == 0
    m = sharding.merge_winners(np.array([rec(0, 0, 0.0, 0, 0), rec(1, 7, 9.0, 0.1, 0.2)], dtype=BEST_DTYPE))
    assert int(m["valid"]) == 1 and int(m["index"]) == 7


@pytest.mark.timeout(300)
def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    allb, merged = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process truth
    names = ["c0_seed0", "c0_seed1", "c0_hazards_40steps", "c0_seed3"]
    for k, n in enumerate(names):
        p_, sc, lin, ang = G.CASES[n]()
        b = ol.oracle_score(p_, sc, lin, ang)[1]
        assert (allb[k][0], allb[k][1]) == (b.valid, b.index), n
    p_, sc, lin, ang = G.CASES["c0_hazards_seed5"]()
    b = ol.oracle_score(p_, sc, lin, ang)[1]
    assert (merged[0], merged[1], merged[4], merged[5]) == (b.valid, b.index, b.v, b.w)
