"""Multi-rank check + timing of the fused winner exchange against NCCL all-gather (run under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/exchange_check.py

Every rank scores its own scenes (seed 1000 + rank * n + k); the records every rank ends up with through the
kernel-epilogue exchange must equal what an NCCL all-gather of the per-rank winners returns, on every rank."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dataclasses
import numpy as np
import torch
import torch.distributed as dist
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scorer import Scorer
from social_force_window_planner_b200._abi import BEST_DTYPE

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 4
wl = dataclasses.replace(S.WORKLOADS["C3"], n_v=32, n_w=32)
scs = [S.make_scene(wl, rank * n_scenes + k) for k in range(n_scenes)]
p = wl.params(); lin, ang = wl.sample_arrays()
st = torch.cuda.Stream(device=dev)
sc = Scorer(lr, st.cuda_stream)
handles = [None] * world
dist.all_gather_object(handles, sc.exchange_export(n_scenes))
sc.exchange_connect(rank, world, handles)
dist.barrier()
nb = BEST_DTYPE.itemsize
with torch.cuda.stream(st):
    sc.upload(p, scs, lin, ang)
    for tick in range(4):
        sc.run()
        sc.exchange_sync()
        fused = sc.exchange_fetch()
        _, mine = sc.download()
        t = torch.from_numpy(mine.view(np.uint8).reshape(-1).copy()).to(dev)
        out = torch.empty(world * n_scenes * nb, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, t)
        ref = out.cpu().numpy().view(BEST_DTYPE).reshape(world, n_scenes)
        assert np.array_equal(fused, ref), (rank, tick)
    # timing: K ticks, device time of (kernel + exchange wait) vs (kernel + NCCL all-gather)
    K = 20
    def timed(fused_path, exchange=True):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        dist.barrier(); torch.cuda.synchronize(dev)
        ev[0].record(st)
        for _ in range(K):
            sc.run()
            if not exchange:
                continue
            if fused_path:
                sc.exchange_sync()
            else:
                ptr = sc._lib.sfw_device_best(sc._ctx)
                class H: pass
                h = H(); h.__cuda_array_interface__ = {"shape": (n_scenes * nb,), "typestr": "|u1", "data": (int(ptr), False), "version": 3, "strides": None}
                dist.all_gather_into_tensor(out, torch.as_tensor(h, device=dev))
        ev[1].record(st)
        torch.cuda.synchronize(dev)
        return ev[0].elapsed_time(ev[1]) / K
    timed(True); timed(False)
    tf, tn, tk = timed(True), timed(False), timed(True, exchange=False)
t = torch.tensor([tf, tn, tk], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"exchange_check ok: world {world}, {n_scenes} scenes/rank; ms per tick (max over ranks): kernel alone {t[2].item():.4f}, kernel with fused epilogue exchange + wait {t[0].item():.4f}, kernel + NCCL all-gather {t[1].item():.4f}")
sc.close()
dist.barrier()
dist.destroy_process_group()
