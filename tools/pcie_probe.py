"""Host<->device copy ceiling of the box at N ranks (what bounds the numpy-in/numpy-out path): pinned H2D alone, D2H
alone and both at once, all ranks at the same time, no kernel.  Run alone or under torchrun:

    python tools/pcie_probe.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py

Buffers have the sizes of one mantis FOV channel pair (1.47 GB up, 2.10 GB down per repetition).  ``--bind 0`` skips
the CPU-affinity binding of ``hostmem.bind_to_gpu`` so that its effect can be seen; ``--stagger-ms`` starts rank r that
many milliseconds x r late.  The same measurement is the ``e2e.floor`` block of every ``bench.py`` line.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--bind", type=int, default=1)
ap.add_argument("--reps", type=int, default=6)
ap.add_argument("--stagger-ms", type=float, default=0.0)
args = ap.parse_args()
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
bound = False
if args.bind and world > 1:
    from shrimpy_b200.hostmem import bind_to_gpu

    bound = bind_to_gpu(local)
dist = None
if world > 1:
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

up, down = 2 * 600 * 300 * 2048 * 2, 2 * 100 * 2048 * 1279 * 4
h_in = torch.empty(up, dtype=torch.uint8).pin_memory()
h_out = torch.empty(down, dtype=torch.uint8).pin_memory()
d_in = torch.empty(up, dtype=torch.uint8, device="cuda")
d_out = torch.empty(down, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h):
    def once():
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()

    once()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    if args.stagger_ms:
        time.sleep(args.stagger_ms * rank / 1e3)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        once()
    t = torch.tensor([(time.perf_counter() - t0) / args.reps], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


t_up, t_down, t_both = run(True, False), run(False, True), run(True, True)
if rank == 0:
    print(json.dumps({
        "n_ranks": world, "numa_bound": bound, "stagger_ms": args.stagger_ms,
        "h2d_alone_gbs_per_rank": round(up / t_up / 1e9, 1), "d2h_alone_gbs_per_rank": round(down / t_down / 1e9, 1),
        "both_ms": round(1e3 * t_both, 2), "both_gbs_per_rank": round((up + down) / t_both / 1e9, 1),
        "both_gbs_all_ranks": round(world * (up + down) / t_both / 1e9, 1),
        "note": "per repetition and rank: 1.47 GB up and 2.10 GB down, like one mantis FOV (2 channels); max over ranks"}))
if dist is not None:
    dist.destroy_process_group()
