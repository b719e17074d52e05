"""BASELINE configs[4] alone: one oversized volume (4000, 300, 2048) uint16, keep_overhang=True, n=1, split along
the scan axis over the ranks of one box (torchrun, one process per GPU).  The measurement is ``bench.py``'s
``scan_split`` block (``tools/bench_blocks.py``); prints its JSON on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/scan_split_bench.py \
        [--transports peer,nccl,vmm] [--shape 4000,300,2048] [--reps 5] [--check 1]
"""

import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch

from tools import bench_blocks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="4000,300,2048")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check", type=int, default=1)
    ap.add_argument("--transports", default="peer,nccl,vmm")
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    try:
        peak = float(json.load(open(Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    res = bench_blocks.scan_split_block(dist, rank, world, local, peak, shape=tuple(int(v) for v in args.shape.split(",")),
                                        reps=args.reps, transports=tuple(args.transports.split(",")), check=bool(args.check))
    if rank == 0:
        print(json.dumps(res), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
