"""BASELINE configs[3]: synthetic plate streamed from OME-Zarr, sharded over the ranks of one box.

Each rank writes ITS OWN share of a synthetic plate to a RAM-backed store (/dev/shm by default, so the
number is the pipeline, not a disk), then streams it through `deskew_plate`.  Per-volume shape and the
number of positions/timepoints are arguments: the full 96 x 10 x (600,300,2048) plate is 707 GB and does
not fit; the default is 8 positions x 2 timepoints of (600, 300, 2048) per rank, uncompressed.
"""

import argparse
import json
import os
import shutil
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np
import torch

from shrimpy_b200 import plate, zarr_io
from shrimpy_b200.settings import DeskewSettings


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="600,300,2048")
    ap.add_argument("--positions", type=int, default=8)
    ap.add_argument("--timepoints", type=int, default=2)
    ap.add_argument("--root", default="/dev/shm/shrimpy_plate")
    ap.add_argument("--depth", type=int, default=3)
    ap.add_argument("--io-threads", type=int, default=6)
    ap.add_argument("--write", type=int, default=0, help="1: also write the deskewed plate back to the store")
    ap.add_argument("--zstd", type=int, default=-1)
    ap.add_argument("--blosc", default="", help="cname (zstd | lz4): write the raw plate as blosc(cname, clevel 1, shuffle) "
                    "inside shards, the acquisition's own layout (mantis_engine.py:474-481)")
    ap.add_argument("--inner-z", type=int, default=50, help="z extent of the inner chunks of a shard (with --blosc)")
    ap.add_argument("--camera-like", type=int, default=0, help="1: shot-noise frames (compress ~2.3x) instead of uniform noise")
    ap.add_argument("--out-z-chunk", type=int, default=10)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        from shrimpy_b200.hostmem import bind_to_gpu
        bind_to_gpu(local)
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    Z, Y, X = (int(v) for v in args.shape.split(","))
    root = Path(args.root) / f"rank{rank}"
    shutil.rmtree(root, ignore_errors=True)
    names = [f"{'ABCDEFGH'[i // 12]}/{i % 12 + 1}/fov0" for i in range(args.positions)]
    zc = min(512, Z)
    blosc = {"cname": args.blosc, "clevel": 1, "shuffle": "shuffle"} if args.blosc else None
    inner = (1, 1, min(args.inner_z, zc), Y, X) if blosc else None
    if inner is not None:
        zc = -(-zc // inner[2]) * inner[2]
    src = zarr_io.create_plate(root / "raw.zarr", names, (args.timepoints, 1, Z, Y, X), (1, 1, zc, Y, X),
                               np.uint16, channel_names=["GFP"], zstd_level=None if args.zstd < 0 else args.zstd,
                               blosc=blosc, shard_inner=inner)
    rng = np.random.default_rng(rank)
    if args.camera_like:
        stack = (400 + rng.poisson(120, size=(Z, Y, X))).astype(np.uint16)
    else:
        stack = rng.integers(100, 60000, size=(Z, Y, X), dtype=np.uint16)
    t0 = time.perf_counter()
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max(1, args.io_threads)) as wpool:
        for pos in src:
            for t in range(args.timepoints):
                pos.array.write_stack(t, 0, stack, pool=wpool)
                stack[0, 0, :8] += 1
    gen_s = time.perf_counter() - t0
    settings = DeskewSettings(ls_angle_deg=30.0, pixel_size_um=0.116, px_to_scan_ratio=0.39, keep_overhang=False,
                              average_n_slices=3)
    dst = plate.create_deskewed_plate(root / "deskewed.zarr", src, settings, z_chunk=args.out_z_chunk) if args.write else None
    plate.deskew_plate(src[:1], settings, depth=1)                       # warm-up (pinned allocs, first launch)
    if world > 1:
        dist.barrier()
    stats = plate.deskew_plate(src, settings, dst, depth=args.depth, io_threads=args.io_threads)
    secs = torch.tensor([stats.seconds], dtype=torch.float64, device="cuda")
    vox = torch.tensor([float(stats.out_voxels)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(secs, op=dist.ReduceOp.MAX)
        dist.all_reduce(vox, op=dist.ReduceOp.SUM)
    if rank == 0:
        d = stats.as_dict()
        print(json.dumps({"config": f"plate stream: {args.positions} pos x {args.timepoints} t x ({Z},{Y},{X}) uint16 per rank",
                          "n_gpus": world, "store": str(args.root),
                          "codec": (f"sharding_indexed(inner z {inner[2]}) -> blosc({args.blosc}, clevel 1, shuffle)" if blosc else
                                    ("zstd" if args.zstd >= 0 else "bytes")), "camera_like": bool(args.camera_like),
                          "write_back": bool(args.write),
                          "job_gvoxel_out_per_s": round(float(vox.item()) / float(secs.item()) / 1e9, 2),
                          "rank0": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items()},
                          "plate_write_s": round(gen_s, 2)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
