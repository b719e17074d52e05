"""The scan-axis split of one volume on REAL GPUs (BASELINE configs[4]), every transport: NCCL send/recv, peer-mapped
symmetric memory, and neighbour pages mapped with cuMemMap (``paged_stack.py``).  Needs >= 2 GPUs in one box, so it is
skipped on the single-GPU test box; ``bench.py --gpus N`` runs the same block at full size.  Each rank checks its columns
bit for bit against the single-GPU window of the full stack's geometry."""

import json
import os
import socket
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _gpus() -> int:
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, shape, out_path):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist

    from tools import bench_blocks

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        res = bench_blocks.scan_split_block(dist, rank, world, rank, 6459.6, shape=shape, reps=2)
        if rank == 0:
            Path(out_path).write_text(json.dumps(res))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs in one box")
@pytest.mark.parametrize("world", [2, 4])
def test_every_transport_reproduces_the_single_gpu_window(tmp_path, world):
    import torch.multiprocessing as mp

    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "scan.json"
    shape = (1500, 24, 512)          # 24 KiB slices: a 2 MiB page holds 85 of them, every rank gets several pages
    mp.spawn(_worker, args=(world, port, shape, str(out)), nprocs=world, join=True)
    res = json.loads(out.read_text())
    assert set(res["transports"]) == {"peer", "nccl", "vmm"}
    for name, entry in res["transports"].items():
        assert "error" not in entry, (name, entry)
        assert entry["bit_equal_to_single_gpu_window"] is True, name
    assert res["halo_mb_max_per_rank"] > 0
