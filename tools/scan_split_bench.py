"""BASELINE configs[4]: one oversized volume (4000, 300, 2048) uint16, keep_overhang=True, n=1, split along
the scan axis over the ranks of one box with an NVLink halo exchange (torchrun, one process per GPU).

Every rank regenerates ALL raw slabs from per-slab seeds (so that it can also compute the single-GPU
answer for its own columns and compare bit for bit), but the timed path only touches its own slab plus
the halo it receives.  Prints one JSON line on rank 0.
"""

import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch
import torch.distributed as dist

import shrimpy_b200 as sb
from shrimpy_b200 import sharding


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="4000,300,2048")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check", type=int, default=1)
    ap.add_argument("--transport", choices=["nccl", "peer", "vmm"], default="nccl",
                    help="halo exchange: NCCL send/recv; device copies out of peer-mapped (symmetric) memory; or none at "
                         "all -- the neighbours' pages mapped next to the rank's own (paged_stack.py), one launch")
    args = ap.parse_args()
    shape = tuple(int(v) for v in args.shape.split(","))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    g = sb.deskew_geometry(shape, 30.0, 0.39, True, 1)
    shards = sharding.plan_scan_split(g, world)
    me = shards[rank]

    def slab_of(r):
        z0, z1 = shards[r].own_z
        gen = torch.Generator(device="cuda").manual_seed(1000 + r)
        return torch.randint(100, 60000, (z1 - z0,) + shape[1:], dtype=torch.int32, device="cuda",
                             generator=gen).to(torch.uint16)

    own = slab_of(rank)
    side = None
    stack = None
    if args.transport == "vmm":
        from shrimpy_b200 import paged_stack

        frame_bytes = shape[1] * shape[2] * 2
        page = paged_stack.PagedStack.granularity(local)
        pshards = paged_stack.plan_paged_split(g, world, frame_bytes, page)
        stack = paged_stack.PagedStack(pshards, rank, shape[1:], torch.uint16, local, page)

        def slices_of(z0, z1):   # the same per-rank slabs as the other transports, cut where the pages are
            parts = []
            for r in range(world):
                a, b = shards[r].own_z
                lo, hi = max(a, z0), min(b, z1)
                if hi > lo:
                    parts.append((own if r == rank else slab_of(r))[lo - a:hi - a])
            return torch.cat(parts, dim=0)

        stack.fill_own(slices_of)
        stack.barrier()          # every rank's pages are in place before anyone reads them
    if args.transport == "peer":
        peer = sharding.PeerSlab(shards, rank, shape[1:], torch.uint16, torch.device("cuda", local))
        peer.tensor.copy_(own)
        own = peer
        torch.cuda.synchronize()
        peer.barrier()          # every rank's slices are in place before anyone pulls
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
    def run():
        if stack is not None:
            return paged_stack.deskew_paged_split(stack, g, pshards[rank])
        return sharding.deskew_scan_split(own, g, shards, rank, peer_stream=side)

    # warm-up (also builds NCCL channels)
    piece = run()
    torch.cuda.synchronize()
    dist.barrier()
    times = []
    for _ in range(args.reps):
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        piece = run()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    ok = True
    if args.check:
        del own
        full = torch.cat([slab_of(r) for r in range(world)], dim=0)
        ref = sb.deskew_zyx(full, 30.0, 0.39, True, 1)
        ok = bool(torch.equal(ref[:, :, me.cols[0]:me.cols[1]], piece))
        flag = torch.tensor([int(ok)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
        if rank == 0 and world == 1:
            pass
    halo = (me.halo_below[1] - me.halo_below[0] + me.halo_above[1] - me.halo_above[0]) * shape[1] * shape[2] * 2
    if stack is not None:
        halo = pshards[rank].remote_bytes      # whole pages of the neighbours mapped into this rank's window
    halos = torch.tensor([halo], dtype=torch.float64, device="cuda")
    dist.all_reduce(halos, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = sorted(times)[len(times) // 2]
        vin, vout = g.algorithmic_bytes
        print(json.dumps({"config": f"scan-axis split {shape} uint16 keep_overhang=True n=1", "n_gpus": world,
                          "transport": args.transport, "ms": round(ms, 3), "gvoxel_out_per_s": round(vout / ms / 1e6, 1),
                          "alg_gbs_total": round((vin * 2 + vout * 4) / ms / 1e6, 1),
                          "max_halo_mb_per_rank": round(float(halos.item()) / 1e6, 1),
                          "matches_single_gpu_bitwise": ok, "out_shape": g.out_shape,
                          "interior_cols_frac": round(sum(s.interior_cols[1] - s.interior_cols[0] for s in shards) / g.out_shape[2], 3)}),
              flush=True)
    dist.barrier()
    if stack is not None:
        del piece
        stack.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
