#!/usr/bin/env python
"""Contract benchmark: deskewed GVoxel/s on the mantis FOV workload (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over one batch: the two channels of one mantis FOV, each a
uint16 (600, 300, 2048) stack deskewed at 30 deg, px_to_scan_ratio 0.39, keep_overhang=False,
average_n_slices=3 into float32 (100, 2048, 1279).

ours       ``value``  device-resident: inputs already in HBM, K steps timed with CUDA events on the
                      launching stream; one deskew_tma_kernel launch per channel.
           ``e2e``    the same metric through the public numpy API (``shrimpy_b200.deskew_data``:
                      pinned host stack in, pinned host result out, H2D and D2H inside the timed
                      region), with the copy-only ``floor`` of the box at this many ranks beside it and
                      the ``pageable`` / ``online`` forms a reference caller gets.
           ``roofline``      algorithmic bytes / measured launch time vs MEASURED_PEAKS.json, burst (the K
                             steps) and ``sustained`` (>= 1000 launches), also on the bytes that must move
           ``scan_split``    BASELINE.json configs[4]: one oversized volume split along the scan axis over
                             the N ranks, halo over NVLink, bit-equal to the single-GPU window
           ``plate``         BASELINE.json configs[3]: 96 x 10 stacks streamed from an OME-Zarr store in RAM
           ``affine_registration`` (N=1)  BASELINE.json configs[2]
           ``cpu_baseline``  (N=1)  the scipy reference form on the box's host cores, bounded sample
reference  the reference's own CPU form (scipy.ndimage.affine_transform + edge-padded mean, X-chunked
           over all host cores like scripts/measure_psf.py:218-249), same metric and config.

N > 1: one process per GPU (torchrun), every rank deskews its own FOVs (sharding over
(position, time, channel) needs no collective); value = all ranks' voxels / max-over-ranks time.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

RAW_SHAPE = (600, 300, 2048)
CHANNELS = 2
ANGLE, RATIO, KEEP, NAVG = 30.0, 0.39, False, 3
WORKLOAD = ("mantis light-sheet FOV: 2 channels x (600 scan, 300 y, 2048 x) uint16, ls_angle 30 deg, "
            "px_to_scan_ratio 0.39, keep_overhang=False, average_n_slices=3 -> float32 (100, 2048, 1279) per channel")
METRIC = "deskewed_gvoxel_per_s"
UNIT = "GVoxel/s"
# the SAME dict in both arms (the driver compares them); what differs between the arms goes under "notes"
CONFIG = {
    "workload": WORKLOAD,
    "l2": "GPU arm: inputs (2 x 737 MB) and outputs (2 x 1048 MB) of a step exceed the 126 MB L2, no flush needed; "
          "not applicable to the CPU arm",
}


def measured_peak():
    try:
        peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture, or None."""
    try:
        return json.load(open(ROOT / "profiles" / "roofline_traffic.json"))["dram_bytes_per_launch"]
    except Exception:
        return None


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """Polls SM clock and throttle reasons of one GPU through NVML while a region runs."""

    _REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
        0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost",
    }

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            phys = int(ids[index]) if index < len(ids) else index
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _poll(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                for bit, name in self._REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline (scipy on the host cores)
# --------------------------------------------------------------------------------------------------
def scipy_pass(columns_per_worker, workers, reps):
    from oracle import cpu_baseline

    walls, out_vox, in_vox = cpu_baseline.time_scipy_deskew(
        RAW_SHAPE[0], RAW_SHAPE[1], columns_per_worker, workers, (ANGLE, RATIO, KEEP, NAVG), reps=reps)
    return walls, out_vox, in_vox


def cpu_baseline_block(budget_s=12.0):
    """Bounded scipy sample for the default run: one pass of ~budget_s over all host cores."""
    from oracle import cpu_baseline

    cores = os.cpu_count() or 1
    rate = cpu_baseline.calibrate_columns_per_second(RAW_SHAPE[0], RAW_SHAPE[1], (ANGLE, RATIO, KEEP, NAVG))
    cols = int(max(1, min(RAW_SHAPE[2] * CHANNELS // cores, rate * budget_s * 0.6)))
    walls, out_vox, in_vox = scipy_pass(cols, cores, 1)
    return {
        "value": out_vox / walls[0] / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": (f"scipy.ndimage.affine_transform(order=1)+mean on {cores} X-chunks of uint16 "
                   f"({RAW_SHAPE[0]},{RAW_SHAPE[1]},{cols}) (= {cores * cols} of the step's "
                   f"{RAW_SHAPE[2] * CHANNELS} raw columns), one process per core, {walls[0]:.2f} s"),
        "gvoxel_in_per_s": in_vox / walls[0] / 1e9,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import cpu_baseline

    cores = os.cpu_count() or 1
    total_passes = args.steps + args.warmup
    rate = cpu_baseline.calibrate_columns_per_second(RAW_SHAPE[0], RAW_SHAPE[1], (ANGLE, RATIO, KEEP, NAVG))
    # bounded sample: the whole run should end within ~2.5 minutes whatever --steps/--warmup are
    per_pass_s = max(0.05, 120.0 / max(1, total_passes))
    cols = int(max(1, min(RAW_SHAPE[2] * CHANNELS // cores, rate * per_pass_s * 0.6)))
    walls, out_vox, in_vox = scipy_pass(cols, cores, total_passes)
    timed = walls[args.warmup:]
    total = float(sum(timed))
    value = out_vox * len(timed) / total / 1e9
    sample = (f"each step = {cores} X-chunks of uint16 ({RAW_SHAPE[0]},{RAW_SHAPE[1]},{cols}) "
              f"(= {cores * cols} of the workload's {RAW_SHAPE[2] * CHANNELS} raw columns per step), "
              f"scipy.ndimage.affine_transform(order=1)+mean, one process per core")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(CONFIG), "notes": {"sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gvoxel_in_per_s": in_vox * len(timed) / total / 1e9,
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def needed_input_voxels(g):
    """Raw voxels the deskew actually addresses: for every tilt row the scan slices between the taps of its first
    and last inside output column (with keep_overhang=False the head and tail of every row's scan range are never
    read -- 16 % of the mantis stack)."""
    Z, Y, X = g.raw_shape
    Xp = g.out_shape[2]
    o0 = np.arange(Y, dtype=np.float64)
    base = g.shift + o0 * g.m00
    first, last = base + 0.0 * g.m02, base + (Xp - 1) * g.m02
    live = (last >= 0) & (first <= Z - 1)
    lo = np.floor(np.clip(first, 0, Z - 1))
    hi = np.minimum(np.floor(np.clip(last, 0, Z - 1)) + 1, Z - 1)
    return int(np.sum((hi - lo + 1)[live])) * X


def run_ours(args):
    import torch

    import shrimpy_b200 as sb
    from shrimpy_b200 import _cabi
    from tools import bench_blocks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; shrimpy_b200 has no CPU path")
    torch.cuda.set_device(local)
    from shrimpy_b200.hostmem import bind_to_gpu

    numa_bound = bind_to_gpu(local) if world > 1 else False     # pinned buffers on the GPU's own NUMA node
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.gpus != world and rank == 0:
        print(f"[bench] --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    blocks = set(args.blocks.split(","))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    g = sb.deskew_geometry(RAW_SHAPE, ANGLE, RATIO, KEEP, NAVG)
    vox_in, vox_out = g.algorithmic_bytes
    alg_bytes = vox_in * 2 + vox_out * 4
    needed_bytes = needed_input_voxels(g) * 2 + vox_out * 4

    # ---- synthetic inputs resident in HBM (seeded per rank and channel) ----------------------------
    gen = torch.Generator(device="cuda").manual_seed(1 + rank)
    raws = [torch.randint(100, 60000, RAW_SHAPE, dtype=torch.int32, device="cuda", generator=gen).to(torch.uint16)
            for _ in range(CHANNELS)]
    outs = [torch.empty(g.out_shape, dtype=torch.float32, device="cuda") for _ in range(CHANNELS)]

    def step():
        for c in range(CHANNELS):
            sb.deskew_zyx(raws[c], ANGLE, RATIO, KEEP, NAVG, out=outs[c])

    def timed_steps(n):
        barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clocks:
            start.record()
            for _ in range(n):
                step()
            stop.record()
            barrier()
        return max_over_ranks(start.elapsed_time(stop)), clocks.summary()

    for _ in range(max(args.warmup, 3)):
        step()
    launches0 = _cabi.launch_count()
    ms_total, clocks = timed_steps(args.steps)
    launches = _cabi.launch_count() - launches0
    value = world * CHANNELS * vox_out * args.steps / (ms_total * 1e-3) / 1e9
    kernel_ms = ms_total / (args.steps * CHANNELS)      # the region holds only these launches, back to back
    # the same launch sustained: >= 1000 launches back to back (the board reaches its power cap after ~0.1 s)
    sus_steps = max(500, args.steps)
    sus_ms, sus_clocks = timed_steps(sus_steps)
    sus_kernel_ms = sus_ms / (sus_steps * CHANNELS)

    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": dict(CONFIG),
        "notes": {"per_gpu": "every rank deskews its own FOV (2 channels) per step; no data-path collective",
                  "coordinates": "float64", "interpolation": "float32"},
        "gvoxel_in_per_s": world * CHANNELS * vox_in * args.steps / (ms_total * 1e-3) / 1e9,
        "roofline": {
            "bound": "hbm", "kernel": "deskew_tma_kernel<uint16,3>", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": peak_src, "traffic": ncu_traffic(),
            "traffic_source": "static: dram__bytes_read.sum + dram__bytes_write.sum of one launch in the committed ncu "
                              "--set full capture (profiles/roofline_traffic.json), not measured in this run",
            "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": kernel_ms,
            "needed_bytes_per_launch": needed_bytes,
            "frac_needed_bytes": needed_bytes / (kernel_ms * 1e-3) / 1e9 / peak,
            "needed_bytes_note": "input voxels the kernel addresses (keep_overhang=False never reads the head and tail of "
                                 "a tilt row's scan range) x 2 + output voxels x 4",
            "sustained": {"launches": sus_steps * CHANNELS, "launch_ms": sus_kernel_ms,
                          "achieved": alg_bytes / (sus_kernel_ms * 1e-3) / 1e9,
                          "frac": alg_bytes / (sus_kernel_ms * 1e-3) / 1e9 / peak,
                          "frac_needed_bytes": needed_bytes / (sus_kernel_ms * 1e-3) / 1e9 / peak, "clocks": sus_clocks},
        },
        "gpu_launches": int(launches),
        "clocks": clocks,
    }

    # ---- end to end through the public numpy API ---------------------------------------------------
    if "e2e" in blocks:
        step()
        torch.cuda.synchronize()
        line["e2e"] = bench_blocks.e2e_block(dist, rank, world, local, raws, outs, (ANGLE, RATIO, KEEP, NAVG),
                                             max(2, min(args.steps, 8)), numa_bound)
    del raws, outs
    torch.cuda.empty_cache()

    # ---- the two multi-GPU rows of BASELINE.json that are not independent replicas ------------------
    if "scan" in blocks:
        try:
            line["scan_split"] = bench_blocks.scan_split_block(dist, rank, world, local, peak,
                                                               transports=tuple(args.scan_transports.split(",")))
        except Exception as exc:      # noqa: BLE001
            line["scan_split"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        torch.cuda.empty_cache()
    if "plate" in blocks:
        try:
            line["plate"] = bench_blocks.plate_block(dist, rank, world, local, positions=args.plate_positions,
                                                     timepoints=args.plate_timepoints)
        except Exception as exc:      # noqa: BLE001
            line["plate"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        torch.cuda.empty_cache()

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    if world == 1 and "affine" in blocks:
        line["affine_registration"] = bench_blocks.affine_block(peak)
    if world == 1 and "cpu" in blocks and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_block()
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the bounded scipy sample (dev runs)")
    ap.add_argument("--blocks", default="e2e,scan,plate,affine,cpu",
                    help="which blocks to run beside the headline (dev runs): e2e,scan,plate,affine,cpu")
    ap.add_argument("--scan-transports", default="peer,nccl,vmm")
    ap.add_argument("--plate-positions", type=int, default=96)
    ap.add_argument("--plate-timepoints", type=int, default=10)
    args = ap.parse_args()
    if args.steps < 1:
        raise SystemExit("--steps must be >= 1")
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
