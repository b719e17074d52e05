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
                      region).
           ``roofline``      algorithmic bytes / measured launch time vs MEASURED_PEAKS.json
           ``cpu_baseline``  the scipy reference form on the box's host cores, bounded sample
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
        "config": {"workload": WORKLOAD, "sample": sample},
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
def run_ours(args):
    import torch

    import shrimpy_b200 as sb
    from shrimpy_b200 import _cabi

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

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    g = sb.deskew_geometry(RAW_SHAPE, ANGLE, RATIO, KEEP, NAVG)
    vox_in, vox_out = g.algorithmic_bytes
    alg_bytes = vox_in * 2 + vox_out * 4

    # ---- synthetic inputs resident in HBM (seeded per rank and channel) ----------------------------
    gen = torch.Generator(device="cuda").manual_seed(1 + rank)
    raws = [torch.randint(100, 60000, RAW_SHAPE, dtype=torch.int32, device="cuda", generator=gen).to(torch.uint16)
            for _ in range(CHANNELS)]
    outs = [torch.empty(g.out_shape, dtype=torch.float32, device="cuda") for _ in range(CHANNELS)]

    def step():
        for c in range(CHANNELS):
            sb.deskew_zyx(raws[c], ANGLE, RATIO, KEEP, NAVG, out=outs[c])

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _cabi.launch_count()
    with ClockSampler(local) as clocks:
        start.record()
        for _ in range(args.steps):
            step()
        stop.record()
        barrier()
    launches = _cabi.launch_count() - launches0
    ms_total = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = float(ms_total.item())
    value = world * CHANNELS * vox_out * args.steps / (ms_total * 1e-3) / 1e9
    kernel_ms = ms_total / (args.steps * CHANNELS)      # the region holds only these launches, back to back

    # ---- end to end through the public numpy API, pinned host buffers ------------------------------
    e2e_steps = max(1, min(args.steps, 8))
    h_raw = [torch.empty(RAW_SHAPE, dtype=torch.uint16).pin_memory() for _ in range(CHANNELS)]
    h_out = [torch.empty(g.out_shape, dtype=torch.float32).pin_memory() for _ in range(CHANNELS)]
    for c in range(CHANNELS):
        h_raw[c].copy_(raws[c])
    np_raw = [t.numpy() for t in h_raw]
    np_out = [t.numpy() for t in h_out]

    def e2e_step():
        for c in range(CHANNELS):
            sb.deskew_data(np_raw[c], ANGLE, RATIO, KEEP, NAVG, device=f"cuda:{local}", out=np_out[c])

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()          # synchronous: returns when the host result is complete
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    e2e_value = world * CHANNELS * vox_out * e2e_steps / e2e_s / 1e9
    e2e_ok = bool(torch.equal(h_out[0], outs[0].cpu()))

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": WORKLOAD,
            "per_gpu": "every rank deskews its own FOV (2 channels) per step; no data-path collective",
            "l2": "inputs (2 x 737 MB) and outputs (2 x 1048 MB) of a step exceed the 126 MB L2; no flush needed",
            "coordinates": "float64", "interpolation": "float32",
        },
        "gvoxel_in_per_s": world * CHANNELS * vox_in * args.steps / (ms_total * 1e-3) / 1e9,
        "roofline": {
            "bound": "hbm", "kernel": "deskew_tma_kernel<uint16,3>", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": peak_src, "traffic": ncu_traffic(),
            "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": kernel_ms,
        },
        "e2e": {
            "value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": CHANNELS * vox_in * 2,
            "d2h_bytes_per_step": CHANNELS * vox_out * 4, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
            "api": "shrimpy_b200.deskew_data(numpy pinned) -> shrimpy_deskew_host (H2D | kernel | D2H on 3 streams)",
            "matches_device_path": e2e_ok, "numa_bound": numa_bound,
        },
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
    }
    if world == 1:
        del raws, outs, h_raw, h_out
        torch.cuda.empty_cache()
        line["affine_registration"] = affine_block(peak)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_block()
    print(json.dumps(line), flush=True)
    return 0


def affine_block(peak):
    """BASELINE.json configs[2] beside the headline: float32 (107, 2048, 2048) resampled with a 4x4 matrix, device
    resident, CUDA events, median of 8 launches after 3 warm-ups (inputs + outputs of 2.8-3.6 GB exceed the L2)."""
    import torch

    from shrimpy_b200 import register

    shape = (107, 2048, 2048)
    vol = torch.randn(shape, device="cuda")
    a, b, c = np.deg2rad([2.0, 1.0, 3.0])
    Rz = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rx = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    Mg = np.eye(4)
    Mg[:3, :3] = Rz @ Ry @ Rx @ np.diag([1.03, 0.97, 1.1])
    Mg[:3, 3] = [0.4, -1.2, 2.3]
    M90 = np.array([[1.0, 0, 0, 3.5], [0, 0, -1.288, 2040.0], [0, 1.288, 0, -20.0], [0, 0, 0, 1]])
    M90t = M90.copy()
    M90t[0, 1:3] = [0.02, -0.015]
    M90t[1, 0], M90t[2, 0] = 0.03, -0.02
    cases = (("in_plane_identity_like", np.eye(4), shape, "affine_stream_kernel"),
             ("in_plane_rot90_x1.288_onto_deskewed_grid", M90, (100, 2048, 1279), "affine_stream_kernel (lanes along o1)"),
             ("general_rot_2_1_3_deg_aniso_scale", Mg, shape, "affine_tilt_kernel"),
             ("rot90_x1.288_with_tilt_onto_deskewed_grid", M90t, (100, 2048, 1279), "affine_tilt_kernel (lanes along o1)"))
    res = {}
    for name, M, oshape, kern in cases:
        out = torch.empty(oshape, device="cuda")
        for _ in range(3):
            register.affine_transform_zyx(vol, M, oshape, out=out)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
        for e0, e1 in ev:
            e0.record()
            register.affine_transform_zyx(vol, M, oshape, out=out)
            e1.record()
        torch.cuda.synchronize()
        ms = float(np.median([e0.elapsed_time(e1) for e0, e1 in ev]))
        nbytes = (vol.numel() + out.numel()) * 4
        res[name] = {"kernel": kern, "out_shape": list(oshape), "ms": ms, "gvoxel_out_per_s": out.numel() / ms / 1e6,
                     "algorithmic_gbs": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak,
                     "inside_fraction": float((out != 0).float().mean())}
        del out
    return {"workload": "affine registration resample of a float32 label-free volume (107,2048,2048) with a 4x4 matrix "
                        "(BASELINE.json configs[2]); algorithmic bytes = (input + output voxels) x 4", "cases": res}


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the bounded scipy sample (dev runs)")
    args = ap.parse_args()
    if args.steps < 1:
        raise SystemExit("--steps must be >= 1")
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
