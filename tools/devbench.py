"""Developer micro-benchmark: device-resident deskew kernel timing (CUDA events, L2-rotating buffers).

Not the contract benchmark (that is bench.py); used to compare kernel variants on a B200.
"""

import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np
import torch

import shrimpy_b200 as sb

PEAK = 6459.6
try:
    PEAK = json.load(open(Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass


def time_deskew(shape, dtype, n, keep, kernel, reps=10, nbuf=2, r=0.39, ang=30.0):
    g = sb.deskew_geometry(shape, ang, r, keep, n)
    gen = torch.Generator(device="cuda").manual_seed(1)
    raws = []
    for _ in range(nbuf):
        t = torch.randint(100, 60000, shape, dtype=torch.int32, device="cuda", generator=gen)
        raws.append(t.to(torch.uint16) if dtype == "u16" else t.to(torch.float32))
        del t
    outs = [torch.empty(g.out_shape, dtype=torch.float32, device="cuda") for _ in range(nbuf)]
    for i in range(3):
        sb.deskew_zyx(raws[i % nbuf], ang, r, keep, n, out=outs[i % nbuf], kernel=kernel)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for i in range(reps):
        ev[i][0].record()
        sb.deskew_zyx(raws[i % nbuf], ang, r, keep, n, out=outs[i % nbuf], kernel=kernel)
        ev[i][1].record()
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    vin, vout = g.algorithmic_bytes
    nbytes = vin * (2 if dtype == "u16" else 4) + vout * 4
    best, med = ms.min(), np.median(ms)
    return dict(shape=shape, dtype=dtype, n=n, keep=keep, kernel=kernel, out=g.out_shape, ms_med=round(float(med), 4),
                ms_min=round(float(best), 4), gbs_med=round(nbytes / med / 1e6, 1), frac=round(nbytes / med / 1e6 / PEAK, 3),
                gvox_out=round(vout / med / 1e6, 1), gvox_in=round(vin / med / 1e6, 1))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="cfg2")
    args = ap.parse_args()
    print("peak", PEAK, torch.cuda.get_device_name(0))
    if "cfg2" in args.cases:
        for dtype in ("u16", "f32"):
            for kernel in ("tma", "direct"):
                print(json.dumps(time_deskew((600, 300, 2048), dtype, 3, False, kernel)), flush=True)
    if "t2" in args.cases:
        for t2 in (32, 64, 128, 256):
            os.environ["SHRIMPY_DESKEW_T2"] = str(t2)
            print("T2", t2, json.dumps(time_deskew((600, 300, 2048), "u16", 3, False, "tma")), flush=True)
        os.environ.pop("SHRIMPY_DESKEW_T2")
    if "n" in args.cases:
        for n in (1, 2, 4):
            for keep in (False, True):
                print(json.dumps(time_deskew((600, 300, 2048), "u16", n, keep, "tma", nbuf=1 if (keep and n == 1) else 2)), flush=True)
    if "cfg1" in args.cases:
        for keep in (False, True):
            print(json.dumps(time_deskew((101, 256, 256), "u16", 1, keep, "tma", reps=50)), flush=True)
