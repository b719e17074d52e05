"""Whole-sector stores into the CONTIGUOUS result -- ``kernel="tma_staged"`` (results staged through shared memory,
16-byte stores on sector boundaries, csrc/deskew.cu) -- against the plain TMA kernel: first bit equality on awkward
shapes (windows that start inside a sector, ragged X tiles, padded rows, every n, both dtypes), then the config-2 /
n=1 / config-5 timings next to the same launches into padded rows.

    python tools/probe/staged_rows_probe.py > gpurun_out/staged_rows_probe.json
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import torch

import shrimpy_b200 as sb
from shrimpy_b200._cabi import ShrimpyB200Error

res = {"equal": {}, "ms": {}}
VARIANTS = ("tma_staged",)
gen = torch.Generator(device="cuda").manual_seed(7)


def stack(shape, dtype=torch.uint16):
    t = torch.randint(100, 60000, shape, dtype=torch.int32, device="cuda", generator=gen)
    return t.to(dtype)


# ---- equality ---------------------------------------------------------------------------------------------------
for shape, r, keep, n, dtype in [((90, 13, 128), 0.39, False, 1, torch.uint16), ((90, 13, 128), 0.39, True, 2, torch.uint16),
                                 ((120, 31, 200), 0.39, True, 3, torch.uint16), ((120, 31, 200), 0.651, False, 4, torch.uint16),
                                 ((75, 17, 96), 0.39, True, 1, torch.float32), ((300, 40, 264), 1.3, True, 1, torch.uint16),
                                 ((600, 30, 512), 0.39, False, 3, torch.uint16), ((200, 9, 100), 0.77, True, 2, torch.float32),
                                 ((260, 11, 2048), 0.39, True, 1, torch.uint16), ((64, 7, 72), 0.39, False, 1, torch.uint16)]:
    raw = stack(shape, dtype)
    g = sb.deskew_geometry(shape, 30.0, r, keep, n)
    want = sb.deskew_zyx(raw, 30.0, r, keep, n, kernel="tma")
    key = f"{shape}_r{r}_keep{int(keep)}_n{n}_{str(dtype)[6:]}"
    P, X, Xp = g.out_shape
    for variant in VARIANTS:
        got = torch.full_like(want, -7.0)
        try:
            sb.deskew_zyx(raw, 30.0, r, keep, n, out=got, kernel=variant)
        except ShrimpyB200Error as exc:                    # a 256-column tile of n rows does not fit: the selector refuses
            if variant == "tma_staged" and "too large" in str(exc):
                res.setdefault("refused", []).append(f"{key}_{variant}")
                continue
            raise
        res["equal"][f"{key}_{variant}"] = bool(torch.equal(got, want))
        padded = sb.empty_deskewed(g, raw.device)
        storage = padded.as_strided((P, X, padded.stride(1)), padded.stride())
        storage.fill_(-7.0)
        sb.deskew_zyx(raw, 30.0, r, keep, n, out=padded, kernel=variant)
        res["equal"][f"{key}_{variant}_padded_rows"] = bool(torch.equal(padded, want) and bool((storage[:, :, Xp:] == -7.0).all()))
        # windows that start inside a sector, written into a strided view of a canvas that must stay untouched elsewhere
        canvas = torch.full((P, X, Xp), -7.0, device="cuda")
        for c0, c1 in ((3, min(Xp, 3 + 61)), (Xp // 3 + 1, Xp - 2), (5, min(Xp, 5 + 248 + 9))):
            if c1 <= c0:
                continue
            _, zr = sb.window_needs(g, 0, P, c0, c1 - c0)
            z0, z1 = (int(zr[0]), int(zr[1])) if zr[1] > zr[0] else (0, 1)
            sb.deskew_window(raw[z0:z1], g, p_begin=0, p_count=P, c_begin=c0, c_count=c1 - c0, y_origin=0, z_origin=z0,
                             out=canvas[:, :, c0:c1], kernel=variant)
            ok = torch.equal(canvas[:, :, c0:c1], want[:, :, c0:c1])
            canvas[:, :, c0:c1] = -7.0
            res["equal"][f"{key}_{variant}_window{c0}_{c1}"] = bool(ok and bool((canvas == -7.0).all()))
        del got, padded, canvas
    del raw, want

# every tile width the host may choose (the env is read per launch): 1, 2, 4 and 8 warps along o2
import os
raw = stack((120, 31, 200))
for T2 in (32, 64, 128, 256):
    os.environ["SHRIMPY_DESKEW_T2"] = str(T2)
    for n in (1, 3):
        want = sb.deskew_zyx(raw, 30.0, 0.39, True, n, kernel="tma")
        for variant in VARIANTS:
            got = torch.full_like(want, -7.0)
            sb.deskew_zyx(raw, 30.0, 0.39, True, n, out=got, kernel=variant)
            res["equal"][f"T2_{T2}_n{n}_{variant}"] = bool(torch.equal(got, want))
os.environ.pop("SHRIMPY_DESKEW_T2")
del raw

# ---- timings ----------------------------------------------------------------------------------------------------
def time_it(raw, keep, n, kernel, out, reps=10):
    try:
        for _ in range(3):
            sb.deskew_zyx(raw, 30.0, 0.39, keep, n, out=out, kernel=kernel)
    except ShrimpyB200Error as exc:
        return f"refused: {exc}"[:80]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        sb.deskew_zyx(raw, 30.0, 0.39, keep, n, out=out, kernel=kernel)
    b.record()
    torch.cuda.synchronize()
    return round(a.elapsed_time(b) / reps, 4)


if all(res["equal"].values()):
    raw = stack((600, 300, 2048))
    for n, keep in ((1, False), (1, True), (2, False), (3, False), (4, False)):
        g = sb.deskew_geometry((600, 300, 2048), 30.0, 0.39, keep, n)
        out = torch.empty(g.out_shape, dtype=torch.float32, device="cuda")
        for kernel in ("tma",) + VARIANTS:
            res["ms"][f"n{n}_keep{int(keep)}_{kernel}_contiguous"] = time_it(raw, keep, n, kernel, out)
        del out
        for kernel in ("tma", "tma_staged"):
            res["ms"][f"n{n}_keep{int(keep)}_{kernel}_padded_rows"] = time_it(raw, keep, n, kernel, sb.empty_deskewed(g, "cuda"))
    del raw
    raw = stack((4000, 300, 2048))                           # config 5 on one GPU: 25.9 GB of output
    g = sb.deskew_geometry((4000, 300, 2048), 30.0, 0.39, True, 1)
    out = torch.empty(g.out_shape, dtype=torch.float32, device="cuda")
    for kernel in ("tma",) + VARIANTS:
        res["ms"][f"config5_{kernel}_contiguous"] = time_it(raw, True, 1, kernel, out, reps=5)
    del out
    for kernel in ("tma", "tma_staged"):
        res["ms"][f"config5_{kernel}_padded_rows"] = time_it(raw, True, 1, kernel, sb.empty_deskewed(g, "cuda"), reps=5)
    # one rank's share of a 2-GPU scan split: half the columns from the whole stack (no halo traffic at all)
    half = torch.empty((g.out_shape[0], g.out_shape[1], 5248), dtype=torch.float32, device="cuda")
    for kernel in ("tma", "tma_staged"):
        def win():
            sb.deskew_window(raw, g, p_begin=0, p_count=g.out_shape[0], c_begin=0, c_count=5248, y_origin=0, z_origin=0,
                             out=half, kernel=kernel)
        for _ in range(2):
            win()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            win()
        b.record()
        torch.cuda.synchronize()
        res["ms"][f"config5_first_5248_columns_{kernel}"] = round(a.elapsed_time(b) / 5, 4)
print(json.dumps(res, indent=1))
sys.exit(0 if all(res["equal"].values()) else 1)
