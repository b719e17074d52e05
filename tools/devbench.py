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
    args.cases = args.cases.split(",")
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
    if "sweep" in args.cases:
        for n, keep in ((1, False), (1, True), (3, False)):
            for t2 in (64, 128, 256):
                os.environ["SHRIMPY_DESKEW_T2"] = str(t2)
                print("T2", t2, json.dumps(time_deskew((600, 300, 2048), "u16", n, keep, "tma", nbuf=1 if (keep and n == 1) else 2)), flush=True)
        os.environ.pop("SHRIMPY_DESKEW_T2")
    if "affine" in args.cases:
        from shrimpy_b200 import register

        def rot(a, b, c):
            a, b, c = np.deg2rad([a, b, c])
            Rz = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
            Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
            Rx = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
            return Rz @ Ry @ Rx

        shape = (107, 2048, 2048)
        vol = torch.randn(shape, device="cuda")
        Mg = np.eye(4)
        Mg[:3, :3] = rot(2.0, 1.0, 3.0) @ np.diag([1.03, 0.97, 1.1])
        Mg[:3, 3] = [0.4, -1.2, 2.3]
        M90 = np.array([[1.0, 0, 0, 3.5], [0, 0, -1.288, 2040.0], [0, 1.288, 0, -20.0], [0, 0, 0, 1]])
        Mi = np.eye(4)
        M90t = M90.copy()
        M90t[0, 1:3] = [0.02, -0.015]
        M90t[1, 0], M90t[2, 0] = 0.03, -0.02
        for name, M, oshape in (("identity", Mi, shape), ("general", Mg, shape), ("rot90", M90, (100, 2048, 1279)),
                                ("rot90tilt", M90t, (100, 2048, 1279))):
            out = torch.empty(oshape, device="cuda")
            if os.environ.get("DEVBENCH_ONCE"):
                register.affine_transform_zyx(vol, M, oshape, out=out)
                torch.cuda.synchronize()
                continue
            for _ in range(3):
                register.affine_transform_zyx(vol, M, oshape, out=out)
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
            for a, b in ev:
                a.record()
                register.affine_transform_zyx(vol, M, oshape, out=out)
                b.record()
            torch.cuda.synchronize()
            ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
            inside = float((out != 0).float().mean())
            nbytes = vol.numel() * 4 + out.numel() * 4
            print(json.dumps(dict(case=name, out=oshape, ms=round(ms, 4), gbs_alg=round(nbytes / ms / 1e6, 1),
                                  frac=round(nbytes / ms / 1e6 / PEAK, 3), gvox_out=round(out.numel() / ms / 1e6, 1),
                                  inside=round(inside, 3))), flush=True)
    if "affodd" in args.cases:
        from shrimpy_b200 import register
        vol = torch.randn((100, 2048, 1279), device="cuda")
        M90t = np.array([[1.0, 0.02, -0.015, 3.5], [0.03, 0, -1.288, 2040.0], [-0.02, 1.288, 0, -20.0], [0, 0, 0, 1]])
        M90 = np.array([[1.0, 0, 0, 3.5], [0, 0, -1.288, 2040.0], [0, 1.288, 0, -20.0], [0, 0, 0, 1]])
        for name, M in (("odd_x_rot90_inverse", np.linalg.inv(M90)), ("odd_x_rot90tilt_inverse", np.linalg.inv(M90t))):
            out = torch.empty((107, 2048, 2048), device="cuda")
            for pad in (True, False):
                register._PAD_MIN_VOXELS = (1 << 22) if pad else (1 << 62)
                for _ in range(2):
                    register.affine_transform_zyx(vol, M, out.shape, out=out)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5):
                    register.affine_transform_zyx(vol, M, out.shape, out=out)
                b.record(); torch.cuda.synchronize()
                print(json.dumps(dict(case=name, padded_rows=pad, ms=round(a.elapsed_time(b) / 5, 4))), flush=True)
        register._PAD_MIN_VOXELS = 1 << 22
    if "afftile" in args.cases:
        from shrimpy_b200 import register
        shape = (107, 2048, 2048)
        vol = torch.randn(shape, device="cuda")
        out = torch.empty(shape, device="cuda")
        Mi = np.eye(4); Mi[:3, 3] = [0.3, 0.4, 0.5]
        for mode in ("tma", "cpasync"):
            if mode == "cpasync":
                os.environ["SHRIMPY_AFFINE_KERNEL"] = "cpasync"
            for tile in ("8,16,64", "4,16,64", "4,8,64", "8,8,64", "4,8,128", "2,16,128", "8,8,128", "4,16,128", "4,32,32", "8,32,32"):
                os.environ["SHRIMPY_AFFINE_TILE"] = tile
                try:
                    for _ in range(2):
                        register.affine_transform_zyx(vol, Mi, shape, out=out)
                    torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(5):
                        register.affine_transform_zyx(vol, Mi, shape, out=out)
                    b.record(); torch.cuda.synchronize()
                    print(mode, tile, round(a.elapsed_time(b) / 5, 3), "ms", flush=True)
                except Exception as e:
                    print(mode, tile, "failed", str(e)[:80])
        os.environ.pop("SHRIMPY_AFFINE_TILE"); os.environ.pop("SHRIMPY_AFFINE_KERNEL", None)
    if "copy" in args.cases:
        n = 1 << 29
        a = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
        b = torch.empty_like(a)
        for _ in range(3):
            b.copy_(a)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for x, y in ev:
            x.record(); b.copy_(a); y.record()
        torch.cuda.synchronize()
        ms = np.array([x.elapsed_time(y) for x, y in ev])
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(200):
            b.copy_(a)
        s1.record(); torch.cuda.synchronize()
        print(json.dumps({"copy_2GiB_read_plus_2GiB_write": True, "best_gbs": round(2 * n * 4 / ms.min() / 1e6, 1),
                          "median_gbs": round(2 * n * 4 / np.median(ms) / 1e6, 1),
                          "sustained_200_back_to_back_gbs": round(200 * 2 * n * 4 / s0.elapsed_time(s1) / 1e6, 1)}), flush=True)
        # write-only and read-only references
        s0.record()
        for _ in range(50):
            b.fill_(1.0)
        s1.record(); torch.cuda.synchronize()
        print(json.dumps({"fill_write_only_gbs": round(50 * n * 4 / s0.elapsed_time(s1) / 1e6, 1)}), flush=True)
    if "flatfield" in args.cases:
        from shrimpy_b200 import flatfield as ffm
        shape = (600, 300, 2048)
        gen = torch.Generator(device="cuda").manual_seed(3)
        raw = torch.randint(2000, 9000, shape, dtype=torch.int32, device="cuda", generator=gen).to(torch.uint16)
        g = sb.deskew_geometry(shape, 30.0, 0.39, False, 3)
        out = torch.empty(g.out_shape, dtype=torch.float32, device="cuda")

        def timeit(fn, reps=10):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        scale = ffm.flat_field_scale(raw)
        res = {"pattern_ms": timeit(lambda: ffm.flat_field_pattern(raw)),
               "scale_total_ms": timeit(lambda: ffm.flat_field_scale(raw)),
               "standalone_flatfield_ms": timeit(lambda: ffm.flat_field_BF(raw), reps=5),
               "fused_deskew_given_scale_ms": timeit(lambda: ffm.deskew_flat_field_zyx(raw, 30.0, 0.39, False, 3, scale=scale, out=out)),
               "fused_total_ms": timeit(lambda: ffm.deskew_flat_field_zyx(raw, 30.0, 0.39, False, 3, out=out)),
               "plain_deskew_ms": timeit(lambda: sb.deskew_zyx(raw, 30.0, 0.39, False, 3, out=out))}
        ff32 = ffm.flat_field_BF(raw)
        res["two_step_ms"] = res["standalone_flatfield_ms"] + timeit(lambda: sb.deskew_zyx(ff32, 30.0, 0.39, False, 3, out=out))
        print(json.dumps({k: round(v, 4) for k, v in res.items()}), flush=True)
    if "reductions" in args.cases:
        from shrimpy_b200 import reductions as red
        vol = torch.rand((100, 2048, 1279), device="cuda") * 1000
        def t_ms(fn, reps=5):
            fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        nbytes = vol.numel() * 4
        r = {"range_ms": t_ms(lambda: red.value_range(vol)), "percentile_ms": t_ms(lambda: red.percentile(vol, 50.0)),
             "com_ms": t_ms(lambda: red.intensity_center_of_mass(vol, 100.0))}
        r["range_gbs"] = nbytes / r["range_ms"] / 1e6; r["com_gbs"] = nbytes / r["com_ms"] / 1e6
        gen = torch.Generator(device="cuda").manual_seed(1)
        raw = torch.randint(100, 60000, (600, 300, 2048), dtype=torch.int32, device="cuda", generator=gen).to(torch.uint16)
        rng2 = torch.empty(2, dtype=torch.float32, device="cuda")
        out = torch.empty((100, 2048, 1279), dtype=torch.float32, device="cuda")
        r["deskew_plain_ms"] = t_ms(lambda: sb.deskew_zyx(raw, 30.0, 0.39, False, 3, out=out), reps=20)
        r["deskew_with_fused_range_ms"] = t_ms(lambda: sb.deskew_zyx(raw, 30.0, 0.39, False, 3, out=out, value_range=rng2), reps=20)
        r["percentile_given_range_ms"] = t_ms(lambda: red.percentile(out, 50.0, value_range=rng2))
        # the kernels alone (C-ABI calls on preallocated buffers, no host round trip)
        from shrimpy_b200 import _cabi
        lib, st = _cabi.lib(), torch.cuda.current_stream().cuda_stream
        hist = torch.empty(256, dtype=torch.int64, device="cuda")
        sums = torch.empty(4, dtype=torch.float64, device="cuda")
        mip = torch.empty((2048, 1279), dtype=torch.float32, device="cuda")
        lo, hi = red.value_range(out)
        r["hist_kernel_ms"] = t_ms(lambda: lib.shrimpy_hist256_device(out.data_ptr(), out.numel(), lo, hi, hist.data_ptr(), st), reps=20)
        r["com_kernel_ms"] = t_ms(lambda: lib.shrimpy_center_of_mass_device(out.data_ptr(), 100, 2048, 1279, 100.0, sums.data_ptr(), st), reps=20)
        r["zmax_kernel_ms"] = t_ms(lambda: lib.shrimpy_zmax_projection_device(out.data_ptr(), 100, 2048, 1279, 100.0, mip.data_ptr(), st), reps=20)
        # a volume whose voxels crowd into a few bins (background-dominated images): same-address shared atomics
        flat = torch.randn((100, 2048, 1279), device="cuda") * 2.0 + 500.0
        r["hist_kernel_crowded_ms"] = t_ms(lambda: lib.shrimpy_hist256_device(flat.data_ptr(), flat.numel(), 0.0, 1000.0, hist.data_ptr(), st), reps=10)
        flat.fill_(500.0)
        r["hist_kernel_constant_ms"] = t_ms(lambda: lib.shrimpy_hist256_device(flat.data_ptr(), flat.numel(), 0.0, 1000.0, hist.data_ptr(), st), reps=10)
        del flat
        mm = torch.empty(2, dtype=torch.float32, device="cuda")
        r["minmax_kernels_ms"] = t_ms(lambda: lib.shrimpy_minmax_device(out.data_ptr(), out.numel(), mm.data_ptr(), st), reps=20)
        for k in ("hist", "com", "zmax"):
            r[f"{k}_kernel_gbs"] = nbytes / r[f"{k}_kernel_ms"] / 1e6
        print(json.dumps({k: round(v, 3) for k, v in r.items()}), flush=True)
