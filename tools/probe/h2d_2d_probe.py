"""Which stage bounds the numpy-in/numpy-out deskew (shrimpy_deskew_host)?  Replays its copy pattern for one config-2
channel without the kernel: the H2D side is Z strided pieces per slab (cudaMemcpy2DAsync out of the (Z, Y, X) stack), the
D2H side one linear copy per slab.  Prints the time of each side alone and of both together, for several slab heights."""
import json
import time

import torch
from cuda.bindings import runtime as rt

Z, Y, X, NAVG, XP = 600, 300, 2048, 3, 1279
YN = Y // NAVG
h_raw = torch.empty((Z, Y, X), dtype=torch.uint16).pin_memory()
h_out = torch.empty((YN, X, XP), dtype=torch.float32).pin_memory()
d_raw = torch.empty(h_raw.numel(), dtype=torch.uint16, device="cuda")
d_out = torch.empty(h_out.numel(), dtype=torch.float32, device="cuda")
s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
H2D, D2H = rt.cudaMemcpyKind.cudaMemcpyHostToDevice, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost


def check(res):
    err = res[0] if isinstance(res, tuple) else res
    if int(err) != 0:
        raise RuntimeError(f"CUDA error {err}")


def run(ps, up, down, reps=4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for p0 in range(0, YN, ps):
            pc = min(ps, YN - p0)
            rows = pc * NAVG
            y0 = Y - (p0 + pc) * NAVG                      # tilt rows are flipped
            if up:
                width = rows * X * 2
                check(rt.cudaMemcpy2DAsync(d_raw.data_ptr(), width, h_raw.data_ptr() + y0 * X * 2, Y * X * 2, width, Z,
                                           H2D, s_up.cuda_stream))
            if down:
                nbytes = pc * X * XP * 4
                check(rt.cudaMemcpyAsync(h_out.data_ptr() + p0 * X * XP * 4, d_out.data_ptr(), nbytes, D2H,
                                         s_down.cuda_stream))
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps


def linear(reps=4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        check(rt.cudaMemcpyAsync(d_raw.data_ptr(), h_raw.data_ptr(), h_raw.numel() * 2, H2D, s_up.cuda_stream))
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps


run(9, True, True, 1)
out = {"bytes_up": h_raw.numel() * 2, "bytes_down": h_out.numel() * 4, "h2d_linear_ms": round(linear(), 2)}
for ps in (1, 3, 9, 25, 50, 100):
    out[f"ps{ps}"] = {"piece_kb": ps * NAVG * X * 2 // 1024, "h2d_ms": round(run(ps, True, False), 2),
                      "d2h_ms": round(run(ps, False, True), 2), "both_ms": round(run(ps, True, True), 2)}
print(json.dumps(out))
