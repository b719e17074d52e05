"""PCIe ceiling of the box: pinned H2D alone, D2H alone and both at once (what bounds the numpy-in/numpy-out path)."""
import json
import time

import torch

n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(2 * n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


run(True, True, 1)
t_h2d, t_d2h, t_both = run(True, False), run(False, True), run(True, True)
print(json.dumps({"h2d_alone_gbs": round(n / t_h2d / 1e9, 1), "d2h_alone_gbs": round(2 * n / t_d2h / 1e9, 1),
                  "both_h2d_gbs": round(n / t_both / 1e9, 1), "both_d2h_gbs": round(2 * n / t_both / 1e9, 1),
                  "note": "both: 1 GiB up and 2 GiB down per repetition, like one FOV channel (0.74 GB up, 1.05 GB down)"}))
