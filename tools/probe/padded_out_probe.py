"""Does the row pitch of the output matter?  Config-2 stack, n = 1 and n = 3: the deskewed rows are 1279 floats (5116 B,
never 16-byte aligned from row to row); the same launch into a buffer whose rows are padded to 1280 floats."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import torch

import shrimpy_b200 as sb
from shrimpy_b200 import _cabi

gen = torch.Generator(device="cuda").manual_seed(1)
raw = torch.randint(100, 60000, (600, 300, 2048), dtype=torch.int32, device="cuda", generator=gen).to(torch.uint16)
lib, st = _cabi.lib(), torch.cuda.current_stream().cuda_stream
res = {}
for n, keep in ((1, False), (1, True), (3, False)):
    g = sb.deskew_geometry((600, 300, 2048), 30.0, 0.39, keep, n)
    P, X, Xp = g.out_shape
    for pitch in (Xp, (Xp + 3) // 4 * 4, (Xp + 31) // 32 * 32):
        out = torch.empty((P, X, pitch), dtype=torch.float32, device="cuda")

        def run():
            _cabi.check(lib.shrimpy_deskew_device(raw.data_ptr(), 0, out.data_ptr(), 600, 300, 2048, Xp, n, g.m00, g.m02, g.shift,
                                                  0.0, 0, 0, X * pitch, pitch, _cabi.KERNELS["tma"], st))
        for _ in range(3):
            run()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            run()
        b.record()
        torch.cuda.synchronize()
        res[f"n{n}_keep{int(keep)}_pitch{pitch}"] = round(a.elapsed_time(b) / 10, 4)
        del out
print(json.dumps(res))
