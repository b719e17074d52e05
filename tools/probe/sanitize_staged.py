"""Small staged-deskew launches for compute-sanitizer (racecheck: the double-buffered stage with one barrier per chunk;
memcheck: the ragged ends / window offsets).  Checks the results against the plain kernel as well."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import torch

import shrimpy_b200 as sb

gen = torch.Generator(device="cuda").manual_seed(5)
ok = True
for shape, r, keep, n, dtype in [((90, 7, 128), 0.39, False, 1, torch.uint16), ((120, 9, 200), 0.39, True, 2, torch.uint16),
                                 ((75, 5, 96), 0.39, True, 1, torch.float32), ((300, 4, 264), 0.39, False, 3, torch.uint16)]:
    raw = torch.randint(100, 60000, shape, dtype=torch.int32, device="cuda", generator=gen).to(dtype)
    want = sb.deskew_zyx(raw, 30.0, r, keep, n, kernel="tma")
    got = torch.full_like(want, -7.0)
    sb.deskew_zyx(raw, 30.0, r, keep, n, out=got, kernel="tma_staged")
    ok &= bool(torch.equal(got, want))
    g = sb.deskew_geometry(shape, 30.0, r, keep, n)
    P, X, Xp = g.out_shape
    c0, c1 = 5, min(Xp, 5 + 256 + 9)
    _, zr = sb.window_needs(g, 0, P, c0, c1 - c0)
    z0, z1 = (int(zr[0]), int(zr[1])) if zr[1] > zr[0] else (0, 1)
    canvas = torch.full((P, X, Xp), -7.0, device="cuda")
    sb.deskew_window(raw[z0:z1], g, p_begin=0, p_count=P, c_begin=c0, c_count=c1 - c0, y_origin=0, z_origin=z0,
                     out=canvas[:, :, c0:c1], kernel="tma_staged")
    ok &= bool(torch.equal(canvas[:, :, c0:c1], want[:, :, c0:c1]))
torch.cuda.synchronize()
print("equal", ok)
sys.exit(0 if ok else 1)
