"""One launch of every hot-path kernel at its BASELINE size, for ncu (round 2 captures under profiles/r02_*).

    python tools/profile_path.py                       # plain run first (must exit 0)
    ncu --set full --clock-control none --import-source on -k regex:"deskew_tma|affine_" -o gpurun_out/prof_path \
        python tools/profile_path.py

Launch order (after one warm-up launch each, skipped with -s in ncu by kernel-name filter being cheap):
  deskew_tma_kernel<uint16,3>           config 2: (600,300,2048) n=3 keep_overhang=False -> (100,2048,1279)
  deskew_tma_staged_kernel<uint16,1>    the same stack, n=1 keep_overhang=False -> (300,2048,1279) contiguous (odd pitch)
  deskew_tma_staged_kernel<uint16,1>    n=1 keep_overhang=True -> (300,2048,1799)
  affine_stream / affine_tilt           config 3: the four matrices of bench.py's affine block
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np
import torch

import shrimpy_b200 as sb
from shrimpy_b200 import register

gen = torch.Generator(device="cuda").manual_seed(1)
raw = torch.randint(100, 60000, (600, 300, 2048), dtype=torch.int32, device="cuda", generator=gen).to(torch.uint16)
for n, keep in ((3, False), (1, False), (1, True)):
    g = sb.deskew_geometry(raw.shape, 30.0, 0.39, keep, n)
    out = torch.empty(g.out_shape, dtype=torch.float32, device="cuda")
    sb.deskew_zyx(raw, 30.0, 0.39, keep, n, out=out)
    torch.cuda.synchronize()
    del out
del raw

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
for M, oshape in ((np.eye(4), shape), (M90, (100, 2048, 1279)), (Mg, shape), (M90t, (100, 2048, 1279))):
    out = torch.empty(oshape, device="cuda")
    register.affine_transform_zyx(vol, M, oshape, out=out)
    torch.cuda.synchronize()
    del out
print("ok")
