"""One launch of every kernel of the rows either side of the deskew (flat-field, tracking reductions), for ncu.

    ncu --set full --clock-control none -k regex:"median_z|pattern_|apply_scale|minmax_kernel|hist256|com_kernel|zmax|deskew_tma" \
        -o gpurun_out/prof_rows python tools/profile_rows.py
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch

import shrimpy_b200 as sb
from shrimpy_b200 import flatfield as ffm
from shrimpy_b200 import reductions as red

shape = (600, 300, 2048)
gen = torch.Generator(device="cuda").manual_seed(3)
raw = torch.randint(2000, 9000, shape, dtype=torch.int32, device="cuda", generator=gen).to(torch.uint16)
scale = ffm.flat_field_scale(raw)                                   # median_z + pattern_sum + pattern_scale
out = ffm.deskew_flat_field_zyx(raw, 30.0, 0.39, False, 3, scale=scale)   # deskew_tma_kernel<.., SCALED>
ff32 = ffm.flat_field_BF(raw)                                       # median, scale, apply_scale
del ff32
vmin, vmax = red.value_range(out)                                   # minmax
bg = red.percentile(out, 50.0)                                      # minmax + hist256
com = red.intensity_center_of_mass(out, bg)                         # com
mip = red.max_projection(out, bg)                                   # zmax
torch.cuda.synchronize()
print("range", vmin, vmax, "background", bg, "centre of mass", com)
