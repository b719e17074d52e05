"""What a plain ``deskew_data(numpy_array)`` call costs on a B200 box, by where its two host buffers live.

bench.py's ``e2e`` leg hands in pinned arrays for both sides (the contract asks for that); a user of the reference's
``deskew_data`` hands in an ordinary pageable array and takes what comes back.  Since the end of round 1 the result of
such a call is allocated from torch's caching pinned-host allocator (``deskew._empty_pinned_result``); this probe puts
numbers on that choice -- written without a GPU at hand, to be run in round 2:

    python tools/probe/host_call_probe.py > gpurun_out/host_call_probe.json
"""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np
import torch

import shrimpy_b200 as sb

SHAPE, ARGS = (600, 300, 2048), (30.0, 0.39, False, 3)
g = sb.deskew_geometry(SHAPE, *ARGS)
rng = np.random.default_rng(1)
pageable_in = rng.integers(100, 60000, SHAPE, dtype=np.uint16)
pinned_in_t = torch.from_numpy(pageable_in).pin_memory()
pinned_in = pinned_in_t.numpy()
pinned_out_t = torch.empty(g.out_shape, dtype=torch.float32).pin_memory()
pageable_out = np.empty(g.out_shape, dtype=np.float32)
pageable_out[:] = 0                                           # touch the pages before timing


def timed(fn, reps=5):
    fn()
    fn()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    return round((time.perf_counter() - t) / reps * 1e3, 2)


res = {
    "pinned_in__pinned_out_ms": timed(lambda: sb.deskew_data(pinned_in, *ARGS, out=pinned_out_t.numpy())),
    "pageable_in__pinned_out_ms": timed(lambda: sb.deskew_data(pageable_in, *ARGS, out=pinned_out_t.numpy())),
    "pageable_in__pageable_out_ms": timed(lambda: sb.deskew_data(pageable_in, *ARGS, out=pageable_out)),
    "pageable_in__default_out_ms": timed(lambda: sb.deskew_data(pageable_in, *ARGS)),
    "pinned_in__default_out_ms": timed(lambda: sb.deskew_data(pinned_in, *ARGS)),
}
# would page-locking the caller's input for the duration of the call pay?  (register + call + unregister)
rt = torch.cuda.cudart()
t = time.perf_counter()
rc = rt.cudaHostRegister(pageable_in.ctypes.data, pageable_in.nbytes, 0)
res["host_register_737MB_ms"] = round((time.perf_counter() - t) * 1e3, 2)
if int(rc) == 0:
    res["registered_in__pinned_out_ms"] = timed(lambda: sb.deskew_data(pageable_in, *ARGS, out=pinned_out_t.numpy()))
    t = time.perf_counter()
    rt.cudaHostUnregister(pageable_in.ctypes.data)
    res["host_unregister_ms"] = round((time.perf_counter() - t) * 1e3, 2)
t = time.perf_counter()
first = torch.empty((1 << 28,), dtype=torch.float32, pin_memory=True)   # 1 GiB that the cache has not seen
res["first_pin_of_1GiB_ms"] = round((time.perf_counter() - t) * 1e3, 2)
res["out_voxels"] = int(np.prod(g.out_shape))
print(json.dumps(res, indent=1))
