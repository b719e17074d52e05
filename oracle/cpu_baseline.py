"""Timed CPU run of the reference-form deskew (scipy) over all host cores.

TEST/BENCH INFRASTRUCTURE ONLY (see ``deskew_oracle.py``; parity unpinned).  This is the
"reference arm" of ``bench.py``: ``scipy.ndimage.affine_transform(order=1)`` + edge-padded mean
on ``raw.astype(float32)``, parallelised the way the reference script does it
(``scripts/measure_psf.py:218-249``): the raw X axis is cut into chunks, every chunk is deskewed
independently (scipy.ndimage is single-threaded, so one process per core) and the pieces would be
concatenated in reverse along output axis -2.

Workers are separate ``python -m oracle.cpu_baseline`` processes (never forks of a parent that may
hold a CUDA context).  Each generates its own seeded chunk, reports ``ready``, and starts a pass
only when the parent writes ``go``; the parent's clock runs from the last ``go`` to the last
``done``, so the number is compute time, not start-up or data generation.
"""

from __future__ import annotations

import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def _worker_main(argv) -> None:
    spec = json.loads(argv[0])
    sys.path.insert(0, str(ROOT))
    from oracle import deskew_oracle as o  # noqa: PLC0415

    rng = np.random.default_rng(spec["seed"])
    raw = rng.integers(100, 60000, size=tuple(spec["shape"]), dtype=np.uint16)
    params = tuple(spec["params"])
    o.deskew_data(raw[:, :, :1], *params)  # import / first-call warm-up
    print("ready", flush=True)
    for line in sys.stdin:
        if line.strip() != "go":
            break
        t0 = time.perf_counter()
        out = o.deskew_data(raw, *params)
        dt = time.perf_counter() - t0
        print(json.dumps({"done": dt, "shape": out.shape, "sum": float(out[::7, ::5, ::3].sum())}), flush=True)


def time_scipy_deskew(Z, Y, x_per_worker, workers, params, reps=1, seed=100):
    """``reps`` timed passes of ``workers`` concurrent chunk deskews of shape (Z, Y, x_per_worker).

    Returns ``(wall seconds per pass, output voxels per pass, input voxels per pass)``.
    """
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    procs = []
    for i in range(workers):
        spec = json.dumps({"shape": [Z, Y, x_per_worker], "params": list(params), "seed": seed + i})
        procs.append(subprocess.Popen([sys.executable, "-m", "oracle.cpu_baseline", spec], cwd=str(ROOT), env=env,
                                      stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, bufsize=1))
    walls, out_shape = [], None
    try:
        for p in procs:
            line = p.stdout.readline().strip()
            if line != "ready":
                raise RuntimeError(f"cpu baseline worker failed to start: {line!r}")
        for _ in range(reps):
            t0 = time.perf_counter()
            for p in procs:
                p.stdin.write("go\n")
                p.stdin.flush()
            for p in procs:
                out_shape = json.loads(p.stdout.readline())["shape"]
            walls.append(time.perf_counter() - t0)
    finally:
        for p in procs:
            try:
                p.stdin.close()
            except Exception:
                pass
        for p in procs:
            try:
                p.wait(timeout=30)
            except Exception:
                p.kill()
    out_vox = int(np.prod(out_shape)) * workers
    in_vox = Z * Y * x_per_worker * workers
    return walls, out_vox, in_vox


def calibrate_columns_per_second(Z, Y, params, columns=4):
    """Single-core scipy throughput in raw X columns per second (for sizing bounded samples)."""
    from oracle import deskew_oracle as o  # noqa: PLC0415

    raw = np.random.default_rng(7).integers(100, 60000, size=(Z, Y, columns), dtype=np.uint16)
    o.deskew_data(raw[:, :, :1], *params)
    t0 = time.perf_counter()
    o.deskew_data(raw, *params)
    return columns / (time.perf_counter() - t0)


if __name__ == "__main__":
    _worker_main(sys.argv[1:])
