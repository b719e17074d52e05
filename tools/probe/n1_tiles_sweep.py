import os, sys, json
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import devbench as d
for k in (1, 2, 3, 5):
    os.environ["SHRIMPY_DESKEW_TILES"] = str(k)
    r = d.time_deskew((600, 300, 2048), "u16", 1, False, "tma")
    print("tiles", k, r["ms_med"], r["ms_min"], flush=True)
os.environ["SHRIMPY_DESKEW_TILES"] = "1"
r = d.time_deskew((4000, 300, 2048), "u16", 1, True, "tma", reps=5, nbuf=1); print("cfg5 tiles 1", r["ms_med"], flush=True)
os.environ["SHRIMPY_DESKEW_TILES"] = "8"
r = d.time_deskew((4000, 300, 2048), "u16", 1, True, "tma", reps=5, nbuf=1); print("cfg5 tiles 8", r["ms_med"], flush=True)
