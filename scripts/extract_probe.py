"""Times allBips on the device (prf_allbips) beside the C++ host extraction (phb_allbips) at a BASELINE shape and
checks that both produce the same arrays.  usage: python scripts/extract_probe.py T n [p_missing] [reps]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phybin_b200.host import Forest  # noqa: E402
from phybin_b200.rfdistance import EXTRACT_BIPS, EXTRACT_CLADES, RFContext  # noqa: E402

T, n = int(sys.argv[1]), int(sys.argv[2])
p_missing = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
mode = EXTRACT_CLADES if p_missing > 0 else EXTRACT_BIPS
f = Forest.synthetic(T, n, 0x5EED0003, p_missing=p_missing)
t0 = time.time()
host = f.raw_clades() if mode == EXTRACT_CLADES else f.all_bips()
t_host = time.time() - t0
t0 = time.time()
host1 = f.raw_clades(threads=1) if mode == EXTRACT_CLADES else f.all_bips(threads=1)
t_host1 = time.time() - t0
ctx = RFContext(0)
best = None
for _ in range(reps):
    t0 = time.time()
    info = ctx.allbips(f, mode)
    info["wall_ms"] = (time.time() - t0) * 1e3
    if best is None or info["wall_ms"] < best["wall_ms"]:
        best = info
dev = ctx.allbips_fetch()
same = all(np.array_equal(a, b) for a, b in zip(dev, host))
arrays = f.tree_arrays()
print(json.dumps({"T": T, "n": n, "mode": "clades" if mode else "bips", "nnz": int(best["nnz"]), "identical_to_host": bool(same),
                  "host_s_all_threads": round(t_host, 3), "host_s_1_thread": round(t_host1, 3), "host_threads": os.cpu_count(),
                  "device_ms_h2d": round(best["ms_h2d"], 3), "device_ms_kernels": round(best["ms_kernels"], 3),
                  "device_wall_ms": round(best["wall_ms"], 3), "input_bytes": int(sum(a.nbytes for a in arrays)),
                  "output_bytes": int(dev[0].nbytes), "launches": int(best["gpu_launches"])}))
ctx.close()
