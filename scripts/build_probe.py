#!/usr/bin/env python
"""Times the build phase (prf_hashrf_load on device-resident keys) on a B200: phases from the library's CUDA events,
best of 6 loads.  usage: python scripts/build_probe.py [workload ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import WORKLOADS
from phybin_b200.host import Forest
from phybin_b200.rfdistance import MEM_DEVICE, RFContext

def main():
    dev = torch.device("cuda", 0)
    ctx = RFContext(0)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    for name in sys.argv[1:] or ["cfg3"]:
        wl = WORKLOADS[name]
        f = Forest.synthetic(wl["T"], wl["n"], wl["seed"], wl["family"], 8, 0.0)
        bips, off = f.all_bips()
        with torch.cuda.stream(ext):
            d_bips = torch.from_numpy(bips.view(np.int64)).to(dev)
            d_off = torch.from_numpy(off.view(np.int64)).to(dev)
        ext.synchronize()
        res = {"workload": name}
        for cfg in (0,):
            best = None
            for _ in range(6):
                st = ctx.hashrf_load(d_bips.data_ptr(), d_off.data_ptr(), wl["T"], f.words, MEM_DEVICE)
                cur = (st["ms_total"], st["ms_dedup"], st["ms_incidence"])
                best = cur if best is None or cur[0] < best[0] else best
            res["build"] = {"load_ms": round(best[0], 4), "dedup_ms": round(best[1], 4), "lists_ms": round(best[2], 4),
                                                   "n_unique": st["n_unique"], "n_hits": st["n_hits"]}
        print(json.dumps(res), flush=True)
    ctx.close()

if __name__ == "__main__":
    main()
