#!/usr/bin/env python
"""One hash class of a multi-GPU build as a single-GPU problem (every tree keeps ~1/world of its bipartitions, all copies
of a bipartition stay together): lets ncu look at the class build's kernels on one GPU.
usage: python scripts/class_probe.py [world] [workload]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import WORKLOADS
from phybin_b200.host import Forest
from phybin_b200.rfdistance import MEM_DEVICE, RFContext

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
wl = WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "cfg3"]
f = Forest.synthetic(wl["T"], wl["n"], wl["seed"], wl["family"], 8, 0.0)
bips, off = f.all_bips()
h = bips[:, 0].copy()
for w in range(1, bips.shape[1]):
    h = (h * np.uint64(0x9E3779B97F4A7C15) + bips[:, w]) & np.uint64(0xFFFFFFFFFFFFFFFF)
keep = ((h >> np.uint64(33)) % np.uint64(world)) == 0
tree_of = np.repeat(np.arange(wl["T"]), np.diff(off.astype(np.int64)))
cnt = np.bincount(tree_of[keep], minlength=wl["T"])
off2 = np.zeros(wl["T"] + 1, dtype=np.uint64)
off2[1:] = np.cumsum(cnt)
bips2 = np.ascontiguousarray(bips[keep])
dev = torch.device("cuda", 0)
ctx = RFContext(0)
ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
with torch.cuda.stream(ext):
    d_b = torch.from_numpy(bips2.view(np.int64)).to(dev)
    d_o = torch.from_numpy(off2.view(np.int64)).to(dev)
ext.synchronize()
best = None
for _ in range(6):
    st = ctx.hashrf_load(d_b.data_ptr(), d_o.data_ptr(), wl["T"], f.words, MEM_DEVICE)
    cur = (st["ms_total"], st["ms_dedup"], st["ms_incidence"])
    best = cur if best is None or cur[0] < best[0] else best
print(json.dumps({"world": world, "class_entries": int(off2[-1]), "load_ms": best[0], "dedup_ms": best[1], "lists_ms": best[2], "n_unique": st["n_unique"]}))
ctx.close()
