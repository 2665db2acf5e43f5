#!/usr/bin/env python
"""Phase times of the multi-GPU build with every rank on ONE GPU (single-process team, PRF_TRACE=1 serialises and
times the phases): separates what the kernels cost from what NVLink costs.  usage: PRF_TRACE=1 python scripts/team_probe.py [world]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bench import WORKLOADS
from phybin_b200.host import Forest
from phybin_b200.rfdistance import MultiRF

world = int(sys.argv[1]) if len(sys.argv) > 1 else 4
wl = WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "cfg3"]
T = int(sys.argv[3]) if len(sys.argv) > 3 else wl["T"] // 4   # a quarter of the trees: the result must fit one GPU's host round trip
f = Forest.synthetic(T, wl["n"], wl["seed"], wl["family"], 8, 0.0)
bips, off = f.all_bips()
m = MultiRF(devices=[0] * world)
for _ in range(3):
    sys.stderr.write("---- call\n")
    out, _, st = m.hashrf(bips, off)
print(st)
m.close()
