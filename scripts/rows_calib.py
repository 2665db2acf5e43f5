#!/usr/bin/env python
"""Calibration of the rows kernel on a B200 (a -DPRF_CALIB build, lib/libphybin_rf_calib.so; results are WRONG by
design): how long the kernel takes with the list walk removed (fill + bulk store only) and with the member loads
kept but the tile updates disabled.
usage: python scripts/rows_calib.py [workload] [--cfgs 0,1,2] [--modes 0,1,2] [--build]
rows_cfg 10 + k (calibration builds only) = the default 13-warp kernel with optimisation bits k: 1 = whole-block tiles
skip the per-member range test, 2 = a row's singles live in a register, 4 = the lane's current list is cached."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phybin_b200 import build
calib_so = os.path.join(build.LIBDIR, "libphybin_rf_calib.so")
build.RF_SO = calib_so
import numpy as np, torch
from bench import WORKLOADS
from phybin_b200 import _native
from phybin_b200.host import Forest
from phybin_b200.rfdistance import MEM_DEVICE, RFContext

def main():
    args = [a for a in sys.argv[1:]]
    def opt(flag, default):
        return [int(x) for x in args[args.index(flag) + 1].split(",")] if flag in args else default
    cfgs, modes = opt("--cfgs", [0, 1, 2]), opt("--modes", [0, 1, 2])
    names = [a for i, a in enumerate(args) if not a.startswith("--") and (i == 0 or args[i - 1] not in ("--cfgs", "--modes"))]
    name = names[0] if names else "cfg3"
    wl = WORKLOADS[name]
    T, n, ed = wl["T"], wl["n"], wl["editdist"]
    dev = torch.device("cuda", 0)
    ctx = RFContext(0)
    L = ctx._L
    L.prf_debug_calib.argtypes = [_native.C.c_void_p, _native.C.c_int]
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    f = Forest.synthetic(T, n, wl["seed"], wl["family"], 8, 0.0)
    bips, off = f.all_bips()
    P = T * (T - 1) // 2
    with torch.cuda.stream(ext):
        d_bips = torch.from_numpy(bips.view(np.int64)).to(dev)
        d_off = torch.from_numpy(off.view(np.int64)).to(dev)
        out = torch.empty(P, dtype=torch.int16, device=dev)
        mask = torch.zeros((P + 31) // 32, dtype=torch.int32, device=dev) if ed >= 0 else None
    ext.synchronize()
    ctx.hashrf_load(d_bips.data_ptr(), d_off.data_ptr(), T, f.words, MEM_DEVICE)
    res = {"workload": name}
    for cfg in cfgs:
        L.prf_debug_rows_cfg(ctx._h, cfg)
        for mode, label in ((0, "full"), (1, "fill_store_only"), (2, "loads_no_updates"), (3, "no_group_loads"), (4, "no_red_instructions"),
                            (5, "bounds_and_scan_only")):
            if mode not in modes:
                continue
            L.prf_debug_calib(ctx._h, mode)
            ms = [ctx.hashrf_compute(0, P, ed, 1, out.data_ptr(), mask.data_ptr() if mask is not None else 0)["ms_matrix"] for _ in range(4)]
            res[f"cfg{cfg}_{label}"] = round(min(ms[1:]), 4)
    print(json.dumps(res))
    ctx.close()

if __name__ == "__main__":
    if not os.path.exists(calib_so) or "--build" in sys.argv:
        build.build_cuda(True, extra_flags=["-DPRF_CALIB"], out=calib_so)
    if "--build-only" not in sys.argv:
        main()
