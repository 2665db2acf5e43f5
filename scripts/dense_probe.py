#!/usr/bin/env python
"""Times the tensor-core variant (dense_gram_kernel, PRF_VARIANT_DENSE forced) beside the sparse one on the similar-trees
family, outputs compared bit for bit on the device.  usage: python scripts/dense_probe.py [workload ...]"""
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
    for name in [a for a in sys.argv[1:] if not a.startswith("--")] or ["cfg2_nni", "cfg3_nni"]:
        wl = WORKLOADS[name]
        T, n, ed = wl["T"], wl["n"], wl["editdist"]
        f = Forest.synthetic(T, n, wl["seed"], wl["family"], 8, 0.0)
        bips, off = f.all_bips()
        P = T * (T - 1) // 2
        with torch.cuda.stream(ext):
            d_bips = torch.from_numpy(bips.view(np.int64)).to(dev)
            d_off = torch.from_numpy(off.view(np.int64)).to(dev)
            out_s = torch.empty(P, dtype=torch.int16, device=dev)
            out_d = torch.empty(P, dtype=torch.int16, device=dev)
        ext.synchronize()
        res = {"workload": name}
        for flip, label in ((1, "flipped"), (0, "unflipped")):  # unflipped: the dense universe of shared bipartitions
            ctx._L.prf_debug_no_flip(ctx._h, 0 if flip else 1)
            st = ctx.hashrf_load(d_bips.data_ptr(), d_off.data_ptr(), T, f.words, MEM_DEVICE)
            row = {"n_shared": st["n_shared"], "n_hits": st["n_hits"]}
            if flip:
                row["sparse_ms"] = min(ctx.hashrf_compute(0, P, ed, 1, out_s.data_ptr(), 0)["ms_matrix"] for _ in range(3))
            try:
                ms = [ctx.hashrf_compute(0, P, ed, 2, out_d.data_ptr(), 0)["ms_matrix"] for _ in range(3)]
                row["dense_ms"] = min(ms[1:])
                kpad = (st["n_shared"] + 127) // 128 * 128
                row["dense_tops"] = 2.0 * kpad * P / (row["dense_ms"] * 1e-3) / 1e12
                row["equal_to_sparse"] = bool(torch.equal(out_s, out_d))
            except Exception as e:  # the dense variant refuses shapes it does not support
                row["dense_error"] = str(e)[:200]
            res[label] = row
        ctx._L.prf_debug_no_flip(ctx._h, 0)
        print(json.dumps(res), flush=True)
        del d_bips, d_off, out_s, out_d
    ctx.close()

if __name__ == "__main__":
    main()
