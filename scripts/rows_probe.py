#!/usr/bin/env python
"""A/B probe of the sparse matrix kernels on a B200: the rows kernel (one warp per row, prf_rows.cuh) in its
instantiations against the round-1 CTA-per-row kernel (prf_sparse.cuh).  Outputs are compared bit for bit
(whole matrix, on the device) and each kernel is timed with the library's own CUDA events.

    python scripts/rows_probe.py [workload ...]      # workloads as in bench.py: cfg2 cfg3 cfg4 cfg3_nni
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from bench import WORKLOADS
from phybin_b200.host import Forest
from phybin_b200.rfdistance import MEM_DEVICE, RFContext


def main():
    names = sys.argv[1:] or ["cfg3"]
    dev = torch.device("cuda", 0)
    ctx = RFContext(0)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    for name in names:
        wl = WORKLOADS[name]
        T, n, ed = wl["T"], wl["n"], wl["editdist"]
        f = Forest.synthetic(T, n, wl["seed"], wl["family"], 8, 0.0)
        bips, off = f.all_bips()
        W = f.words
        P = T * (T - 1) // 2
        with torch.cuda.stream(ext):
            d_bips = torch.from_numpy(bips.view(np.int64)).to(dev)
            d_off = torch.from_numpy(off.view(np.int64)).to(dev)
            out_ref = torch.empty(P, dtype=torch.int16, device=dev)
            out_new = torch.empty(P, dtype=torch.int16, device=dev)
            nw = (P + 31) // 32
            m_ref = torch.zeros(nw, dtype=torch.int32, device=dev) if ed >= 0 else None
            m_new = torch.zeros(nw, dtype=torch.int32, device=dev) if ed >= 0 else None
        ext.synchronize()

        def run(which, cfg, short_max, out, mask, reps=4):
            ctx._L.prf_debug_short_max(ctx._h, short_max)
            ctx._L.prf_debug_sparse_kernel(ctx._h, which, cfg)
            st = ctx.hashrf_load(d_bips.data_ptr(), d_off.data_ptr(), T, W, MEM_DEVICE)
            ms = []
            for _ in range(reps):
                s2 = ctx.hashrf_compute(0, P, ed, 1, out.data_ptr(), mask.data_ptr() if mask is not None else 0)
                ms.append(s2["ms_matrix"])
            return st, min(ms[1:]), ms

        st, t_ref, _ = run(1, 0, 0, out_ref, m_ref)
        alg = 2 * P + (P // 8 if ed >= 0 else 0)
        res = {"workload": name, "T": T, "n": n, "editdist": ed, "n_hits": st["n_hits"], "n_shared": st["n_shared"],
               "round1_kernel_ms": t_ref, "round1_gbs": alg / t_ref / 1e6}
        for cfg in (0, 1, 2, 3):
            for sm in ((8, 0) if cfg == 0 else (8,)):
                st2, t_new, all_ms = run(0, cfg, sm, out_new, m_new)
                same = bool(torch.equal(out_ref, out_new))
                same_mask = bool(torch.equal(m_ref, m_new)) if ed >= 0 else None
                res[f"rows_cfg{cfg}_short{sm}"] = {"ms": t_new, "gbs": alg / t_new / 1e6, "equal": same, "mask_equal": same_mask,
                                                    "load_ms": st2["ms_total"]}
                if not same:
                    diff = (out_ref != out_new).nonzero()
                    res[f"rows_cfg{cfg}_short{sm}"]["n_diff"] = int(diff.numel())
                    res[f"rows_cfg{cfg}_short{sm}"]["first_diff"] = int(diff[0]) if diff.numel() else None
        print(json.dumps(res), flush=True)
        del d_bips, d_off, out_ref, out_new, m_ref, m_new
    ctx.close()


if __name__ == "__main__":
    main()
