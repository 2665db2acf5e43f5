#!/usr/bin/env python
"""Probe of the sparse matrix kernel (prf_rows.cuh) on a B200: its tuning instantiations and the build with / without
singles, timed with the library's own CUDA events; outputs are compared bit for bit with the default configuration
(whole matrix, on the device; parity with the oracle is what pytest -m gpu checks).

    python scripts/rows_probe.py [--default-only] [workload ...]      # workloads as in bench.py: cfg2 cfg3 cfg4 cfg3_nni

--default-only runs just the production configuration (what an ncu capture of rows_kernel should see).
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
    default_only = "--default-only" in sys.argv[1:]
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["cfg3"]
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

        def run(cfg, short_max, out, mask, reps=4):
            ctx._L.prf_debug_short_max(ctx._h, short_max)
            ctx._L.prf_debug_rows_cfg(ctx._h, cfg)
            st = ctx.hashrf_load(d_bips.data_ptr(), d_off.data_ptr(), T, W, MEM_DEVICE)
            ms = []
            for _ in range(reps):
                s2 = ctx.hashrf_compute(0, P, ed, 1, out.data_ptr(), mask.data_ptr() if mask is not None else 0)
                ms.append(s2["ms_matrix"])
            return st, min(ms[1:]), ms

        alg = 2 * P + (P // 8 if ed >= 0 else 0)
        st, t_ref, _ = run(0, 8, out_ref, m_ref)
        res = {"workload": name, "T": T, "n": n, "editdist": ed, "n_hits": st["n_hits"], "n_shared": st["n_shared"],
               "load_ms": st["ms_total"], "dedup_ms": st["ms_dedup"], "lists_ms": st["ms_incidence"],
               "rows_cfg0_short8": {"ms": t_ref, "gbs": alg / t_ref / 1e6}}
        for cfg, sm in () if default_only else ((1, 8), (2, 8), (3, 8)):
            st2, t_new, all_ms = run(cfg, sm, out_new, m_new)
            same = bool(torch.equal(out_ref, out_new))
            same_mask = bool(torch.equal(m_ref, m_new)) if ed >= 0 else None
            res[f"rows_cfg{cfg}_short{sm}"] = {"ms": t_new, "gbs": alg / t_new / 1e6, "equal_to_cfg0": same, "mask_equal": same_mask,
                                                "load_ms": st2["ms_total"]}
        print(json.dumps(res), flush=True)
        del d_bips, d_off, out_ref, out_new, m_ref, m_new
    ctx.close()


if __name__ == "__main__":
    main()
