"""Tuning probe of the agglomeration kernel: the largest component of the NNI family at --editdist D, agglomerated with
different numbers of CTAs (prf_debug_agglo_ctas).  usage: python scripts/agglo_probe.py T n nni D [ctas ...]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phybin_b200.host import Forest, NNI_FAMILY  # noqa: E402
from phybin_b200.rfdistance import RFContext  # noqa: E402

T, n, nni, D = (int(x) for x in sys.argv[1:5])
grids = [int(x) for x in sys.argv[5:]] or [0, 16, 32, 64, 96, 148]
f = Forest.synthetic(T, n, 0x5EED0003, family=NNI_FAMILY, nni_max=max(nni, 1))
ctx = RFContext(0)
ctx.hashrf_load_trees(f)
ctx.hashrf_compute(editdist=D)
labels, _ = ctx.mask_components()
big = np.sort(np.flatnonzero(labels == np.bincount(labels).argmax())).astype(np.uint32)
out = {"T": T, "n": n, "editdist": D, "component": int(len(big)), "runs": []}
ref = {}
for g in grids:
    ctx._ck(ctx._L.prf_debug_agglo_ctas(ctx._h, g))
    for linkage in ("complete", "UPGMA"):
        ctx.agglomerate(big, linkage, D)   # warm-up
        lab, ma, mb, md, info = ctx.agglomerate(big, linkage, D)
        key = (linkage,)
        if key not in ref:
            ref[key] = (ma.copy(), mb.copy(), md.copy())
        same = all(np.array_equal(x, y) for x, y in zip(ref[key], (ma, mb, md)))
        out["runs"].append({"ctas": g, "linkage": linkage, "ms": info["ms_kernels"], "merges": int(info["n_merges"]),
                            "us_per_merge": 1e3 * info["ms_kernels"] / max(int(info["n_merges"]), 1), "same_merges": bool(same)})
print(json.dumps(out))
ctx.close()
