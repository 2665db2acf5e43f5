"""Times the consumers of the matrix at a BASELINE shape: device extraction, matrix + mask, components, flat
clusters for every linkage, strict consensus of every cluster.  usage: python scripts/consumers_probe.py T n nni D [--host] [--dendrogram]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phybin_b200.host import Forest, NNI_FAMILY  # noqa: E402
from phybin_b200.rfdistance import RFContext  # noqa: E402

T, n, nni, D = (int(x) for x in sys.argv[1:5])
f = Forest.synthetic(T, n, 0x5EED0003, family=NNI_FAMILY if nni else 0, nni_max=max(nni, 1))
ctx = RFContext(0)
out = {"T": T, "n": n, "nni": nni, "editdist": D}
for rep in range(3):  # the last pass is the warm one (the stream-ordered memory pool has reached its size)
    t0 = time.time()
    st = ctx.hashrf_load_trees(f)
    t1 = time.time()
    st2 = ctx.hashrf_compute(editdist=D)
    t2 = time.time()
    labels, ncomp = ctx.mask_components()
    t3 = time.time()
    keys, off, info = ctx.consensus(labels)
    t4 = time.time()
    out.update(extract_ms=st["extract"]["ms_kernels"], load_ms=(t1 - t0) * 1e3, matrix_ms=st2["ms_matrix"],
               compute_wall_ms=(t2 - t1) * 1e3, components_ms=(t3 - t2) * 1e3, n_components=int(ncomp),
               consensus_kernels_ms=info["ms_kernels"], consensus_wall_ms=(t4 - t3) * 1e3, consensus_sets=int(info["n_out"]),
               largest_component=int(np.bincount(labels).max()))
big = np.sort(np.flatnonzero(labels == np.bincount(labels).argmax())).astype(np.uint32)
for linkage in ("complete", "UPGMA"):
    ctx.flat_clusters(D, linkage)   # warm-up (memory pool)
    t0 = time.time()
    lab, k = ctx.flat_clusters(D, linkage)
    out[f"flat_{linkage}_ms"] = (time.time() - t0) * 1e3
    out[f"flat_{linkage}_clusters"] = int(k)
    t0 = time.time()
    loc, ma, mb, md, info = ctx.agglomerate(big, linkage, D)
    out[f"largest_component_{linkage}_device_ms"] = info["ms_kernels"]
    out[f"largest_component_{linkage}_merges"] = int(info["n_merges"])
    if "--host" in sys.argv:   # the host restatement on the same inputs (round 1's path): time and equality
        t0 = time.time()
        lab_h, _ = ctx.flat_clusters(D, linkage, device_min=1 << 30)
        out[f"flat_{linkage}_host_ms"] = (time.time() - t0) * 1e3
        out[f"flat_{linkage}_equal_host"] = bool(np.array_equal(lab, lab_h))
if "--dendrogram" in sys.argv:     # the full dendrogram of every tree (dendrogram.txt)
    for linkage in ("single", "complete", "UPGMA"):
        _, ma, mb, md, info = ctx.agglomerate(None, linkage)
        out[f"dendrogram_{linkage}_ms"] = info["ms_kernels"]
        out[f"dendrogram_{linkage}_merges"] = int(info["n_merges"])
print(json.dumps(out))
ctx.close()
