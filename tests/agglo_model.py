"""Sequential model of agglo_kernel's protocol (phybin_b200/csrc/prf_agglo.cuh), used by the CPU tests to check the
DESIGN of the device agglomeration -- row ownership i mod G, nearest-neighbour bookkeeping per owner, the published
per-CTA records, row a's nearest neighbour assembled from the CTAs' shares one step later, dead columns as +inf --
against the host restatement (phb_agglomerate).  It mirrors the kernel phase by phase; it is not a product path."""
import math

import numpy as np

NONE = 0xFFFFFFFF
INF = math.inf


def _less(d1, i1, d2, i2):
    return d1 < d2 or (d1 == d2 and i1 < i2)


def _rescan(D, m, x):
    bd, bj = INF, NONE
    for j in range(x + 1, m):
        if D[x, j] < bd:
            bd, bj = D[x, j], j
    return bd, bj


def agglo_model(full: np.ndarray, linkage: int, thresh: float, G: int):
    """full: m x m symmetric distances.  Returns (merge_a, merge_b, merge_d)."""
    m = full.shape[0]
    D = full.astype(np.float64).copy()
    np.fill_diagonal(D, INF)
    owner = lambda x: x % G
    rows = [[x for x in range(m) if owner(x) == c] for c in range(G)]
    nn = {x: _rescan(D, m, x) for x in range(m)}          # x -> (nnd, nn), held by owner(x)
    act = {x: True for x in range(m)}
    size = [np.ones(m) for _ in range(G)]                  # one copy per CTA
    a_prev = NONE
    share = [(INF, NONE)] * G                              # per-CTA share of row a_prev's nearest neighbour
    merges = []
    step = 0
    while True:
        # publish
        rec = []
        for c in range(G):
            d, i, j = INF, NONE, NONE
            for x in rows[c]:
                if act[x] and x != a_prev and nn[x][1] != NONE and _less(nn[x][0], x, d, i):
                    d, i, j = nn[x][0], x, nn[x][1]
            rec.append((d, i, j, share[c][0], share[c][1]))
        # (grid barrier) every CTA derives the same winner: modelled once
        bd, bi, bj, ad, aj = INF, NONE, NONE, INF, NONE
        for d1, i1, j1, d2, j2 in rec:
            if _less(d1, i1, bd, bi):
                bd, bi, bj = d1, i1, j1
            if _less(d2, j2, ad, aj):
                ad, aj = d2, j2
        if a_prev != NONE:
            nn[a_prev] = (ad, aj)
            if aj != NONE and _less(ad, a_prev, bd, bi):
                bd, bi, bj = ad, a_prev, aj
        if bi == NONE or bd > thresh:
            break
        a, b = bi, bj
        assert a < b
        merges.append((a, b, bd))
        new_share = []
        for c in range(G):
            na, nb = size[c][a], size[c][b]
            size[c][a] += size[c][b]
            pd, pj = INF, NONE
            todo = []
            for x in rows[c]:
                if not act[x]:
                    continue
                if x == b:
                    act[x] = False
                    nn[x] = (INF, NONE)
                    D[a, b] = INF
                    continue
                if x == a:
                    continue
                da, db = D[x, a], D[x, b]
                v = min(da, db) if linkage == 0 else max(da, db) if linkage == 1 else (na * da + nb * db) / (na + nb)
                D[x, a] = v
                D[a, x] = v
                D[x, b] = INF
                if x > a and _less(v, x, pd, pj):
                    pd, pj = v, x
                if x < b:
                    if nn[x][1] in (a, b):
                        todo.append(x)
                    elif x < a and (v < nn[x][0] or (v == nn[x][0] and a < nn[x][1])):
                        nn[x] = (v, a)
            new_share.append((pd, pj))
            for x in todo:
                nn[x] = _rescan(D, m, x)   # reads only row x (own row), except entries of row a never read here (x != a)
        share = new_share
        a_prev = a
        step += 1
    ma = np.array([t[0] for t in merges], dtype=np.uint32)
    mb = np.array([t[1] for t in merges], dtype=np.uint32)
    md = np.array([t[2] for t in merges], dtype=np.float64)
    return ma, mb, md
