"""Python model of tolerant_kernel's per-pair recipe (phybin_b200/csrc/prf_tolerant.cuh) on the arrays the device gets:
CPU tests check the RECIPE -- the structural skip through parent clades, the per-tree sizes and outside sets of the
repeated-label form -- against the oracle's literal naiveDistMatrix.  Not a product path."""
import numpy as np


def _bits(words) -> int:
    return sum(int(w) << (64 * i) for i, w in enumerate(words))


def _norm(m: int, k: int) -> int:
    pc, half = bin(m).count("1"), k // 2
    if pc < half:
        return m
    fl = ((1 << k) - 1) & ~m
    if pc == half and fl != 0:
        u = m | fl
        low = u & -u
        if not (fl & low):
            return m
    return fl


def parents(clades, off):
    """tol_parent_kernel: smallest other clade of the tree containing it, order (size, index); 0xffff = root."""
    par = np.full(len(clades), 0xFFFF, dtype=np.uint32)
    for t in range(len(off) - 1):
        o, n = int(off[t]), int(off[t + 1] - off[t])
        cs = [_bits(clades[o + i]) for i in range(n)]
        pcs = [bin(c).count("1") for c in cs]
        for i in range(n):
            best, best_pc = 0xFFFF, 1 << 62
            for j in range(n):
                above = pcs[j] > pcs[i] or (pcs[j] == pcs[i] and j > i)
                if (cs[i] & ~cs[j]) == 0 and above and pcs[j] < best_pc:
                    best, best_pc = j, pcs[j]
            par[o + i] = best
    return par


def pair_distance(a, b, clades, off, leafsets, par=None, outsets=None, dup=None):
    """d(a, b) as the kernel computes it; par (distinct labels) or outsets + dup = (dup_off, dup_label, dup_extra)."""
    La, Lb = _bits(leafsets[a]), _bits(leafsets[b])
    S = La & Lb
    ns = bin(S).count("1")
    if ns == 0:
        return 0
    k = [ns, ns]
    if dup is not None:
        doff, dlab, dext = dup
        for side, t in enumerate((a, b)):
            for x in range(int(doff[t]), int(doff[t + 1])):
                if (S >> int(dlab[x])) & 1:
                    k[side] += int(dext[x])
    sets = [set(), set()]
    for side, t in enumerate((a, b)):
        full = (La if side == 0 else Lb) == S
        o = int(off[t])
        for i in range(int(off[t + 1] - off[t])):
            m = _bits(clades[o + i]) & S
            if m == 0:
                continue
            if dup is not None:
                if not full and (_bits(outsets[o + i]) & S) == 0:
                    continue
            else:
                pr = int(par[o + i])
                if pr == 0xFFFF:
                    if not full and m == S:
                        continue
                elif (_bits(clades[o + pr]) & S) == m:
                    continue
            key = _norm(m, k[side])
            if bin(key).count("1") > 1:
                sets[side].add(key)
        if k[side] <= 3:
            for l in range(S.bit_length()):
                if (S >> l) & 1:
                    key = _norm(1 << l, k[side])
                    if bin(key).count("1") > 1:
                        sets[side].add(key)
    return len(sets[0] ^ sets[1])
