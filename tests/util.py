"""Shared helpers for the tests: random Newick generators and oracle adapters."""
import random

import numpy as np

from oracle import oracle as orc


def random_multifurcating_newicks(rng: random.Random, T: int, n: int, p_missing: float = 0.0,
                                  unary: float = 0.0, prefix: str = "x") -> str:
    """T random rooted trees on up to n taxa: random multifurcations, optional missing taxa and unary
    interior nodes -- the shapes the reference's tolerant path has to cope with."""
    def one():
        labs = [f"{prefix}{l}" for l in range(n) if rng.random() >= p_missing] or [f"{prefix}0"]
        nodes = labs[:]
        rng.shuffle(nodes)
        while len(nodes) > 1:
            k = min(len(nodes), rng.choice([2, 2, 2, 3, 4]))
            if len(nodes) <= 3 and rng.random() < 0.5:
                k = len(nodes)
            grp = [nodes.pop(rng.randrange(len(nodes))) for _ in range(k)]
            s = "(" + ",".join(grp) + ")"
            if rng.random() < unary:
                s = "(" + s + ")"
            nodes.append(s)
        return nodes[0] + ";"
    return "".join(one() for _ in range(T))


def oracle_packed(forest: "orc.Forest", mode: str, variant: str = "literal") -> np.ndarray:
    """Oracle matrix in the C ABI's packed upper-triangle order."""
    T = forest.num_trees
    flat = forest.naive() if mode == "naive" else forest.hashrf(0, variant)
    return orc.lower_to_packed_upper(flat, T)
