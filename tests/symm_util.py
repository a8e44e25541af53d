"""Symmetry operations for the unfold_w tests: the 48 cubic point-group matrices in the crystal axes of a simple-cubic lattice."""
import itertools

import numpy as np


def cubic_group():
    ops = []
    for perm in itertools.permutations(range(3)):
        for signs in itertools.product((1, -1), repeat=3):
            m = np.zeros((3, 3), dtype=np.int64)
            for r in range(3):
                m[r, perm[r]] = signs[r]
            ops.append(m)
    ident = [i for i, m in enumerate(ops) if np.array_equal(m, np.eye(3, dtype=np.int64))][0]
    ops.insert(0, ops.pop(ident))                      # identity first, as in QE
    invs = np.zeros(len(ops), dtype=np.int32)
    for i, a in enumerate(ops):
        for j, b in enumerate(ops):
            if np.array_equal(a @ b, np.eye(3, dtype=np.int64)):
                invs[i] = j + 1
    return ops, invs


def g_shell_list(nmax):
    """All integer vectors with |m|^2 <= nmax, ordered by length then lexicographically (a closed set under the cubic group)."""
    r = int(np.ceil(np.sqrt(nmax)))
    pts = [m for m in itertools.product(range(-r, r + 1), repeat=3) if sum(x * x for x in m) <= nmax]
    pts.sort(key=lambda m: (sum(x * x for x in m), m))
    return np.array(pts, dtype=np.int64)
