#!/usr/bin/env python
"""The min / max network behind the top-(num_reduced + 1) selection of a beta sample (csrc/k_inner_cem.cuh, beta_sample_fast / icl_topk, num_reduced = 5):
keys arrive four at a time; the four are sorted (5 exchanges), merged into the kept six with the half-cleaner c_i = max(t_i, s_(3-i)) (t_4, t_5 pass) and the
resulting sequence is sorted by a 7-exchange network.  This script
  1. derives the reachable 0-1 patterns of c (t and s ascending 0-1 sequences) and searches, by iterative deepening, for a shortest exchange network that sorts all
     of them -- by the 0-1 principle (every operation is monotone and the input family is closed under thresholding) such a network sorts every reachable input;
  2. checks the complete routine (groups of four + single insertions for the tail) against sorting on random key sets.
usage: topk_network.py [trials]"""
import random
import sys

MERGER = [(0, 4), (1, 5), (0, 2), (1, 3), (0, 1), (2, 3), (4, 5)]      # what the kernel uses (the search's result, reordered into three parallel layers)


def reachable_patterns():
    pats = set()
    for kt in range(7):
        t = [0] * kt + [1] * (6 - kt)
        for ks in range(5):
            s = [0, 0] + [0] * ks + [1] * (4 - ks)
            pats.add(tuple(max(t[i], s[5 - i]) for i in range(6)))
    return pats


def apply(ce, pats):
    i, j = ce
    out = set()
    for p in pats:
        if p[i] > p[j]:
            q = list(p); q[i], q[j] = q[j], q[i]; p = tuple(q)
        out.add(p)
    return frozenset(out)


def is_sorted(p):
    return all(p[i] <= p[i + 1] for i in range(len(p) - 1))


def shortest_network():
    pairs = [(i, j) for i in range(6) for j in range(i + 1, 6)]
    start = frozenset(reachable_patterns())

    def dfs(pats, depth, path, seen):
        if all(is_sorted(p) for p in pats):
            return list(path)
        if depth == 0:
            return None
        for ce in pairs:
            nxt = apply(ce, pats)
            if nxt == pats or (nxt, depth - 1) in seen:
                continue
            seen.add((nxt, depth - 1))
            path.append(ce)
            r = dfs(nxt, depth - 1, path, seen)
            if r:
                return r
            path.pop()
        return None
    for depth in range(1, 10):
        r = dfs(start, depth, [], set())
        if r:
            return r


def ce(a, i, j):
    if a[i] > a[j]:
        a[i], a[j] = a[j], a[i]


def top6(keys, merger=MERGER):
    """the kernel's routine on Python ints: the six largest keys, ascending"""
    t = [0] * 6
    q = 0
    while q + 4 <= len(keys):
        s = list(keys[q:q + 4])
        ce(s, 0, 1); ce(s, 2, 3); ce(s, 0, 2); ce(s, 1, 3); ce(s, 1, 2)
        t = [max(t[0], s[3]), max(t[1], s[2]), max(t[2], s[1]), max(t[3], s[0]), t[4], t[5]]
        for i, j in merger:
            ce(t, i, j)
        q += 4
    for v in keys[q:]:
        t[0] = max(t[0], v)
        for p in range(5):
            ce(t, p, p + 1)
    return t


def check(trials, seed=1):
    rng = random.Random(seed)
    for trial in range(trials):
        n = 25
        keys = rng.sample(range(1, 1 << 20), n) if trial % 2 else [rng.randrange(1, 64) * 32 + m for m in range(n)]     # the second kind: many near ties, distinct by index bits
        assert top6(keys) == sorted(keys)[-6:], keys
    return True


if __name__ == "__main__":
    net = shortest_network()
    print("reachable 0-1 patterns:", len(reachable_patterns()), " shortest network:", len(net), "exchanges", net)
    pats = frozenset(reachable_patterns())
    for c in MERGER:
        pats = apply(c, pats)
    assert all(is_sorted(p) for p in pats), "the kernel's merger does not sort every reachable pattern"
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    check(n)
    print("kernel merger sorts every reachable pattern; top-6 routine == sorted()[-6:] on", n, "random key sets")
