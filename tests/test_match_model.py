"""CPU: the matching kernel's formulation against the reference's, as plain Python models.

evaluate.py:259-270 sorts the candidates by priority (descending, stable: ties keep generation order, i-major then
j-minor) and accepts greedily while both end points are free.  limb_match.cuh instead repeats "take the best remaining
candidate whose end points are free" with the order (priority desc, i*nB+j asc) packed into one key, striking every
candidate that shares an end point with the winner.  Same rows, same order -- including ties and the min(nA, nB) stop."""
import random


def reference_greedy(cands, nA, nB):
    order = sorted(cands, key=lambda c: c[2], reverse=True)  # Python's sort is stable; reverse=True keeps it stable
    used_a, used_b, rows = set(), set(), []
    for i, j, _ in order:
        if i not in used_a and j not in used_b:
            rows.append((i, j))
            used_a.add(i); used_b.add(j)
            if len(rows) >= min(nA, nB):
                break
    return rows


def kernel_rounds(cands, nA, nB):
    alive = {(i, j): pr for i, j, pr in cands}
    rows = []
    while alive and len(rows) < min(nA, nB):
        # key = (priority, then EARLIER generation order wins): the kernel packs ~(i << 16 | j) under the priority bits
        (wi, wj), _ = max(alive.items(), key=lambda kv: (kv[1], -(kv[0][0] * nB + kv[0][1])))
        rows.append((wi, wj))
        alive = {(i, j): pr for (i, j), pr in alive.items() if i != wi and j != wj}
    return rows


def test_round_based_matching_equals_stable_sorted_greedy():
    rnd = random.Random(2024)
    for trial in range(3000):
        nA, nB = rnd.randint(1, 9), rnd.randint(1, 9)
        pairs = [(i, j) for i in range(nA) for j in range(nB) if rnd.random() < 0.6]  # generation order: i-major, j-minor
        levels = rnd.choice([2, 3, 50])                                                # few levels => many exact ties
        cands = [(i, j, rnd.randrange(levels) / levels) for i, j in pairs]
        assert kernel_rounds(cands, nA, nB) == reference_greedy(cands, nA, nB), (trial, cands)
