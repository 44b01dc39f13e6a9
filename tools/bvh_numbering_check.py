"""CPU only.  The reference numbers BVH nodes in the order its recursion allocates them (bvh.cpp:89-97: two children per
split, left subtree before right).  For a level-synchronous (device-side) build that order has to come out of the tree's
shape alone:   left_child(X) = 1 + 2 * (number of interior nodes before X in pre-order),
with the pre-order rank following from subtree interior counts (rank(left) = rank(X) + 1, rank(right) = rank(X) + 1 +
interior(left subtree)).  This script checks the formula on the reference-built trees in tests/golden/."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cobbletrace_b200 as ct  # noqa: E402
from cobbletrace_b200 import host  # noqa: E402


def check(fs, name):
    left, count = fs.node_left.astype(np.int64), fs.node_count
    n = fs.n_nodes
    interior = count == 0
    order = []                                   # any top-down order: here a stack walk
    stack = [0]
    while stack:
        x = stack.pop()
        order.append(x)
        if interior[x]:
            stack += [left[x] + 1, left[x]]
    inner = np.zeros(n, np.int64)                # interior nodes in the subtree, bottom-up
    for x in reversed(order):
        if interior[x]:
            inner[x] = 1 + inner[left[x]] + inner[left[x] + 1]
    rank = np.zeros(n, np.int64)                 # interior nodes before x in pre-order, top-down
    for x in order:
        if interior[x]:
            rank[left[x]] = rank[x] + 1
            rank[left[x] + 1] = rank[x] + 1 + inner[left[x]]
    ok = all(left[x] == 1 + 2 * rank[x] for x in order if interior[x])
    print(f"{name}: {n} nodes, {int(interior.sum())} interior, numbering formula {'holds' if ok else 'FAILS'}")
    return ok


if __name__ == "__main__":
    gold = os.path.join(ROOT, "tests", "golden")
    meta = json.load(open(os.path.join(gold, "golden.json")))
    good = True
    for name, m in meta["scenes"].items():
        fs = ct.load_ctscene(os.path.join(gold, m["file"]))
        if not fs.has_bvh():
            fs = host.HostScene.from_flat(fs).to_flat(with_bvh=True)     # digest-equal to the reference's (tests/test_host.py)
        good &= check(fs, name)
    sys.exit(0 if good else 1)
