"""The reference's in-place partition (bvh.cpp:70-81: `if small(a[i]) i++ else swap(a[i], a[j--])`) has a closed form
that needs only prefix counts -- the basis for a parallel (or device-side) BVH build that stays index-for-index
identical to BuildBVH.  With k = number of smalls:

  left part  [0, k):   a small stays where it is; the m-th LARGE found there (ascending) is replaced by the m-th SMALL
                       of the right part counted from the end;
  right part, written from the end backwards: for m = 0, 1, ...: the m-th left-part large, then the right-part larges
                       lying between the (m-1)-th and the m-th of those smalls (descending positions);
  what is left, the untouched middle [k, hi] (all large): rotated left by one position.

This script checks the closed form against the sequential loop on random arrays (exhaustively small sizes)."""
import itertools
import random


def sequential(a, small):
    a = list(a); i = 0; j = len(a) - 1
    while i <= j:
        if small(a[i]):
            i += 1
        else:
            a[i], a[j] = a[j], a[i]; j -= 1
    return a, i


def closed_form(a, small):
    n = len(a); k = sum(1 for x in a if small(x))
    out = [None] * n
    left_large = [p for p in range(k) if not small(a[p])]
    right_small = [q for q in range(n - 1, k - 1, -1) if small(a[q])]
    for p in range(k):
        if small(a[p]):
            out[p] = a[p]
    for m, p in enumerate(left_large):
        out[p] = a[right_small[m]]
    w = hi = n - 1
    for m, p in enumerate(left_large):
        out[w] = a[p]; w -= 1
        for q in range(hi, right_small[m], -1):
            if not small(a[q]):
                out[w] = a[q]; w -= 1
        hi = right_small[m] - 1
    middle = [a[q] for q in range(k, hi + 1)]
    for t, q in enumerate(range(k, hi + 1)):
        out[q] = middle[(t + 1) % len(middle)]
    return out, k


if __name__ == "__main__":
    small = lambda x: x[0]
    for n in range(0, 13):                                   # every small/large pattern up to 12 elements
        for bits in itertools.product((False, True), repeat=n):
            a = [(b, t) for t, b in enumerate(bits)]
            assert sequential(a, small) == closed_form(a, small), a
    for _ in range(100000):
        n = random.randint(13, 200)
        pr = random.choice([0.05, 0.3, 0.5, 0.7, 0.95])
        a = [(random.random() < pr, t) for t in range(n)]
        assert sequential(a, small) == closed_form(a, small), a
    print("closed form == sequential partition on all cases")
