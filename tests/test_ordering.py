"""BETA ordering / greedy bucket assignment (marius_b200/ordering.py, restating data/ordering.cpp:86-148): the invariants the partition
buffer relies on.  (The reference's own permutations come from libtorch's global generator and are not reproducible.)"""
import itertools

import pytest

from marius_b200 import ordering


@pytest.mark.parametrize("P,c", [(5, 2), (16, 8), (64, 8), (8, 8), (9, 4)])
@pytest.mark.parametrize("seed", [0, 1, 7])
def test_beta_ordering_invariants(P, c, seed):
    states, buckets = ordering.beta_ordering(P, c, seed)
    assert all(len(s) == c and len(set(s)) == c and all(0 <= p < P for p in s) for s in states)
    for a, b in zip(states, states[1:]):  # one partition swapped per step (positions may be permuted: the buffer matches by partition id)
        assert len(set(b) - set(a)) == 1 and len(set(a) - set(b)) == 1
    flat = list(itertools.chain.from_iterable(buckets))
    assert len(flat) == P * P and len(set(flat)) == P * P  # every edge bucket exactly once
    for s, bs in zip(states, buckets):
        assert all(i in s and j in s for i, j in bs)  # ... while both of its partitions are resident
    assert len(buckets) == len(states)
    # the first state gets its c*c buckets, every later state only the buckets of the partition just admitted
    assert len(buckets[0]) == c * c
    assert all(len(b) <= 2 * c - 1 for b in buckets[1:])
