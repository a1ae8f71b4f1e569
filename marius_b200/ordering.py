"""marius_b200.ordering -- the BETA buffer-state ordering and the greedy edge-bucket assignment the partition buffer is driven with
(data/ordering.cpp:86-148 getBetaOrderingHelper / greedyAssignEdgeBucketsToBuffers, with fine_to_coarse_ratio = 1 and no cached
partitions: EdgeBucketOrdering::OLD_BETA).  Host control logic: it decides WHICH partitions are swapped when, the swap itself is
marius_b200/csrc/host/buffer.cpp.  The reference draws its permutations from libtorch's global generator; here they come from the
numpy Generator that is passed in, so an ordering is a pure function of the seed."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def beta_buffer_states(num_partitions: int, buffer_capacity: int, rng: np.random.Generator) -> List[List[int]]:
    """Sequence of buffer states (ordering.cpp:86-130): consecutive states differ in exactly one partition, and every pair of
    partitions is resident together in at least one state."""
    if not (1 < buffer_capacity <= num_partitions):
        raise ValueError("need 1 < buffer_capacity <= num_partitions")
    all_parts = rng.permutation(num_partitions).astype(np.int64)
    in_buffer = all_parts[:buffer_capacity].copy()
    on_disk = all_parts[buffer_capacity:].copy()
    states = [in_buffer.tolist()]
    while on_disk.size >= 1:
        in_buffer = in_buffer[rng.permutation(in_buffer.size)]
        on_disk = on_disk[rng.permutation(on_disk.size)]
        for i in range(on_disk.size):  # the last slot cycles through everything on disk
            admit = on_disk[i]
            on_disk[i] = in_buffer[-1]
            in_buffer[-1] = admit
            states.append(in_buffer.tolist())
        on_disk = on_disk[rng.permutation(on_disk.size)]
        replaced = 0
        for i in range(buffer_capacity - 1):  # then the other slots are refilled from disk
            if i >= on_disk.size:
                break
            replaced += 1
            in_buffer[i] = on_disk[i]
            states.append(in_buffer.tolist())
        on_disk = on_disk[replaced:]
    return states


def greedy_edge_buckets(buffer_states: List[List[int]], num_partitions: int) -> List[List[Tuple[int, int]]]:
    """Every edge bucket (src partition, dst partition) goes to the FIRST buffer state that holds both partitions (ordering.cpp:132-148)."""
    seen = np.zeros((num_partitions, num_partitions), bool)
    out: List[List[Tuple[int, int]]] = []
    for state in buffer_states:
        cur = []
        for s in state:
            for d in state:
                if not seen[s, d]:
                    seen[s, d] = True
                    cur.append((int(s), int(d)))
        out.append(cur)
    if not seen.all():
        raise ValueError("ordering does not cover every edge bucket")
    return out


def beta_ordering(num_partitions: int, buffer_capacity: int, seed: int = 0):
    """(buffer states, edge buckets per buffer state) -- getEdgeBucketOrdering(OLD_BETA) (ordering.cpp:12-20)."""
    rng = np.random.default_rng(seed)
    states = beta_buffer_states(num_partitions, buffer_capacity, rng)
    return states, greedy_edge_buckets(states, num_partitions)
