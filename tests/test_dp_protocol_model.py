"""Index arithmetic of the peer-memory gradient exchange, restated in Python and checked exhaustively for small shapes
(CPU only).  Mirrors make_dp_args / vla_dp_create (csrc/vla_api.cu) and dp_exchange_body (csrc/dp_exchange.cu): shard
ownership, RECV slot indexing inside the per-part regions, RSUM coverage -- for world sizes and buffer lengths that do not
divide evenly, and for the two-part (early decoder exchange) split."""
import itertools

import numpy as np
import pytest


def even_shard(n2, world):
    return ((n2 + world - 1) // world + 1) & ~1            # vla_api.cu: per2 is even (a float4 never straddles two shards)


def simulate(n_floats, world, split2=None):
    """Runs the protocol on integer 'gradients' g[r][i] = value unique per (rank, element).  Returns every rank's RSUM."""
    n2 = n_floats // 2
    alloc_per2 = even_shard(n2, world)                     # vla_dp_create: RECV holds 2 regions of world * per2 words
    recv = [np.full((2 * world * alloc_per2, 2), -1, dtype=np.int64) for _ in range(world)]
    rsum = [np.full((n2, 2), -1, dtype=np.int64) for _ in range(world)]
    g = [np.arange(n_floats, dtype=np.int64).reshape(n2, 2) * 100 + r for r in range(world)]
    parts = [(0, 0, n2)] if split2 is None else [(1, split2, n2), (0, 0, split2)]
    for part, first2, end2 in parts:
        cnt2 = end2 - first2
        if cnt2 <= 0:
            continue
        per = even_shard(cnt2, world)
        assert per <= alloc_per2
        base = part * world * alloc_per2
        # phase A: every rank pushes its values of the other ranks' shards
        for me in range(world):
            for s in range(world):
                if s == me:
                    continue
                for j in range(per):
                    if per * s + j < cnt2:
                        idx = base + per * me + j
                        assert base <= idx < base + world * alloc_per2, "RECV write outside the part's region"
                        assert (recv[s][idx] == -1).all(), "RECV slot written twice"
                        recv[s][idx] = g[me][first2 + per * s + j]
        # phase B: the owner adds in rank order and pushes to every rank
        for me in range(world):
            lo = per * me
            for j in range(max(0, min(per, cnt2 - lo))):
                acc = np.zeros(2, dtype=np.int64)
                for s in range(world):
                    v = g[me][first2 + lo + j] if s == me else recv[me][base + per * s + j]
                    assert (v >= 0).all(), "owner read a RECV word nobody wrote"
                    acc += v
                for r in range(world):
                    assert (rsum[r][first2 + lo + j] == -1).all(), "RSUM word written twice"
                    rsum[r][first2 + lo + j] = acc
    return g, rsum


@pytest.mark.parametrize("world,n_floats", list(itertools.product([1, 2, 3, 4, 5, 8, 16], [4, 8, 12, 36, 100, 1028])))
def test_every_element_is_reduced_exactly_once(world, n_floats):
    g, rsum = simulate(n_floats, world)
    want = sum(g)
    for r in range(world):
        np.testing.assert_array_equal(rsum[r], want)


@pytest.mark.parametrize("world,n_floats,split", [(4, 100, 20), (8, 1028, 512), (5, 36, 2), (8, 64, 60), (16, 12, 4)])
def test_two_part_exchange_covers_the_buffer(world, n_floats, split):
    """Early decoder exchange: float2s [split, n2) on the side stream (part 1), [0, split) with the optimizer (part 0)."""
    g, rsum = simulate(n_floats, world, split2=split // 2)
    want = sum(g)
    for r in range(world):
        np.testing.assert_array_equal(rsum[r], want)
