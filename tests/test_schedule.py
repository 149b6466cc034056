"""CPU: the static graph -> CTA schedule of the step kernel (Engine.balance_ids; ids[i] runs on CTA i % grid in round
i // grid).  The closed form used for one partial second round (the BASELINE case: 256 graphs on 148 SMs) must be a
permutation, pair graphs exactly like longest-processing-time-first does, and keep the single-graph CTAs last (the
kernel's slice owners and the documentation rely on "pairs first")."""
import heapq
from types import SimpleNamespace

import numpy as np
import pytest

from drl_urban_planning_b200.engine import Engine


def cta_loads(order, cost, grid):
    loads = np.zeros(grid)
    for i, gid in enumerate(order):
        loads[i % grid] += cost[gid]
    return loads


def lpt_makespan(cost, grid):
    heap = [(0.0, b) for b in range(grid)]
    for c in sorted(cost, reverse=True):
        load, b = heapq.heappop(heap)
        heapq.heappush(heap, (load + float(c), b))
    return max(l for l, _ in heap)


@pytest.mark.parametrize("count,grid,seed", [(256, 148, 0), (256, 148, 1), (200, 148, 2), (295, 148, 3), (149, 148, 4),
                                             (64, 148, 5), (148, 148, 6), (500, 148, 7), (31, 8, 8)])
def test_balance_ids_is_lpt(count, grid, seed):
    rng = np.random.default_rng(seed)
    n = rng.integers(223, 401, size=count)
    cost = (17 * 5.45 * n + 76 * n + 36000).astype(np.float64)          # the kernel's cost model on HLG-like sizes
    ids = rng.permutation(count)
    order = Engine.balance_ids(SimpleNamespace(grid=grid), ids, cost)
    assert sorted(order.tolist()) == sorted(ids.tolist())                # a permutation of the ids it was given
    loads = cta_loads(order, cost, min(grid, count))
    assert loads.max() <= lpt_makespan(cost[ids], min(grid, count)) * (1 + 1e-12)
    if grid < count <= 2 * grid:                                          # one partial second round
        per_cta = np.bincount(np.arange(count) % grid, minlength=grid)
        assert (per_cta[:count - grid] == 2).all() and (per_cta[count - grid:] == 1).all()
        singles = order[count - grid:grid]                               # CTAs count-grid .. grid-1 hold one graph each,
        # and those are the largest graphs (LPT leaves the longest jobs alone)
        largest = set(np.argsort(-cost[ids], kind="stable")[:2 * grid - count].tolist())
        assert {int(np.where(ids == s)[0][0]) for s in singles} == largest


def test_balance_ids_deterministic_and_stable():
    cost = np.full(256, 1000.0)
    ids = np.arange(256)
    a = Engine.balance_ids(SimpleNamespace(grid=148), ids, cost)
    b = Engine.balance_ids(SimpleNamespace(grid=148), ids, cost)
    assert np.array_equal(a, b)
