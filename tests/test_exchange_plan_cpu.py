"""CommBrick::exchange's compaction order (host logic of the multi-GPU migration, no device needed):
b200md_exchange_plan against a literal restatement of the reference loop (oracle/engine/comm.cpp Comm::exchange,
LAMMPS comm_brick.cpp: "when atom is deleted, fill it in with last atom")."""
import ctypes

import numpy as np
import pytest

import lammps_plugins_b200 as b2


def reference_loop(n, leaves):
    """the sequential loop on an explicit array of atom IDs"""
    ids = list(range(n))
    packed = []
    nlocal, i = n, 0
    while i < nlocal:
        if leaves[ids[i]]:
            packed.append(ids[i])
            ids[i] = ids[nlocal - 1]
            nlocal -= 1
        else:
            i += 1
    return packed, ids[:nlocal]


def plan(n, leavers):
    L = b2.lib()
    lv = np.ascontiguousarray(leavers, dtype=np.int32)
    order = np.zeros(max(len(lv), 1), dtype=np.int32)
    moves = np.zeros(2 * max(len(lv), 1) + 2, dtype=np.int32)
    nm, nl = ctypes.c_int(), ctypes.c_int()
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
    rc = L.b200md_exchange_plan(n, ip(lv), len(lv), ip(order), ip(moves), ctypes.byref(nm), ctypes.byref(nl))
    assert rc == 0
    ids = np.arange(n)
    for k in range(nm.value):
        ids[moves[2 * k]] = moves[2 * k + 1]
    return list(order[:len(lv)]), list(ids[:nl.value])


CASES = [("none", 10, []), ("all", 7, list(range(7))), ("tail", 9, [6, 7, 8]), ("head", 9, [0, 1, 2]),
         ("alternating", 11, list(range(0, 11, 2))), ("single-last", 5, [4]), ("single-first", 5, [0]),
         ("one-atom", 1, [0]), ("empty", 0, []), ("hole-then-tail-run", 12, [2, 9, 10, 11]),
         ("last-two-and-first", 6, [0, 4, 5])]


@pytest.mark.parametrize("name,n,leavers", CASES, ids=[c[0] for c in CASES])
def test_plan_matches_reference_loop(name, n, leavers):
    leaves = np.zeros(max(n, 1), dtype=bool)
    leaves[leavers] = True
    want_packed, want_ids = reference_loop(n, leaves)
    got_packed, got_ids = plan(n, leavers)
    assert got_packed == want_packed
    assert got_ids == want_ids


def test_plan_random():
    rng = np.random.default_rng(5)
    for trial in range(300):
        n = int(rng.integers(1, 200))
        frac = rng.choice([0.02, 0.2, 0.5, 0.9])
        leaves = rng.random(n) < frac
        want_packed, want_ids = reference_loop(n, leaves)
        got_packed, got_ids = plan(n, np.nonzero(leaves)[0])
        assert got_packed == want_packed and got_ids == want_ids, trial
