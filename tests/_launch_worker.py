"""Worker of tests/test_launch_cpu.py: one rank of a world_size-N gloo job (no GPU).  Exercises the launch
plumbing bench.py uses for N > 1 -- brick partition, unique-id broadcast, scalar reductions -- and checks the
partition against the oracle engine's own rank ownership for the same grid."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import support as S  # noqa: E402
from lammps_plugins_b200 import launch, workloads as W  # noqa: E402


def main():
    grp = launch.Group("gloo")
    assert grp.world == int(sys.argv[1]), (grp.world, sys.argv)
    grid = launch.procgrid_for(grp.world)
    for kind in ("rebomos", "aeam"):
        if kind == "rebomos":
            w = W.mos2_bulk(2 * grid[0], grid[1], grid[2])
            lmp = S.make_rebomos_system(S.PORT_SO, (2 * grid[0], grid[1], grid[2]), grid=grid)
        else:
            w = W.fcc_alsi((6 * grid[0], 5 * grid[1], 4 * grid[2]), 0.0, 1)
            lmp = S.make_aeam_system(S.PORT_SO, (6 * grid[0], 5 * grid[1], 4 * grid[2]), grid=grid, si_fraction=0.0)
        mask = launch.my_atoms(w, grid, grp.rank)
        n_mine = int(mask.sum())
        # every atom has exactly one owner
        assert int(grp.reduce_scalar(n_mine, "sum")) == len(w["x"])
        assert grp.reduce_scalar(n_mine, "max") >= len(w["x"]) / grp.world
        # same ownership AND same creation order as the engine rank (create_atoms / replicate on that grid)
        nl = lmp.get_int("nlocal", grp.rank)
        assert nl == n_mine, (kind, nl, n_mine)
        if kind == "rebomos":       # replicate keeps global IDs; create_atoms on R ranks numbers them rank by rank
            assert np.array_equal(np.sort(lmp.tag(grp.rank)[:nl]), np.sort(w["tag"][mask])), kind
        a, b = lmp.x(grp.rank, nl).copy(), w["x"][mask]
        a = a[np.lexsort((a[:, 2].round(6), a[:, 1].round(6), a[:, 0].round(6)))]
        b = b[np.lexsort((b[:, 2].round(6), b[:, 1].round(6), b[:, 0].round(6)))]
        assert np.allclose(a, b, rtol=0, atol=1e-9), kind
        lmp.close()
    # the 128-byte id travels intact from rank 0
    want = bytes((7 * i + 3) % 256 for i in range(128))
    got = launch.nccl_unique_id(grp, lambda: want)
    assert got == want
    assert grp.reduce_scalar(float(grp.rank), "max") == grp.world - 1
    grp.barrier()
    print("LAUNCH_OK rank %d of %d grid %s" % (grp.rank, grp.world, grid), flush=True)
    grp.close()


if __name__ == "__main__":
    main()
