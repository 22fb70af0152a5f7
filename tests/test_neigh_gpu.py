"""GPU neighbor build vs the host build of the oracle engine (LAMMPS-core NBinStandard +
NPairFullBin[Ghost] restated): rows must be BIT-EXACT -- same members in the same order."""
import numpy as np
import pytest

import support as S

pytestmark = pytest.mark.gpu


def check_lists(ctx, lmp, ghost_rows):
    snap = S.snapshot(lmp)
    nt = lmp.get_int("ntypes")
    ctx.neigh_build(lmp.b200_box(), nt, lmp.cutneighsq(False), lmp.cutneighsq(True), snap["nlocal"], snap["nghost"],
                    snap["x"], snap["type"], ghost_rows, snap["skin"])
    num, off, val = ctx.neigh_download()
    nrows = snap["inum"] + snap["gnum"]
    assert len(num) == nrows
    ref_num = np.diff(snap["off"])
    bad = np.nonzero(num != ref_num)[0]
    assert bad.size == 0, "row length differs first at row %d: %d vs %d" % (bad[0], num[bad[0]], ref_num[bad[0]])
    assert np.array_equal(off, snap["off"])
    assert np.array_equal(val, snap["val"] & 0x1FFFFFFF)
    return snap


REBO_CASES = [
    dict(id="bulk288", replicate=(1, 1, 1), displace=0.0),
    dict(id="bulk288-d0.4", replicate=(1, 1, 1), displace=0.4),
    dict(id="rep3x2x2-d0.2", replicate=(3, 2, 2), displace=0.2),
]


@pytest.mark.parametrize("case", REBO_CASES, ids=[c["id"] for c in REBO_CASES])
def test_rebomos_full_bin_ghost(ctx, oracle_built, case):
    """pair build full/bin/ghost, triclinic box, ghost rows with the shorter cutghost+skin cutoffs"""
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), case["replicate"], displace=case["displace"])
    lmp.setup(1, 2)
    snap = check_lists(ctx, lmp, ghost_rows=1)
    if case["id"] == "bulk288":
        assert snap["off"][288] == 142848          # FullNghs of log.rebomos-bulk.1:79
        assert snap["nghost"] == 4285              # Nghost    of log.rebomos-bulk.1:75
    lmp.close()


AEAM_CASES = [
    dict(id="fcc4", cells=(4, 4, 4), si=0.0075, displace=0.0),
    dict(id="fcc6x5x7-d0.3", cells=(6, 5, 7), si=0.1, displace=0.3),
]


@pytest.mark.parametrize("case", AEAM_CASES, ids=[c["id"] for c in AEAM_CASES])
def test_aeam_full_bin(ctx, oracle_built, case):
    """pair build full/bin/atomonly, orthogonal box, per-type-pair cutoffs, owned rows only"""
    lmp = S.make_aeam_system(S.oracle_plugin("aeam"), case["cells"], si_fraction=case["si"], displace=case["displace"])
    lmp.setup(1, 2)
    check_lists(ctx, lmp, ghost_rows=0)
    lmp.close()


def test_device_list_feeds_compute(ctx, oracle_built):
    """forces from a device-built list == forces from the uploaded host list"""
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), (2, 1, 1), displace=0.2)
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
    ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
    f0, e0, v0 = ctx.rebomos_compute(snap["nlocal"], snap["nghost"], snap["x"], snap["type"], snap["tag"])
    ctx.neigh_build(lmp.b200_box(), 2, lmp.cutneighsq(False), lmp.cutneighsq(True), snap["nlocal"], snap["nghost"],
                    snap["x"], snap["type"], 1, snap["skin"])
    f1, e1, v1 = ctx.rebomos_compute(snap["nlocal"], snap["nghost"], snap["x"], snap["type"], snap["tag"])
    assert S.rel_err(f1, f0) < 1e-13 and abs(e1 - e0) < 1e-10 * abs(e0)
    lmp.close()
