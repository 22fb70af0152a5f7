"""Domain decomposition on the device: the brick grid, CommBrick::exchange (migration),
borders, forward/reverse halo with REAL partner ranks.  The ranks are contexts of this process on cuda:0
driven by one host thread each (loopback transport, b200md_system_comm_init_local) -- NCCL refuses two
ranks on one device, and the round-end GPU tests run on a one-GPU box; the NCCL transport differs only in
how a buffer travels (xfer_sendrecv in system.cu) and is exercised by `bench.py --gpus N`.

Checked against the oracle engine run on the SAME grid (thread-ranks), as the reference's own
log.rebomos-bulk.1 vs .4 do for 1x1x1 vs 2x2x1."""
import json
import os
import threading

import numpy as np
import pytest

import lammps_plugins_b200 as b2
import support as S

pytestmark = pytest.mark.gpu


def aeam_tables():
    t = S.load_aeam_fixture()
    return {k: t[k] for k in ("nelements", "nnonangular", "nrho", "drho", "nr", "dr", "cut", "frho", "rhor", "z2r")}


class Ranks:
    """R device systems of one loopback group; every call runs on all ranks concurrently."""

    def __init__(self, grid, style):
        self.grid = tuple(grid)
        self.R = grid[0] * grid[1] * grid[2]
        self.style = style
        self.group = b2.lib().b200md_local_group_create(self.R)
        assert self.group > 0
        self.ctx = [b2.Context(0) for _ in range(self.R)]

    def each(self, fn):
        out = [None] * self.R
        err = []

        def work(r):
            try:
                out[r] = fn(r, self.ctx[r])
            except Exception as e:      # noqa: BLE001 - re-raised below
                err.append((r, e))

        th = [threading.Thread(target=work, args=(r,)) for r in range(self.R)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if err:
            raise AssertionError("rank %d: %r" % err[0])
        return out

    def create(self, lmp, sort_every=1000):
        """hand every engine rank's owned atoms (as created, before setup) to the matching device rank"""
        d = lmp.box()
        mass, units = lmp.mass(), lmp.units()
        ntypes, skin, dt = lmp.get_int("ntypes"), lmp.get_double("skin"), lmp.get_double("dt")
        atoms = []
        for r in range(self.R):
            nl = lmp.get_int("nlocal", r)
            atoms.append((lmp.x(r, nl).copy(), lmp.v(r).copy(), lmp.type(r)[:nl].copy(), lmp.tag(r)[:nl].copy()))

        def fn(r, c):
            if self.style == "rebomos":
                c.rebomos_init(S.rebomos_params_struct(), [0, 1])
            else:
                c.aeam_init(aeam_tables())
            c.comm_init_local(self.group, self.R, r)
            box = b2.make_box(d["boxlo"], d["boxhi"], d["xy"], d["xz"], d["yz"], triclinic=d["triclinic"])
            x, v, t, g = atoms[r]
            c.system_create(self.style, ntypes, mass, box, x, v, t, g, skin, dt, units, procgrid=self.grid, rank=r,
                            sort_every=sort_every)

        self.each(fn)

    def create_from_arrays(self, x, v, typ, tag, box_d, mass, units, ntypes, skin, dt, sort_every=1000):
        """partition one global set of atoms over the grid (launch.my_atoms, as bench.py does for N > 1)"""
        from lammps_plugins_b200 import workloads as W
        owner = W.brick_owner(x, box_d["boxlo"], box_d["boxhi"], box_d["xy"], box_d["xz"], box_d["yz"], self.grid)

        def fn(r, c):
            if self.style == "rebomos":
                c.rebomos_init(S.rebomos_params_struct(), [0, 1])
            else:
                c.aeam_init(aeam_tables())
            c.comm_init_local(self.group, self.R, r)
            box = b2.make_box(box_d["boxlo"], box_d["boxhi"], box_d["xy"], box_d["xz"], box_d["yz"],
                              triclinic=box_d["triclinic"])
            m = owner == r
            c.system_create(self.style, ntypes, mass, box, x[m], v[m], typ[m], tag[m], skin, dt, units,
                            procgrid=self.grid, rank=r, sort_every=sort_every)

        self.each(fn)

    def gather_by_tag(self, natoms):
        """owned x, v, f of all ranks indexed by atom ID - 1, plus how many ranks own each atom"""
        X, V, F = np.zeros((natoms, 3)), np.zeros((natoms, 3)), np.zeros((natoms, 3))
        cnt = np.zeros(natoms, dtype=np.int64)
        for c in self.ctx:
            st = c.system_download()
            n = st["nlocal"]
            idx = st["tag"][:n] - 1
            X[idx], V[idx], F[idx] = st["x"][:n], st["v"], st["f"][:n]
            np.add.at(cnt, idx, 1)
        return X, V, F, cnt

    def run(self, steps, thermo_every):
        self.each(lambda r, c: c.system_run(steps, thermo_every))

    def close(self):
        for c in self.ctx:
            c.close()


def test_golden_log_rebomos_bulk_4_ranks(oracle_built):
    """in.rebomos-bulk on a 2x2x1 grid of device ranks reproduces log.rebomos-bulk.4:54-56 digit for digit,
    with the log's per-rank atom and ghost counts (72 owned; 2768..2775 ghosts)."""
    gold = json.load(open(os.path.join(S.GOLDEN, "log_rebomos_bulk.json")))["log.rebomos-bulk.4"]
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), grid=(2, 2, 1))
    rk = Ranks((2, 2, 1), "rebomos")
    rk.create(lmp)
    rk.run(20, 10)
    for r in range(4):
        rows = rk.ctx[r].system_thermo_rows()
        assert [q["step"] for q in rows] == [0, 10, 20]
        for q, g in zip(rows, gold["thermo"]):
            assert (S.fmt8(q["temp"]), S.fmt8(q["press"]), S.fmt8(q["pe"]), S.fmt8(q["ke"])) == \
                   (S.fmt8(g[1]), S.fmt8(g[2]), S.fmt8(g[3]), S.fmt8(g[4]))
    sizes = [c.system_sizes() for c in rk.ctx]
    assert [s["nlocal"] for s in sizes] == [72] * 4
    ng = [s["nghost"] for s in sizes]
    assert (np.mean(ng), max(ng), min(ng)) == tuple(gold["nghost_ave_max_min"])
    assert all(s["nbuild"] == 0 and s["natoms"] == 288 for s in sizes)
    rk.close()
    lmp.close()


@pytest.mark.parametrize("style,grid,rep", [("rebomos", (2, 2, 1), (2, 2, 1)), ("rebomos", (2, 1, 2), (2, 1, 2)),
                                            ("aeam", (2, 2, 2), (6, 6, 6)), ("aeam", (3, 1, 2), (9, 5, 6))])
def test_setup_state_matches_engine_per_rank(oracle_built, style, grid, rep):
    """after setup every device rank holds the same owned + ghost atoms in the same order as the engine rank,
    coordinates bit-identical; forces after the reverse halo agree to 1e-10"""
    if style == "rebomos":
        lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), rep, grid=grid, displace=0.1)
    else:
        lmp = S.make_aeam_system(S.oracle_plugin("aeam"), rep, grid=grid, si_fraction=0.05, displace=0.1)
    rk = Ranks(grid, style)
    rk.create(lmp)
    got = [c.system_download() for c in rk.ctx]
    lmp.setup(1, 2)
    snaps = [S.snapshot(lmp, r) for r in range(rk.R)]
    lmp.compute(1, 2, reverse=True)
    fmax = max(float(np.abs(lmp.f(r)[:snaps[r]["nlocal"]]).max()) for r in range(rk.R))
    for r in range(rk.R):
        g, s = got[r], snaps[r]
        assert (g["nlocal"], g["nghost"]) == (s["nlocal"], s["nghost"])
        assert np.array_equal(g["tag"], s["tag"]) and np.array_equal(g["type"], s["type"])
        assert np.array_equal(g["x"], s["x"])
        nl = s["nlocal"]
        assert float(np.abs(g["f"][:nl] - lmp.f(r)[:nl]).max()) < 1e-10 * fmax
    row, ref = rk.ctx[0].system_thermo_rows()[0], lmp.thermo()[0]
    assert abs(row["pe"] - ref["pe"]) < 1e-12 * abs(ref["pe"])
    assert abs(row["press"] - ref["press"]) < 1e-9 * max(1.0, abs(ref["press"]))
    rk.close()
    lmp.close()


@pytest.mark.parametrize("style,grid", [("rebomos", (2, 2, 1)), ("aeam", (2, 2, 2)), ("aeam", (3, 1, 1))])
def test_nve_with_migration_tracks_engine(oracle_built, style, grid):
    """hot NVE run with a small skin: lists are rebuilt many times and atoms change owner.  Same number of
    rebuilds as the engine on the same grid, every rank ends with the engine rank's atoms IN THE SAME ORDER
    (CommBrick::exchange replayed exactly), thermo within 1e-7, no atom lost."""
    steps = 60
    if style == "rebomos":
        lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), (2, 2, 1), grid=grid,
                                    extra=["velocity all create 3000.0 4928459", "neighbor 0.5 bin"])
    else:
        lmp = S.make_aeam_system(S.oracle_plugin("aeam"), (6, 6, 6), grid=grid, si_fraction=0.02,
                                 extra=["velocity all create 3000.0 1082337", "neighbor 0.4 bin"])
    rk = Ranks(grid, style)
    rk.create(lmp)
    natoms = lmp.get_int("natoms")
    rk.run(steps, 20)
    lmp.commands(["thermo 20", "fix 1 all nve", "run %d" % steps])
    sizes = [c.system_sizes() for c in rk.ctx]
    assert sum(s["nlocal"] for s in sizes) == natoms
    moved = sum(s["nmigrated"] for s in sizes)
    print(style, grid, "builds", sizes[0]["nbuild"], lmp.get_int("nbuild"), "migrated", moved)
    assert moved > 0, "workload too cold: nothing migrated"
    assert sizes[0]["nbuild"] == lmp.get_int("nbuild") and sizes[0]["nbuild"] > 0
    for r in range(rk.R):
        g = rk.ctx[r].system_download()
        nl = lmp.get_int("nlocal", r)
        assert g["nlocal"] == nl
        assert np.array_equal(g["tag"][:nl], lmp.tag(r)[:nl]), "owned-atom order differs on rank %d" % r
    rows, ref = rk.ctx[0].system_thermo_rows(), lmp.thermo()
    for q, g in zip(rows, ref):
        assert q["step"] == g["step"]
        assert abs(q["pe"] - g["pe"]) < 1e-7 * abs(g["pe"])
        assert abs(q["ke"] - g["ke"]) < 1e-6 * max(abs(g["ke"]), 1.0)
    rk.close()
    lmp.close()


def global_state(lmp):
    nl = lmp.get_int("nlocal")
    return dict(x=lmp.x(0, nl).copy(), v=lmp.v().copy(), typ=lmp.type()[:nl].copy(), tag=lmp.tag()[:nl].copy(),
                box_d=lmp.box(), mass=lmp.mass(), units=lmp.units(), ntypes=lmp.get_int("ntypes"),
                skin=lmp.get_double("skin"), dt=lmp.get_double("dt"))


def test_thermo_invariant_across_grids(oracle_built):
    """SURVEY.md 8(e) invariance: the same atoms and velocities on 1x1x1, 2x1x1, 2x2x1 and 2x2x2 device grids
    give the same thermo output to >= 8 digits (as log.rebomos-bulk.1 vs .4)"""
    lmp = S.make_aeam_system(S.oracle_plugin("aeam"), (6, 6, 6), si_fraction=0.02,
                             extra=["velocity all create 863.0 1082337"])
    st = global_state(lmp)
    lmp.close()
    tables = {}
    for grid in [(1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2)]:
        rk = Ranks(grid, "aeam")
        rk.create_from_arrays(**st)
        rk.run(30, 10)
        tables[grid] = rk.ctx[0].system_thermo_rows()
        rk.close()
    base = tables[(1, 1, 1)]
    for grid, rows in tables.items():
        for q, g in zip(rows, base):
            assert S.fmt8(q["pe"]) == S.fmt8(g["pe"]) and S.fmt8(q["temp"]) == S.fmt8(g["temp"]), grid
            assert abs(q["press"] - g["press"]) < 1e-8 * max(abs(g["press"]), 1.0), grid


@pytest.mark.parametrize("style", ["rebomos", "aeam"])
def test_lockstep_1_rank_vs_4_ranks_through_migration(oracle_built, style):
    """the same hot system on 1 and on 2x2x1 device ranks, advanced step by step: forces of every atom agree
    to 1e-10 at every step -- before, at and after the reneighboring steps at which atoms change owner --
    every atom is owned exactly once, and the two runs rebuild on the same steps"""
    if style == "rebomos":
        lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), (2, 2, 1),
                                    extra=["velocity all create 3000.0 4928459", "neighbor 0.5 bin"])
    else:
        lmp = S.make_aeam_system(S.oracle_plugin("aeam"), (6, 6, 6), si_fraction=0.02,
                                 extra=["velocity all create 3000.0 1082337", "neighbor 0.4 bin"])
    st = global_state(lmp)
    natoms = len(st["tag"])
    lmp.close()
    a, b = Ranks((1, 1, 1), style), Ranks((2, 2, 1), style)
    a.create_from_arrays(**st)
    b.create_from_arrays(**st)
    worst = 0.0
    for step in range(40):
        a.run(1, 0)
        b.run(1, 0)
        Xa, Va, Fa, ca = a.gather_by_tag(natoms)
        Xb, Vb, Fb, cb = b.gather_by_tag(natoms)
        assert (cb == 1).all(), "an atom is owned by %s ranks at step %d" % (set(cb.tolist()), step)
        assert a.ctx[0].system_sizes()["nbuild"] == b.ctx[0].system_sizes()["nbuild"]
        worst = max(worst, float(np.abs(Fa - Fb).max() / np.abs(Fa).max()), float(np.abs(Va - Vb).max()))
    moved = sum(c.system_sizes()["nmigrated"] for c in b.ctx)
    print(style, "migrated", moved, "builds", b.ctx[0].system_sizes()["nbuild"], "worst", worst)
    assert moved > 0 and worst < 1e-10
    a.close()
    b.close()
