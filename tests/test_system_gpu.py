"""GPU-resident MD system (b200md_system_*): ghosts, lists, NVE loop and thermo entirely on the device,
checked against (a) the golden thermo table of log.rebomos-bulk.1 and (b) the oracle engine."""
import json
import os

import numpy as np
import pytest

import lammps_plugins_b200 as b2
import support as S

pytestmark = pytest.mark.gpu


def start_system(ctx, lmp, style, sort_every=1000):
    """hand the engine's owned atoms (as created, before any setup) to the device system"""
    nl = lmp.get_int("nlocal")
    assert lmp.get_int("nghost") == 0
    d = lmp.box()
    box = b2.make_box(d["boxlo"], d["boxhi"], d["xy"], d["xz"], d["yz"], triclinic=d["triclinic"])
    ctx.system_create(style, lmp.get_int("ntypes"), lmp.mass(), box, lmp.x(0, nl).copy(), lmp.v().copy(),
                      lmp.type()[:nl].copy(), lmp.tag()[:nl].copy(), lmp.get_double("skin"), lmp.get_double("dt"),
                      lmp.units(), sort_every=sort_every)


def aeam_tables():
    t = S.load_aeam_fixture()
    return {k: t[k] for k in ("nelements", "nnonangular", "nrho", "drho", "nr", "dr", "cut", "frho", "rhor", "z2r")}


def test_golden_log_rebomos_bulk(ctx, oracle_built):
    """in.rebomos-bulk run on the device reproduces log.rebomos-bulk.1:54-56 to all 8 printed digits,
    with 4285 ghosts and 0 neighbor list builds in 20 steps."""
    gold = json.load(open(os.path.join(S.GOLDEN, "log_rebomos_bulk.json")))["log.rebomos-bulk.1"]
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"))
    ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
    start_system(ctx, lmp, "rebomos")
    ctx.system_run(20, 10)
    rows = ctx.system_thermo_rows()
    assert [r["step"] for r in rows] == [0, 10, 20]
    for r, g in zip(rows, gold["thermo"]):
        print(r["step"], S.fmt8(r["temp"]), S.fmt8(r["press"]), S.fmt8(r["pe"]), S.fmt8(r["ke"]))
        assert S.fmt8(r["temp"]) == S.fmt8(g[1])
        assert S.fmt8(r["press"]) == S.fmt8(g[2])
        assert S.fmt8(r["pe"]) == S.fmt8(g[3])
        assert S.fmt8(r["ke"]) == S.fmt8(g[4])
        assert S.fmt8(r["vol"]) == S.fmt8(g[6])
    sz = ctx.system_sizes()
    assert sz["nlocal"] == 288 and sz["nghost"] == int(gold["nghost_ave_max_min"][0])
    assert sz["nbuild"] == gold["builds"] == 0
    lmp.close()


@pytest.mark.parametrize("style,rep", [("rebomos", (2, 2, 1)), ("aeam", (5, 4, 6))])
def test_setup_state_matches_engine_bitwise(ctx, oracle_built, style, rep):
    """After setup the device holds the same atoms in the same order as the host engine: owned order after
    Atom::sort, ghost order after CommBrick::borders, coordinates bit-identical (incl. the lamda round trip)."""
    if style == "rebomos":
        lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), rep, displace=0.1)
        ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
    else:
        lmp = S.make_aeam_system(S.oracle_plugin("aeam"), rep, si_fraction=0.05, displace=0.1)
        ctx.aeam_init(aeam_tables())
    start_system(ctx, lmp, style)
    got = ctx.system_download()
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    assert got["nlocal"] == snap["nlocal"] and got["nghost"] == snap["nghost"]
    assert np.array_equal(got["tag"], snap["tag"])
    assert np.array_equal(got["type"], snap["type"])
    assert np.array_equal(got["x"], snap["x"])          # bit-exact
    # forces after reverse comm, energy, pressure
    lmp.compute(1, 2, reverse=True)
    nl = snap["nlocal"]
    assert S.rel_err(got["f"][:nl], lmp.f()[:nl]) < 1e-10
    row = ctx.system_thermo_rows()[0]
    ref = lmp.thermo()[0]
    assert abs(row["pe"] - ref["pe"]) < 1e-12 * abs(ref["pe"])
    assert abs(row["press"] - ref["press"]) < 1e-9 * max(1.0, abs(ref["press"]))
    lmp.close()


@pytest.mark.parametrize("style", ["rebomos", "aeam"])
def test_nve_run_with_reneighboring_tracks_engine(ctx, oracle_built, style):
    """A hot NVE run that rebuilds its lists several times: same number of rebuilds as the host engine,
    thermo agreeing to 1e-7 relative (trajectories diverge slowly through rounding), energy conserved."""
    steps = 60
    if style == "rebomos":
        lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), (2, 1, 1),
                                    extra=["velocity all create 2000.0 4928459", "neighbor 0.5 bin"])
        ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
    else:
        lmp = S.make_aeam_system(S.oracle_plugin("aeam"), (5, 5, 5), si_fraction=0.02,
                                 extra=["velocity all create 2000.0 1082337", "neighbor 0.4 bin"])
        ctx.aeam_init(aeam_tables())
    start_system(ctx, lmp, style)
    ctx.system_run(steps, 20)
    rows = ctx.system_thermo_rows()
    lmp.commands(["thermo 20", "fix 1 all nve", "run %d" % steps])
    ref = lmp.thermo()
    sz = ctx.system_sizes()
    print(style, "builds", sz["nbuild"], lmp.get_int("nbuild"), "dangerous", sz["ndanger"])
    assert sz["nbuild"] == lmp.get_int("nbuild") and sz["nbuild"] > 0
    for r, g in zip(rows, ref):
        print(r["step"], r["temp"], g["temp"], r["pe"], g["pe"], r["press"], g["press"])
        assert r["step"] == g["step"]
        assert abs(r["pe"] - g["pe"]) < 1e-7 * abs(g["pe"])
        assert abs(r["ke"] - g["ke"]) < 1e-6 * max(abs(g["ke"]), 1.0)
    e0 = rows[0]["pe"] + rows[0]["ke"]
    e1 = rows[-1]["pe"] + rows[-1]["ke"]
    eref = ref[-1]["pe"] + ref[-1]["ke"]
    assert abs((e1 - e0) - (eref - (ref[0]["pe"] + ref[0]["ke"]))) < 1e-6 * abs(e0)
    lmp.close()


@pytest.mark.parametrize("style", ["rebomos", "aeam"])
def test_two_level_list_refresh_matches_single_level(ctx, oracle_built, style):
    """default inner margin = skin/2: inner lists are re-derived from the master list whenever an atom moved more
    than margin/2, master rebuilds stay on LAMMPS' schedule; the trajectory equals the run whose inner lists
    carry the full skin (no refresh) to rounding"""
    if style == "rebomos":
        lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), (2, 1, 1), extra=["velocity all create 1500.0 4928459"])
        ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
    else:
        lmp = S.make_aeam_system(S.oracle_plugin("aeam"), (5, 5, 5), si_fraction=0.02,
                                 extra=["velocity all create 1500.0 1082337"])
        ctx.aeam_init(aeam_tables())
    out = {}
    for label, margin in (("full", int(1000 * lmp.get_double("skin"))), ("half", 0)):
        ctx.set_option("margin", margin)
        start_system(ctx, lmp, style)
        ctx.system_run(80, 20)
        out[label] = (ctx.system_thermo_rows(), ctx.system_sizes(), ctx.system_download())
    ctx.set_option("margin", 0)
    (ra, sa, da), (rb, sb, db) = out["full"], out["half"]
    print(style, "builds", sa["nbuild"], sb["nbuild"], "inner refreshes", sa["ninner"], sb["ninner"])
    assert sa["ninner"] == 0 and sb["ninner"] > 0 and sa["nbuild"] == sb["nbuild"]
    for q, g in zip(rb, ra):
        assert abs(q["pe"] - g["pe"]) < 1e-11 * abs(g["pe"]) and abs(q["ke"] - g["ke"]) < 1e-9 * max(g["ke"], 1.0)
        assert abs(q["press"] - g["press"]) < 1e-8 * max(abs(g["press"]), 1.0)
    assert S.rel_err(db["f"][:db["nlocal"]], da["f"][:da["nlocal"]]) < 1e-9
    lmp.close()


def test_three_level_list_matches_two_level(ctx, oracle_built):
    """third list level (rebomos): the force kernels stream tight rows (rcut + margin_tight) that are re-derived from the
    inner rows -- not from the master rows -- whenever an atom moved margin_tight/2.  A hot run with the level switched
    on (several tight derives, inner refreshes and master rebuilds) equals the run without it to rounding."""
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), (2, 2, 1), extra=["velocity all create 1500.0 4928459"])
    ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
    out = {}
    for label, mt in (("off", 0), ("on", 300)):
        ctx.set_option("margin_tight", mt)
        t0 = ctx.counter("tight_refreshes")
        start_system(ctx, lmp, "rebomos")
        ctx.system_run(150, 25)
        out[label] = (ctx.system_thermo_rows(), ctx.system_sizes(), ctx.system_download(), ctx.counter("tight_refreshes") - t0)
    ctx.set_option("margin_tight", 400)
    (ra, sa, da, ta), (rb, sb, db, tb) = out["off"], out["on"]
    print("builds", sa["nbuild"], sb["nbuild"], "inner refreshes", sa["ninner"], sb["ninner"], "tight derives", ta, tb)
    assert ta == 0 and tb > sb["nbuild"] + sb["ninner"] + 1      # derived on its own schedule, not only after rebuilds
    assert sa["nbuild"] == sb["nbuild"] and sa["ninner"] == sb["ninner"] and sb["ninner"] > 0
    for q, g in zip(rb, ra):
        assert abs(q["pe"] - g["pe"]) < 1e-11 * abs(g["pe"]) and abs(q["ke"] - g["ke"]) < 1e-9 * max(g["ke"], 1.0)
        assert abs(q["press"] - g["press"]) < 1e-8 * max(abs(g["press"]), 1.0)
    assert S.rel_err(db["f"][:db["nlocal"]], da["f"][:da["nlocal"]]) < 1e-9
    assert np.array_equal(db["tag"][:db["nlocal"]], da["tag"][:da["nlocal"]])
    lmp.close()


@pytest.mark.parametrize("style", ["rebomos", "aeam"])
def test_run_loop_options_do_not_change_the_trajectory(ctx, oracle_built, style):
    """round-2 run-loop changes -- the second half kick riding in the next step's integrate launch ("fuse_integrate"),
    the reneighbor vote through a mapped host word instead of copy + synchronise ("peer_vote"), master rebuilds in one
    stencil walk over clipped, bin-sorted candidates ("one_pass_neigh"), the self halos as one gather through a ghost ->
    owned-atom map ("flat_halo"; the forward halos -- the reverse fold stays staged in deterministic mode) -- reorder no arithmetic: in deterministic mode
    (no atomics anywhere) a hot run with rebuilds is BITWISE the same with each of them off."""
    if style == "rebomos":
        lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), (2, 2, 1),
                                    extra=["velocity all create 2000.0 4928459", "neighbor 0.5 bin"])
        ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
    else:
        lmp = S.make_aeam_system(S.oracle_plugin("aeam"), (5, 5, 5), si_fraction=0.02,
                                 extra=["velocity all create 2000.0 1082337", "neighbor 0.4 bin"])
        ctx.aeam_init(aeam_tables())
    ctx.set_option("deterministic", 1)
    runs = {}
    try:
        for label, opts in (("default", {}), ("plain", {"fuse_integrate": 0, "peer_vote": 0, "one_pass_neigh": 0, "flat_halo": 0})):
            for k, v in opts.items():
                ctx.set_option(k, v)
            start_system(ctx, lmp, style)
            ctx.system_run(90, 15)
            runs[label] = (ctx.system_thermo_rows(), ctx.system_sizes(), ctx.system_download())
    finally:
        for k in ("fuse_integrate", "peer_vote", "one_pass_neigh", "flat_halo"):
            ctx.set_option(k, 1)
        ctx.set_option("deterministic", 0)
    (ra, sa, da), (rb, sb, db) = runs["default"], runs["plain"]
    print(style, "builds", sa["nbuild"], sb["nbuild"])
    assert sa["nbuild"] == sb["nbuild"] and sa["nbuild"] >= 1
    for q, g in zip(ra, rb):
        assert q["pe"] == g["pe"] and q["ke"] == g["ke"] and q["press"] == g["press"], (q, g)
    for key in ("x", "v", "f"):
        assert np.array_equal(da[key][:da["nlocal"]], db[key][:db["nlocal"]]), key
    assert np.array_equal(da["tag"][:da["nlocal"]], db["tag"][:db["nlocal"]])
    lmp.close()


@pytest.mark.parametrize("tstop", [863.0, 500.0], ids=["constant-T", "ramp"])
def test_device_nvt_tracks_engine(ctx, oracle_built, tstop):
    """`fix nvt` in the GPU-resident loop (USER-AEAM/sample.in:23, Nose-Hoover chain restated from FixNH) against the
    engine's fix nvt driving the reference plugin: sample.in on 8^3 cells, 200 steps -- thermo rows and the thermostat's
    share of the conserved quantity agree to 1e-8; also with a temperature ramp (Tstart != Tstop)."""
    pot = os.path.join(S.potential_dir(), "AlSi.aeam")
    lmp = S.MiniLmp((1, 1, 1))
    lmp.command("plugin load " + S.oracle_plugin("aeam"))
    for c in S.input_script("sample.in"):
        w = c.split()
        if w[0] == "run":
            continue
        if w[0] == "region":
            c = "region MeSi block 0 8 0 8 0 8"
        if w[0] == "pair_coeff":
            c = "pair_coeff * * %s Al Si" % pot
        if w[0] == "fix":
            c = "fix 1 all nvt temp 863.0 %g 0.1" % tstop
        if w[0] == "thermo":
            c = "thermo 50"
        lmp.command(c)
    ctx.aeam_init(aeam_tables())
    start_system(ctx, lmp, "aeam")
    ctx.system_set_nvt(863.0, tstop, 0.1)
    ctx.system_run(200, 50)
    rows = ctx.system_thermo_rows()
    nh = ctx.system_nh_energy()
    ctx.system_set_nvt(0.0, 0.0, 0.0)
    lmp.command("run 200")
    ref = lmp.thermo()
    assert [r["step"] for r in rows] == [g["step"] for g in ref] == [0, 50, 100, 150, 200]
    for r, g in zip(rows, ref):
        print(r["step"], r["temp"], g["temp"], r["pe"], g["pe"])
        assert abs(r["temp"] - g["temp"]) < 1e-8 * max(g["temp"], 1.0)
        assert abs(r["pe"] - g["pe"]) < 1e-9 * abs(g["pe"])
        assert abs(r["press"] - g["press"]) < 1e-7 * max(abs(g["press"]), 1.0)
    assert abs(rows[0]["temp"] - 863.0) < 1e-9 and abs(rows[-1]["temp"] - 863.0) > 50.0    # the run is not trivial
    assert abs(nh - lmp.get_double("nh_energy")) < 1e-8 * max(abs(nh), 1.0)
    lmp.close()


def test_energy_drift_1000_nve_steps_matches_reference(ctx, oracle_built):
    """north star: "energy drift over 1000 NVE steps matching the reference's".  The shipped 288-atom cell at 300 K,
    dt = 1 fs, 1000 steps on the device and in the engine with the reference plugin.  Early on the two runs are the same
    trajectory (thermo equal to 1e-9); over the whole run the total-energy excursion and the end-to-end drift of the
    device run equal the reference's (same integrator error, no extra energy source or sink)."""
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), extra=["velocity all create 300.0 4928459"])
    ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
    start_system(ctx, lmp, "rebomos")
    ctx.system_run(1000, 50)
    rows = ctx.system_thermo_rows()
    lmp.commands(["thermo 50", "fix 1 all nve", "run 1000"])
    ref = lmp.thermo()
    assert [r["step"] for r in rows] == [g["step"] for g in ref] and len(rows) == 21
    et = np.array([r["pe"] + r["ke"] for r in rows])
    er = np.array([g["pe"] + g["ke"] for g in ref])
    ke_scale = max(g["ke"] for g in ref)
    for r, g in zip(rows[:5], ref[:5]):          # first 200 steps: same trajectory
        assert abs(r["pe"] - g["pe"]) < 1e-9 * abs(g["pe"]) and abs(r["ke"] - g["ke"]) < 1e-7 * ke_scale
    exc_t, exc_r = np.abs(et - et[0]).max(), np.abs(er - er[0]).max()
    drift_t, drift_r = et[-1] - et[0], er[-1] - er[0]
    print("max excursion  device %.3e  reference %.3e eV;  end-to-end drift  device %.3e  reference %.3e eV;  KE %.3f eV"
          % (exc_t, exc_r, drift_t, drift_r, ke_scale))
    assert exc_r < 1e-2 * ke_scale                      # the reference itself: 0.5 % of KE at dt = 1 fs (measured 5.3e-2 eV)
    assert abs(exc_t - exc_r) < 0.05 * exc_r + 1e-9
    assert abs(drift_t - drift_r) < 0.05 * exc_r + 1e-9
    assert np.abs(et - er).max() < 0.05 * exc_r + 1e-9  # the whole E_tot(t) curve lies on the reference's
    lmp.close()
