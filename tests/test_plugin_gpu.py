"""The drop-in claim, end to end: the SAME host application (mini LAMMPS engine) runs the SAME input with
`plugin load` pointing at the reference plugin or at the B200 plugin."""
import json
import os

import numpy as np
import pytest

import support as S

pytestmark = pytest.mark.gpu


def run_script(plugin, cmds, grid=(1, 1, 1)):
    lmp = S.MiniLmp(grid)
    lmp.command("plugin load " + plugin)
    lmp.commands(cmds)
    return lmp


@pytest.mark.parametrize("grid", [(1, 1, 1), (2, 2, 1)])
@pytest.mark.parametrize("neigh", ["device", "host"])
def test_in_rebomos_bulk_through_the_b200_plugin(oracle_built, grid, neigh, monkeypatch):
    """shipped input, B200 plugin: thermo equals log.rebomos-bulk.1 / .4 to all printed digits"""
    monkeypatch.setenv("B200MD_NEIGH", neigh)
    gold = json.load(open(os.path.join(S.GOLDEN, "log_rebomos_bulk.json")))["log.rebomos-bulk.1"]
    pot = os.path.join(S.potential_dir(), "MoS.REBO.set5b")
    cmds = [("pair_coeff * * %s M S" % pot) if c.startswith("pair_coeff") else c for c in S.input_script("in.rebomos-bulk")]
    lmp = run_script(S.B200_REBOMOS_SO, cmds, grid)
    for r, g in zip(lmp.thermo(), gold["thermo"]):
        assert (S.fmt8(r["temp"]), S.fmt8(r["press"]), S.fmt8(r["pe"]), S.fmt8(r["ke"])) == \
               (S.fmt8(g[1]), S.fmt8(g[2]), S.fmt8(g[3]), S.fmt8(g[4]))
    lmp.close()


@pytest.mark.parametrize("pin", ["0", "1"])
@pytest.mark.parametrize("overwrite", ["0", "1"])
def test_plugin_force_return_modes(oracle_built, overwrite, pin, monkeypatch):
    """The host class lets the library WRITE atom->f when it is known to be zero on entry (top-level pair style, no
    pre_force fix; B200MD_F_OVERWRITE=1, the default) and ADDS otherwise (forced here with =0): same golden rows."""
    monkeypatch.setenv("B200MD_F_OVERWRITE", overwrite)
    monkeypatch.setenv("B200MD_PIN_HOST", pin)          # page-locking of the host application's x/f blocks is opt-in
    gold = json.load(open(os.path.join(S.GOLDEN, "log_rebomos_bulk.json")))["log.rebomos-bulk.1"]
    pot = os.path.join(S.potential_dir(), "MoS.REBO.set5b")
    cmds = [("pair_coeff * * %s M S" % pot) if c.startswith("pair_coeff") else c for c in S.input_script("in.rebomos-bulk")]
    lmp = run_script(S.B200_REBOMOS_SO, cmds)
    for r, g in zip(lmp.thermo(), gold["thermo"]):
        assert (S.fmt8(r["temp"]), S.fmt8(r["press"]), S.fmt8(r["pe"]), S.fmt8(r["ke"])) == \
               (S.fmt8(g[1]), S.fmt8(g[2]), S.fmt8(g[3]), S.fmt8(g[4]))
    lmp.close()


@pytest.mark.parametrize("grid", [(1, 1, 1), (2, 1, 2)])
def test_aeam_sample_b200_plugin_vs_reference_plugin(oracle_built, grid):
    """sample.in-like run (fcc Al + Si, 863 K, NVE): B200 plugin tracks the reference plugin"""
    cmds = S.aeam_commands((6, 6, 6), 0.02) + ["velocity all create 863.0 1082337", "fix 1 all nve", "thermo 10", "run 30"]
    ref = run_script(S.oracle_plugin("aeam"), cmds, grid)
    got = run_script(S.B200_AEAM_SO, cmds, grid)
    assert got.get_int("nbuild") == ref.get_int("nbuild")
    for r, g in zip(got.thermo(), ref.thermo()):
        assert r["step"] == g["step"]
        assert abs(r["pe"] - g["pe"]) < 1e-9 * abs(g["pe"])
        assert abs(r["temp"] - g["temp"]) < 1e-7 * max(g["temp"], 1.0)
        assert abs(r["press"] - g["press"]) < 1e-6 * max(abs(g["press"]), 1.0)
    ref.close()
    got.close()


@pytest.mark.parametrize("grid", [(2, 2, 2)])
def test_sample_in_as_shipped_b200_plugin_vs_reference_plugin(oracle_built, grid):
    """BASELINE configs[1]: USER-AEAM/sample.in AS SHIPPED -- 20^3 fcc cells = 32 000 atoms, 0.75 % Si through
    `set type/fraction ... 7683797`, `velocity all create 863.0 1082337` (LAMMPS' RanPark streams), `fix nvt`
    (Nose-Hoover chain), 400 steps, thermo 100 -- run literally by the engine with the reference plugin and with the
    B200 plugin (only the potential path is rewritten).  The reference ships no log for it; the check is that
    switching the plugin changes nothing: same rebuild count, thermo rows equal to 1e-8."""
    pot = os.path.join(S.potential_dir(), "AlSi.aeam")
    cmds = [("pair_coeff * * %s Al Si" % pot) if c.startswith("pair_coeff") else c for c in S.input_script("sample.in")]
    ref = run_script(S.oracle_plugin("aeam"), cmds, grid)
    got = run_script(S.B200_AEAM_SO, cmds, grid)
    assert ref.get_int("natoms") == 32000
    assert got.get_int("nbuild") == ref.get_int("nbuild")
    rr, gg = ref.thermo(), got.thermo()
    assert [r["step"] for r in rr] == [0, 100, 200, 300, 400] == [r["step"] for r in gg]
    assert abs(rr[0]["temp"] - 863.0) < 1e-9
    for r, g in zip(gg, rr):
        for key in ("temp", "pe", "etotal", "press"):
            assert abs(r[key] - g[key]) < 1e-8 * max(abs(g[key]), 1.0), (r["step"], key, r[key], g[key])
    print("sample.in thermo (step temp etotal pe press):",
          [(g["step"], round(g["temp"], 3), round(g["etotal"], 4), round(g["pe"], 4), round(g["press"], 2)) for g in rr])
    ref.close()
    got.close()


def test_forces_lockstep_along_reference_trajectory(ctx, oracle_built):
    """lock-step parity (chaotic divergence makes free-running comparison meaningless):
    the reference drives the trajectory; at every sampled step the CUDA path is fed the reference's own
    positions, ghosts and neighbor list and must return forces within 1e-10 relative, energy within 1e-12."""
    a = S.make_rebomos_system(S.oracle_plugin("rebomos"), (2, 1, 1),
                              extra=["velocity all create 1200.0 777", "fix 1 all nve", "neighbor 0.8 bin"])
    ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
    worst = 0.0
    for step in range(6):
        a.command("run 7")                       # advances 7 steps (rebuilding lists when needed)
        snap = S.snapshot(a)
        a.compute(1, 2, reverse=True)
        nl = snap["nlocal"]
        f_ref, e_ref = a.f()[:nl].copy(), a.get_double("eng_vdwl")
        ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
        f, e, v = ctx.rebomos_compute(nl, snap["nghost"], snap["x"], snap["type"], snap["tag"])
        f = S.fold_ghost_forces(f, snap["swaps"], nl)
        worst = max(worst, S.rel_err(f, f_ref))
        assert S.rel_err(f, f_ref) < 1e-10
        assert abs(e - e_ref) < 1e-12 * abs(e_ref)
    print("worst relative force error along the trajectory: %.3e" % worst)
    a.close()


@pytest.mark.parametrize("style", ["rebomos", "aeam"])
def test_per_atom_tallies_through_the_plugin(oracle_built, style):
    """compute pe/atom / stress/atom path: the host application asks the pair style for ENERGY_ATOM | VIRIAL_ATOM;
    the B200 plugin fills Pair::eatom / Pair::vatom like the reference plugin does (owned + ghost entries)."""
    out = {}
    for which in ("ref", "b200"):
        if style == "rebomos":
            so = S.oracle_plugin("rebomos") if which == "ref" else S.B200_REBOMOS_SO
            lmp = S.make_rebomos_system(so, (2, 1, 1), displace=0.2)
        else:
            so = S.oracle_plugin("aeam") if which == "ref" else S.B200_AEAM_SO
            lmp = S.make_aeam_system(so, (5, 5, 5), si_fraction=0.1, displace=0.2)
        lmp.setup(1, 2)
        lmp.compute(1 | 2, 2 | 4, reverse=False)
        nl, nall = lmp.get_int("nlocal"), lmp.nall()
        swaps = lmp.swaps()
        ea = lmp._arr("eatom", 0, nall, np.float64).copy()
        va = lmp._arr("vatom", 0, nall, np.float64, 6).copy()
        for s in reversed(swaps):           # what compute pe/atom does: reverse-communicate the ghost shares
            if s["recvnum"]:
                np.add.at(ea, s["sendlist"], ea[s["firstrecv"]:s["firstrecv"] + s["recvnum"]])
                np.add.at(va, s["sendlist"], va[s["firstrecv"]:s["firstrecv"] + s["recvnum"]])
        out[which] = (ea[:nl], va[:nl], lmp.get_double("eng_vdwl"))
        lmp.close()
    (ea0, va0, e0), (ea1, va1, e1) = out["ref"], out["b200"]
    assert abs(e1 - e0) < 1e-12 * abs(e0)
    assert S.rel_err(ea1, ea0) < 1e-10 and S.rel_err(va1, va0) < 1e-10
    assert abs(ea1.sum() - ea0.sum()) < 1e-10 * abs(e0)
    if style == "rebomos":
        # the reference's REBOMoS tallies are complete: per-atom energies add up to eng_vdwl.  Its AEAM tallies are
        # not (pair_aeam.cpp:295-300 gives angular atoms F/3, :393 tallies the pair term with evdwl = 0 ...), so for
        # aeam the check is "the same per-atom values as the reference", asserted above.
        assert abs(ea1.sum() - e0) < 1e-10 * abs(e0)
