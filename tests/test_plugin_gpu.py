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


def test_forces_per_step_lockstep(oracle_built):
    """lock-step: at every step of a reference-driven trajectory the B200 plugin, fed the same positions,
    returns forces within 1e-10 relative"""
    a = S.make_rebomos_system(S.oracle_plugin("rebomos"), (2, 1, 1), extra=["velocity all create 600.0 777", "fix 1 all nve"])
    b = S.make_rebomos_system(S.B200_REBOMOS_SO, (2, 1, 1), extra=["velocity all create 600.0 777", "fix 1 all nve"])
    a.setup(1, 2)
    b.setup(1, 2)
    nl = a.get_int("nlocal")
    for step in range(5):
        a.command("run 4")
        # copy the reference trajectory's state into the B200-driven engine and recompute there
        b.x()[:nl] = a.x()[:nl]
        b.forward_comm()
        b.compute(1, 2, reverse=True)
        a.compute(1, 2, reverse=True)
        assert S.rel_err(b.f()[:nl], a.f()[:nl]) < 1e-10
        assert abs(b.get_double("eng_vdwl") - a.get_double("eng_vdwl")) < 1e-12 * abs(a.get_double("eng_vdwl"))
    a.close()
    b.close()
