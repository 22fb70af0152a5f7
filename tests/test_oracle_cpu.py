"""CPU tests of the oracle: the reference pair styles compiled verbatim (oracle/_ref) inside the mini
LAMMPS engine reproduce every golden number the reference ships (log.rebomos-bulk.1 / .4), and the plain
restatement (oracle/port) agrees with the verbatim build.  No GPU involved."""
import json
import os

import numpy as np
import pytest

import support as S

GOLD = json.load(open(os.path.join(S.GOLDEN, "log_rebomos_bulk.json")))

needs_ref = pytest.mark.skipif(not os.path.exists(S.REF_REBOMOS_SO), reason="oracle/_ref not built (no reference tree)")


def run_bulk(plugin, grid):
    lmp = S.MiniLmp(grid)
    lmp.command("plugin load " + plugin)
    pot = os.path.join(S.potential_dir(), "MoS.REBO.set5b")
    for c in S.input_script("in.rebomos-bulk"):       # the shipped input, line by line
        if c.startswith("pair_coeff"):
            c = "pair_coeff * * %s M S" % pot
        lmp.command(c)
    return lmp


@pytest.mark.parametrize("which,grid", [("1", (1, 1, 1)), ("4", (2, 2, 1))])
@pytest.mark.parametrize("impl", ["ref", "port"])
def test_log_goldens(oracle_built, which, grid, impl):
    plugin = S.REF_REBOMOS_SO if impl == "ref" else S.PORT_SO
    if not os.path.exists(plugin):
        pytest.skip(plugin + " not built")
    g = GOLD["log.rebomos-bulk." + which]
    assert tuple(g["procgrid"]) == grid
    lmp = run_bulk(plugin, grid)
    rows = lmp.thermo()
    assert len(rows) == len(g["thermo"]) == 3
    for r, ref in zip(rows, g["thermo"]):
        assert r["step"] == int(ref[0])
        for key, col in (("temp", 1), ("press", 2), ("pe", 3), ("ke", 4), ("vol", 6)):
            assert S.fmt8(r[key]) == S.fmt8(ref[col]), (key, r[key], ref[col])
    nprocs = grid[0] * grid[1] * grid[2]
    nlocal = [lmp.get_int("nlocal", r) for r in range(nprocs)]
    nghost = [lmp.get_int("nghost", r) for r in range(nprocs)]
    full = [int(lmp.neigh_csr(r)[0][lmp.get_int("inum", r)]) for r in range(nprocs)]
    assert max(nlocal) == g["nlocal_ave_max_min"][1] and min(nlocal) == g["nlocal_ave_max_min"][2]
    assert max(nghost) == g["nghost_ave_max_min"][1] and min(nghost) == g["nghost_ave_max_min"][2]
    assert abs(np.mean(nghost) - g["nghost_ave_max_min"][0]) < 1e-9
    assert sum(full) == g["total_neighbors"] and abs(np.mean(full) - g["fullnghs_ave"]) < 1e-9
    assert lmp.get_int("nbuild") == g["builds"]
    lmp.close()


@needs_ref
@pytest.mark.parametrize("style", ["rebomos", "aeam"])
def test_port_matches_verbatim_reference(oracle_built, style):
    """same inputs, same engine: forces on owned AND ghost atoms (before reverse comm), energy, both virial
    flavours (fdotr and explicit tally) agree to rounding."""
    if not os.path.exists(S.PORT_SO):
        pytest.skip("port not built")
    out = []
    for plugin in (S.oracle_plugin(style), S.PORT_SO):
        if style == "rebomos":
            lmp = S.make_rebomos_system(plugin, (2, 1, 1), displace=0.3)
        else:
            lmp = S.make_aeam_system(plugin, (5, 5, 5), si_fraction=0.15, displace=0.25)
        lmp.setup(1, 2)
        lmp.compute(1, 2)
        f = lmp.f().copy()
        e = lmp.get_double("eng_vdwl")
        v = np.array([lmp.get_double("virial%d" % k) for k in range(6)])
        lmp.compute(1, 1)        # VIRIAL_PAIR: explicit ev_tally / v_tally path
        v2 = np.array([lmp.get_double("virial%d" % k) for k in range(6)])
        out.append((f, e, v, v2))
        lmp.close()
    (f0, e0, v0, w0), (f1, e1, v1, w1) = out
    assert S.rel_err(f1, f0) < 1e-12
    assert abs(e1 - e0) < 1e-13 * abs(e0)
    assert S.rel_err(v1, v0) < 1e-11 and S.rel_err(w1, w0) < 1e-11
    assert S.rel_err(w0, v0) < 1e-9          # tally virial == fdotr virial


def test_engine_rejects_what_lammps_rejects(oracle_built):
    lmp = S.MiniLmp()
    lmp.command("plugin load " + S.oracle_plugin("rebomos"))
    lmp.commands(S.rebomos_bulk_commands()[:-2])       # up to create_atoms/mass, without pair_style/pair_coeff
    lmp.command("pair_style rebomos")
    with pytest.raises(S.LammpsError, match="Illegal pair_style command"):
        lmp.command("pair_style rebomos 1.0")
    lmp.command("pair_style rebomos")
    with pytest.raises(S.LammpsError, match="Incorrect args for pair coefficients"):
        lmp.command("pair_coeff * * %s M" % os.path.join(S.potential_dir(), "MoS.REBO.set5b"))
    with pytest.raises(S.LammpsError, match="Incorrect args for pair coefficients"):
        lmp.command("pair_coeff * * %s M Xx" % os.path.join(S.potential_dir(), "MoS.REBO.set5b"))
    with pytest.raises(S.LammpsError):
        lmp.command("pair_coeff * * /nonexistent/file M S")
    lmp.close()


@pytest.mark.skipif(not S.have_reference_tree(), reason="reference tree not mounted")
def test_fixture_potentials_equal_reference_files():
    """The potentials re-emitted from tests/golden parse to bit-identical doubles as the reference files."""
    ref = [float(ln.split()[0]) for ln in open(os.path.join(S.REFERENCE, "USER-REBOMOS", "MoS.REBO.set5b"))
           if ln.strip() and not ln.startswith("#")]
    mine = [float(ln.split()[0]) for ln in open(os.path.join(S.potential_dir(), "MoS.REBO.set5b"))
            if ln.strip() and not ln.startswith("#")]
    assert ref == mine and len(ref) == 61
    lines = open(os.path.join(S.REFERENCE, "USER-AEAM", "AlSi.aeam")).read().splitlines()
    mine = open(os.path.join(S.potential_dir(), "AlSi.aeam")).read().splitlines()
    assert lines[11].split() == mine[11].split()
    for a, b in zip(lines[12:18], mine[12:18]):
        assert [float(v) for v in a.split()[:3]] == [float(v) for v in b.split()[:3]]
    va = np.array([float(v) for ln in lines[18:] for v in ln.split()])
    vb = np.array([float(v) for ln in mine[18:] for v in ln.split()])
    assert np.array_equal(va, vb) and len(va) == 90000


def test_workload_generators_match_engine(oracle_built):
    from lammps_plugins_b200 import workloads as W
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), (2, 3, 2))
    n = lmp.get_int("nlocal")
    w = W.mos2_bulk(2, 3, 2)
    assert n == len(w["x"]) == 288 * 12
    assert np.array_equal(lmp.type()[:n], w["type"]) and np.array_equal(lmp.tag()[:n], w["tag"])
    b = lmp.box()
    assert np.allclose(b["boxhi"], w["boxhi"], rtol=1e-15) and abs(b["xy"] - w["xy"]) < 1e-12
    own = W.brick_owner(w["x"], w["boxlo"], w["boxhi"], w["xy"], 0.0, 0.0, (2, 2, 1))
    assert set(own) == {0, 1, 2, 3} and np.bincount(own).sum() == n
    lmp.close()
    l2 = S.MiniLmp()
    l2.command("plugin load " + S.oracle_plugin("aeam"))
    l2.commands(S.aeam_commands((3, 4, 5), 0.0))
    n = l2.get_int("nlocal")
    f = W.fcc_alsi((3, 4, 5), 0.0)
    assert n == len(f["x"]) and np.array_equal(l2.x(0, n), f["x"]) and np.array_equal(l2.tag()[:n], f["tag"])
    l2.close()


def test_sample_in_literal_nvt_conserves_the_nose_hoover_quantity(oracle_built):
    """USER-AEAM/sample.in runs LITERALLY in the engine (set type/fraction and velocity create
    with LAMMPS' RanPark streams, fix nvt = Nose-Hoover chain restated from FixNH) -- here on a 6^3-cell box so that
    the CPU suite stays short; the GPU suite runs the shipped 20^3 box with both plugins.  The reference ships no log
    for this input (parity unpinned by goldens), so the integrator is pinned by its own invariant: etotal + thermostat
    energy is conserved to O(dt^2) while the thermostat moves tens of eV in and out of the system."""
    err = {}
    for scale in (1.0, 0.5):
        lmp = S.MiniLmp((1, 1, 1))
        lmp.command("plugin load " + S.oracle_plugin("aeam"))
        pot = os.path.join(S.potential_dir(), "AlSi.aeam")
        for c in S.input_script("sample.in"):
            w = c.split()
            if w[0] == "region":
                c = "region MeSi block 0 6 0 6 0 6"
            if w[0] == "pair_coeff":
                c = "pair_coeff * * %s Al Si" % pot
            if w[0] == "thermo":
                c = "thermo %d" % int(20 / scale)
            if w[0] == "timestep":
                c = "timestep %g" % (0.001 * scale)
            if w[0] == "run":
                c = "run %d" % int(200 / scale)
            lmp.command(c)
        rows = lmp.thermo()
        assert len(rows) == 11
        assert abs(rows[0]["temp"] - 863.0) < 1e-9          # velocity create scales to exactly the requested T
        flow = rows[-1]["etotal"] - rows[0]["etotal"]        # eta = eta_dot = 0 at step 0
        err[scale] = abs(flow + lmp.get_double("nh_energy"))
        assert abs(flow) > 10.0                              # the thermostat really acts (eV)
        assert err[scale] < 2e-3 * abs(flow)
        lmp.close()
    assert 3.0 < err[1.0] / err[0.5] < 5.5                   # second-order integrator
