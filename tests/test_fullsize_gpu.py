"""Parity at BASELINE.json's FULL sizes (configs[2]: 995 904 MoS2 atoms; configs[3]: 2 048 000 Al/Si atoms), where the
oracle cannot run in test time, through size-independent properties of the domain:

* intensive quantities of a perfect replicated crystal equal those of the shipped cell, whose values are pinned by the
  reference's golden log (PotEng/atom and Press of log.rebomos-bulk.1 step 0) or by the known answer of the potential
  file (perfect fcc Al);
* Newton's third law: the forces of a closed periodic system sum to zero, also for a thermally disordered one;
* translation invariance: shifting every atom by the same vector (re-wrapped into the box, which changes the atom
  ordering, the ghost set and every neighbor row) changes neither the energy nor any atom's force;
* the plugin-mode entry point (host buffers, pipelined transfers) and the GPU-resident loop agree on the same atoms.
"""
import json
import os

import numpy as np
import pytest

import lammps_plugins_b200 as b2
import support as S
from lammps_plugins_b200 import workloads as W

pytestmark = pytest.mark.gpu


def make_system(ctx, w, style, x=None):
    box = b2.make_box(w["boxlo"], w["boxhi"], w["xy"], w["xz"], w["yz"], triclinic=w["triclinic"])
    xx = w["x"] if x is None else x
    ctx.system_create(style, w["ntypes"], w["mass"], box, xx, np.zeros_like(xx), w["type"], w["tag"], w["skin"], 0.001,
                      b2.METAL_UNITS, sort_every=1000)
    ctx.system_run(0, 1)
    return ctx.system_thermo_rows()[0]


def forces_by_tag(ctx):
    st = ctx.system_download()
    nl = st["nlocal"]
    f = np.empty((nl, 3))
    f[st["tag"][:nl] - 1] = st["f"][:nl]
    return f


def test_rebomos_million_atom_crystal_equals_the_golden_cell(ctx):
    gold = json.load(open(os.path.join(S.GOLDEN, "log_rebomos_bulk.json")))["log.rebomos-bulk.1"]
    w = W.mos2_bulk(14, 13, 19)
    w["skin"] = 2.0
    n = len(w["x"])
    assert n == 995904
    ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
    row = make_system(ctx, w, "rebomos")
    pe0, press0 = gold["thermo"][0][3], gold["thermo"][0][2]          # -2061.6112 eV per 288 atoms, 28799.53 bar
    assert S.fmt8(row["pe"] / (n // 288)) == S.fmt8(pe0)
    assert S.fmt8(row["press"]) == S.fmt8(press0)
    f0 = forces_by_tag(ctx)
    assert np.abs(f0.sum(axis=0)).max() < 1e-9 * np.abs(f0).max() * np.sqrt(n)
    # the crystal is periodic with the shipped cell: every image of a cell atom feels the same force
    per_cell = f0.reshape(n // 288, 288, 3)
    assert np.abs(per_cell - per_cell[0]).max() < 1e-10 * max(np.abs(f0).max(), 1.0)

    # thermal disorder + rigid shift: new ordering, new ghosts, new rows -- same physics
    rng = np.random.default_rng(2024)
    xd = w["x"] + rng.normal(scale=0.05, size=w["x"].shape)
    rowd = make_system(ctx, w, "rebomos", xd)
    fd = forces_by_tag(ctx)
    assert np.abs(fd.sum(axis=0)).max() < 1e-9 * np.abs(fd).max() * np.sqrt(n)
    shift = np.array([3.7, -11.3, 5.9])
    rows = make_system(ctx, w, "rebomos", xd + shift)
    fs = forces_by_tag(ctx)
    assert abs(rows["pe"] - rowd["pe"]) < 1e-11 * abs(rowd["pe"])
    assert S.rel_err(fs, fd) < 1e-10

    # plugin-mode entry point (host x/f, pipelined upload/download) on the atoms the resident system holds
    st = ctx.system_download()
    nl, ng = st["nlocal"], st["nghost"]
    c2 = b2.Context(0)
    try:
        c2.rebomos_init(S.rebomos_params_struct(), [0, 1])
        P = S.rebomos_params_struct()
        cs, cg, cmax = W.rebomos_neighbor_cutoffs(list(P.rcmax), [0, 1], w["skin"])
        box = W.single_rank_box(w, cmax)
        x = c2.pinned_array((nl + ng, 3))
        f = c2.pinned_array((nl + ng, 3))
        x[:] = st["x"]
        c2.neigh_build(box, w["ntypes"], cs, cg, nl, ng, x, st["type"], 1, w["skin"])
        c2.set_option("f_overwrite", 1)
        for _ in range(2):                      # second call takes the pipelined path
            f[:] = 0.0
            _, e, _ = c2.rebomos_compute(nl, ng, x, st["type"], st["tag"], 1, 0, f=f)
        assert c2.counter("pipelined_calls") == 1
        assert abs(e - rows["pe"]) < 1e-11 * abs(rows["pe"])
        # ghost forces folded onto their owners by tag (the host application's reverse communication)
        ft = np.zeros((n, 3))
        np.add.at(ft, st["tag"][:nl + ng] - 1, f)
        assert S.rel_err(ft, fs) < 1e-10
    finally:
        c2.close()


def test_aeam_two_million_atom_crystal_known_answer(ctx):
    t = S.load_aeam_fixture()
    ctx.aeam_init({k: t[k] for k in ("nelements", "nnonangular", "nrho", "drho", "nr", "dr", "cut", "frho", "rhor", "z2r")})
    w = W.fcc_alsi((80, 80, 80), 0.0, 1)
    w["skin"] = 1.0
    n = len(w["x"])
    assert n == 2048000
    row = make_system(ctx, w, "aeam")
    assert abs(row["pe"] / n - (-3.4106573819)) < 1e-9          # perfect fcc Al
    f0 = forces_by_tag(ctx)
    assert np.abs(f0).max() < 1e-10

    # alloy (0.75 % angular Si atoms) with thermal disorder: momentum conservation and translation invariance
    w = W.fcc_alsi((80, 80, 80), 0.0075, 7683797)
    w["skin"] = 1.0
    rng = np.random.default_rng(7)
    xd = w["x"] + rng.normal(scale=0.08, size=w["x"].shape)
    rowd = make_system(ctx, w, "aeam", xd)
    fd = forces_by_tag(ctx)
    assert np.abs(fd.sum(axis=0)).max() < 1e-9 * np.abs(fd).max() * np.sqrt(n)
    rows = make_system(ctx, w, "aeam", xd + np.array([1.3, 2.9, -0.7]))
    fs = forces_by_tag(ctx)
    assert abs(rows["pe"] - rowd["pe"]) < 1e-11 * abs(rowd["pe"])
    assert S.rel_err(fs, fd) < 1e-10
