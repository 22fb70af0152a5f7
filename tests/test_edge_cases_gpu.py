"""Edge cases of the force path through the C ABI, each against the oracle (reference sources compiled verbatim) on
identical inputs: empty and nearly empty inputs, isolated atoms (empty neighbor rows), molecules with 1, 2, 3 bonds per
center (degenerate bond-pair tables), free surfaces (ragged rows next to vacuum), one element only."""
import os

import numpy as np
import pytest

import support as S

pytestmark = pytest.mark.gpu

FTOL, ETOL = 1.0e-10, 1.0e-12


def aeam_tables():
    t = S.load_aeam_fixture()
    return {k: t[k] for k in ("nelements", "nnonangular", "nrho", "drho", "nr", "dr", "cut", "frho", "rhor", "z2r")}


def rebomos_engine(cmds):
    lmp = S.MiniLmp((1, 1, 1))
    lmp.command("plugin load " + S.oracle_plugin("rebomos"))
    pot = os.path.join(S.potential_dir(), "MoS.REBO.set5b")
    lmp.commands(["units metal"] + cmds + ["mass 1 95.95", "mass 2 32.065", "pair_style rebomos",
                                            "pair_coeff * * %s M S" % pot])
    return lmp


def aeam_engine(cmds):
    lmp = S.MiniLmp((1, 1, 1))
    lmp.command("plugin load " + S.oracle_plugin("aeam"))
    pot = os.path.join(S.potential_dir(), "AlSi.aeam")
    lmp.commands(["units metal"] + cmds + ["pair_style aeam", "pair_coeff * * %s Al Si" % pot, "neighbor 1.0 bin"])
    return lmp


def compare(ctx, lmp, style, label):
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    lmp.compute(1, 2, reverse=True)
    nl = snap["nlocal"]
    f_ref = lmp.f()[:nl].copy()
    e_ref = lmp.get_double("eng_vdwl")
    v_ref = np.array([lmp.get_double("virial%d" % k) for k in range(6)])
    ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
    fn = ctx.rebomos_compute if style == "rebomos" else ctx.aeam_compute
    f, e, v = fn(nl, snap["nghost"], snap["x"], snap["type"], snap["tag"], 1, 2)
    f = S.fold_ghost_forces(f, snap["swaps"], nl)
    print("%s: nlocal %d nghost %d E %.12g max|f| %.4g ferr %.2e" %
          (label, nl, snap["nghost"], e_ref, np.abs(f_ref).max() if nl else 0.0, S.rel_err(f, f_ref)))
    assert S.rel_err(f, f_ref) < FTOL, label
    assert abs(e - e_ref) <= ETOL * max(abs(e_ref), 1.0), label
    assert S.rel_err(v, v_ref) < FTOL, label
    lmp.close()


@pytest.mark.parametrize("style", ["rebomos", "aeam"])
def test_empty_input(ctx, style):
    """no owned atoms, no ghosts, empty list: a rank whose brick is vacuum.  Zero energy, no error."""
    if style == "rebomos":
        ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
        fn = ctx.rebomos_compute
    else:
        ctx.aeam_init(aeam_tables())
        fn = ctx.aeam_compute
    ctx.set_neighbor_csr(0, 0, np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), 2.0)
    f, e, v = fn(0, 0, np.zeros((0, 3)), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32), 1, 2)
    assert e == 0.0 and not np.any(v) and f.shape == (0, 3)


# lattice constant 40 A: one molecule per cell, molecules never see each other (cutoffs ~13 A)
MOLECULES = {
    "isolated-atoms": ("lattice custom 40.0 a1 1 0 0 a2 0 1 0 a3 0 0 1 basis 0 0 0 basis 0.5 0.5 0.5", "basis 1 1 basis 2 2"),
    "MoS-dimer": ("lattice custom 40.0 a1 1 0 0 a2 0 1 0 a3 0 0 1 basis 0.2 0.2 0.2 basis 0.26 0.2 0.2", "basis 1 1 basis 2 2"),
    "S2-dimer": ("lattice custom 40.0 a1 1 0 0 a2 0 1 0 a3 0 0 1 basis 0.2 0.2 0.2 basis 0.25 0.2 0.2", "basis 1 2 basis 2 2"),
    "Mo2-dimer-in-LJ-range": ("lattice custom 40.0 a1 1 0 0 a2 0 1 0 a3 0 0 1 basis 0.2 0.2 0.2 basis 0.33 0.2 0.2", "basis 1 1 basis 2 1"),
    "SMoS-bent": ("lattice custom 40.0 a1 1 0 0 a2 0 1 0 a3 0 0 1 basis 0.2 0.2 0.2 basis 0.255 0.22 0.2 basis 0.15 0.235 0.21",
                  "basis 1 1 basis 2 2 basis 3 2"),
    "MoS3-pyramid": ("lattice custom 40.0 a1 1 0 0 a2 0 1 0 a3 0 0 1 basis 0.2 0.2 0.2 basis 0.255 0.2 0.22 "
                     "basis 0.17 0.25 0.22 basis 0.17 0.15 0.225", "basis 1 1 basis 2 2 basis 3 2 basis 4 2"),
    "Mo3-triangle-60deg": ("lattice custom 40.0 a1 1 0 0 a2 0 1 0 a3 0 0 1 basis 0.2 0.2 0.2 basis 0.279 0.2 0.2 "
                           "basis 0.2395 0.2684 0.2", "basis 1 1 basis 2 1 basis 3 1"),
}


@pytest.mark.parametrize("name", list(MOLECULES))
def test_rebomos_molecules(ctx, oracle_built, name):
    """0, 1, 2 and 3 bonds per center: bond-pair tables with 0, 1 and 3 entries, the 60-degree blend case with a
    single pair, a dimer that only interacts through the LJ window, atoms with empty rows."""
    lat, basis = MOLECULES[name]
    lmp = rebomos_engine([lat, "region box block 0 2 0 1 0 1", "create_box 2 box", "create_atoms 1 box " + basis])
    ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
    compare(ctx, lmp, "rebomos", name)


def test_rebomos_slab_with_vacuum(ctx, oracle_built):
    """MoS2 layers below 20 A of vacuum: free surfaces, ragged rows, ghost set only in x and y"""
    cmds = []
    for c in S.input_script("in.rebomos-bulk"):
        w = c.split()
        if w[0] == "region":
            cmds += ["region box prism 0 4 0 8 0 2.5 -2.0 0.0 0.0", "region slab block -100 100 -100 100 0 1"]
        elif w[0] == "create_atoms":
            cmds.append(c.replace("create_atoms 2 box", "create_atoms 2 region slab"))
        elif w[0] in ("lattice", "create_box"):
            cmds.append(c)
    lmp = rebomos_engine(cmds + ["displace_atoms all random 0.1 0.1 0.1 4242"])
    ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
    compare(ctx, lmp, "rebomos", "slab")


@pytest.mark.parametrize("name,cmds", [
    ("isolated-Al", ["lattice sc 30.0", "region box block 0 2 0 2 0 2", "create_box 2 box", "create_atoms 1 box"]),
    ("isolated-Si-angular", ["lattice sc 30.0", "region box block 0 2 0 2 0 2", "create_box 2 box", "create_atoms 2 box"]),
    ("AlSi-dimer", ["lattice custom 30.0 a1 1 0 0 a2 0 1 0 a3 0 0 1 basis 0.2 0.2 0.2 basis 0.29 0.2 0.2",
                    "region box block 0 1 0 2 0 1", "create_box 2 box", "create_atoms 1 box basis 1 1 basis 2 2"]),
    ("Si3-angular-triplet", ["lattice custom 30.0 a1 1 0 0 a2 0 1 0 a3 0 0 1 basis 0.2 0.2 0.2 basis 0.28 0.2 0.2 "
                             "basis 0.23 0.27 0.21", "region box block 0 1 0 1 0 2", "create_box 2 box",
                             "create_atoms 2 box"]),
    ("pure-Si-diamond", ["lattice diamond 5.43", "region box block 0 3 0 3 0 3", "create_box 2 box", "create_atoms 2 box",
                         "displace_atoms all random 0.1 0.1 0.1 99"]),
    ("Al-slab-with-vacuum", ["lattice fcc 4.045", "region box block 0 4 0 4 0 9", "region slab block 0 4 0 4 0 3",
                             "create_box 2 box", "create_atoms 1 region slab",
                             "set region slab type/fraction 2 0.1 12345", "displace_atoms all random 0.1 0.1 0.1 7"]),
])
def test_aeam_sparse_and_single_element_systems(ctx, oracle_built, name, cmds):
    """rho = 0 (the reference's Fptmp = 0 branch, pair_aeam.cpp:128,329-332), every atom angular, angular centers with
    one and two neighbors, free surfaces"""
    lmp = aeam_engine(cmds)
    ctx.aeam_init(aeam_tables())
    compare(ctx, lmp, "aeam", name)


@pytest.mark.parametrize("style", ["rebomos", "aeam"])
def test_empty_rank_null_pointers(ctx, style):
    """What the host class of an EMPTY rank passes when LAMMPS has not even allocated atom->x / atom->f: NULL arrays
    with nlocal = nghost = 0 (the library used to reject NULL)."""
    import ctypes
    if style == "rebomos":
        ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
        fn = ctx.L.b200md_rebomos_compute
    else:
        ctx.aeam_init(aeam_tables())
        fn = ctx.L.b200md_aeam_compute
    ctx.set_neighbor_csr(0, 0, np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), 2.0)
    eng = ctypes.c_double(7.0)
    vir = np.full(6, 7.0)
    rc = fn(ctx.h, 0, 0, None, None, None, 1, 2, None, ctypes.byref(eng), vir.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    assert rc == 0, ctx.last_error()
    assert eng.value == 0.0 and not vir.any()


def test_neighbor_list_through_a_permuted_ilist(ctx, oracle_built):
    """LAMMPS indexes numneigh[] / firstneigh[] by atom index and lists the atoms with a row in ilist[]
    (b200md_set_neighbor_list_ilist): a permuted ilist gives the same forces as the identity; a list that misses an
    owned atom (skip list) is refused."""
    import ctypes
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), displace=0.1)
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
    nl, ng = snap["nlocal"], snap["nghost"]
    ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
    f0, e0, v0 = ctx.rebomos_compute(nl, ng, snap["x"], snap["type"], snap["tag"])
    rows = snap["inum"] + snap["gnum"]
    num = np.diff(snap["off"]).astype(np.int32)
    val = np.ascontiguousarray(snap["val"], dtype=np.int32)
    PI = ctypes.POINTER(ctypes.c_int)
    first = (PI * rows)()
    base = val.ctypes.data
    for i in range(rows):
        first[i] = ctypes.cast(base + 4 * int(snap["off"][i]), PI)
    rng = np.random.default_rng(5)
    ilist = np.concatenate([rng.permutation(snap["inum"]), snap["inum"] + rng.permutation(snap["gnum"])]).astype(np.int32)
    rc = ctx.L.b200md_set_neighbor_list_ilist(ctx.h, snap["inum"], snap["gnum"], ilist.ctypes.data_as(PI),
                                              num.ctypes.data_as(PI), first, ctypes.c_double(snap["skin"]))
    assert rc == 0, ctx.last_error()
    f1, e1, v1 = ctx.rebomos_compute(nl, ng, snap["x"], snap["type"], snap["tag"])
    assert S.rel_err(f1, f0) < 1e-13 and abs(e1 - e0) < 1e-12 * abs(e0)
    bad = ilist.copy()
    bad[0] = bad[1]                                  # an owned atom without a row, another one twice
    rc = ctx.L.b200md_set_neighbor_list_ilist(ctx.h, snap["inum"], snap["gnum"], bad.ctypes.data_as(PI),
                                              num.ctypes.data_as(PI), first, ctypes.c_double(snap["skin"]))
    assert rc == -2 and "skip lists" in ctx.last_error()
    lmp.close()


@pytest.mark.parametrize("style", ["rebomos", "aeam"])
def test_explicit_pair_virial_matches_the_reference_tally(oracle_built, style):
    """vflag = VIRIAL_PAIR (what PairHybrid hands its sub-styles, and what every caller gets now that the classes set
    no_virial_fdotr_compute): the reference accumulates the global virial through ev_tally / v_tally3 / v_tally2
    (pair_rebomos.cpp:444,554,710,725; pair_aeam.cpp:393,472); the B200 plugin returns the device sums.  Same numbers."""
    out = {}
    for which in ("ref", "b200"):
        if style == "rebomos":
            so = S.oracle_plugin("rebomos") if which == "ref" else S.B200_REBOMOS_SO
            lmp = S.make_rebomos_system(so, (2, 1, 1), displace=0.15)
        else:
            so = S.oracle_plugin("aeam") if which == "ref" else S.B200_AEAM_SO
            lmp = S.make_aeam_system(so, (5, 5, 5), si_fraction=0.1, displace=0.15)
        res = {}
        for vflag in (1, 2):                 # VIRIAL_PAIR, VIRIAL_FDOTR
            lmp.setup(1, vflag)
            res[vflag] = np.array([lmp.get_double("virial%d" % k) for k in range(6)])
        out[which] = res
        lmp.close()
    scale = np.abs(out["ref"][2]).max()
    assert np.abs(out["b200"][1] - out["ref"][1]).max() < 1e-10 * scale
    assert np.abs(out["b200"][2] - out["ref"][2]).max() < 1e-10 * scale
    assert np.abs(out["ref"][1] - out["ref"][2]).max() < 1e-9 * scale     # the two routes agree in the reference itself


@pytest.mark.parametrize("style", ["rebomos", "aeam"])
def test_plugin_on_a_grid_with_empty_ranks(oracle_built, style):
    """A slab under vacuum on a 1x1x4 brick grid: the upper ranks own no atoms and (far enough from the slab) see no
    ghosts.  Pair::compute of those ranks must return quietly and the thermo must equal the reference's."""
    rows = {}
    for which in ("ref", "b200"):
        lmp = S.MiniLmp((1, 1, 4))
        if style == "rebomos":
            lmp.command("plugin load " + (S.oracle_plugin("rebomos") if which == "ref" else S.B200_REBOMOS_SO))
            pot = os.path.join(S.potential_dir(), "MoS.REBO.set5b")
            for c in S.input_script("in.rebomos-bulk"):
                w = c.split()
                if w[0] == "region":
                    lmp.commands(["region box prism 0 4 0 8 0 8 -2.0 0.0 0.0", "region slab block -100 100 -100 100 0 1"])
                elif w[0] == "create_atoms":
                    lmp.command(c.replace("create_atoms 2 box", "create_atoms 2 region slab"))
                elif w[0] == "pair_coeff":
                    lmp.command("pair_coeff * * %s M S" % pot)
                elif w[0] not in ("thermo_style", "thermo", "fix", "run"):
                    lmp.command(c)
        else:
            lmp.command("plugin load " + (S.oracle_plugin("aeam") if which == "ref" else S.B200_AEAM_SO))
            pot = os.path.join(S.potential_dir(), "AlSi.aeam")
            lmp.commands(["units metal", "atom_style atomic", "boundary p p p", "lattice fcc 4.045",
                          "region box block 0 4 0 4 0 24", "region slab block 0 4 0 4 0 3", "create_box 2 box",
                          "create_atoms 1 region slab", "pair_style aeam", "pair_coeff * * %s Al Si" % pot,
                          "neighbor 1.0 bin", "set region slab type/fraction 2 0.05 4711"])
        lmp.commands(["velocity all create 300.0 4928459", "fix 1 all nve", "thermo 5", "run 10"])
        empty = [r for r in range(4) if lmp.get_int("nlocal", r) + lmp.get_int("nghost", r) == 0]
        assert empty, "the test needs at least one rank without owned and ghost atoms"
        rows[which] = lmp.thermo()
        lmp.close()
    for r, g in zip(rows["b200"], rows["ref"]):
        assert abs(r["pe"] - g["pe"]) < 1e-9 * abs(g["pe"])
        assert abs(r["press"] - g["press"]) < 1e-6 * max(abs(g["press"]), 1.0)
        assert abs(r["temp"] - g["temp"]) < 1e-7 * max(g["temp"], 1.0)
