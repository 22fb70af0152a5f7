"""GPU parity of the REBOMoS force path against the oracle (reference sources compiled verbatim,
oracle/_ref, or the port when _ref is absent) on identical inputs.  All calls go through the C ABI.

Tolerance: forces 1e-10 relative to the largest force component (north star), energy 1e-12 relative,
virial 1e-10 relative to the largest virial component.  Neighbor sub-lists: bit-exact."""
import numpy as np
import pytest

import support as S

pytestmark = pytest.mark.gpu

FTOL = 1.0e-10
ETOL = 1.0e-12


def oracle_forces(lmp):
    """owned-atom forces after reverse comm, energy, fdotr virial from the oracle engine"""
    lmp.compute(1, 2, reverse=True)
    nlocal = lmp.get_int("nlocal")
    f = lmp.f()[:nlocal].copy()
    return f, lmp.get_double("eng_vdwl"), np.array([lmp.get_double("virial%d" % k) for k in range(6)])


def gpu_forces(ctx, snap, eflag=1, vflag=2):
    ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
    f, e, v = ctx.rebomos_compute(snap["nlocal"], snap["nghost"], snap["x"], snap["type"], snap["tag"], eflag, vflag)
    return S.fold_ghost_forces(f, snap["swaps"], snap["nlocal"]), e, v


def init_ctx(ctx):
    ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])


CASES = [
    dict(id="bulk288", replicate=(1, 1, 1), displace=0.0),
    dict(id="bulk288-d0.05", replicate=(1, 1, 1), displace=0.05),
    dict(id="bulk288-d0.3", replicate=(1, 1, 1), displace=0.3),     # S-S pairs enter the switching window
    dict(id="rep2x2x1-d0.15", replicate=(2, 2, 1), displace=0.15),
    dict(id="rep2x1x2-d0.6", replicate=(2, 1, 2), displace=0.6),     # strongly disordered: LJ taper regime
]


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_forces_energy_virial(ctx, oracle_built, case):
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), case["replicate"], displace=case["displace"])
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    f_ref, e_ref, v_ref = oracle_forces(lmp)
    init_ctx(ctx)
    f, e, v = gpu_forces(ctx, snap)
    ferr = S.rel_err(f, f_ref)
    eerr = abs(e - e_ref) / abs(e_ref)
    verr = S.rel_err(v, v_ref)
    print("\n%s: nlocal %d nghost %d  max|f| %.4g  ferr %.3e  E %.10f eerr %.3e  verr %.3e  short %d lj %d"
          % (case["id"], snap["nlocal"], snap["nghost"], np.abs(f_ref).max(), ferr, e, eerr, verr,
             ctx.counter("short_entries"), ctx.counter("lj_entries")))
    assert ferr < FTOL
    assert eerr < ETOL
    assert verr < FTOL
    lmp.close()


def test_golden_step0_energy(ctx, oracle_built):
    """PotEng of log.rebomos-bulk.1 step 0 straight from the CUDA path: -2061.6112 (8 digits)."""
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"))
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    init_ctx(ctx)
    f, e, v = gpu_forces(ctx, snap)
    assert S.fmt8(e) == "-2061.6112"
    # Press(0) = trace(virial)/(3V) * nktv2p = 28799.53 bar
    vol = 5922.4926
    row = lmp.thermo()[0]
    press = (v[0] + v[1] + v[2]) / 3.0 / row["vol"] * lmp.get_double("nktv2p")
    assert S.fmt8(press) == "28799.53"
    lmp.close()


def test_no_energy_no_virial_flags(ctx, oracle_built):
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), displace=0.1)
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    f_ref, _, _ = oracle_forces(lmp)
    init_ctx(ctx)
    f, e, v = gpu_forces(ctx, snap, eflag=0, vflag=0)
    assert e == 0.0 and not v.any()
    assert S.rel_err(f, f_ref) < FTOL
    lmp.close()


def test_forces_accumulate(ctx, oracle_built):
    """f is accumulated into, like atom->f (pair_rebomos.cpp:436-441 use +=)."""
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), displace=0.1)
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    init_ctx(ctx)
    ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
    nall = snap["nlocal"] + snap["nghost"]
    f0, _, _ = ctx.rebomos_compute(snap["nlocal"], snap["nghost"], snap["x"], snap["type"], snap["tag"])
    pre = np.full((nall, 3), 1.5)
    f1, _, _ = ctx.rebomos_compute(snap["nlocal"], snap["nghost"], snap["x"], snap["type"], snap["tag"], f=pre.copy())
    assert np.allclose(f1 - 1.5, f0, rtol=0, atol=1e-12)
    lmp.close()


def test_inner_list_margin(ctx, oracle_built):
    """A smaller inner-list margin (two-level Verlet list) gives the same forces and rebuilds itself
    when atoms have moved more than margin/2 since the inner build."""
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), displace=0.05)
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    init_ctx(ctx)
    ctx.set_option("margin", 500)    # 0.5 A
    f, e, v = gpu_forces(ctx, snap)
    f_ref, e_ref, v_ref = oracle_forces(lmp)
    assert S.rel_err(f, f_ref) < FTOL
    n0 = ctx.counter("inner_rebuilds")
    # move owned atoms by up to 0.4 A (> margin/2, < skin/2) keeping the master list valid
    rng = np.random.default_rng(7)
    x = lmp.x()
    nlocal = snap["nlocal"]
    x[:nlocal] += rng.uniform(-0.4, 0.4, size=(nlocal, 3)) / np.sqrt(3.0)
    lmp.forward_comm()
    snap2 = S.snapshot(lmp)
    f_ref, e_ref, v_ref = oracle_forces(lmp)
    f2, e2, v2 = ctx.rebomos_compute(snap2["nlocal"], snap2["nghost"], snap2["x"], snap2["type"], snap2["tag"])
    f2 = S.fold_ghost_forces(f2, snap2["swaps"], nlocal)
    assert ctx.counter("inner_rebuilds") == n0 + 1
    assert S.rel_err(f2, f_ref) < FTOL
    assert abs(e2 - e_ref) / abs(e_ref) < ETOL
    ctx.set_option("margin", 0)
    lmp.close()


def test_rebo_sublist_bit_exact(ctx, oracle_built):
    """REBO_neigh parity (pair_rebomos.cpp:281-352): sub-lists of owned AND ghost atoms identical in
    content and order to a direct restatement of the filter over the oracle's full list; nM/nS to 1e-14."""
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), displace=0.3)
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    init_ctx(ctx)
    ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
    num, rows, nM, nS = ctx.rebomos_neigh(snap["nlocal"], snap["nghost"], snap["x"], snap["type"], stride=24)
    P = S.rebomos_params_struct()
    rcmin = np.array(P.rcmin[:]).reshape(2, 2)
    rcmax = np.array(P.rcmax[:]).reshape(2, 2)
    x, el = snap["x"], snap["type"] - 1
    off, val = snap["off"], snap["val"]
    nall = snap["nlocal"] + snap["nghost"]
    worst = 0.0
    for i in range(nall):
        js = val[off[i]:off[i + 1]] & 0x1FFFFFFF
        d = x[i] - x[js]
        rsq = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
        keep = rsq < (rcmax[el[i], el[js]] * rcmax[el[i], el[js]])
        ref = js[keep]
        assert num[i] == len(ref), i
        assert np.array_equal(rows[i, :num[i]], ref), i
        r = np.sqrt(rsq[keep])
        t = (r - rcmin[el[i], el[ref]]) / (rcmax[el[i], el[ref]] - rcmin[el[i], el[ref]])
        w = np.where(t <= 0, 1.0, np.where(t >= 1, 0.0, 0.5 * (1 + np.cos(t * np.pi))))
        worst = max(worst, abs(nM[i] - w[el[ref] == 0].sum()), abs(nS[i] - w[el[ref] == 1].sum()))
    assert worst < 1e-13
    lmp.close()


@pytest.mark.parametrize("case", [CASES[1], CASES[4]], ids=["bulk288-d0.05", "rep2x1x2-d0.6"])
def test_deterministic_mode(ctx, oracle_built, case):
    """option "deterministic": bond forces go through the (center, slot) table + fixed-order gather instead of
    FP64 atomics.  Same parity bar against the oracle; ghost AND owned forces bitwise identical between
    repeated calls; agrees with the atomic path to rounding."""
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), case["replicate"], displace=case["displace"])
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    f_ref, e_ref, v_ref = oracle_forces(lmp)
    init_ctx(ctx)
    ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
    args = (snap["nlocal"], snap["nghost"], snap["x"], snap["type"], snap["tag"], 1, 2)
    fa, ea, va = ctx.rebomos_compute(*args)
    try:
        ctx.set_option("deterministic", 1)
        runs = [ctx.rebomos_compute(*args) for _ in range(4)]
    finally:
        ctx.set_option("deterministic", 0)
    for f, e, v in runs[1:]:
        assert np.array_equal(f, runs[0][0]), "deterministic forces differ between calls"
        # global sums too: block partial sums are added in fixed point (integer atomics are associative)
        assert e == runs[0][1] and np.array_equal(v, runs[0][2]), "deterministic energy / virial differ between calls"
    f, e, v = runs[0]
    assert S.rel_err(S.fold_ghost_forces(f, snap["swaps"], snap["nlocal"]), f_ref) < FTOL
    assert abs(e - e_ref) / abs(e_ref) < ETOL and S.rel_err(v, v_ref) < FTOL
    assert S.rel_err(f, fa) < 1e-13
    lmp.close()


def test_plugin_mode_tight_rows(ctx, oracle_built):
    """Plugin mode keeps tight rows (rcut + 0.4 A) of its own: the force kernels stream them, the displacement check of
    every call covers them, they are re-derived AFTER a call in which an atom came within 80 % of their limit (the next
    upload waits for that), and a call that finds an atom beyond the limit recomputes with rows derived on the spot.
    Every call must equal the plain path (no pipeline, no tight rows) on the same positions."""
    import lammps_plugins_b200 as b2
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), (2, 2, 1), displace=0.1)
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    nl, ng = snap["nlocal"], snap["nghost"]
    init_ctx(ctx)
    plain = b2.Context(0)
    plain.rebomos_init(S.rebomos_params_struct(), [0, 1])
    for c in (ctx, plain):
        c.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
    ctx.set_option("d2h_min_atoms", 0)
    ctx.set_option("d2h_chunks", 2)
    ctx.set_option("h2d_chunks", 3)
    plain.set_option("h2d_chunks", 1)
    plain.set_option("d2h_chunks", 1)
    plain.set_option("margin_tight", 0)
    try:
        rng = np.random.default_rng(11)
        x = snap["x"].copy()
        redo0 = None
        for call in range(8):
            if call == 0:
                pass
            elif call in (1, 2):
                x = x + rng.uniform(-0.02, 0.02, x.shape)            # small moves: tight rows stay
            elif call == 3:
                x[11] += (0.17, 0.0, 0.0)                             # > 0.8 * 0.2: derived after this call
            elif call == 4:
                x[11] += (0.10, 0.0, 0.0)                             # 0.27 from the ORIGINAL derive, 0.10 from the new one
            elif call == 5:
                x[23] += (0.0, 0.3, 0.0)                              # beyond margin_t/2 = 0.2 at once: recompute (level 1)
            else:
                x = x + rng.uniform(-0.03, 0.03, x.shape)
            f, e, v = ctx.rebomos_compute(nl, ng, x, snap["type"], snap["tag"], 1, 2)
            fq, eq, vq = plain.rebomos_compute(nl, ng, x, snap["type"], snap["tag"], 1, 2)
            assert S.rel_err(f, fq) < 1e-12 and abs(e - eq) < 1e-12 * abs(eq) and S.rel_err(v, vq) < 1e-11, call
            if call == 4:
                redo0 = ctx.counter("pipelined_redos")
            if call == 5:
                assert ctx.counter("pipelined_redos") == redo0 + 1
        assert ctx.counter("pipelined_calls") >= 6 and ctx.counter("tight_refreshes") >= 3
    finally:
        ctx.set_option("d2h_chunks", 4)
        ctx.set_option("h2d_chunks", 6)
        ctx.set_option("d2h_min_atoms", 65536)
        plain.close()
    lmp.close()


def test_many_rebo_neighbors_overflow_path(ctx, oracle_built):
    """a compressed cell gives S atoms more than 8 REBO neighbors: those centers leave the narrow launch through
    the overflow list and are evaluated by the wide one -- same parity bar"""
    lmp = S.MiniLmp()
    lmp.command("plugin load " + S.oracle_plugin("rebomos"))
    lmp.commands([c.replace("lattice custom 1.0", "lattice custom 0.8") for c in S.rebomos_bulk_commands((2, 1, 2))])
    lmp.command("displace_atoms all random 0.05 0.05 0.05 12345")
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    f_ref, e_ref, v_ref = oracle_forces(lmp)
    init_ctx(ctx)
    f, e, v = gpu_forces(ctx, snap)
    num, rows, nM, nS = ctx.rebomos_neigh(snap["nlocal"], snap["nghost"], snap["x"], snap["type"])
    styp = np.asarray(snap["type"])[:snap["nlocal"]] == 2
    print("max REBO neighbors: S %d, Mo %d" % (num[:snap["nlocal"]][styp].max(), num[:snap["nlocal"]][~styp].max()))
    assert num[:snap["nlocal"]][styp].max() > 8
    assert S.rel_err(f, f_ref) < FTOL and abs(e - e_ref) / abs(e_ref) < ETOL
    lmp.close()


def fold_ghost_rows(a, swaps, nlocal):
    """reverse communication of a per-atom array with any number of columns (what compute pe/atom and
    stress/atom do with Pair::eatom / vatom): ghost rows added to the atoms they image, swaps in reverse"""
    a = np.array(a, dtype=np.float64, copy=True)
    for s in reversed(swaps):
        if s["recvnum"]:
            np.add.at(a, s["sendlist"], a[s["firstrecv"]:s["firstrecv"] + s["recvnum"]])
    return a[:nlocal]


@pytest.mark.parametrize("case", [CASES[1], CASES[2], CASES[4]], ids=["bulk288-d0.05", "bulk288-d0.3", "rep2x1x2-d0.6"])
def test_per_atom_energy_and_virial(ctx, oracle_built, case):
    """Pair::eatom / Pair::vatom (ENERGY_ATOM | VIRIAL_ATOM): the device distributes energy and virial over atoms the
    way the reference's tallies do -- ev_tally halves, v_tally3 thirds, v_tally2 halves -- so after folding ghost
    shares into their owners every atom's energy and 6 virial components agree with the reference plugin."""
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), case["replicate"], displace=case["displace"])
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    nl, nall = snap["nlocal"], snap["nlocal"] + snap["nghost"]
    lmp.compute(1 | 2, 2 | 4, reverse=True)             # ENERGY_GLOBAL|ENERGY_ATOM, VIRIAL_FDOTR|VIRIAL_ATOM
    e_ref = lmp.get_double("eng_vdwl")
    ea_ref = fold_ghost_rows(lmp._arr("eatom", 0, nall, np.float64), snap["swaps"], nl)
    va_ref = fold_ghost_rows(lmp._arr("vatom", 0, nall, np.float64, 6), snap["swaps"], nl)
    f_ref = lmp.f()[:nl].copy()
    init_ctx(ctx)
    ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
    f, e, v, ea, va = ctx.rebomos_compute_peratom(snap["nlocal"], snap["nghost"], snap["x"], snap["type"], snap["tag"])
    ea, va = fold_ghost_rows(ea, snap["swaps"], nl), fold_ghost_rows(va, snap["swaps"], nl)
    print("\\n%s: sum eatom %.10f (E %.10f)  max|eatom| %.4f err %.2e  max|vatom| %.4f err %.2e"
          % (case["id"], ea.sum(), e_ref, np.abs(ea_ref).max(), np.abs(ea - ea_ref).max(), np.abs(va_ref).max(),
             np.abs(va - va_ref).max()))
    assert abs(ea.sum() - e_ref) < 1e-11 * abs(e_ref) and abs(e - e_ref) < ETOL * abs(e_ref)
    assert S.rel_err(ea, ea_ref) < 1e-10
    assert S.rel_err(va, va_ref) < 1e-10
    assert S.rel_err(S.fold_ghost_forces(f, snap["swaps"], nl), f_ref) < FTOL
    # per-atom virial sums to the global virial the same call reports
    assert S.rel_err(va.sum(axis=0), v) < 1e-10
    lmp.close()


@pytest.mark.parametrize("overwrite", [0, 1], ids=["accumulate", "overwrite"])
@pytest.mark.parametrize("chunks", [2, 5, 16])
def test_ranged_force_return_equals_single_copy(ctx, oracle_built, overwrite, chunks):
    """Plugin-mode pipelining: forces of finished atom-index ranges travel to the host while the remaining LJ ranges
    still compute (d2h_chunks > 1).  Forced on for a small system here; results must equal the single-copy path
    bit for bit per owned atom up to the order of the two adds on f (<= 1 ulp), energy/virial to rounding, in both the
    accumulate (Pair::compute semantics: f +=) and the overwrite mode."""
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), (2, 2, 1), displace=0.15)
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    f_ref, e_ref, v_ref = oracle_forces(lmp)
    init_ctx(ctx)
    ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
    args = (snap["nlocal"], snap["nghost"], snap["x"], snap["type"], snap["tag"], 1, 2)
    nall = snap["nlocal"] + snap["nghost"]
    base = np.full((nall, 3), 0.25)                  # what is in f before the call
    ctx.set_option("f_overwrite", overwrite)
    ctx.set_option("d2h_chunks", 1)
    f1, e1, v1 = ctx.rebomos_compute(*args, f=base.copy())
    ctx.set_option("d2h_min_atoms", 0)
    ctx.set_option("d2h_chunks", chunks)
    f2, e2, v2 = ctx.rebomos_compute(*args, f=base.copy())
    ctx.set_option("f_overwrite", 0)
    ctx.set_option("d2h_chunks", 4)
    ctx.set_option("d2h_min_atoms", 65536)
    assert S.rel_err(f2, f1) < 1e-14
    assert abs(e2 - e1) < 1e-13 * abs(e1) and S.rel_err(v2, v1) < 1e-13
    off = 0.0 if overwrite else 0.25
    f = S.fold_ghost_forces(f2 - off, snap["swaps"], snap["nlocal"])
    assert S.rel_err(f, f_ref) < FTOL
    assert abs(e2 - e_ref) < ETOL * abs(e_ref)
    lmp.close()


def test_lj_center_pairs_equal_single_rows(ctx, oracle_built):
    """LJ over pairs of neighboring centers sharing one union row (default) against one row per center
    (lj_pairs = 0) and against the oracle, on a chemically scrambled cell whose element counts are ODD (a center
    without partner) and whose consecutive same-element centers are far apart in places (union ~ 2 spheres)."""
    for seed in range(11, 40):
        lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), (2, 1, 1), displace=0.1,
                                    extra=["set group all type/fraction 2 0.15 %d" % seed])
        lmp.setup(1, 2)
        snap = S.snapshot(lmp)
        nmo = int((snap["type"][:snap["nlocal"]] == 1).sum())
        if nmo % 2 == 1 and (snap["nlocal"] - nmo) % 2 == 1:
            break
        lmp.close()
    else:
        pytest.skip("no seed with odd element counts")
    f_ref, e_ref, v_ref = oracle_forces(lmp)
    init_ctx(ctx)
    out = {}
    for pairs in (1, 0):
        ctx.set_option("lj_pairs", pairs)
        out[pairs] = gpu_forces(ctx, snap)
        f, e, v = out[pairs]
        assert S.rel_err(f, f_ref) < FTOL and abs(e - e_ref) < ETOL * abs(e_ref) and S.rel_err(v, v_ref) < FTOL
    ctx.set_option("lj_pairs", 1)
    assert S.rel_err(out[1][0], out[0][0]) < 1e-12
    lmp.close()


@pytest.mark.parametrize("overwrite", [0, 1], ids=["accumulate", "overwrite"])
def test_pipelined_upload_equals_plain_path(ctx, oracle_built, overwrite):
    """Plugin-mode pipelining of the position upload (pieces of x land while the bond-order launches of earlier center
    ranges run; the inner lists are used speculatively and the step is recomputed when the displacement check, which
    needs all positions, says they were stale).  Forced on for a small system: (1) same forces as the plain path and
    the oracle, (2) after a small move, (3) after a move beyond margin/2, which must take the recompute path."""
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), (2, 2, 1), displace=0.15)
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    f_ref, e_ref, v_ref = oracle_forces(lmp)
    nl, ng = snap["nlocal"], snap["nghost"]
    init_ctx(ctx)
    ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
    ctx.set_option("f_overwrite", overwrite)
    ctx.set_option("d2h_min_atoms", 0)
    ctx.set_option("d2h_chunks", 2)
    ctx.set_option("h2d_chunks", 3)
    base = 0.0 if overwrite else 0.5

    def run(x):
        f, e, v = ctx.rebomos_compute(nl, ng, x, snap["type"], snap["tag"], 1, 2, f=np.full((nl + ng, 3), base))
        return f - base, e, v

    try:
        x0 = snap["x"].copy()
        f0, e0, v0 = run(x0)                        # derives the inner lists: plain path
        p0, r0 = ctx.counter("pipelined_calls"), ctx.counter("pipelined_redos")
        f1, e1, v1 = run(x0)                        # pipelined
        assert ctx.counter("pipelined_calls") == p0 + 1 and ctx.counter("pipelined_redos") == r0
        # in a cell this small range 0 names atoms of the last piece: they travel ahead of the pieces as stragglers
        # (at full size: the neighbors of atoms wrapped through a periodic face since the last sort)
        assert ctx.counter("upload_stragglers") > 0
        assert S.rel_err(f1, f0) < 1e-14 and abs(e1 - e0) < 1e-13 * abs(e0) and S.rel_err(v1, v0) < 1e-13
        assert S.rel_err(S.fold_ghost_forces(f1.copy(), snap["swaps"], nl), f_ref) < FTOL
        assert abs(e1 - e_ref) < ETOL * abs(e_ref)
        rng = np.random.default_rng(5)
        x1 = x0 + rng.uniform(-0.05, 0.05, x0.shape)   # inside the margin: lists stay
        f2, e2, v2 = run(x1)
        assert ctx.counter("pipelined_redos") == r0
        x2 = x1.copy()
        x2[7] += (0.45, 0.45, 0.0)                  # |d| = 0.64 > margin/2 = 0.5 (< skin/2: the master list is still good)
        f3, e3, v3 = run(x2)
        assert ctx.counter("pipelined_redos") == r0 + 1
        ctx.set_option("h2d_chunks", 1)             # plain path (re-derives the lists at the positions it is given)
        for x, (fp, ep, vp) in ((x1, (f2, e2, v2)), (x2, (f3, e3, v3))):
            fq, eq, vq = run(x)
            assert S.rel_err(fp, fq) < 1e-13 and abs(ep - eq) < 1e-13 * abs(eq) and S.rel_err(vp, vq) < 1e-12
    finally:
        ctx.set_option("f_overwrite", 0)
        ctx.set_option("d2h_chunks", 4)
        ctx.set_option("h2d_chunks", 6)
        ctx.set_option("d2h_min_atoms", 65536)
        lmp.close()
