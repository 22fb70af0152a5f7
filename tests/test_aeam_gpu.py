"""GPU parity of the AEAM force path against the oracle (USER-AEAM/pair_aeam.cpp compiled verbatim,
or the port) on identical inputs, through the C ABI.

AEAM ships no golden output: parity is pinned to the reference SOURCE compiled
verbatim; derived known answers (perfect fcc Al) are checked as well.
Tolerances: forces 1e-10 of the largest force component, energy 1e-12, virial 1e-10 (relative)."""
import numpy as np
import pytest

import support as S

pytestmark = pytest.mark.gpu

FTOL = 1.0e-10
ETOL = 1.0e-12


def aeam_tables():
    t = S.load_aeam_fixture()
    return dict(nelements=t["nelements"], nnonangular=t["nnonangular"], nrho=t["nrho"], drho=t["drho"],
                nr=t["nr"], dr=t["dr"], cut=t["cut"], frho=t["frho"], rhor=t["rhor"], z2r=t["z2r"])


def oracle_forces(lmp):
    lmp.compute(1, 2, reverse=True)
    nlocal = lmp.get_int("nlocal")
    return (lmp.f()[:nlocal].copy(), lmp.get_double("eng_vdwl"),
            np.array([lmp.get_double("virial%d" % k) for k in range(6)]))


def gpu_forces(ctx, snap, eflag=1, vflag=2):
    ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
    f, e, v = ctx.aeam_compute(snap["nlocal"], snap["nghost"], snap["x"], snap["type"], snap["tag"], eflag, vflag)
    return S.fold_ghost_forces(f, snap["swaps"], snap["nlocal"]), e, v


CASES = [
    dict(id="fcc-pure-d0.05", cells=(4, 4, 4), si=0.0, displace=0.05),
    dict(id="al-0.75pctSi-d0.1", cells=(6, 6, 6), si=0.0075, displace=0.1),
    dict(id="al-20pctSi-d0.2", cells=(5, 5, 5), si=0.2, displace=0.2),      # Si-Si pairs: CutDec paths
    dict(id="si-rich-60pct-d0.3", cells=(4, 4, 4), si=0.6, displace=0.3),
    dict(id="al-5pctSi-d0.5", cells=(5, 4, 6), si=0.05, displace=0.5),
]


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_forces_energy_virial(ctx, oracle_built, case):
    lmp = S.make_aeam_system(S.oracle_plugin("aeam"), case["cells"], si_fraction=case["si"], displace=case["displace"])
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    f_ref, e_ref, v_ref = oracle_forces(lmp)
    ctx.aeam_init(aeam_tables())
    f, e, v = gpu_forces(ctx, snap)
    ferr, eerr, verr = S.rel_err(f, f_ref), abs(e - e_ref) / abs(e_ref), S.rel_err(v, v_ref)
    nsi = int((snap["type"][:snap["nlocal"]] == 2).sum())
    print("\n%s: nlocal %d (Si %d) nghost %d max|f| %.4g ferr %.3e E %.10f eerr %.3e verr %.3e"
          % (case["id"], snap["nlocal"], nsi, snap["nghost"], np.abs(f_ref).max(), ferr, e, eerr, verr))
    assert ferr < FTOL
    assert eerr < ETOL
    assert verr < FTOL
    lmp.close()


def test_perfect_fcc_known_answer(ctx, oracle_built):
    """Perfect fcc Al, a = 4.045: rho = 1.0404330586, E/atom = -3.4106573819 eV,
    forces vanish by symmetry."""
    lmp = S.make_aeam_system(S.oracle_plugin("aeam"), (4, 4, 4), si_fraction=0.0)
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    ctx.aeam_init(aeam_tables())
    f, e, v = gpu_forces(ctx, snap)
    rho, fp = ctx.aeam_rho_fp(snap["nlocal"])
    assert abs(e / snap["nlocal"] - (-3.4106573819)) < 1e-9
    assert np.allclose(rho, 1.0404330586, rtol=0, atol=1e-9)
    assert np.abs(f).max() < 1e-10
    lmp.close()


def test_two_phase_equals_one_shot(ctx, oracle_built):
    """density phase -> (host halo exchange of fp, here by atom ID) -> force phase == one-shot compute."""
    lmp = S.make_aeam_system(S.oracle_plugin("aeam"), (5, 5, 5), si_fraction=0.1, displace=0.15)
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    ctx.aeam_init(aeam_tables())
    f1, e1, v1 = gpu_forces(ctx, snap)
    nl, ng = snap["nlocal"], snap["nghost"]
    rho, fp = ctx.aeam_density_phase(nl, ng, snap["x"], snap["type"])
    owner = np.zeros(snap["tag"].max() + 1, dtype=np.int64)
    owner[snap["tag"][:nl]] = np.arange(nl)
    rho[nl:] = rho[owner[snap["tag"][nl:]]]      # what comm->forward_comm(this) does on one rank
    fp[nl:] = fp[owner[snap["tag"][nl:]]]
    f2, e2, v2 = ctx.aeam_force_phase(rho, fp)
    f2 = S.fold_ghost_forces(f2, snap["swaps"], nl)
    assert S.rel_err(f2, f1) < 1e-13
    assert abs(e2 - e1) < 1e-9 * abs(e1)
    assert S.rel_err(v2, v1) < 1e-12
    try:    # the force phase with its download pipelined (forced on for this small system)
        ctx.set_option("d2h_min_atoms", 0)
        ctx.set_option("d2h_chunks", 3)
        rho, fp = ctx.aeam_density_phase(nl, ng, snap["x"], snap["type"])
        rho[nl:] = rho[owner[snap["tag"][nl:]]]
        fp[nl:] = fp[owner[snap["tag"][nl:]]]
        p0 = ctx.counter("pipelined_calls")
        f3, e3, v3 = ctx.aeam_force_phase(rho, fp)
        assert ctx.counter("pipelined_calls") == p0 + 1
        f3 = S.fold_ghost_forces(f3, snap["swaps"], nl)
        assert S.rel_err(f3, f1) < 1e-13 and abs(e3 - e1) < 1e-9 * abs(e1) and S.rel_err(v3, v1) < 1e-12
    finally:
        ctx.set_option("d2h_min_atoms", 65536)
        ctx.set_option("d2h_chunks", 4)
    lmp.close()


def _interpolate_np(n, delta, f):
    """numpy restatement of PairAEAM::interpolate (pair_aeam.cpp:915-942); rows 1..n, elementwise IEEE ops"""
    s = np.zeros((n + 1, 7))
    s[1:, 6] = f
    s[1, 5] = s[2, 6] - s[1, 6]
    s[2, 5] = 0.5 * (s[3, 6] - s[1, 6])
    s[n - 1, 5] = 0.5 * (s[n, 6] - s[n - 2, 6])
    s[n, 5] = s[n, 6] - s[n - 1, 6]
    m = np.arange(3, n - 1)
    s[m, 5] = ((s[m - 2, 6] - s[m + 2, 6]) + 8.0 * (s[m + 1, 6] - s[m - 1, 6])) / 12.0
    m = np.arange(1, n)
    s[m, 4] = 3.0 * (s[m + 1, 6] - s[m, 6]) - 2.0 * s[m, 5] - s[m + 1, 5]
    s[m, 3] = s[m, 5] + s[m + 1, 5] - 2.0 * (s[m + 1, 6] - s[m, 6])
    s[n, 4] = 0.0
    s[n, 3] = 0.0
    s[1:, 2] = s[1:, 5] / delta
    s[1:, 1] = 2.0 * s[1:, 4] / delta
    s[1:, 0] = 3.0 * s[1:, 3] / delta
    return s


def test_spline_tables_bit_identical(ctx):
    """The 7-coefficient tables built inside the library equal array2spline's bit for bit."""
    t = S.load_aeam_fixture()
    ctx.aeam_init(aeam_tables())
    nel = t["nelements"]
    for i in range(nel):
        got = ctx.aeam_get_spline(0, i, t["nrho"][i])
        assert np.array_equal(got, _interpolate_np(t["nrho"][i], t["drho"][i], t["frho"][i]))
    for i in range(nel):
        for j in range(nel):
            got = ctx.aeam_get_spline(1, i * nel + j, int(t["nr"][i, j]))
            assert np.array_equal(got, _interpolate_np(int(t["nr"][i, j]), float(t["dr"][i, j]), t["rhor"][i][j]))
    k = 0
    for i in range(nel):
        for j in range(i + 1):
            got = ctx.aeam_get_spline(2, k, int(t["nr"][i, j]))
            assert np.array_equal(got, _interpolate_np(int(t["nr"][i, j]), float(t["dr"][i, j]), t["z2r"][i][j]))
            k += 1


def test_flags_and_accumulate(ctx, oracle_built):
    lmp = S.make_aeam_system(S.oracle_plugin("aeam"), (4, 4, 4), si_fraction=0.05, displace=0.1)
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    ctx.aeam_init(aeam_tables())
    ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
    nl, ng = snap["nlocal"], snap["nghost"]
    f0, e0, v0 = ctx.aeam_compute(nl, ng, snap["x"], snap["type"], snap["tag"], 0, 0)
    assert e0 == 0.0 and not v0.any()
    f1, _, _ = ctx.aeam_compute(nl, ng, snap["x"], snap["type"], snap["tag"], 1, 2, f=np.full((nl + ng, 3), 2.0))
    assert np.allclose(f1 - 2.0, f0, rtol=0, atol=1e-12)
    lmp.close()


@pytest.mark.parametrize("overwrite", [0, 1], ids=["accumulate", "overwrite"])
def test_pipelined_force_download_equals_plain_path(ctx, oracle_built, overwrite):
    """Plugin mode: the force download is pipelined with the pair kernel (angular kernel first, ghost forces leave, then
    ranges of centers through the RANGED instance of the force kernel, every range's forces behind its launch).  Forced
    on for a small system with angular atoms: same forces, energy and virial as the plain path and as the oracle."""
    lmp = S.make_aeam_system(S.oracle_plugin("aeam"), (5, 5, 5), si_fraction=0.2, displace=0.2)
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    f_ref, e_ref, v_ref = oracle_forces(lmp)
    ctx.aeam_init(aeam_tables())
    ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
    nl, ng = snap["nlocal"], snap["nghost"]
    base = 0.0 if overwrite else 1.5

    def run(eflag, vflag):
        f, e, v = ctx.aeam_compute(nl, ng, snap["x"], snap["type"], snap["tag"], eflag, vflag, f=np.full((nl + ng, 3), base))
        return f - base, e, v

    try:
        ctx.set_option("f_overwrite", overwrite)
        f0, e0, v0 = run(1, 2)                       # plain path (below d2h_min_atoms)
        ctx.set_option("d2h_min_atoms", 0)
        ctx.set_option("d2h_chunks", 3)
        p0 = ctx.counter("pipelined_calls")
        f1, e1, v1 = run(1, 2)
        f2, _, _ = run(0, 0)
        assert ctx.counter("pipelined_calls") == p0 + 2
        assert S.rel_err(f1, f0) < 1e-13 and abs(e1 - e0) < 1e-13 * abs(e0) and S.rel_err(v1, v0) < 1e-12
        assert S.rel_err(f2, f0) < 1e-13
        assert S.rel_err(S.fold_ghost_forces(f1.copy(), snap["swaps"], nl), f_ref) < FTOL
        assert abs(e1 - e_ref) < ETOL * abs(e_ref) and S.rel_err(v1, v_ref) < FTOL
    finally:
        ctx.set_option("f_overwrite", 0)
        ctx.set_option("d2h_min_atoms", 65536)
        ctx.set_option("d2h_chunks", 4)
    lmp.close()


def test_pipelined_upload_equals_plain_path(ctx, oracle_built):
    """Plugin mode: the position upload is pipelined with the density pass (pieces of x land while the density launches of
    earlier center ranges run; dependences and stragglers from the master rows), the inner rows are used speculatively,
    re-derived one call ahead of need ("soon": an atom at 80 % of margin/2) and the call is recomputed when they were
    stale after all.  Forced on for a small system: every call equals the plain path at the same positions."""
    lmp = S.make_aeam_system(S.oracle_plugin("aeam"), (5, 5, 5), si_fraction=0.2, displace=0.2)
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    f_ref, e_ref, v_ref = oracle_forces(lmp)
    ctx.aeam_init(aeam_tables())
    ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
    nl, ng = snap["nlocal"], snap["nghost"]

    def run(x):
        return ctx.aeam_compute(nl, ng, x, snap["type"], snap["tag"], 1, 2)

    try:
        ctx.set_option("d2h_min_atoms", 0)
        ctx.set_option("d2h_chunks", 2)
        ctx.set_option("h2d_chunks", 3)
        x0 = snap["x"].copy()
        run(x0)                                     # plain upload: derives the inner rows, sets the pipeline up
        p0, r0, i0 = ctx.counter("pipelined_calls"), ctx.counter("pipelined_redos"), ctx.counter("inner_rebuilds")
        f1, e1, v1 = run(x0)                        # pipelined upload
        assert ctx.counter("upload_stragglers") > 0    # a cell this small: range 0 names atoms of the last piece
        assert S.rel_err(S.fold_ghost_forces(f1.copy(), snap["swaps"], nl), f_ref) < FTOL
        assert abs(e1 - e_ref) < ETOL * abs(e_ref) and S.rel_err(v1, v_ref) < FTOL
        rng = np.random.default_rng(11)
        x1 = x0 + rng.uniform(-0.03, 0.03, x0.shape)                   # inside 80 % of margin/2 = 0.2: rows stay
        f2, e2, v2 = run(x1)
        assert ctx.counter("inner_rebuilds") == i0 and ctx.counter("pipelined_redos") == r0
        x2 = x1.copy()
        x2[5] = x0[5] + (0.156, 0.156, 0.0)         # |d| = 0.2206: "soon" (> 0.8 x 0.25), not stale (< margin/2 = 0.25)
        f3, e3, v3 = run(x2)
        assert ctx.counter("pipelined_redos") == r0 and ctx.counter("inner_rebuilds") == i0 + 1     # deferred re-derive
        x3 = x2.copy()
        x3[9] += (0.3, 0.3, 0.0)                    # |d| = 0.42 > margin/2 (< skin/2: the master list is still good)
        f4, e4, v4 = run(x3)
        assert ctx.counter("pipelined_redos") == r0 + 1
        assert ctx.counter("pipelined_calls") == p0 + 4
        ctx.set_option("h2d_chunks", 1)             # plain upload again
        ctx.set_option("d2h_chunks", 1)
        for x, (fp, ep, vp) in ((x1, (f2, e2, v2)), (x2, (f3, e3, v3)), (x3, (f4, e4, v4))):
            fq, eq, vq = run(x)
            assert S.rel_err(fp, fq) < 1e-12 and abs(ep - eq) < 1e-12 * abs(eq) and S.rel_err(vp, vq) < 1e-11
    finally:
        ctx.set_option("d2h_min_atoms", 65536)
        ctx.set_option("d2h_chunks", 4)
        ctx.set_option("h2d_chunks", 6)
    lmp.close()


def fold_ghost_rows(a, swaps, nlocal):
    a = np.array(a, dtype=np.float64, copy=True)
    for s in reversed(swaps):
        if s["recvnum"]:
            np.add.at(a, s["sendlist"], a[s["firstrecv"]:s["firstrecv"] + s["recvnum"]])
    return a[:nlocal]


@pytest.mark.parametrize("case", [CASES[1], CASES[3]], ids=[CASES[1]["id"], CASES[3]["id"]])
def test_per_atom_energy_and_virial(ctx, oracle_built, case):
    """Pair::eatom / Pair::vatom of pair_style aeam: embedding energy to the atom (one third for angular atoms,
    pair_aeam.cpp:295-300), phi/2 per visit (:389), ev_tally halves (:393), ev_tally3 thirds (:472) -- compared per
    atom with the reference plugin after folding ghost shares into their owners."""
    lmp = S.make_aeam_system(S.oracle_plugin("aeam"), case["cells"], si_fraction=case["si"], displace=case["displace"])
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    nl, nall = snap["nlocal"], snap["nlocal"] + snap["nghost"]
    lmp.compute(1 | 2, 2 | 4, reverse=True)
    ea_ref = fold_ghost_rows(lmp._arr("eatom", 0, nall, np.float64), snap["swaps"], nl)
    va_ref = fold_ghost_rows(lmp._arr("vatom", 0, nall, np.float64, 6), snap["swaps"], nl)
    f_ref = lmp.f()[:nl].copy()
    ctx.aeam_init(aeam_tables())
    ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
    f, e, v, ea, va = ctx.aeam_compute_peratom(snap["nlocal"], snap["nghost"], snap["x"], snap["type"], snap["tag"])
    ea, va = fold_ghost_rows(ea, snap["swaps"], nl), fold_ghost_rows(va, snap["swaps"], nl)
    print("\\n%s: sum eatom %.8f  max|eatom| %.4f err %.2e  max|vatom| %.4f err %.2e"
          % (case["id"], ea.sum(), np.abs(ea_ref).max(), np.abs(ea - ea_ref).max(), np.abs(va_ref).max(),
             np.abs(va - va_ref).max()))
    assert S.rel_err(ea, ea_ref) < 1e-10
    assert S.rel_err(va, va_ref) < 1e-10
    assert S.rel_err(S.fold_ghost_forces(f, snap["swaps"], nl), f_ref) < 1e-10
    lmp.close()


@pytest.mark.parametrize("case", [CASES[1], CASES[2], CASES[3]], ids=[CASES[1]["id"], CASES[2]["id"], CASES[3]["id"]])
def test_cluster_rows_equal_single_rows(ctx, oracle_built, case):
    """The three row forms agree: one row per center with f' handed from the density to the force pass (default,
    aeam_cluster = 2), the round-1 kernels (0), and the cluster form (1: 4 consecutive centers share one union row) in
    its lane layouts, with and without index-sorted rows."""
    lmp = S.make_aeam_system(S.oracle_plugin("aeam"), case["cells"], si_fraction=case["si"], displace=case["displace"])
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    f_ref, e_ref, v_ref = oracle_forces(lmp)
    ctx.aeam_init(aeam_tables())
    res = {}
    try:
        for name, opts in (("single", dict(aeam_cluster=0)), ("cluster", dict(aeam_cluster=1, aeam_sort_rows=0)),
                           ("cluster-sorted", dict(aeam_cluster=1, aeam_sort_rows=1)),
                           ("cluster-lpe1", dict(aeam_cluster=1, aeam_sort_rows=0, aeam_variant=11)),
                           ("cluster-lpe2", dict(aeam_cluster=1, aeam_sort_rows=0, aeam_variant=33)),
                           ("single-df", dict(aeam_cluster=2, aeam_variant=0))):
            for k, v in opts.items():
                ctx.set_option(k, v)
            res[name] = gpu_forces(ctx, snap)
            ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
            fo, eo, vo = ctx.aeam_compute(snap["nlocal"], snap["nghost"], snap["x"], snap["type"], snap["tag"], 0, 0)
            assert S.rel_err(S.fold_ghost_forces(fo, snap["swaps"], snap["nlocal"]), res[name][0]) < 1e-13
    finally:
        ctx.set_option("aeam_cluster", 2)
        ctx.set_option("aeam_sort_rows", 0)
        ctx.set_option("aeam_variant", 0)
    for name, (f, e, v) in res.items():
        print("\n%s %s: ferr vs reference %.2e" % (case["id"], name, S.rel_err(f, f_ref)))
        assert S.rel_err(f, f_ref) < FTOL and abs(e - e_ref) < ETOL * abs(e_ref) and S.rel_err(v, v_ref) < FTOL
        assert S.rel_err(f, res["single"][0]) < 1e-12
        assert abs(e - res["single"][1]) < 1e-12 * abs(e_ref)
    lmp.close()


@pytest.mark.parametrize("case", [CASES[1], CASES[3]], ids=[CASES[1]["id"], CASES[3]["id"]])
def test_deterministic_mode(ctx, oracle_built, case):
    """option "deterministic" for pair_style aeam: the angular triplet forces (pair_aeam.cpp:438-473) leave the warp once
    per staged neighbor and are added in fixed point (integer atomics: order-independent), the global energy / virial
    sums likewise.  Forces (owned and ghost), energy and virial are bitwise identical over repeated calls, meet the
    parity bar against the reference, and agree with the default path to rounding."""
    lmp = S.make_aeam_system(S.oracle_plugin("aeam"), case["cells"], si_fraction=case["si"], displace=case["displace"])
    lmp.setup(1, 2)
    snap = S.snapshot(lmp)
    f_ref, e_ref, v_ref = oracle_forces(lmp)
    ctx.aeam_init(aeam_tables())
    ctx.set_neighbor_csr(snap["inum"], snap["gnum"], snap["off"], snap["val"], snap["skin"])
    args = (snap["nlocal"], snap["nghost"], snap["x"], snap["type"], snap["tag"], 1, 2)
    fa, ea, va = ctx.aeam_compute(*args)
    try:
        ctx.set_option("deterministic", 1)
        runs = [ctx.aeam_compute(*args) for _ in range(4)]
    finally:
        ctx.set_option("deterministic", 0)
    for f, e, v in runs[1:]:
        assert np.array_equal(f, runs[0][0]), "deterministic forces differ between calls"
        assert e == runs[0][1] and np.array_equal(v, runs[0][2]), "deterministic energy / virial differ between calls"
    f, e, v = runs[0]
    assert S.rel_err(S.fold_ghost_forces(f, snap["swaps"], snap["nlocal"]), f_ref) < FTOL
    assert abs(e - e_ref) / abs(e_ref) < ETOL and S.rel_err(v, v_ref) < FTOL
    assert S.rel_err(f, fa) < 1e-12
    lmp.close()
