"""Direct oracle parity at BASELINE.json's FULL sizes: the reference plugin (USER-REBOMOS / USER-AEAM sources compiled
verbatim, oracle/_ref) inside the engine on thread-ranks of the box's host cores computes ONE Pair::compute() of
configs[2] (995 904 MoS2 atoms, thermally displaced) and of configs[3] (2 048 000 fcc Al atoms with 0.75 % Si,
displaced); the device path gets the same atoms and must return the same per-atom forces (matched by atom ID), total
energy and virial.  Tolerances as everywhere: forces 1e-10 of the largest force component, energy 1e-12 relative,
virial 1e-10 of its largest component.  Also printed: the worst per-atom RELATIVE error |df_i| / |f_i| over atoms with
|f_i| > 1e-3 max|f| (the max-norm alone would hide errors on atoms with small forces).

Reference: R:USER-REBOMOS/pair_rebomos.cpp:102-111, R:USER-AEAM/pair_aeam.cpp:110-479."""
import os

import numpy as np
import pytest

import lammps_plugins_b200 as b2
import support as S

pytestmark = pytest.mark.gpu


def thread_grid():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for cand, g in ((16, (4, 2, 2)), (8, (2, 2, 2)), (4, (2, 2, 1)), (2, (2, 1, 1))):
        if n >= cand:
            return g
    return (1, 1, 1)


def gather_reference(lmp, natoms):
    """owned atoms of every thread-rank by atom ID: x, type, f after reverse communication; summed energy / virial"""
    R = lmp.nprocs
    X, F = np.zeros((natoms, 3)), np.zeros((natoms, 3))
    T = np.zeros(natoms, dtype=np.int32)
    seen = np.zeros(natoms, dtype=np.int64)
    for r in range(R):
        nl = lmp.get_int("nlocal", r)
        idx = lmp.tag(r)[:nl] - 1
        X[idx], F[idx], T[idx] = lmp.x(r, nl), lmp.f(r, nl), lmp.type(r)[:nl]
        np.add.at(seen, idx, 1)
    assert (seen == 1).all()
    # eng_vdwl / virial are per-rank tallies (LAMMPS sums them in compute pe / pressure)
    e = sum(lmp.get_double("eng_vdwl", r) for r in range(R))
    v = np.array([sum(lmp.get_double("virial%d" % k, r) for r in range(R)) for k in range(6)])
    return X, T, F, e, v


def report(name, f, f_ref):
    fmax = np.abs(f_ref).max()
    ferr = np.abs(f - f_ref).max() / fmax
    mag = np.linalg.norm(f_ref, axis=1)
    big = mag > 1e-3 * mag.max()
    rel = np.linalg.norm(f - f_ref, axis=1)[big] / mag[big]
    print("\n%s: %d atoms, max|f| %.4g, max-norm error %.3e, worst per-atom relative error %.3e over the %d atoms with "
          "|f_i| > 1e-3 max|f|" % (name, len(f), fmax, ferr, rel.max(), int(big.sum())))
    return ferr, float(rel.max())


def device_forces(ctx, style, lmp, X, T, natoms):
    d = lmp.box()
    box = b2.make_box(d["boxlo"], d["boxhi"], d["xy"], d["xz"], d["yz"], triclinic=d["triclinic"])
    tag = np.arange(1, natoms + 1, dtype=np.int32)
    ctx.system_create(style, lmp.get_int("ntypes"), lmp.mass(), box, X, np.zeros_like(X), T, tag, lmp.get_double("skin"),
                      0.001, lmp.units(), sort_every=1000)
    row = ctx.system_thermo_rows()[0]
    st = ctx.system_download()
    nl = st["nlocal"]
    assert nl == natoms
    f = np.empty((natoms, 3))
    f[st["tag"][:nl] - 1] = st["f"][:nl]
    return f, row["pe"], np.array(row["virial"])


def test_rebomos_995904_atoms_against_the_reference_plugin(ctx, oracle_built):
    grid = thread_grid()
    lmp = S.make_rebomos_system(S.oracle_plugin("rebomos"), replicate=(14, 13, 19), grid=grid, displace=0.08, seed=30018)
    natoms = lmp.get_int("natoms")
    assert natoms == 995904
    lmp.setup(1, 2)
    lmp.compute(1, 2, reverse=True)
    X, T, f_ref, e_ref, v_ref = gather_reference(lmp, natoms)
    ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
    f, e, v = device_forces(ctx, "rebomos", lmp, X, T, natoms)
    ferr, rel = report("rebomos 14x13x19, displaced 0.08 A, reference on %dx%dx%d thread-ranks" % grid, f, f_ref)
    print("E %.6f vs %.6f (rel %.2e); virial rel %.2e" % (e, e_ref, abs(e - e_ref) / abs(e_ref),
                                                             np.abs(v - v_ref).max() / np.abs(v_ref).max()))
    assert ferr < 1e-10 and rel < 1e-8
    assert abs(e - e_ref) < 1e-12 * abs(e_ref)
    assert np.abs(v - v_ref).max() < 1e-10 * np.abs(v_ref).max()
    lmp.close()


def test_aeam_2048000_atoms_against_the_reference_plugin(ctx, oracle_built):
    grid = thread_grid()
    lmp = S.make_aeam_system(S.oracle_plugin("aeam"), (80, 80, 80), grid=grid, si_fraction=0.0075, displace=0.12, seed=86318)
    natoms = lmp.get_int("natoms")
    assert natoms == 2048000
    lmp.setup(1, 2)
    lmp.compute(1, 2, reverse=True)
    X, T, f_ref, e_ref, v_ref = gather_reference(lmp, natoms)
    assert 0.006 < (T == 2).mean() < 0.009
    t = S.load_aeam_fixture()
    ctx.aeam_init({k: t[k] for k in ("nelements", "nnonangular", "nrho", "drho", "nr", "dr", "cut", "frho", "rhor", "z2r")})
    f, e, v = device_forces(ctx, "aeam", lmp, X, T, natoms)
    ferr, rel = report("aeam 80^3 cells, 0.75 %% Si, displaced 0.12 A, reference on %dx%dx%d thread-ranks" % grid, f, f_ref)
    print("E %.6f vs %.6f (rel %.2e); virial rel %.2e" % (e, e_ref, abs(e - e_ref) / abs(e_ref),
                                                             np.abs(v - v_ref).max() / np.abs(v_ref).max()))
    assert ferr < 1e-10 and rel < 1e-8
    assert abs(e - e_ref) < 1e-12 * abs(e_ref)
    assert np.abs(v - v_ref).max() < 1e-10 * np.abs(v_ref).max()
    lmp.close()
