"""Synthetic workload generators for the benchmark driver and tests.

They restate, in numpy, what the two shipped inputs do with LAMMPS commands, so that the GPU-resident
system can be fed the BASELINE configurations without a LAMMPS executable:

* ``mos2_bulk(nx, ny, nz)``  -- ``lattice custom`` + ``region prism 0 4 0 8 0 1 -2 0 0`` + ``create_atoms``
  of ``USER-REBOMOS/in.rebomos-bulk:3-22`` (288-atom triclinic 2H-MoS2 cell), followed by
  ``replicate nx ny nz`` (image order ix, iy, iz with iz fastest; atom IDs offset per image).
* ``fcc_alsi(cells, si_fraction, seed)`` -- ``lattice fcc 4.045`` + ``region block`` + ``create_atoms`` of
  ``USER-AEAM/sample.in:7-10``, with a numpy RNG choosing the Si sites (LAMMPS' ``set type/fraction``
  RNG is restated only in the test engine; SURVEY.md 8(d) config 2 allows the generated equivalent).
* ``maxwell_velocities`` -- Gaussian velocities at temperature T, zero total momentum, rescaled to T
  exactly with dof = 3N-3 (metal units).

Pure host-side data generation; no force or neighbor arithmetic lives here.
"""
from __future__ import annotations

import numpy as np

# in.rebomos-bulk:3-13
MOS2_A1 = np.array([3.1903157234, 0.0, 0.0])
MOS2_A2 = np.array([-1.5964590311, 2.7651481541, 0.0])
MOS2_A3 = np.array([0.0, 0.0, 13.9827680588])
MOS2_BASIS = np.array([
    [0.0, 0.0, 3.0 / 4.0],
    [0.0, 0.0, 1.0 / 4.0],
    [2.0 / 3.0, 1.0 / 3.0, 0.862008989],
    [1.0 / 3.0, 2.0 / 3.0, 0.137990996],
    [1.0 / 3.0, 2.0 / 3.0, 0.362008989],
    [2.0 / 3.0, 1.0 / 3.0, 0.637991011],
])
MOS2_BASIS_TYPE = np.array([1, 1, 2, 2, 2, 2], dtype=np.int32)   # create_atoms ... basis 1 1 basis 2 1 basis 3..6 2
MOS2_ORIGIN = np.array([0.1, 0.1, 0.1])
MOS2_MASS = np.array([0.0, 95.95, 32.065])
ALSI_MASS = np.array([0.0, 27.0, 28.0])

BOLTZ = 8.617343e-5
MVV2E = 1.0364269e-4


def _lattice_spacings(a1, a2, a3):
    """Lattice::setup: spacings = bounding box of the unit cell"""
    corners = np.array([i * a1 + j * a2 + k * a3 for k in (0, 1) for j in (0, 1) for i in (0, 1)])
    return corners.max(axis=0) - corners.min(axis=0)


def mos2_unit_box():
    """Box of the shipped input: region prism 0 4 0 8 0 1 -2.0 0 0 in lattice units."""
    sp = _lattice_spacings(MOS2_A1, MOS2_A2, MOS2_A3)
    boxlo = np.zeros(3)
    boxhi = np.array([4 * sp[0], 8 * sp[1], 1 * sp[2]])
    xy = -2.0 * sp[0]
    return boxlo, boxhi, xy, sp


def mos2_cell():
    """The 288 atoms of in.rebomos-bulk in create_atoms order (k, j, i, basis), IDs 1..288."""
    boxlo, boxhi, xy, sp = mos2_unit_box()
    prd = boxhi - boxlo
    shift = sp * MOS2_ORIGIN
    pts, types = [], []
    # generous loop bounds; membership is decided in lamda space exactly like CreateAtoms::loop_lattice
    for k in range(-1, 3):
        for j in range(-2, 12):
            for i in range(-8, 12):
                for m in range(6):
                    f = np.array([i, j, k], dtype=np.float64) + MOS2_BASIS[m]
                    x = f[0] * MOS2_A1 + f[1] * MOS2_A2 + f[2] * MOS2_A3 + shift
                    lam1 = (x[1] - boxlo[1]) / prd[1]
                    lam0 = (x[0] - boxlo[0]) / prd[0] - xy / (prd[0] * prd[1]) * (x[1] - boxlo[1])
                    lam2 = (x[2] - boxlo[2]) / prd[2]
                    if 0.0 <= lam0 < 1.0 - 2e-6 and 0.0 <= lam1 < 1.0 - 2e-6 and 0.0 <= lam2 < 1.0 - 2e-6:
                        pts.append(x)
                        types.append(MOS2_BASIS_TYPE[m])
    x = np.array(pts)
    t = np.array(types, dtype=np.int32)
    assert len(x) == 288, len(x)
    return x, t, np.arange(1, 289, dtype=np.int32)


def mos2_bulk(nx=1, ny=1, nz=1):
    """Replicated MoS2 bulk: returns dict(x, type, tag, boxlo, boxhi, xy, xz, yz, mass)."""
    x0, t0, g0 = mos2_cell()
    boxlo, boxhi, xy, _ = mos2_unit_box()
    prd = boxhi - boxlo
    n0 = len(x0)
    nimg = nx * ny * nz
    x = np.empty((nimg * n0, 3))
    t = np.empty(nimg * n0, dtype=np.int32)
    g = np.empty(nimg * n0, dtype=np.int64)
    c = 0
    for ix in range(nx):
        for iy in range(ny):
            for iz in range(nz):
                sl = slice(c * n0, (c + 1) * n0)
                x[sl, 0] = x0[:, 0] + ix * prd[0] + iy * xy
                x[sl, 1] = x0[:, 1] + iy * prd[1]
                x[sl, 2] = x0[:, 2] + iz * prd[2]
                t[sl] = t0
                g[sl] = g0 + (iz * ny * nx + iy * nx + ix) * n0
                c += 1
    assert g.max() < 2 ** 31
    new_hi = boxlo + prd * np.array([nx, ny, nz])
    return dict(x=x, type=t, tag=g.astype(np.int32), boxlo=boxlo, boxhi=new_hi, xy=ny * xy, xz=0.0, yz=0.0,
                mass=MOS2_MASS.copy(), ntypes=2, triclinic=1)


def fcc_alsi(cells=(20, 20, 20), si_fraction=0.0075, seed=7683797, a=4.045):
    """fcc Al with a random fraction of Si (type 2); create_atoms order (k, j, i, basis), IDs 1..N."""
    nx, ny, nz = cells
    basis = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5]])
    k, j, i, m = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), np.arange(4), indexing="ij")
    cell = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(np.float64)
    x = (cell + basis[m.ravel()]) * a
    n = len(x)
    rng = np.random.default_rng(seed)
    t = np.where(rng.random(n) < si_fraction, 2, 1).astype(np.int32)
    return dict(x=x, type=t, tag=np.arange(1, n + 1, dtype=np.int32), boxlo=np.zeros(3),
                boxhi=np.array([nx * a, ny * a, nz * a]), xy=0.0, xz=0.0, yz=0.0, mass=ALSI_MASS.copy(),
                ntypes=2, triclinic=0)


def maxwell_velocities(types, mass, temperature, seed):
    """N(0, sqrt(kB T / m)) velocities (metal units: A/ps), momentum zeroed, rescaled to T with dof 3N-3."""
    n = len(types)
    rng = np.random.default_rng(seed)
    m = np.asarray(mass)[types]
    v = rng.normal(size=(n, 3)) * np.sqrt(BOLTZ * temperature / (m * MVV2E))[:, None]
    v -= (v * m[:, None]).sum(axis=0) / m.sum()
    dof = 3 * n - 3
    t_now = (m[:, None] * v * v).sum() * MVV2E / (dof * BOLTZ)
    if t_now > 0:
        v *= np.sqrt(temperature / t_now)
    return v


def brick_owner(x, boxlo, boxhi, xy, xz, yz, procgrid):
    """Rank owning each atom under the uniform brick decomposition (lamda space, x fastest)."""
    prd = np.asarray(boxhi) - np.asarray(boxlo)
    d = x - np.asarray(boxlo)
    l1 = (d[:, 1] - yz / prd[2] * d[:, 2]) / prd[1]
    l2 = d[:, 2] / prd[2]
    l0 = (d[:, 0] - xy * l1 - xz * l2) / prd[0]
    lam = np.stack([l0, l1, l2], axis=1)
    lam -= np.floor(lam)
    idx = np.minimum((lam * np.asarray(procgrid)).astype(np.int64), np.asarray(procgrid) - 1)
    return idx[:, 2] * procgrid[1] * procgrid[0] + idx[:, 1] * procgrid[0] + idx[:, 0]


# ---------------------------------------------------------------------------------------------------------
# what LAMMPS-core would hand to a pair style on ONE rank: neighbor cutoffs and the ghost cutoff of the box
def rebomos_neighbor_cutoffs(rcmax_2x2, type_map, skin):
    """(cutneighsq, cutneighghostsq, cutneighmax) as Neighbor::init builds them for pair_style rebomos:
    init_one returns cut3rebo = 3*rcmax[Mo][Mo] for every pair, cutghost[i][j] = rcmax[map i][map j]."""
    nt = len(type_map)
    cs = np.zeros((nt + 1, nt + 1))
    cg = np.zeros((nt + 1, nt + 1))
    cut3rebo = 3.0 * rcmax_2x2[0]
    cmax = 0.0
    for i in range(1, nt + 1):
        for j in range(1, nt + 1):
            cn = float(np.sqrt(cut3rebo * cut3rebo)) + skin
            cs[i, j] = cn * cn
            g = rcmax_2x2[type_map[i - 1] * 2 + type_map[j - 1]] + skin
            cg[i, j] = g * g
            cmax = max(cmax, cn)
    return cs, cg, cmax


def aeam_neighbor_cutoffs(cut, skin):
    """Neighbor::init for pair_style aeam: init_one(i,j) = cut[i-1][j-1] for i <= j, mirrored."""
    cut = np.asarray(cut)
    nt = cut.shape[0]
    cs = np.zeros((nt + 1, nt + 1))
    cmax = 0.0
    for i in range(1, nt + 1):
        for j in range(1, nt + 1):
            a, b = (i, j) if i <= j else (j, i)
            c = float(cut[a - 1, b - 1])
            cn = float(np.sqrt(c * c)) + skin
            cs[i, j] = cn * cn
            cmax = max(cmax, cn)
    return cs, cs.copy(), cmax


def single_rank_box(w, cutneighmax):
    """b200md_box of a 1x1x1 decomposition: sub-domain = whole box, Comm::cutghost per CommBrick::setup."""
    from . import make_box
    b = make_box(w["boxlo"], w["boxhi"], w["xy"], w["xz"], w["yz"], triclinic=w["triclinic"])
    prd = np.asarray(w["boxhi"]) - np.asarray(w["boxlo"])
    if w["triclinic"]:
        h_inv = [1.0 / prd[0], 1.0 / prd[1], 1.0 / prd[2], -w["yz"] / (prd[1] * prd[2]),
                 (w["yz"] * w["xy"] - prd[1] * w["xz"]) / (prd[0] * prd[1] * prd[2]), -w["xy"] / (prd[0] * prd[1])]
        cg = [cutneighmax * np.sqrt(h_inv[0] ** 2 + h_inv[5] ** 2 + h_inv[4] ** 2),
              cutneighmax * np.sqrt(h_inv[1] ** 2 + h_inv[3] ** 2), cutneighmax * h_inv[2]]
        lo, hi = [0.0] * 3, [1.0] * 3
    else:
        cg = [cutneighmax] * 3
        lo, hi = list(w["boxlo"]), list(w["boxhi"])
    for d in range(3):
        b.sublo[d], b.subhi[d], b.cutghost[d] = lo[d], hi[d], cg[d]
    b.cutneighmax = cutneighmax
    return b


# ---------------------------------------------------------------------------------------------------------
# The same workloads as LAMMPS input commands (for a host application that loads the plugin: LAMMPS itself, or the
# mini engine the benchmark's CPU arm and drop-in leg use).  They restate the two shipped inputs:
# USER-REBOMOS/in.rebomos-bulk:1-22 and USER-AEAM/sample.in:1-19.
def rebomos_bulk_script(potential_file, replicate=(1, 1, 1), cells=None):
    """in.rebomos-bulk up to (not including) thermo/fix/run.  `replicate` appends LAMMPS' replicate command;
    `cells` = (rx, ry, rz) instead builds one box of rx x ry x rz shipped cells (region prism scaled, tilt kept
    proportional) -- what the CPU arm uses so that a brick grid divides the box evenly."""
    b = "basis 0.0000000000 0.000000000 $(3.0/4.0) basis 0.0000000000 0.000000000 $(1.0/4.0) " \
        "basis $(2.0/3.0) $(1.0/3.0) 0.862008989 basis $(1.0/3.0) $(2.0/3.0) 0.137990996 " \
        "basis $(1.0/3.0) $(2.0/3.0) 0.362008989 basis $(2.0/3.0) $(1.0/3.0) 0.637991011"
    region = "region box prism 0 4 0 8 0 1 -2.0 0.0 0.0"
    if cells is not None:
        rx, ry, rz = cells
        region = "region box prism 0 %d 0 %d 0 %d %g 0.0 0.0" % (4 * rx, 8 * ry, rz, -2.0 * ry)
    cmds = ["units metal",
            "lattice custom 1.0 a1 3.1903157234 0.0000000000 0.0000000000 a2 -1.5964590311 2.7651481541 0.0000000000 "
            "a3 0.0000000000 0.0000000000 13.9827680588 " + b + " origin 0.1 0.1 0.1",
            region, "create_box 2 box",
            "create_atoms 2 box basis 1 1 basis 2 1 basis 3 2 basis 4 2 basis 5 2 basis 6 2"]
    if tuple(replicate) != (1, 1, 1):
        cmds.append("replicate %d %d %d" % tuple(replicate))
    cmds += ["mass 1 95.95", "mass 2 32.065", "pair_style rebomos", "pair_coeff * * %s M S" % potential_file]
    return cmds


def aeam_script(potential_file, cells=(20, 20, 20), si_fraction=0.0075, seed=7683797, skin=1.0):
    """sample.in up to (not including) fix/thermo/velocity/run"""
    return ["units metal", "atom_style atomic", "dimension 3", "boundary p p p", "lattice fcc 4.045",
            "region MeSi block 0 %d 0 %d 0 %d" % tuple(cells), "create_box 2 MeSi", "create_atoms 1 region MeSi",
            "pair_style aeam", "pair_coeff * * %s Al Si" % potential_file, "neighbor %g bin" % skin,
            "neigh_modify every 1 delay 1 check yes", "set region MeSi type/fraction 2 %g %d" % (si_fraction, seed)]
