"""Test support: ctypes driver of the minilmp engine (oracle side), potential-file
writers from the committed fixtures, and helpers to compare the CUDA path against
the oracle.  Nothing here is imported by the product package."""
from __future__ import annotations

import ctypes
import json
import math
import os
import subprocess
import tempfile
from ctypes import POINTER, c_char_p, c_double, c_int, c_longlong, c_void_p

import numpy as np

TESTS = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(TESTS)
GOLDEN = os.path.join(TESTS, "golden")
ORACLE = os.path.join(REPO, "oracle")
REFERENCE = os.environ.get("B200MD_REFERENCE", "/root/reference")

MINILMP_SO = os.path.join(ORACLE, "libminilmp.so")
REF_REBOMOS_SO = os.path.join(ORACLE, "_ref", "rebomosplugin.so")
REF_AEAM_SO = os.path.join(ORACLE, "_ref", "aeamplugin.so")
PORT_SO = os.path.join(ORACLE, "portplugin.so")
B200_REBOMOS_SO = os.path.join(REPO, "lammps_plugins_b200", "rebomosplugin.so")
B200_AEAM_SO = os.path.join(REPO, "lammps_plugins_b200", "aeamplugin.so")


def have_reference_tree():
    return os.path.isdir(os.path.join(REFERENCE, "USER-REBOMOS"))


def oracle_plugin(style):
    """The oracle plugin of a pair style: the reference sources compiled verbatim (oracle/_ref, built where
    /root/reference exists and shipped with the snapshot).  A test that needs the reference FAILS when it is missing;
    the port is only used where a test names it (B200MD_ALLOW_PORT=1 lets a box without _ref run the suite anyway)."""
    return reference_plugin(style, allow_port=os.environ.get("B200MD_ALLOW_PORT") == "1")


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE, "all"], check=True)


# ----------------------------------------------------------------------------- minilmp
sys_path_added = REPO not in __import__("sys").path and __import__("sys").path.insert(0, REPO)
from oracle.minilmp import LammpsError, MiniLmp, reference_plugin  # noqa: E402,F401


def fold_ghost_forces(f, swaps, nlocal):
    """Reverse communication on one rank with self-swaps only (1x1x1 grid): add each ghost's force to
    the atom it was copied from, swaps in reverse order (CommBrick::reverse_comm).  Returns f[:nlocal]."""
    f = np.array(f, dtype=np.float64, copy=True)
    for s in reversed(swaps):
        if s["recvnum"] == 0:
            continue
        first = s["firstrecv"]
        np.add.at(f, s["sendlist"], f[first:first + s["recvnum"]])
    return f[:nlocal]


def fmt8(v):
    """LAMMPS thermo prints %.8g"""
    return "%.8g" % v


# ----------------------------------------------------------------------------- potentials from fixtures
def load_rebomos_fixture():
    d = json.load(open(os.path.join(GOLDEN, "rebomos_set5b.json")))
    return d["header"], [(n, float(v)) for n, v in d["params"]]


def write_rebomos_file(path):
    """Emit the fixture in the 'value name' format PotentialFileReader::next_double parses
    (reference: USER-REBOMOS/MoS.REBO.set5b:1-65)."""
    header, params = load_rebomos_fixture()
    with open(path, "w") as fh:
        fh.write(header.rstrip("\n") + "\n")
        fh.write("# re-emitted from tests/golden/rebomos_set5b.json\n\n")
        for name, v in params:
            fh.write("%-24s %s\n" % (repr(v), name))
    return path


def rebomos_params_struct():
    """b200md_rebomos_params filled the way PairREBOMoS::read_file does (pair_rebomos.cpp:964-1066)."""
    from lammps_plugins_b200 import RebomosParams
    _, params = load_rebomos_fixture()
    v = [p[1] for p in params]
    it = iter(v)
    nx = lambda: next(it)
    P = RebomosParams()

    def sym(a, mm, ms, ss):
        a[0], a[1], a[2], a[3] = mm, ms, ms, ss

    rcmin = (nx(), nx(), nx()); rcmax = (nx(), nx(), nx())
    Q = (nx(), nx(), nx()); al = (nx(), nx(), nx()); A = (nx(), nx(), nx())
    B = (nx(), nx(), nx()); be = (nx(), nx(), nx())
    sym(P.rcmin, *rcmin); sym(P.rcmax, *rcmax); sym(P.Q, *Q); sym(P.alpha, *al); sym(P.A, *A)
    sym(P.BIJc, *B); sym(P.Beta, *be)
    for o in range(7):
        P.b[o][0] = nx()
    for o in range(7):
        P.bg[o][0] = nx()
    for o in range(7):
        P.b[o][1] = nx()
    for o in range(7):
        P.bg[o][1] = nx()
    for o in range(4):
        P.a[o][0] = nx()
    for o in range(4):
        P.a[o][1] = nx()
    eps_mm, eps_ss, sig_mm, sig_ss = nx(), nx(), nx(), nx()
    sig_ms = (sig_mm + sig_ss) / 2
    eps_ms = math.sqrt(eps_mm * eps_ss)
    sym(P.sigma, sig_mm, sig_ms, sig_ss)
    sym(P.epsilon, eps_mm, eps_ms, eps_ss)
    sym(P.rcLJmin, *rcmin)
    sym(P.rcLJmax, 2.5 * sig_mm, 2.5 * sig_ms, 2.5 * sig_ss)
    return P


def load_aeam_fixture():
    z = np.load(os.path.join(GOLDEN, "alsi_aeam.npz"))
    nel = int(z["nelements"])
    nrho = [int(v) for v in z["nrho"]]
    nr = np.array(z["nr"]).reshape(nel, nel)
    vals = z["values"]
    k = 0
    frho, rhor, z2r = [], [[None] * nel for _ in range(nel)], [[None] * nel for _ in range(nel)]
    for i in range(nel):
        frho.append(vals[k:k + nrho[i]].copy()); k += nrho[i]
    for i in range(nel):
        for j in range(nel):
            rhor[i][j] = vals[k:k + nr[i, j]].copy(); k += nr[i, j]
    for i in range(nel):
        for j in range(i + 1):
            z2r[i][j] = vals[k:k + nr[i, j]].copy(); k += nr[i, j]
    assert k == len(vals)
    return dict(nelements=nel, nnonangular=int(z["nnonangular"]), nangular=int(z["nangular"]),
                names=[str(s) for s in z["names"]], nrho=nrho, drho=[float(v) for v in z["drho"]],
                mass=[float(v) for v in z["mass"]], nr=nr, dr=np.array(z["dr"]).reshape(nel, nel),
                cut=np.array(z["cut"]).reshape(nel, nel), frho=frho, rhor=rhor, z2r=z2r,
                header=[str(s) for s in z["header"]], values=vals)


def write_aeam_file(path):
    """Emit the fixture in the AEAM setfl-like format PairAEAM::read_file parses
    (reference: USER-AEAM/AlSi.aeam:1-18 + values 5 per line; pair_aeam.cpp:627-746)."""
    t = load_aeam_fixture()
    nel = t["nelements"]
    with open(path, "w") as fh:
        for ln in t["header"][:11]:
            fh.write(ln.rstrip("\n") + "\n")
        fh.write("%d %d %d %s\n" % (nel, t["nnonangular"], t["nangular"], " ".join(t["names"])))
        for i in range(nel):
            fh.write("\t%d\t%s\t%s\t%s\n" % (t["nrho"][i], repr(t["drho"][i]), repr(t["mass"][i]), t["names"][i]))
        for i in range(nel):
            for j in range(nel):
                fh.write("\t%d\t%s\t%s\t%s %s\n" % (t["nr"][i, j], repr(float(t["dr"][i, j])),
                                                   repr(float(t["cut"][i, j])), t["names"][i], t["names"][j]))
        vals = t["values"]
        for k in range(0, len(vals), 5):
            fh.write("\t".join(repr(float(v)) for v in vals[k:k + 5]) + "\n")
    return path


_tmpdir = None


def potential_dir():
    """Directory holding MoS.REBO.set5b and AlSi.aeam re-emitted from the fixtures."""
    global _tmpdir
    if _tmpdir is None:
        _tmpdir = tempfile.mkdtemp(prefix="b200md_pot_")
        write_rebomos_file(os.path.join(_tmpdir, "MoS.REBO.set5b"))
        write_aeam_file(os.path.join(_tmpdir, "AlSi.aeam"))
    return _tmpdir


# ----------------------------------------------------------------------------- standard systems
def input_script(name):
    return json.load(open(os.path.join(GOLDEN, "input_scripts.json")))[name]


def rebomos_bulk_commands(replicate=(1, 1, 1), pot=None):
    """The shipped in.rebomos-bulk up to (not including) thermo/fix/run, optionally replicated."""
    pot = pot or os.path.join(potential_dir(), "MoS.REBO.set5b")
    cmds = []
    for c in input_script("in.rebomos-bulk"):
        w = c.split()
        if w[0] in ("thermo_style", "thermo", "fix", "run"):
            continue
        if w[0] == "pair_coeff":
            c = "pair_coeff * * %s M S" % pot
        if w[0] == "mass" and replicate != (1, 1, 1) and not any(x.startswith("replicate") for x in cmds):
            cmds.append("replicate %d %d %d" % tuple(replicate))
        cmds.append(c)
    return cmds


def make_rebomos_system(plugin, replicate=(1, 1, 1), grid=(1, 1, 1), displace=0.0, seed=12345, extra=()):
    lmp = MiniLmp(grid)
    lmp.command("plugin load " + plugin)
    lmp.commands(rebomos_bulk_commands(replicate))
    if displace > 0.0:
        lmp.command("displace_atoms all random %g %g %g %d" % (displace, displace, displace, seed))
    lmp.commands(extra)
    return lmp


def aeam_commands(cells=(4, 4, 4), si_fraction=0.0075, seed=7683797, pot=None, skin=1.0):
    pot = pot or os.path.join(potential_dir(), "AlSi.aeam")
    return [
        "units metal", "atom_style atomic", "dimension 3", "boundary p p p",
        "lattice fcc 4.045",
        "region MeSi block 0 %d 0 %d 0 %d" % tuple(cells),
        "create_box 2 MeSi", "create_atoms 1 region MeSi",
        "pair_style aeam", "pair_coeff * * %s Al Si" % pot,
        "neighbor %g bin" % skin, "neigh_modify every 1 delay 1 check yes",
        "set region MeSi type/fraction 2 %g %d" % (si_fraction, seed),
    ]


def make_aeam_system(plugin, cells=(4, 4, 4), grid=(1, 1, 1), si_fraction=0.0075, displace=0.0, seed=4711,
                     extra=()):
    lmp = MiniLmp(grid)
    lmp.command("plugin load " + plugin)
    lmp.commands(aeam_commands(cells, si_fraction))
    if displace > 0.0:
        lmp.command("displace_atoms all random %g %g %g %d" % (displace, displace, displace, seed))
    lmp.commands(extra)
    return lmp


def snapshot(lmp, rank=0):
    """Copy of everything a pair compute consumes on one rank."""
    nlocal, nghost = lmp.get_int("nlocal", rank), lmp.get_int("nghost", rank)
    off, val = lmp.neigh_csr(rank)
    return dict(nlocal=nlocal, nghost=nghost, x=lmp.x(rank).copy(), type=lmp.type(rank).copy(),
                tag=lmp.tag(rank).copy(), inum=lmp.get_int("inum", rank), gnum=lmp.get_int("gnum", rank),
                off=off, val=val, skin=lmp.get_double("skin", rank), swaps=lmp.swaps(rank), box=lmp.box())


def rel_err(a, b):
    """max |a-b| / max|b| -- the '1e-10 relative' of the north star is relative to the largest force."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = float(np.max(np.abs(b))) if b.size else 0.0
    if den == 0.0:
        den = 1.0
    return float(np.max(np.abs(a - b))) / den if a.size else 0.0
