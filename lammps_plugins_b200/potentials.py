"""Readers of the two potential-file formats, for callers that drive the C ABI directly (bench.py, tests, scripts).

They follow the host pair classes -- and therefore the reference -- line for line:

* ``read_rebomos(path)``: ``PairREBOMoS::read_file`` (reference USER-REBOMOS/pair_rebomos.cpp:857-1066;
  here host/pair_rebomos.cpp ``read_file``).  Line 1 is the header (DATE/UNITS tags); every following non-blank,
  non-comment line contributes its FIRST token (the trailing name such as ``rcmin_MM`` is ignored, like
  ``PotentialFileReader::next_double``); 61 values in the order rcmin MM MS SS, rcmax, Q, alpha, A, BIJc, Beta,
  b0..b6 (Mo), bg0..bg6 (Mo), b0..b6 (S), bg0..bg6 (S), a0..a3 (Mo), a0..a3 (S), epsilon MM SS, sigma MM SS;
  mixed LJ terms sigma_MS = (sigma_M + sigma_S)/2, epsilon_MS = sqrt(eps_M eps_S), rcLJmin = rcmin,
  rcLJmax = 2.5 sigma (:1048-1056).
* ``read_aeam(path)``: ``PairAEAM::read_file`` (reference USER-AEAM/pair_aeam.cpp:627-746): 11 comment lines, line 12
  ``nelements nnonangular nangular names...``, one ``nrho drho mass`` line per element, one ``nr dr cut`` line per
  ordered element pair, then the tabulated values in free format: frho per element, rhor per ordered pair, z2r for
  j <= i.

The package ships value-identical copies of the two files the reference distributes (``potentials/MoS.REBO.set5b``,
``potentials/AlSi.aeam``); ``default_path`` names them.
"""
from __future__ import annotations

import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def default_path(name: str) -> str:
    """Path of a potential file shipped with the package ("MoS.REBO.set5b" or "AlSi.aeam")."""
    p = os.path.join(HERE, "potentials", name)
    if not os.path.exists(p):
        raise FileNotFoundError("potential file %s is not shipped with the package" % name)
    return p


def rebomos_values(path: str):
    """The numeric entries of a REBOMoS parameter file in file order (header line and comments skipped)."""
    vals = []
    with open(path) as fh:
        fh.readline()                                   # header: DATE / UNITS tags
        for ln in fh:
            ln = ln.split("#", 1)[0].strip()
            if ln:
                vals.append(float(ln.split()[0]))       # ValueError on a non-numeric token, like TokenizerException
    if len(vals) < 61:
        raise ValueError("%s: premature end of file (REBOMoS needs 61 parameters, found %d)" % (path, len(vals)))
    return vals


def read_rebomos(path: str | None = None):
    """b200md_rebomos_params filled the way PairREBOMoS::read_file fills the class members."""
    from . import RebomosParams
    it = iter(rebomos_values(path or default_path("MoS.REBO.set5b")))
    nx = lambda: next(it)
    P = RebomosParams()

    def sym(a, mm, ms, ss):
        a[0], a[1], a[2], a[3] = mm, ms, ms, ss

    rcmin = (nx(), nx(), nx())
    sym(P.rcmin, *rcmin)
    for arr in (P.rcmax, P.Q, P.alpha, P.A, P.BIJc, P.Beta):
        sym(arr, nx(), nx(), nx())
    for elem in (0, 1):
        for o in range(7):
            P.b[o][elem] = nx()
        for o in range(7):
            P.bg[o][elem] = nx()
    for elem in (0, 1):
        for o in range(4):
            P.a[o][elem] = nx()
    eps_mm, eps_ss, sig_mm, sig_ss = nx(), nx(), nx(), nx()
    sig_ms = (sig_mm + sig_ss) / 2
    eps_ms = math.sqrt(eps_mm * eps_ss)
    sym(P.sigma, sig_mm, sig_ms, sig_ss)
    sym(P.epsilon, eps_mm, eps_ms, eps_ss)
    sym(P.rcLJmin, *rcmin)
    sym(P.rcLJmax, 2.5 * sig_mm, 2.5 * sig_ms, 2.5 * sig_ss)
    return P


def read_aeam(path: str | None = None):
    """Tables of an AEAM file as the dict ``Context.aeam_init`` takes (plus names and masses)."""
    path = path or default_path("AlSi.aeam")
    with open(path) as fh:
        lines = fh.read().split("\n")
    if len(lines) < 13:
        raise ValueError("%s: not an AEAM potential file" % path)
    head = lines[11].split()
    nel, nna, nang = int(head[0]), int(head[1]), int(head[2])
    names = head[3:3 + nel]
    if nna + nang != nel or len(names) != nel:
        raise ValueError("%s: line 12 must read `nelements nnonangular nangular names...`" % path)
    k = 12
    nrho, drho, mass = [], [], []
    for _ in range(nel):
        w = lines[k].split()
        k += 1
        nrho.append(int(w[0]))
        drho.append(float(w[1]))
        mass.append(float(w[2]))
    nr = np.zeros((nel, nel), dtype=np.int64)
    dr = np.zeros((nel, nel))
    cut = np.zeros((nel, nel))
    for i in range(nel):
        for j in range(nel):
            w = lines[k].split()
            k += 1
            nr[i, j], dr[i, j], cut[i, j] = int(w[0]), float(w[1]), float(w[2])
    body = "\n".join(ln.split("#", 1)[0] for ln in lines[k:])
    vals = np.array(body.split(), dtype=np.float64)
    need = sum(nrho) + int(nr.sum()) + sum(int(nr[i, j]) for i in range(nel) for j in range(i + 1))
    if len(vals) < need:
        raise ValueError("%s: premature end of file (%d tabulated values, need %d)" % (path, len(vals), need))
    p = 0
    frho = []
    for i in range(nel):
        frho.append(vals[p:p + nrho[i]].copy())
        p += nrho[i]
    rhor = [[None] * nel for _ in range(nel)]
    for i in range(nel):
        for j in range(nel):
            rhor[i][j] = vals[p:p + nr[i, j]].copy()
            p += int(nr[i, j])
    z2r = [[None] * nel for _ in range(nel)]
    for i in range(nel):
        for j in range(i + 1):
            z2r[i][j] = vals[p:p + nr[i, j]].copy()
            p += int(nr[i, j])
    return dict(nelements=nel, nnonangular=nna, nangular=nang, names=names, nrho=nrho, drho=drho, mass=mass, nr=nr,
                dr=dr, cut=cut, frho=frho, rhor=rhor, z2r=z2r)


AEAM_INIT_KEYS = ("nelements", "nnonangular", "nrho", "drho", "nr", "dr", "cut", "frho", "rhor", "z2r")


def aeam_init_tables(path: str | None = None):
    """exactly the keys Context.aeam_init reads"""
    t = read_aeam(path)
    return {k: t[k] for k in AEAM_INIT_KEYS}
