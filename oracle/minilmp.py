"""ctypes driver of the mini LAMMPS engine (oracle/engine -> libminilmp.so): the HOST APPLICATION stand-in that loads a
pair-style plugin -- the reference's (oracle/_ref), the port, or the B200 one -- and runs LAMMPS input commands.

Test infrastructure: used by tests/, __graft_entry__.smoke() and bench.py's CPU legs; bench.py's drop-in leg uses it as
the host application around the B200 plugin (the pair style on that path is the product's, nothing of the oracle's pair
code is loaded there).  Never imported by the package."""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_int, c_longlong, c_void_p

import numpy as np

ORACLE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(ORACLE)
MINILMP_SO = os.path.join(ORACLE, "libminilmp.so")
REF_REBOMOS_SO = os.path.join(ORACLE, "_ref", "rebomosplugin.so")
REF_AEAM_SO = os.path.join(ORACLE, "_ref", "aeamplugin.so")
PORT_SO = os.path.join(ORACLE, "portplugin.so")
B200_REBOMOS_SO = os.path.join(REPO, "lammps_plugins_b200", "rebomosplugin.so")
B200_AEAM_SO = os.path.join(REPO, "lammps_plugins_b200", "aeamplugin.so")


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE, "all"], check=True)


def reference_plugin(style, allow_port=False):
    """oracle/_ref/<style>plugin.so = the reference sources compiled verbatim.  It is built where /root/reference exists
    and travels with the snapshot; when it is missing the caller gets an error, or -- only if it says so -- the port."""
    so = REF_REBOMOS_SO if style == "rebomos" else REF_AEAM_SO
    if os.path.exists(so):
        return so
    if allow_port and os.path.exists(PORT_SO):
        return PORT_SO
    raise FileNotFoundError("%s is missing: the verbatim build of the reference plugin (make -C oracle, needs "
                            "/root/reference) must be present; a parity check must not degrade to the port" % so)


# ----------------------------------------------------------------------------- minilmp
_ml = None


def _minilmp():
    global _ml
    if _ml is None:
        if not os.path.exists(MINILMP_SO):
            build_oracle()
        L = ctypes.CDLL(MINILMP_SO, mode=ctypes.RTLD_GLOBAL)
        L.minilmp_open.restype = c_void_p
        L.minilmp_open.argtypes = [c_int, c_int, c_int]
        L.minilmp_close.argtypes = [c_void_p]
        L.minilmp_last_error.restype = c_char_p
        L.minilmp_last_error.argtypes = [c_void_p]
        L.minilmp_command.argtypes = [c_void_p, c_char_p]
        L.minilmp_file.argtypes = [c_void_p, c_char_p]
        L.minilmp_setup.argtypes = [c_void_p, c_int, c_int]
        L.minilmp_compute.argtypes = [c_void_p, c_int, c_int, c_int]
        L.minilmp_forward_comm.argtypes = [c_void_p]
        L.minilmp_set_flags.argtypes = [c_void_p, c_int, c_int]
        L.minilmp_get_int.restype = c_longlong
        L.minilmp_get_int.argtypes = [c_void_p, c_int, c_char_p]
        L.minilmp_get_double.restype = c_double
        L.minilmp_get_double.argtypes = [c_void_p, c_int, c_char_p]
        L.minilmp_get_ptr.restype = c_void_p
        L.minilmp_get_ptr.argtypes = [c_void_p, c_int, c_char_p]
        L.minilmp_neigh_total.restype = c_longlong
        L.minilmp_neigh_total.argtypes = [c_void_p, c_int]
        L.minilmp_neigh_csr.argtypes = [c_void_p, c_int, POINTER(c_longlong), POINTER(c_int)]
        L.minilmp_swap_info.argtypes = [c_void_p, c_int, c_int, POINTER(c_int)]
        L.minilmp_swap_sendlist.argtypes = [c_void_p, c_int, c_int, POINTER(c_int)]
        L.minilmp_thermo_count.argtypes = [c_void_p]
        L.minilmp_thermo_row.argtypes = [c_void_p, c_int, POINTER(c_double)]
        L.minilmp_thermo_clear.argtypes = [c_void_p]
        L.minilmp_nprocs.argtypes = [c_void_p]
        _ml = L
    return _ml


class LammpsError(RuntimeError):
    pass


class MiniLmp:
    """A minilmp instance (px*py*pz thread ranks).  Mirrors the spirit of LAMMPS' python module:
    ``lmp.command("pair_style rebomos")``."""

    def __init__(self, grid=(1, 1, 1)):
        self.L = _minilmp()
        self.h = self.L.minilmp_open(*grid)
        self.nprocs = grid[0] * grid[1] * grid[2]

    def close(self):
        if self.h:
            self.L.minilmp_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise LammpsError(self.L.minilmp_last_error(self.h).decode())

    def command(self, line):
        self._chk(self.L.minilmp_command(self.h, line.encode()))

    def commands(self, lines):
        for ln in lines:
            self.command(ln)

    def file(self, path):
        self._chk(self.L.minilmp_file(self.h, path.encode()))

    def setup(self, eflag=1, vflag=2):
        self._chk(self.L.minilmp_setup(self.h, eflag, vflag))

    def compute(self, eflag=1, vflag=2, reverse=False):
        self._chk(self.L.minilmp_compute(self.h, eflag, vflag, 1 if reverse else 0))

    def forward_comm(self):
        self._chk(self.L.minilmp_forward_comm(self.h))

    def get_int(self, name, rank=0):
        v = self.L.minilmp_get_int(self.h, rank, name.encode())
        if v == -999999:
            raise KeyError(name)
        return int(v)

    def get_double(self, name, rank=0):
        return float(self.L.minilmp_get_double(self.h, rank, name.encode()))

    def _arr(self, name, rank, n, dtype, cols=None):
        p = self.L.minilmp_get_ptr(self.h, rank, name.encode())
        if not p or n == 0:
            return np.zeros((0, cols) if cols else (0,), dtype=dtype)
        ct = c_double if dtype == np.float64 else c_int
        total = n * (cols or 1)
        a = np.ctypeslib.as_array(ctypes.cast(p, POINTER(ct)), shape=(total,))
        return a.reshape(n, cols) if cols else a

    def nall(self, rank=0):
        return self.get_int("nlocal", rank) + self.get_int("nghost", rank)

    def x(self, rank=0, n=None):
        return self._arr("x", rank, self.nall(rank) if n is None else n, np.float64, 3)

    def v(self, rank=0):
        return self._arr("v", rank, self.get_int("nlocal", rank), np.float64, 3)

    def f(self, rank=0, n=None):
        return self._arr("f", rank, self.nall(rank) if n is None else n, np.float64, 3)

    def type(self, rank=0):
        return self._arr("type", rank, self.nall(rank), np.int32)

    def tag(self, rank=0):
        return self._arr("tag", rank, self.nall(rank), np.int32)

    def mass(self):
        nt = self.get_int("ntypes")
        return self._arr("mass", 0, nt + 1, np.float64).copy()

    def neigh_csr(self, rank=0):
        nrows = self.get_int("inum", rank) + self.get_int("gnum", rank)
        tot = int(self.L.minilmp_neigh_total(self.h, rank))
        off = np.zeros(nrows + 1, dtype=np.int64)
        val = np.zeros(max(tot, 1), dtype=np.int32)
        self.L.minilmp_neigh_csr(self.h, rank, off.ctypes.data_as(POINTER(c_longlong)),
                                 val.ctypes.data_as(POINTER(c_int)))
        return off, val[:tot]

    def swaps(self, rank=0):
        """Halo plan of one rank: list of dicts (sendnum, recvnum, firstrecv, sendproc, recvproc, pbc_flag, pbc, sendlist)."""
        out = []
        for s in range(self.get_int("nswap", rank)):
            info = (c_int * 12)()
            self.L.minilmp_swap_info(self.h, rank, s, info)
            sl = np.zeros(max(info[0], 1), dtype=np.int32)
            self.L.minilmp_swap_sendlist(self.h, rank, s, sl.ctypes.data_as(POINTER(c_int)))
            out.append(dict(sendnum=info[0], recvnum=info[1], firstrecv=info[2], sendproc=info[3],
                            recvproc=info[4], pbc_flag=info[5], pbc=list(info[6:12]), sendlist=sl[: info[0]]))
        return out

    def thermo(self):
        rows = []
        buf = (c_double * 13)()
        for i in range(self.L.minilmp_thermo_count(self.h)):
            self.L.minilmp_thermo_row(self.h, i, buf)
            rows.append(dict(step=int(buf[0]), temp=buf[1], press=buf[2], pe=buf[3], ke=buf[4], etotal=buf[5],
                             vol=buf[6], virial=np.array(buf[7:13])))
        return rows

    def thermo_clear(self):
        self.L.minilmp_thermo_clear(self.h)

    def cutneighsq(self, ghost=False):
        nt = self.get_int("ntypes")
        return self._arr("cutneighghostsq" if ghost else "cutneighsq", 0, (nt + 1) * (nt + 1), np.float64).copy()

    def b200_box(self, rank=0):
        """b200md_box of this rank's sub-domain (lamda bounds if triclinic), as Neighbor/Comm see it."""
        from lammps_plugins_b200 import make_box
        d = self.box()
        b = make_box(d["boxlo"], d["boxhi"], d["xy"], d["xz"], d["yz"], triclinic=d["triclinic"])
        for k in range(3):
            b.sublo[k] = self.get_double("sublo%d" % k, rank)
            b.subhi[k] = self.get_double("subhi%d" % k, rank)
            b.cutghost[k] = self.get_double("cutghost%d" % k, rank)
        b.cutneighmax = self.get_double("cutneighmax", rank)
        return b

    def units(self):
        return {k: self.get_double(k) for k in ("boltz", "mvv2e", "ftm2v", "nktv2p")}

    def box(self):
        d = {k: self.get_double(k) for k in ("xy", "xz", "yz")}
        d["boxlo"] = [self.get_double("boxlo%d" % k) for k in range(3)]
        d["boxhi"] = [self.get_double("boxhi%d" % k) for k in range(3)]
        d["triclinic"] = self.get_int("triclinic")
        return d


