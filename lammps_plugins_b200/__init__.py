"""lammps_plugins_b200 -- B200-native (sm_100a) force path for the LAMMPS pair styles
``rebomos`` and ``aeam``.

This package is a thin ctypes binding of the C ABI declared in ``include/b200md.h``
(implemented by ``libb200md.so``: hand-written CUDA, built in-tree by
``__graft_entry__.build()`` or ``make -C lammps_plugins_b200``).  There is no
Python or CPU implementation behind it: importing works anywhere, but every
compute call needs the CUDA library and a B200; a missing library raises
``B200MDError`` instead of falling back.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_longlong, c_void_p

import numpy as np

__all__ = ["B200MDError", "lib", "Context", "RebomosParams", "PKG_DIR", "REPO_DIR", "library_path"]

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)

ENERGY_GLOBAL = 1
VIRIAL_PAIR = 1
VIRIAL_FDOTR = 2


class B200MDError(RuntimeError):
    pass


def library_path() -> str:
    return os.path.join(PKG_DIR, "libb200md.so")


class RebomosParams(ctypes.Structure):
    """Mirror of ``b200md_rebomos_params`` (include/b200md.h)."""

    _fields_ = [
        ("rcmin", c_double * 4), ("rcmax", c_double * 4),
        ("Q", c_double * 4), ("alpha", c_double * 4), ("A", c_double * 4),
        ("BIJc", c_double * 4), ("Beta", c_double * 4),
        ("b", (c_double * 2) * 7), ("bg", (c_double * 2) * 7), ("a", (c_double * 2) * 4),
        ("rcLJmin", c_double * 4), ("rcLJmax", c_double * 4),
        ("epsilon", c_double * 4), ("sigma", c_double * 4),
    ]


class AeamTables(ctypes.Structure):
    """Mirror of ``b200md_aeam_tables``."""

    _fields_ = [
        ("nelements", c_int), ("nnonangular", c_int),
        ("nrho", POINTER(c_int)), ("drho", POINTER(c_double)),
        ("nr", POINTER(c_int)), ("dr", POINTER(c_double)), ("cut", POINTER(c_double)),
        ("frho", POINTER(POINTER(c_double))), ("rhor", POINTER(POINTER(c_double))),
        ("z2r", POINTER(POINTER(c_double))),
    ]


class Box(ctypes.Structure):
    """Mirror of ``b200md_box``."""

    _fields_ = [
        ("triclinic", c_int),
        ("boxlo", c_double * 3), ("boxhi", c_double * 3),
        ("xy", c_double), ("xz", c_double), ("yz", c_double),
        ("sublo", c_double * 3), ("subhi", c_double * 3),
        ("cutghost", c_double * 3), ("cutneighmax", c_double),
    ]


class SystemDesc(ctypes.Structure):
    """Mirror of ``b200md_system_desc``."""

    _fields_ = [
        ("style", c_int), ("ntypes", c_int), ("mass", POINTER(c_double)), ("box", Box),
        ("procgrid", c_int * 3), ("rank", c_int), ("skin", c_double), ("dt", c_double),
        ("ftm2v", c_double), ("mvv2e", c_double), ("boltz", c_double), ("nktv2p", c_double),
        ("sort_every", c_int),
    ]


_lib = None

# every symbol include/b200md.h declares: (name, restype, argtypes)
_PD = POINTER(c_double)
_PI = POINTER(c_int)
_PL = POINTER(c_int64)
SYMBOLS = [
    ("b200md_create", c_int, [c_int, POINTER(c_void_p)]),
    ("b200md_destroy", None, [c_void_p]),
    ("b200md_last_error", c_char_p, [c_void_p]),
    ("b200md_version", c_int, []),
    ("b200md_device_count", c_int, []),
    ("b200md_rebomos_init", c_int, [c_void_p, POINTER(RebomosParams), c_int, _PI]),
    ("b200md_aeam_init", c_int, [c_void_p, POINTER(AeamTables)]),
    ("b200md_aeam_get_spline", c_int, [c_void_p, c_int, c_int, _PD, c_int]),
    ("b200md_set_neighbor_list", c_int, [c_void_p, c_int, c_int, _PI, POINTER(_PI), c_double]),
    ("b200md_set_neighbor_list_ilist", c_int, [c_void_p, c_int, c_int, _PI, _PI, POINTER(_PI), c_double]),
    ("b200md_set_neighbor_csr", c_int, [c_void_p, c_int, c_int, _PL, _PI, c_double]),
    ("b200md_neigh_build", c_int, [c_void_p, POINTER(Box), c_int, _PD, _PD, c_int, c_int, _PD, _PI, c_int, c_double]),
    ("b200md_neigh_size", c_int, [c_void_p, _PI, _PL]),
    ("b200md_neigh_download", c_int, [c_void_p, _PI, _PL, _PI]),
    ("b200md_rebomos_compute", c_int, [c_void_p, c_int, c_int, _PD, _PI, _PI, c_int, c_int, _PD, _PD, _PD]),
    ("b200md_rebomos_compute_peratom", c_int, [c_void_p, c_int, c_int, _PD, _PI, _PI, c_int, c_int, _PD, _PD, _PD, _PD, _PD]),
    ("b200md_rebomos_neigh", c_int, [c_void_p, c_int, c_int, _PD, _PI, c_int, _PI, _PI, _PD, _PD]),
    ("b200md_aeam_compute", c_int, [c_void_p, c_int, c_int, _PD, _PI, _PI, c_int, c_int, _PD, _PD, _PD]),
    ("b200md_aeam_density_phase", c_int, [c_void_p, c_int, c_int, _PD, _PI, _PD, _PD]),
    ("b200md_aeam_force_phase", c_int, [c_void_p, _PD, _PD, c_int, c_int, _PD, _PD, _PD]),
    ("b200md_aeam_compute_peratom", c_int, [c_void_p, c_int, c_int, _PD, _PI, _PI, c_int, c_int, _PD, _PD, _PD, _PD, _PD]),
    ("b200md_aeam_force_phase_peratom", c_int, [c_void_p, _PD, _PD, c_int, c_int, _PD, _PD, _PD, _PD, _PD]),
    ("b200md_aeam_get_rho_fp", c_int, [c_void_p, c_int, _PD, _PD]),
    ("b200md_set_option", c_int, [c_void_p, c_char_p, c_longlong]),
    ("b200md_get_counter", c_longlong, [c_void_p, c_char_p]),
    ("b200md_last_kernel_ms", c_double, [c_void_p, c_char_p]),
    ("b200md_kernel_stats", c_int, [c_void_p, c_int, c_char_p, c_int, _PD, POINTER(c_longlong)]),
    ("b200md_kernel_stats_reset", c_int, [c_void_p]),
    ("b200md_stream", c_void_p, [c_void_p]),
    ("b200md_event_record", c_int, [c_void_p, c_int]),
    ("b200md_event_elapsed_ms", c_double, [c_void_p, c_int, c_int]),
    ("b200md_measure_peaks", c_int, [c_void_p, _PD, _PD]),
    ("b200md_copy_probe", c_int, [c_void_p, c_void_p, ctypes.c_size_t, c_int, _PD]),
    ("b200md_host_alloc", c_void_p, [ctypes.c_size_t]),
    ("b200md_host_free", None, [c_void_p]),
    ("b200md_host_register", c_int, [c_void_p, ctypes.c_size_t]),
    ("b200md_host_unregister", c_int, [c_void_p]),
    ("b200md_system_create", c_int, [c_void_p, POINTER(SystemDesc), c_int, _PD, _PD, _PI, _PI]),
    ("b200md_nccl_unique_id", c_int, [c_void_p]),
    ("b200md_system_comm_init", c_int, [c_void_p, c_void_p, c_int, c_int]),
    ("b200md_local_group_create", c_int, [c_int]),
    ("b200md_system_comm_init_local", c_int, [c_void_p, c_int, c_int, c_int]),
    ("b200md_exchange_plan", c_int, [c_int, _PI, c_int, _PI, _PI, _PI, _PI]),
    ("b200md_system_run", c_int, [c_void_p, c_int, c_int]),
    ("b200md_system_set_nvt", c_int, [c_void_p, c_double, c_double, c_double]),
    ("b200md_system_nh_energy", c_double, [c_void_p]),
    ("b200md_system_thermo", c_int, [c_void_p, _PD]),
    ("b200md_system_thermo_count", c_int, [c_void_p]),
    ("b200md_system_thermo_row", c_int, [c_void_p, c_int, _PD]),
    ("b200md_system_sizes", c_int, [c_void_p, POINTER(c_longlong)]),
    ("b200md_system_download", c_int, [c_void_p, _PD, _PD, _PD, _PI, _PI]),
]


def lib():
    """Load ``libb200md.so`` (once).  Raises B200MDError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise B200MDError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    try:
        L = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    except OSError as e:
        raise B200MDError(f"cannot load {path}: {e}") from e
    for name, res, args in SYMBOLS:
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(_PD)


def _ip(a):
    return a.ctypes.data_as(_PI)


class Context:
    """One CUDA context/stream of the force library (``b200md_ctx``)."""

    def __init__(self, device: int = 0):
        self.L = lib()
        h = c_void_p()
        rc = self.L.b200md_create(device, ctypes.byref(h))
        if rc != 0:
            raise B200MDError(self.L.b200md_last_error(None).decode())
        self.h = h
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.b200md_destroy(self.h)
            self.h = None
            for p in self._keep:
                self.L.b200md_host_free(p)
            self._keep = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def copy_probe(self, host_array, to_device):
        """ms of one plain H2D (to_device) / D2H copy of a host array"""
        ms = c_double()
        self._check(self.L.b200md_copy_probe(self.h, host_array.ctypes.data_as(c_void_p), host_array.nbytes,
                                             1 if to_device else 0, ctypes.byref(ms)))
        return ms.value

    def last_error(self):
        return self.L.b200md_last_error(self.h).decode()

    def _check(self, rc):
        if rc != 0:
            raise B200MDError(f"[{rc}] " + self.L.b200md_last_error(self.h).decode())

    # -- options / counters
    def set_option(self, name, value):
        self._check(self.L.b200md_set_option(self.h, name.encode(), int(value)))

    def counter(self, name):
        return int(self.L.b200md_get_counter(self.h, name.encode()))

    def kernel_ms(self, name):
        return float(self.L.b200md_last_kernel_ms(self.h, name.encode()))

    def kernel_stats(self, reset=False):
        """{kernel name: (total_ms, launches)} accumulated while option sync_timing is on"""
        out = {}
        i = 0
        name = ctypes.create_string_buffer(64)
        tot = c_double()
        cnt = c_longlong()
        while self.L.b200md_kernel_stats(self.h, i, name, 64, ctypes.byref(tot), ctypes.byref(cnt)) == 0:
            out[name.value.decode()] = (tot.value, cnt.value)
            i += 1
        if reset:
            self.L.b200md_kernel_stats_reset(self.h)
        return out

    def event_record(self, slot):
        self._check(self.L.b200md_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b):
        return float(self.L.b200md_event_elapsed_ms(self.h, a, b))

    def measure_peaks(self):
        """(FP64 TFLOP/s of a DFMA-saturating kernel, GB/s of a 2 GiB device copy) on this device"""
        tf, gb = c_double(), c_double()
        self._check(self.L.b200md_measure_peaks(self.h, ctypes.byref(tf), ctypes.byref(gb)))
        return tf.value, gb.value

    def pinned_array(self, shape, dtype=np.float64):
        """numpy array over page-locked host memory (freed with the context)"""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self.L.b200md_host_alloc(n)
        if not p:
            raise B200MDError("cudaMallocHost failed")
        self._keep.append(p)
        buf = (ctypes.c_char * n).from_address(p)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    # -- potentials
    def rebomos_init(self, params: RebomosParams, type_map):
        m = np.ascontiguousarray(np.concatenate([[-1], np.asarray(type_map, dtype=np.int32)]), dtype=np.int32)
        self._check(self.L.b200md_rebomos_init(self.h, ctypes.byref(params), len(type_map), _ip(m)))

    def aeam_init(self, tab: dict):
        """tab: nelements, nnonangular, nrho[], drho[], nr[][], dr[][], cut[][], frho[i], rhor[i][j], z2r[i][j]."""
        nel = int(tab["nelements"])
        nrho = np.ascontiguousarray(tab["nrho"], dtype=np.int32)
        drho = np.ascontiguousarray(tab["drho"], dtype=np.float64)
        nr = np.ascontiguousarray(tab["nr"], dtype=np.int32).reshape(-1)
        dr = np.ascontiguousarray(tab["dr"], dtype=np.float64).reshape(-1)
        cut = np.ascontiguousarray(tab["cut"], dtype=np.float64).reshape(-1)
        frho = [np.ascontiguousarray(tab["frho"][i], dtype=np.float64) for i in range(nel)]
        rhor = [np.ascontiguousarray(tab["rhor"][i][j], dtype=np.float64) for i in range(nel) for j in range(nel)]
        z2r = [np.ascontiguousarray(tab["z2r"][i][j] if j <= i else np.zeros(1), dtype=np.float64)
               for i in range(nel) for j in range(nel)]
        PP = POINTER(c_double)
        fr = (PP * nel)(*[_dp(a) for a in frho])
        rh = (PP * (nel * nel))(*[_dp(a) for a in rhor])
        zz = (PP * (nel * nel))(*[_dp(a) for a in z2r])
        t = AeamTables(nel, int(tab["nnonangular"]), _ip(nrho), _dp(drho), _ip(nr), _dp(dr), _dp(cut), fr, rh, zz)
        keep = (nrho, drho, nr, dr, cut, frho, rhor, z2r, fr, rh, zz)
        self._check(self.L.b200md_aeam_init(self.h, ctypes.byref(t)))
        del keep

    def aeam_get_spline(self, kind, index, nrows):
        out = np.zeros((nrows + 1, 7))
        self._check(self.L.b200md_aeam_get_spline(self.h, kind, index, _dp(out), nrows))
        return out

    # -- neighbor lists
    def set_neighbor_csr(self, inum, gnum, offsets, values, skin):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        values = np.ascontiguousarray(values, dtype=np.int32)
        self._check(self.L.b200md_set_neighbor_csr(self.h, inum, gnum, offsets.ctypes.data_as(_PL), _ip(values),
                                                   float(skin)))

    def neigh_build(self, box: Box, ntypes, cutneighsq, cutneighghostsq, nlocal, nghost, x, type_, ghost_rows, skin):
        cs = np.ascontiguousarray(cutneighsq, dtype=np.float64)
        cg = np.ascontiguousarray(cutneighghostsq, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        type_ = np.ascontiguousarray(type_, dtype=np.int32)
        self._check(self.L.b200md_neigh_build(self.h, ctypes.byref(box), ntypes, _dp(cs), _dp(cg), nlocal, nghost,
                                              _dp(x), _ip(type_), int(ghost_rows), float(skin)))

    def neigh_download(self):
        nrows = c_int()
        nent = c_int64()
        self._check(self.L.b200md_neigh_size(self.h, ctypes.byref(nrows), ctypes.byref(nent)))
        num = np.zeros(nrows.value, dtype=np.int32)
        off = np.zeros(nrows.value + 1, dtype=np.int64)
        val = np.zeros(max(nent.value, 1), dtype=np.int32)
        self._check(self.L.b200md_neigh_download(self.h, _ip(num), off.ctypes.data_as(_PL), _ip(val)))
        return num, off, val[: nent.value]

    # -- compute
    def rebomos_compute(self, nlocal, nghost, x, type_, tag, eflag=1, vflag=2, f=None):
        nall = nlocal + nghost
        x = np.ascontiguousarray(x, dtype=np.float64)
        type_ = np.ascontiguousarray(type_, dtype=np.int32)
        tag = np.ascontiguousarray(tag, dtype=np.int32)
        if f is None:
            f = np.zeros((nall, 3))
        eng = c_double()
        vir = np.zeros(6)
        self._check(self.L.b200md_rebomos_compute(self.h, nlocal, nghost, _dp(x), _ip(type_), _ip(tag), eflag, vflag,
                                                  _dp(f), ctypes.byref(eng), _dp(vir)))
        return f, eng.value, vir

    def rebomos_compute_peratom(self, nlocal, nghost, x, type_, tag, eflag=3, vflag=6):
        """returns f, eng, virial, eatom[nall], vatom[nall, 6]"""
        nall = nlocal + nghost
        x = np.ascontiguousarray(x, dtype=np.float64)
        type_ = np.ascontiguousarray(type_, dtype=np.int32)
        tag = np.ascontiguousarray(tag, dtype=np.int32)
        f, ea, va = np.zeros((nall, 3)), np.zeros(nall), np.zeros((nall, 6))
        eng = c_double()
        vir = np.zeros(6)
        self._check(self.L.b200md_rebomos_compute_peratom(self.h, nlocal, nghost, _dp(x), _ip(type_), _ip(tag), eflag,
                                                          vflag, _dp(f), ctypes.byref(eng), _dp(vir), _dp(ea), _dp(va)))
        return f, eng.value, vir, ea, va

    def rebomos_neigh(self, nlocal, nghost, x, type_, stride=32):
        nall = nlocal + nghost
        x = np.ascontiguousarray(x, dtype=np.float64)
        type_ = np.ascontiguousarray(type_, dtype=np.int32)
        num = np.zeros(nall, dtype=np.int32)
        rows = np.full((nall, stride), -1, dtype=np.int32)
        nM = np.zeros(nall)
        nS = np.zeros(nall)
        self._check(self.L.b200md_rebomos_neigh(self.h, nlocal, nghost, _dp(x), _ip(type_), stride, _ip(num),
                                                _ip(rows), _dp(nM), _dp(nS)))
        return num, rows, nM, nS

    def aeam_compute(self, nlocal, nghost, x, type_, tag, eflag=1, vflag=2, f=None):
        nall = nlocal + nghost
        x = np.ascontiguousarray(x, dtype=np.float64)
        type_ = np.ascontiguousarray(type_, dtype=np.int32)
        tag = np.ascontiguousarray(tag, dtype=np.int32)
        if f is None:
            f = np.zeros((nall, 3))
        eng = c_double()
        vir = np.zeros(6)
        self._check(self.L.b200md_aeam_compute(self.h, nlocal, nghost, _dp(x), _ip(type_), _ip(tag), eflag, vflag,
                                               _dp(f), ctypes.byref(eng), _dp(vir)))
        return f, eng.value, vir

    def aeam_compute_peratom(self, nlocal, nghost, x, type_, tag, eflag=3, vflag=6):
        """returns f, eng, virial, eatom[nall], vatom[nall, 6]"""
        nall = nlocal + nghost
        x = np.ascontiguousarray(x, dtype=np.float64)
        type_ = np.ascontiguousarray(type_, dtype=np.int32)
        tag = np.ascontiguousarray(tag, dtype=np.int32)
        f, ea, va = np.zeros((nall, 3)), np.zeros(nall), np.zeros((nall, 6))
        eng = c_double()
        vir = np.zeros(6)
        self._check(self.L.b200md_aeam_compute_peratom(self.h, nlocal, nghost, _dp(x), _ip(type_), _ip(tag), eflag, vflag,
                                                       _dp(f), ctypes.byref(eng), _dp(vir), _dp(ea), _dp(va)))
        return f, eng.value, vir, ea, va

    def aeam_density_phase(self, nlocal, nghost, x, type_, rho=None, fp=None):
        """rho, fp: optional caller-owned output arrays of nlocal + nghost doubles (e.g. pinned_array), as a host
        application hands its own atom->rho / atom->fp"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        type_ = np.ascontiguousarray(type_, dtype=np.int32)
        if rho is None:
            rho = np.zeros(nlocal + nghost)
        if fp is None:
            fp = np.zeros(nlocal + nghost)
        self._check(self.L.b200md_aeam_density_phase(self.h, nlocal, nghost, _dp(x), _ip(type_), _dp(rho), _dp(fp)))
        return rho, fp

    def aeam_force_phase(self, rho_all, fp_all, eflag=1, vflag=2, f=None):
        rho_all = np.ascontiguousarray(rho_all, dtype=np.float64)
        fp_all = np.ascontiguousarray(fp_all, dtype=np.float64)
        if f is None:
            f = np.zeros((len(rho_all), 3))
        eng = c_double()
        vir = np.zeros(6)
        self._check(self.L.b200md_aeam_force_phase(self.h, _dp(rho_all), _dp(fp_all), eflag, vflag, _dp(f),
                                                   ctypes.byref(eng), _dp(vir)))
        return f, eng.value, vir

    def aeam_rho_fp(self, nlocal):
        rho = np.zeros(nlocal)
        fp = np.zeros(nlocal)
        self._check(self.L.b200md_aeam_get_rho_fp(self.h, nlocal, _dp(rho), _dp(fp)))
        return rho, fp

    # -- GPU-resident system
    def system_create(self, style, ntypes, mass, box: Box, x, v, type_, tag, skin, dt, units, procgrid=(1, 1, 1),
                      rank=0, sort_every=1000):
        """style: 'rebomos' | 'aeam'; mass: 1-based list (index 0 unused); units: dict ftm2v, mvv2e, boltz, nktv2p"""
        mass = np.ascontiguousarray(mass, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        type_ = np.ascontiguousarray(type_, dtype=np.int32)
        tag = np.ascontiguousarray(tag, dtype=np.int32)
        d = SystemDesc()
        d.style = 0 if style == "rebomos" else 1
        d.ntypes = ntypes
        d.mass = _dp(mass)
        d.box = box
        d.procgrid[0], d.procgrid[1], d.procgrid[2] = procgrid
        d.rank = rank
        d.skin = skin
        d.dt = dt
        d.ftm2v, d.mvv2e, d.boltz, d.nktv2p = units["ftm2v"], units["mvv2e"], units["boltz"], units["nktv2p"]
        d.sort_every = sort_every
        vp = None
        if v is not None:
            v = np.ascontiguousarray(v, dtype=np.float64)
            vp = _dp(v)
        self._check(self.L.b200md_system_create(self.h, ctypes.byref(d), len(type_), _dp(x), vp, _ip(type_), _ip(tag)))

    def system_run(self, nsteps, thermo_every=0):
        self._check(self.L.b200md_system_run(self.h, int(nsteps), int(thermo_every)))

    def system_set_nvt(self, t_start, t_stop, t_period):
        """fix nvt temp t_start t_stop t_period for the resident loop (t_period <= 0: back to NVE)"""
        self._check(self.L.b200md_system_set_nvt(self.h, float(t_start), float(t_stop), float(t_period)))

    def system_nh_energy(self):
        return float(self.L.b200md_system_nh_energy(self.h))

    def system_thermo_rows(self):
        rows = []
        buf = np.zeros(12)
        for i in range(self.L.b200md_system_thermo_count(self.h)):
            self._check(self.L.b200md_system_thermo_row(self.h, i, _dp(buf)))
            rows.append(dict(step=int(buf[0]), temp=buf[1], press=buf[2], pe=buf[3], ke=buf[4], vol=buf[5],
                             virial=buf[6:12].copy()))
        return rows

    def system_sizes(self):
        out = (c_longlong * 8)()
        self._check(self.L.b200md_system_sizes(self.h, out))
        return dict(nlocal=out[0], nghost=out[1], nbuild=out[2], ndanger=out[3], nmigrated=out[4], natoms=out[5],
                    ninner=out[6], noverlap=out[7])

    def comm_init_nccl(self, id128: bytes, nranks, rank):
        self._check(self.L.b200md_system_comm_init(self.h, id128, nranks, rank))

    def comm_init_local(self, group, nranks, rank):
        self._check(self.L.b200md_system_comm_init_local(self.h, group, nranks, rank))

    def system_download(self):
        sz = self.system_sizes()
        nall = sz["nlocal"] + sz["nghost"]
        x = np.zeros((nall, 3))
        v = np.zeros((sz["nlocal"], 3))
        f = np.zeros((nall, 3))
        type_ = np.zeros(nall, dtype=np.int32)
        tag = np.zeros(nall, dtype=np.int32)
        self._check(self.L.b200md_system_download(self.h, _dp(x), _dp(v), _dp(f), _ip(type_), _ip(tag)))
        return dict(x=x, v=v, f=f, type=type_, tag=tag, **sz)


METAL_UNITS = dict(boltz=8.617343e-5, mvv2e=1.0364269e-4, ftm2v=1.0 / 1.0364269e-4, nktv2p=1.6021765e6)


def make_box(boxlo, boxhi, xy=0.0, xz=0.0, yz=0.0, triclinic=None):
    b = Box()
    b.triclinic = int(triclinic if triclinic is not None else (xy != 0.0 or xz != 0.0 or yz != 0.0))
    for d in range(3):
        b.boxlo[d] = boxlo[d]
        b.boxhi[d] = boxhi[d]
    b.xy, b.xz, b.yz = xy, xz, yz
    return b
