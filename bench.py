#!/usr/bin/env python
"""Benchmark driver: atom-timesteps/s of the GPU-resident REBOMoS and AEAM force paths.

  python bench.py --gpus N --steps K --warmup W            # B200 arm (one JSON line on rank 0)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference pair style compiled
                                                           # verbatim (oracle/_ref) on the host cores

A "step" is one NVE timestep of the whole job: integrate, ghost exchange, (re)neighboring when an atom moved more
than skin/2, force computation, reverse exchange.  The top-level record is BASELINE.json configs[2] / [4]: the shipped
MoS2 cell replicated 14x13x19 = 995 904 atoms per GPU (weak scaling), 300 K, dt 1 fs.  `workloads.aeam` is configs[3]
(fcc Al + 0.75 % Si, 80^3 cells = 2 048 000 atoms per GPU, 863 K) with the same keys; at N > 1 `workloads.aeam_strong`
is the same FIXED 2 048 000-atom system decomposed over the N GPUs (strong scaling, configs[3] as written).
`--workload rebomos|aeam` measures one of them alone.

Keys beyond the base contract: roofline (dominant kernel: its OWN algorithmic bytes and flops, live CUDA-event time;
per_kernel table; whole_step = the headline fractions), cpu_baseline, e2e (plugin-mode C-ABI call with pinned HOST
buffers and moving atoms, H2D/D2H inside the timed region), e2e_plugin (N = 1: the same through `plugin load` ->
PairREBOMoS::compute inside a host application), clocks, gpu_launches, kernels_ms_per_step, neighbor (rebuilds in the
timed region, a forced master rebuild timed on its own, a long run with its natural rebuilds).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

# Reference formulation, SURVEY.md 8(d): one pass over the full neighbor row + x, f, type, tag; flops as the
# reference writes them.  Used for `whole_step` only (the headline fractions); the per-kernel entries use the rows
# and arithmetic of the kernels as built.
ALGO = {
    "rebomos": dict(bytes=2040.0, flops=19.0e3),
    "aeam": dict(bytes=756.0, flops=6.9e3),
}
FP64_NOMINAL_TFLOPS = 37.0
LONG_RUN_STEPS = {"rebomos": 1000, "aeam": 300}
# kernel templates, each with the launches that belong to it (one per center element; "_ev" = the energy/virial
# instances that run on thermo steps only -- one step per system_run call -- and are averaged in)
GROUPS = {"rebomos": {"lj": ["lj_mo", "lj_s", "lj_mo_ev", "lj_s_ev"],
                      "rebo_center": ["rebo_center_mo", "rebo_center_s", "rebo_center_mo_ev", "rebo_center_s_ev",
                                      "rebo_center_overflow", "rebo_gather"]},
          "aeam": {"aeam_force": ["aeam_force", "aeam_force_ev"], "aeam_density": ["aeam_density"],
                   "aeam_angular": ["aeam_force_ang", "aeam_force_ang_ev", "aeam_density_ang"]}}


def measured_peaks():
    p = os.path.join(HERE, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def procgrid_for(n):
    from lammps_plugins_b200.launch import procgrid_for as pf
    return pf(n)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons through NVML in-process (a sample costs microseconds and takes no driver-wide lock;
    spawning nvidia-smi during a timed region stalled the GPU on some boxes).  Sampled every 20 ms for the whole run;
    every timed region names its window, and a window shorter than the sampling period takes the samples around it."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.nv = self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv = None

    def sample(self):
        try:
            if self.nv is not None:
                nv, h = self.nv, self.h
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                smax = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.samples.append((time.time(), sm, smax, r))
            else:
                q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                self.samples.append((time.time(), float(f[0]), float(f[1]), int(f[2], 16)))
        except Exception:
            pass

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.02 if self.nv is not None else 1.0)

    def summary(self, t0, t1):
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        where = "inside the timed region"
        if not inside:      # a region shorter than the sampling period: the samples right before and after it
            before = [s for s in self.samples if s[0] < t0][-2:]
            after = [s for s in self.samples if s[0] > t1][:2]
            inside = before + after
            where = "right before and after the timed region (it is shorter than the 20 ms sampling period)"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock sample available"]}
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = sorted({name for s in inside for b, name in bits.items() if s[3] & b})
        return {"sm_mhz": float(np.median([s[1] for s in inside])), "sm_max_mhz": inside[0][2], "reasons": reasons,
                "samples": len(inside), "source": "nvml in-process, " + where}


# ----------------------------------------------------------------------------------------------- workloads
def make_workload(kind, rep, nranks, scaling="weak"):
    from lammps_plugins_b200 import workloads as W
    grid = procgrid_for(nranks)
    mult = grid if scaling == "weak" else (1, 1, 1)
    if kind == "rebomos":
        r = rep or (14, 13, 19)
        full = (r[0] * mult[0], r[1] * mult[1], r[2] * mult[2])
        w = W.mos2_bulk(*full)
        w["v"] = W.maxwell_velocities(w["type"], w["mass"], 300.0, 12345)
        w["name"] = "rebomos MoS2 bulk (in.rebomos-bulk cell) replicated %dx%dx%d" % full
        w["skin"], w["dt"] = 2.0, 0.001
    else:
        r = rep or (80, 80, 80)
        full = (r[0] * mult[0], r[1] * mult[1], r[2] * mult[2])
        w = W.fcc_alsi(full, 0.0075, 7683797)
        w["v"] = W.maxwell_velocities(w["type"], w["mass"], 863.0, 1082337)
        w["name"] = "aeam fcc Al-0.75%%Si a=4.045 %dx%dx%d cells" % full
        w["skin"], w["dt"] = 1.0, 0.001
    w["grid"] = grid
    w["cells"] = full
    return w


def init_potential(ctx, kind):
    """parameters through the package's own readers of the two file formats (lammps_plugins_b200/potentials.py)"""
    from lammps_plugins_b200 import potentials as P
    if kind == "rebomos":
        ctx.rebomos_init(P.read_rebomos(), [0, 1])
    else:
        ctx.aeam_init(P.aeam_init_tables())


def neighbor_cutoffs(kind, skin):
    from lammps_plugins_b200 import potentials as P
    from lammps_plugins_b200 import workloads as W
    if kind == "rebomos":
        return W.rebomos_neighbor_cutoffs(list(P.read_rebomos().rcmax), [0, 1], skin)
    return W.aeam_neighbor_cutoffs(P.read_aeam()["cut"], skin)


# ----------------------------------------------------------------------------------------------- per-kernel model
def kernel_model(kind, ctx, atoms_per_gpu):
    """Own algorithmic bytes per launch set of the kernels as built: the rows a kernel streams (entries x 4 B, counted on
    the device) plus what it reads and writes per atom.  Flops of the formulation come from the committed ncu capture
    (DFMA x 2 + DADD + DMUL per step, profiles/r02_kernel_model.json) -- they are a property of the SASS, not of the run."""
    n = atoms_per_gpu
    m = {}
    if kind == "rebomos":
        lj = ctx.counter("lj_entries_tight")
        if lj < 0:
            lj = ctx.counter("lj_entries")
        sh = ctx.counter("short_entries_tight")
        if sh < 0:
            sh = ctx.counter("short_entries_owned")
        m["lj"] = {"rows_entries": lj, "bytes": 4.0 * lj + n * (32 + 24 + 16),
                   "bytes_note": "union rows of center pairs (int32) + x of the centers (32 B) + f written (24 B) + row header"}
        m["rebo_center"] = {"rows_entries": sh, "bytes": 4.0 * max(sh, 0) + n * (32 + 24 + 8),
                            "bytes_note": "short rows (int32) + x (32 B) + f (24 B, atomics on neighbors not counted) + count"}
    else:
        e = ctx.counter("aeam_entries")
        m["aeam_density"] = {"rows_entries": e, "bytes": (4.0 + 8.0) * e + n * (32 + 8 + 12),
                             "bytes_note": "rows (int32) + f'(r) stored per entry (8 B) + x (32 B) + rho (8 B) + row header"}
        m["aeam_force"] = {"rows_entries": e, "bytes": (4.0 + 8.0) * e + n * (32 + 24 + 12),
                           "bytes_note": "rows (int32) + f'(r) read per entry (8 B) + x (32 B) + f (24 B) + row header"}
        m["aeam_angular"] = {"rows_entries": None, "bytes": None, "bytes_note": "0.75 % of the atoms; latency-bound"}
    return m


def load_profile_model():
    p = os.path.join(HERE, "profiles", "r02_kernel_model.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return {}
    return {}


# ----------------------------------------------------------------------------------------------- one workload
def measure(kind, args, grp, sampler, scaling="weak", legs=("e2e", "cpu", "long", "plugin")):
    import lammps_plugins_b200 as b2
    from lammps_plugins_b200 import launch

    rank, world, local_rank = grp.rank, grp.world, grp.local_rank
    w = make_workload(kind, args.rep if kind == args.rep_kind else None, world, scaling)
    natoms = len(w["x"])
    grid = w["grid"]
    mine = launch.my_atoms(w, grid, rank) if world > 1 else slice(None)

    ctx = b2.Context(local_rank)
    init_potential(ctx, kind)
    launch.join_system(ctx, grp)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))

    box = b2.make_box(w["boxlo"], w["boxhi"], w["xy"], w["xz"], w["yz"], triclinic=w["triclinic"])
    t_setup = time.time()
    ctx.system_create(kind, w["ntypes"], w["mass"], box, w["x"][mine], w["v"][mine], w["type"][mine], w["tag"][mine],
                      w["skin"], w["dt"], b2.METAL_UNITS, procgrid=grid, rank=rank, sort_every=1000)
    t_setup = time.time() - t_setup
    sz0 = ctx.system_sizes()
    barrier = grp.barrier

    # equilibration, untimed and part of the setup: the lattice starts perfect with Maxwell velocities, so the first
    # steps are not the steady state; a forced master rebuild afterwards lets every list buffer reach its steady-state
    # size (rows grow ~20 % from the cold lattice to the hot one) before anything is timed
    ctx.system_run(args.equil, 0)
    ctx.set_option("force_rebuild", 1)
    ctx.system_run(3, 0)

    ctx.system_run(args.warmup, 0)
    barrier()
    l0 = ctx.counter("kernel_launches")
    s0 = ctx.system_sizes()
    t_wall0 = time.time()
    ctx.event_record(0)
    ctx.system_run(args.steps, 0)
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1)
    t_wall1 = time.time()
    barrier()
    launches = ctx.counter("kernel_launches") - l0
    s1 = ctx.system_sizes()
    builds, inner = s1["nbuild"] - s0["nbuild"], s1["ninner"] - s0["ninner"]
    thermo = ctx.system_thermo_rows()
    ms = grp.reduce_scalar(ms, "max")
    value = natoms * args.steps / (ms * 1e-3)
    ms_per_step = ms / args.steps

    # ---- per-kernel device times: a second, shorter pass of the same loop with a CUDA-event pair around every launch
    # (kept out of the timed region: the event records themselves cost ~2 us per launch)
    ksteps = max(10, min(args.steps, 50))
    ctx.set_option("sync_timing", 1)
    ctx.kernel_stats(reset=True)
    ctx.event_record(2)
    ctx.system_run(ksteps, 0)
    ctx.event_record(3)
    ms_kpass = ctx.event_elapsed_ms(2, 3)
    kstats = ctx.kernel_stats()
    ctx.set_option("sync_timing", 0)
    barrier()

    # ---- a master rebuild on its own: one ordinary step, then one step forced through the reneighboring path
    ctx.event_record(4)
    ctx.system_run(1, 0)
    ctx.event_record(5)
    ctx.set_option("force_rebuild", 1)
    ctx.system_run(1, 0)
    ctx.event_record(6)
    plain_ms = grp.reduce_scalar(ctx.event_elapsed_ms(4, 5), "max")
    forced_ms = grp.reduce_scalar(ctx.event_elapsed_ms(5, 6), "max")
    rebuild_ms = max(forced_ms - plain_ms, 0.0)
    barrier()

    # ---- the long run (BASELINE configs[2]: "NVE 1000 steps"): natural master rebuilds inside the timed region
    long_rec = None
    if "long" in legs and args.long_steps != 0:
        nl = args.long_steps if args.long_steps > 0 else LONG_RUN_STEPS[kind]
        b0 = ctx.system_sizes()
        barrier()
        ctx.event_record(6)
        ctx.system_run(nl, 0)
        ctx.event_record(7)
        ms_long = grp.reduce_scalar(ctx.event_elapsed_ms(6, 7), "max")
        barrier()
        b1 = ctx.system_sizes()
        nb = b1["nbuild"] - b0["nbuild"]
        long_rec = {"steps": nl, "ms_per_step": ms_long / nl, "value": natoms * nl / (ms_long * 1e-3),
                    "master_rebuilds": nb, "inner_list_refreshes": b1["ninner"] - b0["ninner"],
                    "steps_between_rebuilds_measured": (nl / nb) if nb else None}
    # amortised over 1000 steps: ordinary steps + the measured rebuild cost at the measured rebuild frequency
    between = long_rec["steps_between_rebuilds_measured"] if long_rec else None
    per_step_no_rebuild = (ms - builds * rebuild_ms) / args.steps if builds else ms_per_step
    if between:
        amort_ms = per_step_no_rebuild + rebuild_ms / between
    elif long_rec:      # no rebuild within the long run: at most one per `steps` of it
        amort_ms = per_step_no_rebuild + rebuild_ms / long_rec["steps"]
    else:
        amort_ms = None

    atoms_per_gpu_max = int(grp.reduce_scalar(sz0["nlocal"], "max"))
    ghosts_per_gpu_max = int(grp.reduce_scalar(sz0["nghost"], "max"))
    migrated = int(grp.reduce_scalar(ctx.system_sizes()["nmigrated"], "sum"))
    model = kernel_model(kind, ctx, sz0["nlocal"])
    master_entries = ctx.counter("master_entries")
    tight = ctx.counter("tight_refreshes")
    p2p_ex = ctx.counter("p2p_exchanges")

    # ---- e2e: plugin-mode C-ABI call with pinned HOST buffers on every rank at once
    e2e = None
    if "e2e" in legs and not args.no_e2e:
        barrier()
        part = run_e2e(ctx, kind, w, args, local_rank, rank, world)
        t_max = grp.reduce_scalar(part["seconds"], "max")
        atoms_steps = grp.reduce_scalar(part["nlocal"] * part["steps"], "sum")
        h2d = grp.reduce_scalar(part["h2d_bytes_per_step"], "sum")
        d2h = grp.reduce_scalar(part["d2h_bytes_per_step"], "sum")
        e2e = {"value": atoms_steps / t_max, "unit": "atom-steps/s", "steps": part["steps"],
               "ms_per_step": t_max / part["steps"] * 1e3, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "path": part["path"] + ("; %d ranks concurrently" % world if world > 1 else ""),
               "frames": part["frames"], "pipelined_calls": part["pipelined_calls"], "pipelined_redos": part["pipelined_redos"],
               "neighbor_handover_ms": grp.reduce_scalar(part["handover_ms"], "max"),
               "h2d_ms_alone": grp.reduce_scalar(part["h2d_ms"], "max"), "d2h_ms_alone": grp.reduce_scalar(part["d2h_ms"], "max"),
               "checksum_f": part["checksum_f"],
               "ms_per_call_median": grp.reduce_scalar(part["ms_per_call_median"], "max"),
               "ms_per_call_min": grp.reduce_scalar(part["ms_per_call_min"], "max"),
               "ms_per_call_max": grp.reduce_scalar(part["ms_per_call_max"], "max"),
               "tight_row_derives_total": part["tight_row_derives"]}
        if between or long_rec:
            every = between or long_rec["steps"]
            e2e["value_with_handover_amortised"] = atoms_steps / (t_max + part["steps"] * 1e-3 * e2e["neighbor_handover_ms"] / every)
            e2e["handover_every_steps"] = every
    ctx_counters = {"tight_row_derives_total": tight, "atoms_migrated_total": migrated}
    try:
        fp64_peak, hbm_here = ctx.measure_peaks()
    except Exception:
        fp64_peak, hbm_here = None, None
    ctx.close()
    if rank != 0:
        return None

    # ---- roofline: per kernel its own bytes and flops, bound = the larger fraction; whole_step = headline
    hbm_peak, peak_src = measured_peaks()
    per_step = {k: v[0] / ksteps for k, v in kstats.items()}
    per_launch = {k: v[0] / max(v[1], 1) for k, v in kstats.items()}
    groups = GROUPS[kind]
    gtime = {g: sum(per_step.get(k, 0.0) for k in ks) for g, ks in groups.items()}
    prof = load_profile_model().get(kind, {})
    per_kernel = {}
    for g, t_ms in gtime.items():
        mk, pk = model.get(g, {}), prof.get(g, {})
        rec = {"ms_per_step": round(t_ms, 5), "share_of_step": t_ms / ms_per_step if ms_per_step else None,
               "own_bytes_per_step": mk.get("bytes"), "own_bytes_note": mk.get("bytes_note"), "rows_entries": mk.get("rows_entries")}
        if mk.get("bytes") and t_ms > 0:
            rec["hbm_gbs"] = mk["bytes"] / (t_ms * 1e-3) / 1e9
            rec["hbm_frac"] = rec["hbm_gbs"] / hbm_peak
        # FP64 side: the share of the FP64 pipe's issue rate the kernel's instructions take (ncu sm__inst_executed_pipe_fp64 of the
        # committed capture of the same kernel on the same workload: a property of the SASS and the rows, not of this run)
        if pk.get("fp64_pipe_pct_ncu") is not None:
            rec["fp64_frac"] = pk["fp64_pipe_pct_ncu"] / 100.0
            if fp64_peak:
                rec["fp64_tflops_equiv"] = rec["fp64_frac"] * fp64_peak
        for k in ("dram_bytes_per_step_ncu", "fp64_pipe_pct_ncu", "issue_active_pct_ncu", "l1_data_pipe_pct_ncu", "lanes_active_ncu",
                  "gathered_sectors_per_step_ncu", "ms_per_step_ncu"):
            if k in pk:
                rec[k] = pk[k]
        fr = [(rec.get("hbm_frac") or 0.0, "hbm"), (rec.get("fp64_frac") or 0.0, "fp64")]
        rec["bound"] = max(fr)[1] if max(fr)[0] > 0 else None
        per_kernel[g] = rec
    dom = max(gtime, key=gtime.get)
    d = per_kernel[dom]
    step_gbs = ALGO[kind]["bytes"] * natoms / world / (ms_per_step * 1e-3) / 1e9
    step_tflops = ALGO[kind]["flops"] * natoms / world / (ms_per_step * 1e-3) / 1e12
    use_fp64 = d.get("bound") == "fp64"
    roofline = {"bound": "hbm" if not use_fp64 else "fp64", "kernel": dom,
                "achieved": d.get("fp64_tflops_equiv") if use_fp64 else d.get("hbm_gbs"),
                "peak": fp64_peak if use_fp64 else hbm_peak, "unit": "TFLOP/s" if use_fp64 else "GB/s",
                "frac": d.get("fp64_frac") if use_fp64 else d.get("hbm_frac"),
                "traffic": d.get("dram_bytes_per_step_ncu"),
                "peak_source": ("DFMA-saturating kernel timed in place (MEASURED_PEAKS.json has no FP64 entry)" if use_fp64 else peak_src),
                "achieved_note": "per kernel: HBM side = its OWN algorithmic bytes (the rows it streams + x/f) / live kernel time; FP64 side "
                                 "= share of the FP64 pipe its instructions take (ncu, profiles/r02_kernel_model.json); bound = the larger. "
                                 "These gather kernels are limited by neither: see l1_data_pipe_pct_ncu / issue_active_pct_ncu in per_kernel. "
                                 "`whole_step` holds the headline fractions (reference formulation)",
                "kernel_ms_per_step": gtime[dom], "kernel_share_of_step": gtime[dom] / ms_per_step,
                "launches_per_step": len([k for k in groups[dom] if k in per_step]),
                "hbm_copy_measured_here_gbs": hbm_here, "per_kernel": per_kernel,
                "whole_step": {"hbm_gbs": step_gbs, "hbm_frac": step_gbs / hbm_peak, "fp64_tflops": step_tflops,
                               "fp64_peak_measured_tflops": fp64_peak,
                               "fp64_frac": (step_tflops / fp64_peak) if fp64_peak else None,
                               "fp64_frac_of_nominal": step_tflops / FP64_NOMINAL_TFLOPS,
                               "algorithmic_bytes_per_atom_step": ALGO[kind]["bytes"],
                               "algorithmic_flops_per_atom_step": ALGO[kind]["flops"],
                               "note": "reference formulation (SURVEY 8d): bytes of one pass over the full neighbor row + x, f, "
                                       "type, tag; flops as the reference writes them; FP64 peak = DFMA-saturating kernel on this GPU"}}

    # ---- CPU baseline (reference sources compiled verbatim) on the host cores
    cpu = None
    if "cpu" in legs and world == 1 and not args.no_cpu:
        try:
            cpu = run_cpu_reference(kind, args.cpu_seconds)
        except Exception as e:      # the GPU number stands on its own; say why the CPU leg is missing
            cpu = {"value": None, "unit": "atom-steps/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}

    # ---- the drop-in leg: plugin load -> Pair::compute inside a host application (one rank)
    plugin = None
    if "plugin" in legs and world == 1 and not args.no_plugin and not args.rep:
        try:
            plugin = run_plugin_leg(kind, args)
        except Exception as e:
            plugin = {"value": None, "error": repr(e)}

    rec = {
        "metric": "atom-timesteps/s", "value": value, "unit": "atom-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "pair_style": kind, "atoms": natoms, "atoms_per_gpu": atoms_per_gpu_max,
                   "ghosts_per_gpu": ghosts_per_gpu_max, "parallelism": "brick %dx%dx%d" % grid, "ensemble": "NVE dt=1fs",
                   "skin": w["skin"], "equilibration_steps_untimed": args.equil + 3,
                   "l2": "working set (master rows %.2f GB, streamed rows %.2f GB per step) >> 126 MB L2; no flush needed"
                   % (master_entries * 4 / 1e9, sum((m.get("rows_entries") or 0) for m in model.values()) * 4 / 1e9)},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_plugin": plugin,
        "clocks": sampler.summary(t_wall0, t_wall1) if sampler else None,
        "gpu_launches": launches,
        "halo": {"peer_memory_exchanges": p2p_ex,
                 "transport": "cuda-ipc peer windows over NVLink" if p2p_ex > 0 else ("nccl send/recv" if world > 1 else "self (periodic images)")},
        "kernels_ms_per_step": {k: round(v, 5) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1])},
        "kernel_groups_ms_per_step": {g: round(v, 5) for g, v in gtime.items()},
        "kernels_ms_per_launch": {k: round(v, 5) for k, v in sorted(per_launch.items(), key=lambda kv: -kv[1])[:12]},
        "kernel_timing_pass": {"steps": ksteps, "ms_per_step_with_event_pairs": ms_kpass / ksteps},
        "neighbor": dict({"rebuilds_in_timed_region": builds, "inner_list_refreshes_in_timed_region": inner,
                          "master_rebuild_ms": rebuild_ms, "step_without_rebuild_ms": plain_ms,
                          "steps_between_rebuilds_measured": between, "long_run": long_rec,
                          "value_amortised_1000": (natoms / (amort_ms * 1e-3)) if amort_ms else None,
                          "ms_per_step_amortised_1000": amort_ms, "setup_s": t_setup}, **ctx_counters),
        "thermo_last": {k: (float(v) if not isinstance(v, np.ndarray) else None) for k, v in thermo[-1].items() if k != "virial"},
    }
    if cpu and cpu.get("value"):
        rec["cpu_baseline"]["value_per_core"] = cpu["value"] / max(cpu["cores"], 1)
    return rec


# ----------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    from lammps_plugins_b200 import launch
    grp = launch.Group()
    rank, world, local_rank = grp.rank, grp.world, grp.local_rank
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run (one rank per GPU)" % args.gpus)
    sampler = None
    if rank == 0:
        sampler = ClockSampler(local_rank)
        sampler.start()
    main_kind = "aeam" if args.workload == "aeam" else "rebomos"
    line = measure(main_kind, args, grp, sampler)
    if args.workload == "all":
        subs = {}
        sub = measure("aeam", args, grp, sampler)
        if rank == 0:
            subs["aeam"] = sub
        if world > 1:
            sub = measure("aeam", args, grp, sampler, scaling="strong", legs=("long",))
            if rank == 0:
                subs["aeam_strong"] = sub
        if rank == 0:
            line["workloads"] = subs
    if sampler:
        sampler.stop_flag = True
    if rank == 0:
        emit(line)
    grp.close()


def run_e2e(ctx_sys, kind, w, args, device=0, rank=0, world=1):
    """Plugin mode: per step H2D x (pinned) -> forces on the device -> D2H f (pinned).  The atoms MOVE: the calls walk
    through successive frames of the resident trajectory (taken with the same ghost set, i.e. between two master
    rebuilds), so the displacement check and the speculative use of the inner lists see real motion.  The neighbor list
    is built on the device from the first frame, outside the timed region; its cost is reported as
    neighbor_handover_ms and amortised over the measured rebuild interval in value_with_handover_amortised."""
    import lammps_plugins_b200 as b2
    from lammps_plugins_b200 import workloads as W
    frames = []
    nb0 = ctx_sys.system_sizes()["nbuild"]
    st = ctx_sys.system_download()
    nl, ng = st["nlocal"], st["nghost"]
    nall = nl + ng
    ctx = b2.Context(device)
    init_potential(ctx, kind)
    for k in range(args.e2e_frames):
        buf = ctx.pinned_array((nall, 3))
        buf[:] = st["x"]
        frames.append(buf)
        if k + 1 < args.e2e_frames:
            ctx_sys.system_run(1, 0)
            if ctx_sys.system_sizes()["nbuild"] != nb0:
                break
            st2 = ctx_sys.system_download()
            if st2["nlocal"] != nl or st2["nghost"] != ng:
                break
            st = dict(st, x=st2["x"])
    st = dict(st, type=st["type"], tag=st["tag"])
    cs, cg, cmax = neighbor_cutoffs(kind, w["skin"])
    box = W.single_rank_box(w, cmax)
    if world > 1:       # this rank's brick (lamda bounds if triclinic), numbered x fastest like b200md_system_desc
        g = w["grid"]
        loc = (rank % g[0], (rank // g[0]) % g[1], rank // (g[0] * g[1]))
        for d in range(3):
            lo, hi = (0.0, 1.0) if w["triclinic"] else (w["boxlo"][d], w["boxhi"][d])
            box.sublo[d] = lo + (hi - lo) * (loc[d] / g[d])
            box.subhi[d] = lo + (hi - lo) * ((loc[d] + 1) / g[d]) if loc[d] < g[d] - 1 else hi
    f = ctx.pinned_array((nall, 3))
    typ, tag = st["type"], st["tag"]
    t0 = time.perf_counter()
    ctx.neigh_build(box, w["ntypes"], cs, cg, nl, ng, frames[0], typ, 1 if kind == "rebomos" else 0, w["skin"])
    handover_ms = (time.perf_counter() - t0) * 1e3
    ctx.set_option("f_overwrite", 1)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    # two-phase form: the host application's own rho / fp arrays (page-locked, as with B200MD_PIN_HOST=1)
    rho_all, fp_all = ctx.pinned_array((nall,)), ctx.pinned_array((nall,))
    rho_all[:] = 1.0
    fp_all[:] = 0.0

    def one(x):
        if kind == "rebomos":
            ctx.rebomos_compute(nl, ng, x, typ, tag, 0, 0, f=f)
        elif world == 1:
            ctx.aeam_compute(nl, ng, x, typ, tag, 0, 0, f=f)
        else:
            # two-phase form (PairAEAM::compute with the host's halo in between); the ghost fp values a host halo
            # would deliver are held at placeholders here -- same work, same transfers
            ctx.aeam_density_phase(nl, ng, x, typ, rho=rho_all, fp=fp_all)
            ctx.aeam_force_phase(rho_all, fp_all, 0, 0, f=f)

    order = list(range(len(frames))) + list(range(len(frames) - 2, 0, -1))      # forth and back along the trajectory
    for k in range(3):
        one(frames[order[k % len(order)]])
    # the two transfers on their own (what the PCIe link needs with nothing to overlap with)
    h2d_ms = min(ctx.copy_probe(frames[0], True) for _ in range(3))
    d2h_ms = min(ctx.copy_probe(f, False) for _ in range(3))
    h0, d0 = ctx.counter("h2d_bytes"), ctx.counter("d2h_bytes")
    p0, r0 = ctx.counter("pipelined_calls"), ctx.counter("pipelined_redos")
    n = max(5, min(args.steps, args.e2e_steps))
    per_call = []
    t0 = time.perf_counter()
    for k in range(n):
        ta = time.perf_counter()
        one(frames[order[(k + 3) % len(order)]])
        per_call.append((time.perf_counter() - ta) * 1e3)
    dt = time.perf_counter() - t0
    out = {"value": nl * n / dt, "unit": "atom-steps/s", "steps": n, "ms_per_step": dt / n * 1e3, "seconds": dt, "nlocal": nl,
           "ms_per_call_median": float(np.median(per_call)), "ms_per_call_min": float(np.min(per_call)),
           "ms_per_call_max": float(np.max(per_call)), "tight_row_derives": ctx.counter("tight_refreshes"),
           "h2d_bytes_per_step": (ctx.counter("h2d_bytes") - h0) // n, "d2h_bytes_per_step": (ctx.counter("d2h_bytes") - d0) // n,
           "path": "b200md_%s_compute via ctypes, pinned host x/f, atoms moving along %d frames of the resident trajectory, "
                   "device-built neighbor list reused" % (kind, len(frames)),
           "frames": len(frames), "pipelined_calls": ctx.counter("pipelined_calls") - p0,
           "pipelined_redos": ctx.counter("pipelined_redos") - r0, "handover_ms": handover_ms, "h2d_ms": h2d_ms, "d2h_ms": d2h_ms,
           "checksum_f": float(np.abs(f[:nl]).sum())}
    ctx.close()
    return out


# ----------------------------------------------------------------------------------------------- drop-in leg
def run_plugin_leg(kind, args):
    """`plugin load <B200 plugin>` + the workload's input commands + `run N` inside a host application, in a process of
    its own.  The host application is the mini engine of this repo (oracle/engine: Verlet, CommBrick, Neighbor -- what
    LAMMPS provides around Pair::compute()); the pair style is lammps_plugins_b200/<style>plugin.so and nothing of the
    oracle's pair code is loaded.  Reported: the engine's own timers (Pair = time inside Pair::compute(), i.e. the
    plugin including its transfers; Loop = the whole run loop on ONE host core) and the library's counters."""
    out = {}
    for neigh in (("device", "host") if kind == "rebomos" else ("device",)):
        stats = tempfile.NamedTemporaryFile(prefix="b200md_stats_", suffix=".jsonl", delete=False)
        stats.close()
        env = dict(os.environ, B200MD_PIN_HOST="1", B200MD_NEIGH=neigh, B200MD_STATS_FILE=stats.name, B200MD_DEVICE="0")
        cmd = [sys.executable, os.path.abspath(__file__), "--_plugin-child", kind, "--plugin-steps", str(args.plugin_steps)]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
        if r.returncode != 0:
            out[neigh] = {"error": (r.stderr or r.stdout)[-400:]}
            continue
        rec = json.loads(r.stdout.strip().splitlines()[-1])
        try:
            rec["library_counters"] = json.loads(open(stats.name).read().strip().splitlines()[-1])
        except Exception:
            rec["library_counters"] = None
        os.unlink(stats.name)
        out[neigh] = rec
    best = out.get("device", {})
    return {"value": best.get("value_pair"), "unit": "atom-steps/s", "what": "atoms x steps / time inside Pair::compute() "
            "(the plugin: neighbor hand-overs, H2D, kernels, D2H), B200MD_NEIGH=device; the run loop around it is a single "
            "host core of the stand-in application", "by_list_source": out}


def plugin_child(kind, nsteps):
    """child process of run_plugin_leg"""
    from oracle import minilmp as M
    from lammps_plugins_b200 import potentials as P
    from lammps_plugins_b200 import workloads as W
    lmp = M.MiniLmp((1, 1, 1))
    if kind == "rebomos":
        lmp.command("plugin load " + M.B200_REBOMOS_SO)
        lmp.commands(W.rebomos_bulk_script(P.default_path("MoS.REBO.set5b"), cells=(14, 13, 19)))
        lmp.command("neighbor 2.0 bin")
        lmp.command("velocity all create 300.0 12345")
    else:
        lmp.command("plugin load " + M.B200_AEAM_SO)
        lmp.commands(W.aeam_script(P.default_path("AlSi.aeam"), cells=(80, 80, 80)))
        lmp.command("velocity all create 863.0 1082337")
    natoms = lmp.get_int("natoms")
    t0 = time.time()
    lmp.commands(["fix 1 all nve", "thermo 0", "run 5"])          # setup + first steps (allocations, list hand-over)
    t_first = time.time() - t0
    nb0 = lmp.get_int("nbuild")
    lmp.command("run %d" % nsteps)
    t_loop, t_pair = lmp.get_double("time_loop"), lmp.get_double("time_pair")
    rec = {"atoms": natoms, "steps": nsteps, "loop_ms_per_step": t_loop / nsteps * 1e3, "pair_ms_per_step": t_pair / nsteps * 1e3,
           "value_pair": natoms * nsteps / t_pair, "value_loop": natoms * nsteps / t_loop,
           "neighbor_builds_in_run": lmp.get_int("nbuild") - nb0, "first_run_s": t_first,
           "list_source": os.environ.get("B200MD_NEIGH", "device")}
    lmp.close()     # destroys the pair style: the stats file is written now
    sys.stdout.write(json.dumps(rec) + "\n")


# ----------------------------------------------------------------------------------------------- CPU arm
def run_cpu_reference(kind, seconds, steps=None):
    """The reference pair style compiled verbatim (oracle/_ref) inside the mini LAMMPS engine, domain-decomposed over
    thread-ranks on the host cores (stand-in for mpirun: no MPI here)."""
    from oracle import minilmp as M
    from lammps_plugins_b200 import potentials as P
    from lammps_plugins_b200 import workloads as W
    from lammps_plugins_b200.launch import procgrid_for as pf
    ncores = host_cores()
    nranks = 1
    for cand in (64, 48, 32, 27, 24, 16, 12, 8, 6, 4, 2, 1):
        if cand <= ncores:
            nranks = cand
            break
    grid = pf(nranks)
    try:
        plugin = M.reference_plugin(kind)
        knd = "reference"
    except FileNotFoundError:
        plugin, knd = M.reference_plugin(kind, allow_port=True), "port"
    lmp = M.MiniLmp(grid)
    lmp.command("plugin load " + plugin)
    # size the sample from a per-core rate guess, then report what was actually run
    if kind == "rebomos":
        rate = 3.0e4 * nranks
        nsteps = steps or 20
        cells = max(1.0, rate * seconds / nsteps / 288.0)
        r = max(1, round(cells ** (1 / 3)))
        rx, ry, rz = max(r, grid[0]), max(r, grid[1]), max(r, grid[2])
        lmp.commands(W.rebomos_bulk_script(P.default_path("MoS.REBO.set5b"), cells=(rx, ry, rz)))
        lmp.command("velocity all create 300.0 12345")
        sample = "MoS2 bulk %dx%dx%d cells" % (rx, ry, rz)
    else:
        rate = 1.5e5 * nranks
        nsteps = steps or 20
        n = max(grid[0] * 4, round((rate * seconds / nsteps / 4.0) ** (1 / 3)))
        lmp.commands(W.aeam_script(P.default_path("AlSi.aeam"), cells=(n, n, n)))
        lmp.command("velocity all create 863.0 1082337")
        sample = "fcc Al-0.75%%Si %dx%dx%d cells" % (n, n, n)
    natoms = lmp.get_int("natoms")
    lmp.commands(["fix 1 all nve", "thermo 0", "run %d" % nsteps])
    t = lmp.get_double("time_loop")
    out = {"value": natoms * nsteps / t, "unit": "atom-steps/s", "cores": nranks, "kind": knd,
           "sample": "%s = %d atoms, %d NVE steps, %dx%dx%d thread-ranks (brick decomposition, -O2), loop %.2f s, pair %.0f%%"
           % (sample, natoms, nsteps, grid[0], grid[1], grid[2], t, 100.0 * lmp.get_double("time_pair") / t),
           "value_per_core": natoms * nsteps / t / nranks, "host_cores_visible": ncores}
    lmp.close()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kinds = ["rebomos", "aeam"] if args.workload == "all" else [args.workload]
    recs = {}
    for kind in kinds:
        times, last = [], None
        for it in range(args.warmup + args.steps):
            last = run_cpu_reference(kind, args.cpu_seconds / max(args.steps, 1) / len(kinds), steps=10)
            if it >= args.warmup:
                times.append(last["value"])
        v = float(np.mean(times))
        recs[kind] = {"impl": "reference", "metric": "atom-timesteps/s", "value": v, "unit": "atom-steps/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": "%s (bounded sample of the B200 arm's lattice)" % kind, "pair_style": kind},
                      "cpu_baseline": dict(last, value=v, value_per_core=v / max(last["cores"], 1)),
                      "e2e": {"value": v, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line = recs[kinds[0]]
    if len(kinds) > 1:
        line["workloads"] = {k: recs[k] for k in kinds[1:]}
    emit(line)


_REAL_STDOUT = None


def protect_stdout():
    """Everything but the one JSON line goes to stderr -- including C-level prints of libraries (NCCL writes its
    version banner to fd 1 when NCCL_DEBUG is set on the box)."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "rebomos", "aeam"])
    ap.add_argument("--rep", type=int, nargs=3, default=None, help="per-GPU replication (rebomos cells / fcc cells) of --workload")
    ap.add_argument("--equil", type=int, default=40, help="untimed equilibration steps before the warm-up (part of the setup)")
    ap.add_argument("--long-steps", type=int, default=-1, help="steps of the long run (-1: 1000 rebomos / 300 aeam, 0: skip)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-plugin", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=50)
    ap.add_argument("--e2e-frames", type=int, default=8)
    ap.add_argument("--plugin-steps", type=int, default=30)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (tuning experiments)")
    ap.add_argument("--_plugin-child", dest="plugin_child", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.plugin_child:
        plugin_child(args.plugin_child, args.plugin_steps)
        return
    protect_stdout()
    args.rep_kind = "aeam" if args.workload == "aeam" else "rebomos"
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        if args.steps > 5:
            args.steps, args.warmup = 3, 1          # each CPU "step" is a bounded multi-second sample
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
