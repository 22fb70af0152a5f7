"""Digest of `ncu --set full` captures -> profiles/r02_kernel_model.json (read by bench.py) + a markdown table.

  ncu -i gpurun_out/r02b_rebomos.ncu-rep --page raw --csv > /tmp/rb_raw.csv
  ncu -i gpurun_out/r02b_aeam.ncu-rep    --page raw --csv > /tmp/ra_raw.csv
  python tools/ncu_model.py /tmp/rb_raw.csv /tmp/ra_raw.csv > profiles/r02_ncu_table.md

The first launch of every kernel template in a capture is taken (captures are of steady-state steps)."""
import csv
import json
import os
import sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}

GROUPS = {
    "rebomos": {
        "rebo_center": ["rebo_center_kernel"],
        "lj": ["lj_pair_kernel"],
        "derive_tight": ["derive_tight_short_kernel", "derive_tight_lj_kernel"],
    },
    "aeam": {
        "aeam_force": ["aeam_force_df_kernel"],
        "aeam_density": ["aeam_density_kernel"],
        "aeam_angular": ["aeam_force_ang_kernel", "aeam_density_ang_kernel"],
        "rebuild": ["neigh_rows_kernel", "aeam_build_inner_kernel"],
    },
}


def read(path):
    rows = list(csv.reader(open(path)))
    h, u = rows[0], rows[1]

    def val(r, k, scale=False):
        if k not in h:
            return None
        try:
            v = float(r[h.index(k)])
        except ValueError:
            return None
        return v * SCALE.get(u[h.index(k)], 1.0) if scale else v

    out, seen = [], set()
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        short = name.split("(")[0].replace("void ", "")
        if short in seen:
            continue
        seen.add(short)
        out.append({
            "kernel": short,
            "ms": val(r, "gpu__time_duration.sum", True),
            "registers": val(r, "launch__registers_per_thread"),
            "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "fp64_pipe_pct": val(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "l1_data_pipe_pct": val(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "lanes_active": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            "dram_bytes": (val(r, "dram__bytes_read.sum", True) or 0.0) + (val(r, "dram__bytes_write.sum", True) or 0.0),
            "gathered_sectors": val(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
            "inst_executed": val(r, "smsp__inst_executed.sum"),
        })
    return out


def main():
    launches = []
    for p in sys.argv[1:]:
        launches += read(p)
    model = {"_source": "ncu --set full --clock-control none --import-source on, steady-state steps of the default workloads "
                        "(995 904 MoS2 atoms / 2 048 000 Al-Si atoms), captures gpurun_out/r02b_rebomos.ncu-rep and "
                        "r02b_aeam.ncu-rep (round 2, final kernels; digest by tools/ncu_model.py); percentages are "
                        "time-weighted over the launches of a kernel template; fp64_pipe_pct = sm__inst_executed_pipe_fp64 "
                        "(fraction of the FP64 pipe's issue rate)"}
    for kind, groups in GROUPS.items():
        model[kind] = {}
        for gname, prefixes in groups.items():
            ls = [l for l in launches if any(l["kernel"].startswith(p) for p in prefixes)]
            if not ls:
                continue
            t = sum(l["ms"] for l in ls)

            def w(key):
                return round(sum((l[key] or 0.0) * l["ms"] for l in ls) / t, 1)

            model[kind][gname] = {
                "ms_per_step_ncu": round(t, 5), "fp64_pipe_pct_ncu": w("fp64_pipe_pct"),
                "issue_active_pct_ncu": w("issue_active_pct"), "l1_data_pipe_pct_ncu": w("l1_data_pipe_pct"),
                "lanes_active_ncu": w("lanes_active"), "dram_bytes_per_step_ncu": sum(l["dram_bytes"] for l in ls),
                "gathered_sectors_per_step_ncu": sum(l["gathered_sectors"] or 0.0 for l in ls),
                "launches": [{k: (round(v, 5) if isinstance(v, float) else v) for k, v in l.items()} for l in ls],
            }
    json.dump(model, open(os.path.join(HERE, "profiles", "r02_kernel_model.json"), "w"), indent=1)
    print("| launch | ms | regs | warps active | FP64 pipe | issue active | L1 data pipe | lanes active | DRAM bytes | warp instructions |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for l in launches:
        print("| `%s` | %.3f | %d | %.0f %% | %.1f %% | %.1f %% | %.1f %% | %.1f | %.0f MB | %.0f M |" % (
            l["kernel"], l["ms"], l["registers"], l["warps_active_pct"], l["fp64_pipe_pct"], l["issue_active_pct"] or 0.0,
            l["l1_data_pipe_pct"], l["lanes_active"], l["dram_bytes"] / 1e6, (l["inst_executed"] or 0.0) / 1e6))


if __name__ == "__main__":
    main()
