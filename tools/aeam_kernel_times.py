"""diagnostic: AEAM pair kernels timed (a) inside the resident loop, (b) in plugin-mode calls (PCIe copies in between)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import lammps_plugins_b200 as b2
from lammps_plugins_b200 import workloads as W
kind = "aeam"
w = bench.make_workload(kind, None, 1)
ctx = b2.Context(0)
bench.init_potential(ctx, kind)
box = b2.make_box(w["boxlo"], w["boxhi"], w["xy"], w["xz"], w["yz"], triclinic=w["triclinic"])
ctx.system_create(kind, w["ntypes"], w["mass"], box, w["x"], w["v"], w["type"], w["tag"], w["skin"], w["dt"],
                  b2.METAL_UNITS, procgrid=(1, 1, 1), rank=0, sort_every=1000)
ctx.system_run(43, 0)
def stats(c, n):
    ks = c.kernel_stats()
    return {k: round(v[0] / max(v[1], 1), 4) for k, v in ks.items() if k in ("aeam_density", "aeam_force", "aeam_force_ang", "aeam_embed")}
ctx.set_option("sync_timing", 1); ctx.kernel_stats(reset=True)
ctx.system_run(10, 0)
print("resident loop, 10 steps back to back :", stats(ctx, 10), flush=True)
ctx.kernel_stats(reset=True)
for k in range(10):
    ctx.system_run(1, 0); time.sleep(0.05)
print("resident loop, 50 ms idle between steps:", stats(ctx, 10), flush=True)
ctx.set_option("sync_timing", 0)
st = ctx.system_download()
nl, ng = st["nlocal"], st["nghost"]
c2 = b2.Context(0); bench.init_potential(c2, kind)
x = c2.pinned_array((nl + ng, 3)); x[:] = st["x"]
f = c2.pinned_array((nl + ng, 3))
cs, cg, cmax = bench.neighbor_cutoffs(kind, w["skin"])
bx = W.single_rank_box(w, cmax)
c2.neigh_build(bx, w["ntypes"], cs, cg, nl, ng, x, st["type"], 0, w["skin"])
c2.set_option("f_overwrite", 1)
for k in range(3): c2.aeam_compute(nl, ng, x, st["type"], st["tag"], 0, 0, f=f)
c2.set_option("sync_timing", 1); c2.kernel_stats(reset=True)
for k in range(10): c2.aeam_compute(nl, ng, x, st["type"], st["tag"], 0, 0, f=f)
print("plugin-mode calls                     :", stats(c2, 10), flush=True)
