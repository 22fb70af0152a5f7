"""diagnostic: plugin-mode call before / after a master rebuild and a long run of the resident loop (found the straggler problem: one wrapped atom made the first range wait for the whole upload)"""
import sys, os, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import lammps_plugins_b200 as b2
from lammps_plugins_b200 import workloads as W
kind = "rebomos"
w = bench.make_workload(kind, None, 1)
ctx = b2.Context(0)
bench.init_potential(ctx, kind)
box = b2.make_box(w["boxlo"], w["boxhi"], w["xy"], w["xz"], w["yz"], triclinic=w["triclinic"])
ctx.system_create(kind, w["ntypes"], w["mass"], box, w["x"], w["v"], w["type"], w["tag"], w["skin"], w["dt"],
                  b2.METAL_UNITS, procgrid=(1, 1, 1), rank=0, sort_every=1000)
ctx.system_run(43, 0)
def show(tag):
    st = ctx.system_download()
    nl, ng = st["nlocal"], st["nghost"]
    c2 = b2.Context(0)
    bench.init_potential(c2, kind)
    x = c2.pinned_array((nl + ng, 3)); x[:] = st["x"]
    f = c2.pinned_array((nl + ng, 3))
    cs, cg, cmax = bench.neighbor_cutoffs(kind, w["skin"])
    bx = W.single_rank_box(w, cmax)
    c2.neigh_build(bx, w["ntypes"], cs, cg, nl, ng, x, st["type"], 1, w["skin"])
    c2.set_option("f_overwrite", 1)
    for k in range(3): c2.rebomos_compute(nl, ng, x, st["type"], st["tag"], 0, 0, f=f)
    t = time.perf_counter()
    for k in range(20): c2.rebomos_compute(nl, ng, x, st["type"], st["tag"], 0, 0, f=f)
    dt = (time.perf_counter() - t) / 20 * 1e3
    c2.set_option("sync_timing", 1); c2.kernel_stats(reset=True)
    for k in range(5): c2.rebomos_compute(nl, ng, x, st["type"], st["tag"], 0, 0, f=f)
    ks = c2.kernel_stats()
    # index locality: how far (in atom index) the first/last thirds of the tag order sit
    xs = st["x"][:nl]
    print(tag, "ms/call %.3f" % dt, "nghost", ng, "z of atoms 0, n/2, n-1:", xs[0, 2].round(1), xs[nl // 2, 2].round(1), xs[-1, 2].round(1),
          "| y:", xs[0, 1].round(1), xs[nl // 2, 1].round(1), xs[-1, 1].round(1), flush=True)
    print("   ", {k: (round(v[0] / 5, 3), v[1] // 5) for k, v in sorted(ks.items(), key=lambda kv: -kv[1][0])[:12]}, flush=True)
    c2.close()
show("fresh")
ctx.set_option("force_rebuild", 1)
ctx.system_run(2, 0)
show("after a forced rebuild")
ctx.system_run(1000, 0)
show("after +1000 steps")
