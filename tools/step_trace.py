"""Per-step wall-clock trace of the GPU-resident loop (diagnostic): which steps are slow and what happened in them."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import lammps_plugins_b200 as b2

kind = sys.argv[1] if len(sys.argv) > 1 else "aeam"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 80
opts = sys.argv[3:]
w = bench.make_workload(kind, None, 1)
ctx = b2.Context(0)
bench.init_potential(ctx, kind)
for kv in opts:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
box = b2.make_box(w["boxlo"], w["boxhi"], w["xy"], w["xz"], w["yz"], triclinic=w["triclinic"])
t0 = time.time()
ctx.system_create(kind, w["ntypes"], w["mass"], box, w["x"], w["v"], w["type"], w["tag"], w["skin"], w["dt"],
                  b2.METAL_UNITS, procgrid=(1, 1, 1), rank=0, sort_every=1000)
print("setup %.3f s" % (time.time() - t0))
ctx.system_run(5, 0)
prev = ctx.system_sizes()
rows = []
for s in range(nsteps):
    t = time.perf_counter()
    ctx.system_run(1, 0)
    dt = (time.perf_counter() - t) * 1e3
    cur = ctx.system_sizes()
    rows.append((s, dt, cur["nbuild"] - prev["nbuild"], cur["ninner"] - prev["ninner"]))
    prev = cur
d = np.array([r[1] for r in rows])
print("median %.3f ms, mean %.3f ms, max %.3f ms" % (np.median(d), d.mean(), d.max()))
for r in rows:
    if r[1] > 1.5 * np.median(d) or r[2] or r[3]:
        print("step %3d  %8.3f ms  master rebuild %d  inner refresh %d" % r)
