"""Multi-GPU check of the NCCL transport (run under torchrun on an N-GPU box; not collected by pytest, the
round-end GPU tests run on one GPU and cover the same code through the loopback transport):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/nccl_check.py

A hot system with a small skin is advanced on N ranks (halo over ncclSend/ncclRecv, migration at every
rebuild); rank 0 also runs the SAME atoms on one rank.  Checks: atoms conserved, migration happened, same
number of rebuilds, thermo equal to 1e-9 relative at every 10th step, total energy conserved."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import lammps_plugins_b200 as b2  # noqa: E402
import support as S  # noqa: E402
from lammps_plugins_b200 import launch, workloads as W  # noqa: E402


def tables():
    t = S.load_aeam_fixture()
    return {k: t[k] for k in ("nelements", "nnonangular", "nrho", "drho", "nr", "dr", "cut", "frho", "rhor", "z2r")}


def main():
    grp = launch.Group()
    grid = launch.procgrid_for(grp.world)
    ok = True
    for style in ("aeam", "rebomos"):
        if style == "aeam":
            w = W.fcc_alsi((12 * grid[0], 12 * grid[1], 12 * grid[2]), 0.02, 7683797)
            w["v"] = W.maxwell_velocities(w["type"], w["mass"], 2500.0, 1082337)
            skin, steps = 0.4, 60
        else:
            w = W.mos2_bulk(3 * grid[0], 2 * grid[1], 2 * grid[2])
            w["v"] = W.maxwell_velocities(w["type"], w["mass"], 3000.0, 12345)
            skin, steps = 0.5, 60
        box = b2.make_box(w["boxlo"], w["boxhi"], w["xy"], w["xz"], w["yz"], triclinic=w["triclinic"])

        def system(ctx, mask, g, rank):
            if style == "aeam":
                ctx.aeam_init(tables())
            else:
                ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
            ctx.system_create(style, w["ntypes"], w["mass"], box, w["x"][mask], w["v"][mask], w["type"][mask],
                              w["tag"][mask], skin, 0.001, b2.METAL_UNITS, procgrid=g, rank=rank, sort_every=1000)

        ctx = b2.Context(grp.local_rank)
        if style == "aeam":
            ctx.aeam_init(tables())
        else:
            ctx.rebomos_init(S.rebomos_params_struct(), [0, 1])
        launch.join_system(ctx, grp)
        system(ctx, launch.my_atoms(w, grid, grp.rank) if grp.world > 1 else slice(None), grid, grp.rank)
        ctx.system_run(steps, 10)
        p2p = ctx.counter("p2p_exchanges")
        sz = ctx.system_sizes()
        nl_sum = int(grp.reduce_scalar(sz["nlocal"], "sum"))
        moved = int(grp.reduce_scalar(sz["nmigrated"], "sum"))
        rows = ctx.system_thermo_rows()
        if grp.rank == 0:
            one = b2.Context(grp.local_rank)
            system(one, slice(None), (1, 1, 1), 0)
            one.system_run(steps, 10)
            ref = one.system_thermo_rows()
            worst = 0.0
            for q, g in zip(rows, ref):
                worst = max(worst, abs(q["pe"] - g["pe"]) / abs(g["pe"]), abs(q["ke"] - g["ke"]) / max(abs(g["ke"]), 1.0),
                            abs(q["press"] - g["press"]) / max(abs(g["press"]), 1.0))
            e0, e1 = rows[0]["pe"] + rows[0]["ke"], rows[-1]["pe"] + rows[-1]["ke"]
            good = (nl_sum == len(w["x"]) and sz["natoms"] == len(w["x"]) and moved > 0 and worst < 1e-9
                    and sz["nbuild"] == one.system_sizes()["nbuild"] and abs(e1 - e0) < 1e-3 * abs(rows[0]["ke"] + 1.0))
            print("NCCL_CHECK %s ranks %d grid %s atoms %d migrated %d builds %d/%d worst_rel %.2e drift %.3e "
                  "peer-memory exchanges %d %s"
                  % (style, grp.world, grid, len(w["x"]), moved, sz["nbuild"], one.system_sizes()["nbuild"], worst,
                     e1 - e0, p2p, "OK" if good else "FAIL"), flush=True)
            ok = ok and good
            one.close()
        grp.barrier()
        ctx.close()
    grp.close()
    if grp.rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
