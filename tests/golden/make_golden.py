#!/usr/bin/env python
"""Regenerate the committed fixtures under tests/golden/ from the reference tree.

Run HERE (the build container, where /root/reference exists); the GPU box has no
reference tree and only reads the generated files.

  potentials : the numbers of MoS.REBO.set5b / AlSi.aeam, stored as JSON / npz (data, not source);
               tests re-emit them in the on-disk formats the host classes parse
               (tests/support.py: write_rebomos_file, write_aeam_file) and check, when the reference
               tree is present, that the re-emitted files parse to bit-identical doubles.
  log goldens: the thermo table and neighbor statistics of log.rebomos-bulk.1 / .4
  force fixtures (--forces): inputs + outputs of the reference pair styles compiled verbatim
               (oracle/_ref) on small perturbed configurations.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("B200MD_REFERENCE", "/root/reference")


def make_rebomos():
    lines = open(os.path.join(REF, "USER-REBOMOS", "MoS.REBO.set5b")).read().splitlines()
    header = lines[0]
    params = []
    for ln in lines[1:]:
        body = ln.split("#")[0].split()
        if not body:
            continue
        params.append([body[1] if len(body) > 1 else "", repr(float(body[0]))])
    assert len(params) == 61, len(params)
    json.dump({"source": "USER-REBOMOS/MoS.REBO.set5b", "header": header, "params": params},
              open(os.path.join(HERE, "rebomos_set5b.json"), "w"), indent=0)


def make_aeam():
    lines = open(os.path.join(REF, "USER-AEAM", "AlSi.aeam")).read().splitlines()
    head = lines[:11]
    w = lines[11].split()
    nel, nnon, nang = int(w[0]), int(w[1]), int(w[2])
    names = w[3:3 + nel]
    nrho, drho, mass = [], [], []
    k = 12
    for i in range(nel):
        t = lines[k].split(); k += 1
        nrho.append(int(t[0])); drho.append(float(t[1])); mass.append(float(t[2]))
    nr, dr, cut = [], [], []
    for i in range(nel * nel):
        t = lines[k].split(); k += 1
        nr.append(int(t[0])); dr.append(float(t[1])); cut.append(float(t[2]))
    vals = np.array([float(v) for ln in lines[k:] for v in ln.split("#")[0].split()])
    need = sum(nrho) + sum(nr) + sum(nr[i * nel + j] for i in range(nel) for j in range(i + 1))
    assert len(vals) == need, (len(vals), need)
    np.savez_compressed(os.path.join(HERE, "alsi_aeam.npz"), header=np.array(head), nelements=nel,
                        nnonangular=nnon, nangular=nang, names=np.array(names), nrho=np.array(nrho),
                        drho=np.array(drho), mass=np.array(mass), nr=np.array(nr), dr=np.array(dr),
                        cut=np.array(cut), values=vals)


def make_logs():
    out = {}
    for tag in ("1", "4"):
        txt = open(os.path.join(REF, "USER-REBOMOS", "log.rebomos-bulk." + tag)).read()
        rows = []
        grab = False
        for ln in txt.splitlines():
            if ln.strip().startswith("Step"):
                grab = True
                continue
            if grab:
                if ln.startswith("Loop time"):
                    break
                rows.append([float(v) for v in ln.split()])
        d = {"thermo_columns": ["step", "temp", "press", "pe", "ke", "cellgamma", "vol"], "thermo": rows}
        d["nghost_ave_max_min"] = [float(v) for v in re.search(r"Nghost:\s+(\S+) ave\s+(\S+) max\s+(\S+) min", txt).groups()]
        d["nlocal_ave_max_min"] = [float(v) for v in re.search(r"Nlocal:\s+(\S+) ave\s+(\S+) max\s+(\S+) min", txt).groups()]
        d["fullnghs_ave"] = float(re.search(r"FullNghs:\s+(\S+) ave", txt).group(1))
        d["total_neighbors"] = int(re.search(r"Total # of neighbors = (\d+)", txt).group(1))
        d["builds"] = int(re.search(r"Neighbor list builds = (\d+)", txt).group(1))
        d["katom_step_per_s"] = float(re.search(r"(\S+) katom-step/s", txt).group(1))
        d["procgrid"] = [int(v) for v in re.search(r"(\d+) by (\d+) by (\d+) MPI processor grid", txt).groups()]
        out["log.rebomos-bulk." + tag] = d
    json.dump(out, open(os.path.join(HERE, "log_rebomos_bulk.json"), "w"), indent=1)
    # the two shipped inputs, as command lists (comments stripped, continuations joined)
    scripts = {}
    for name, path in (("in.rebomos-bulk", "USER-REBOMOS/in.rebomos-bulk"), ("sample.in", "USER-AEAM/sample.in")):
        acc, cmds = "", []
        for ln in open(os.path.join(REF, path)).read().splitlines():
            ln = ln.split("#")[0].rstrip()
            if ln.endswith("&"):
                acc += ln[:-1] + " "
                continue
            acc += ln
            if acc.strip():
                cmds.append(" ".join(acc.split()))
            acc = ""
        scripts[name] = cmds
    json.dump(scripts, open(os.path.join(HERE, "input_scripts.json"), "w"), indent=1)


if __name__ == "__main__":
    make_rebomos()
    make_aeam()
    make_logs()
    if "--forces" in sys.argv:
        sys.path.insert(0, os.path.dirname(HERE))
        import support
        support.make_force_fixtures(HERE)
    print("golden fixtures written to", HERE)
