"""N > 1 host-side path on CPU: world_size-2 (and 4) `gloo` jobs launched exactly like the driver launches
bench.py (python -m torch.distributed.run ... --master-addr 127.0.0.1)."""
import os
import socket
import subprocess
import sys

import pytest

import support as S
from lammps_plugins_b200 import launch

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_ranks_partition_and_share_id(oracle_built, world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(HERE, "_launch_worker.py"), str(world)]
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("LAUNCH_OK") == world, r.stdout[-3000:]


def test_procgrid_choices():
    assert [launch.procgrid_for(n) for n in (1, 2, 4, 8)] == [(1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2)]
    for n in (3, 6, 12, 16):
        g = launch.procgrid_for(n)
        assert g[0] * g[1] * g[2] == n


def test_reference_arm_only_rank0_prints(oracle_built):
    """bench.py --impl reference under torchrun: rank 0 runs and prints one JSON line, the other rank exits 0"""
    import json
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(os.path.dirname(HERE), "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0", "--cpu-seconds", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0


def test_bench_reference_arm_prints_one_contract_line():
    """bench.py --impl reference (the CPU arm the driver runs next to the B200 arm): exactly one JSON line on stdout
    with the contract's keys, `impl: reference`, a cpu_baseline describing the run and an e2e block with zero
    transfer bytes."""
    import json
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(repo, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--cpu-seconds", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "atom-timesteps/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
