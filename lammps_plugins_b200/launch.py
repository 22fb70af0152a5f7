"""One-process-per-GPU launch plumbing for the GPU-resident system (bench.py and multi-GPU drivers).

``torch.distributed`` is used for exactly three things, none of them on the data path: agreeing on the NCCL
unique id the library's own communicator is built from (``b200md_system_comm_init``), barriers around timed
regions, and max/sum reductions of a few host scalars.  Halo exchange, migration and thermo reductions run
inside ``libb200md.so`` on its own NCCL communicator (system.cu).  Everything here also works on the
``gloo`` backend without a GPU, which is how the CPU tests cover it (tests/test_launch_cpu.py).
"""
from __future__ import annotations

import os

import numpy as np

from . import workloads as W

GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def procgrid_for(nranks: int):
    """Brick grid for n ranks: the SURVEY 8(e) grids for 1/2/4/8, else the most cubic factorisation."""
    if nranks in GRIDS:
        return GRIDS[nranks]
    best = (nranks, 1, 1)
    for a in range(1, nranks + 1):
        if nranks % a:
            continue
        for b in range(1, nranks // a + 1):
            if (nranks // a) % b:
                continue
            t = tuple(sorted((a, b, nranks // a // b), reverse=True))
            if max(t) - min(t) < max(best) - min(best):
                best = t
    return best


def dist_env():
    """(rank, world_size, local_rank) from the torchrun environment (1 process: 0, 1, 0)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


class Group:
    """Thin wrapper over a torch.distributed process group (or nothing when world_size == 1)."""

    def __init__(self, backend: str | None = None):
        self.rank, self.world, self.local_rank = dist_env()
        self.dist = None
        self.device = "cpu"
        if self.world > 1:
            import torch
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                self.device = "cuda"
            if not dist.is_initialized():
                dist.init_process_group(backend)
            self.dist = dist
            self.backend = backend

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
            if self.device == "cuda":
                import torch
                torch.cuda.synchronize()

    def broadcast_bytes(self, payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
        """rank `src` provides `payload` (nbytes long); every rank returns it"""
        if self.dist is None:
            return bytes(payload)
        import torch
        buf = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        if self.rank == src:
            assert payload is not None and len(payload) == nbytes
            buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
        self.dist.broadcast(buf, src)
        return bytes(buf.cpu().numpy().tobytes())

    def reduce_scalar(self, value: float, op: str = "max") -> float:
        if self.dist is None:
            return float(value)
        import torch
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()


def nccl_unique_id(group: Group, make_id) -> bytes:
    """128-byte ncclUniqueId created on rank 0 by `make_id()` (b200md_nccl_unique_id) and shared with all ranks"""
    raw = make_id() if group.rank == 0 else None
    return group.broadcast_bytes(raw, 128, 0)


def my_atoms(w: dict, grid, rank: int):
    """Boolean mask of the atoms of workload `w` that rank `rank` owns under the uniform brick decomposition
    (lamda space, ranks numbered x fastest -- the numbering of b200md_system_desc.procgrid)."""
    owner = W.brick_owner(w["x"], w["boxlo"], w["boxhi"], w["xy"], w["xz"], w["yz"], grid)
    return owner == rank


def join_system(ctx, group: Group):
    """Give `ctx` its NCCL communicator for the GPU-resident system (no-op for one rank)."""
    if group.world == 1:
        return
    import ctypes

    def make_id():
        raw = ctypes.create_string_buffer(128)
        ctx._check(ctx.L.b200md_nccl_unique_id(raw))
        return raw.raw

    ctx.comm_init_nccl(nccl_unique_id(group, make_id), group.world, group.rank)
