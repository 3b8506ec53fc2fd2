"""torch.distributed plumbing around the multi-GPU C ABI (md_comm_init / md_download_local).

One process per GPU (torchrun).  PyTorch is only the rendezvous here: it broadcasts the NCCL unique id the library
needs and gathers per-rank results for tests; the halo exchange and the per-step all-gather of the reduction sums run
inside the library on its own NCCL communicator.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from .solver import Solver


def plan_decomposition(n, box, r_list, nranks, rank):
    """Host-only view of the slab a rank owns (md_plan_decomposition): dict(x_lo, x_hi, left, right, capacity)."""
    box = np.ascontiguousarray(box, dtype=np.float64)
    lo, hi, cap = C.c_double(), C.c_double(), C.c_int64()
    left, right = C.c_int(), C.c_int()
    rc = _ffi.lib().md_plan_decomposition(int(n), box.ctypes.data_as(C.c_void_p), float(r_list), int(nranks),
                                          int(rank), C.byref(lo), C.byref(hi), C.byref(left), C.byref(right),
                                          C.byref(cap))
    if rc != _ffi.MD_OK:
        raise _ffi.MdError(rc, "md_plan_decomposition")
    return {"x_lo": lo.value, "x_hi": hi.value, "left": left.value, "right": right.value, "capacity": cap.value}


def owner_of(x, box_x, nranks):
    """Rank that owns coordinate x: min(R-1, int(frac(x / Lx) * R)) — the library's rule (k_flag_owned)."""
    f = np.asarray(x, dtype=np.float64) / box_x
    f = f - np.floor(f)
    return np.minimum((f * float(nranks)).astype(np.int64), nranks - 1)


def init_solver_comm(solver: Solver, group=None):
    """Rank 0 creates the NCCL unique id, everyone receives it through torch.distributed, md_comm_init follows."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [Solver.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    solver.comm_init(rank, world, box[0])
    return rank, world


def gather_by_id(local: dict, n_global: int, group=None, dst=0):
    """Reassembles per-rank `download_local()` dicts into full arrays in upload order on rank `dst` (None elsewhere).
    Checks that every upload index is owned by exactly one rank."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    parts = [None] * world if rank == dst else None
    dist.gather_object(local, parts, dst=dst, group=group)
    if rank != dst:
        return None
    out = {}
    seen = np.zeros(n_global, dtype=np.int64)
    for key in ("position", "velocity", "force"):
        out[key] = np.zeros((n_global, 3))
    for key in ("potential", "temp"):
        out[key] = np.zeros(n_global)
    for p in parts:
        ids = p["ids"]
        seen[ids] += 1
        for key in ("position", "velocity", "force", "potential", "temp"):
            out[key][ids] = p[key]
    if not np.all(seen == 1):
        raise RuntimeError(f"decomposition lost or duplicated atoms: {np.sum(seen == 0)} missing, "
                           f"{np.sum(seen > 1)} duplicated")
    out["box"] = parts[0]["box"]
    out["owned_per_rank"] = [len(p["ids"]) for p in parts]
    return out
