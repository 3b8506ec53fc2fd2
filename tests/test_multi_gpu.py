"""Multi-GPU (slab decomposition) tests.  The GPU part launches tests/dist_worker.py under torchrun on 2 GPUs;
the CPU part covers the host-side logic (slab plan, ownership rule, gather) with a world_size-2 gloo group."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_plan_decomposition_host_only():
    from moldyn_b200 import distributed as mdd
    from moldyn_b200 import MdError
    box = [33.38339, 33.38339, 33.38339]
    edges = []
    for r in range(4):
        p = mdd.plan_decomposition(1000, box, 1.709, 4, r)
        assert p["left"] == (r + 3) % 4 and p["right"] == (r + 1) % 4
        assert p["capacity"] >= 250
        edges.append((p["x_lo"], p["x_hi"]))
    assert edges[0][0] == 0.0 and edges[-1][1] == box[0]
    for a, b in zip(edges, edges[1:]):
        assert a[1] == b[0]
    with pytest.raises(MdError) as e:   # slabs narrower than 2.1 x (r_cut + skin)
        mdd.plan_decomposition(1000, box, 1.709, 16, 0)
    assert e.value.code == 8
    # ownership rule partitions every coordinate, including the faces and out-of-box values
    x = np.array([0.0, 8.3458475, 8.345847499999, 33.38338999, -0.1, 33.5, 16.691695])
    assert list(mdd.owner_of(x, box[0], 4)) == [0, 1, 0, 3, 3, 0, 2]


def _gloo_worker(rank, world, port, n):
    import torch.distributed as dist
    from moldyn_b200 import distributed as mdd
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    pos, vel = rng.uniform(0, 10, (n, 3)), rng.normal(size=(n, 3))
    mine = np.nonzero(mdd.owner_of(pos[:, 0], 10.0, world) == rank)[0][::-1].copy()   # any local order
    local = {"ids": mine, "position": pos[mine], "velocity": vel[mine], "force": pos[mine] * 2.0,
             "potential": pos[mine, 0], "temp": vel[mine, 1], "box": np.array([10.0, 10.0, 10.0])}
    got = mdd.gather_by_id(local, n)
    if rank == 0:
        assert np.array_equal(got["position"], pos) and np.array_equal(got["velocity"], vel)
        assert np.array_equal(got["force"], pos * 2.0) and np.array_equal(got["temp"], vel[:, 1])
        assert sum(got["owned_per_rank"]) == n
    else:
        assert got is None
    # a lost atom must be detected
    if rank == 1:
        for k in local:
            if k != "box":
                local[k] = local[k][1:]
    try:
        mdd.gather_by_id(local, n)
        ok = rank != 0
    except RuntimeError:
        ok = rank == 0
    assert ok
    dist.destroy_process_group()


def test_gather_by_id_world_size_2_gloo():
    import torch.multiprocessing as mp
    mp.spawn(_gloo_worker, args=(2, free_port(), 501), nprocs=2, join=True)


@pytest.mark.gpu
def test_two_gpu_slab_decomposition_matches_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert "DIST_OK" in r.stdout, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
