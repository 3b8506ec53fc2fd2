"""Multi-GPU parity worker — run under torchrun (one process per GPU), see tests/test_multi_gpu.py.

Every rank builds the same seeded state, the library decomposes it into x-slabs, and rank 0 compares the gathered
result with the CPU oracle: exact-mode forces and NVE trajectories bit-identical (pair sets and ascending-index sums
do not depend on the decomposition), thermostat/barostat runs within 1e-8."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import moldyn_b200 as md  # noqa: E402
from moldyn_b200 import distributed as mdd  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from helpers import dense_gas, gas, liquid  # noqa: E402

DT = 0.002


def run_case(name, o, exact, n_steps, thermostat=None, barostat=None, rank=0):
    lj = orc.LennardJones()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    with md.Solver(device=local, exact=exact) as s:
        mdd.init_solver_comm(s)
        s.upload_arrays(o.pos, o.vel, o.mass, o.box)
        s.update_force()
        got0 = mdd.gather_by_id(s.download_local(), o.n)
        gth = (md.Thermostat.Berendsen(thermostat[0]), thermostat[1]) if thermostat else None
        gba = (md.Barostat.Berendsen(barostat[0], barostat[1]), barostat[2]) if barostat else None
        s.step(n_steps // 2, DT, thermostat=gth, barostat=gba)
        s.step(n_steps - n_steps // 2, DT, thermostat=gth, barostat=gba)
        got1 = mdd.gather_by_id(s.download_local(), o.n)
        m = s.macro()
        stats = s.stats()
    all_stats = [None] * dist.get_world_size() if rank == 0 else None
    dist.gather_object(stats, all_stats, dst=0)
    if rank != 0:
        return
    ref = o.copy()
    orc.update_force(lj, ref, mode="cells")
    if exact:
        assert np.array_equal(got0["force"], ref.force), name
        assert np.array_equal(got0["potential"], ref.pot) and np.array_equal(got0["temp"], ref.vir), name
    else:
        scale = max(np.abs(ref.force).max(), 1.0)
        assert np.abs(got0["force"] - ref.force).max() <= 1e-10 * scale * 100, name
    oth = orc.Thermostat(orc.Thermostat.BERENDSEN, *thermostat) if thermostat else None
    oba = orc.Barostat(*barostat) if barostat else None
    orc.step(lj, ref, DT, thermostat=oth, barostat=oba, mode="cells", n_steps=n_steps)
    if exact and not thermostat and not barostat:
        assert np.array_equal(got1["position"], ref.pos), name
        assert np.array_equal(got1["velocity"], ref.vel), name
        assert np.array_equal(got1["force"], ref.force), name
    else:
        dx = np.abs(got1["position"] - ref.pos)
        dx = np.minimum(dx, np.abs(dx - ref.box))
        assert dx.max() <= 1e-8 and np.abs(got1["velocity"] - ref.vel).max() <= 1e-8, name
        assert np.abs(got1["box"] / ref.box - 1.0).max() <= 1e-12, name
    om = orc.macro(ref)
    for key in ("kinetic", "thermal", "potential", "temperature", "pressure"):
        assert abs(m[key] - om[key]) <= 1e-9 * max(1.0, abs(om[key])), (name, key, m[key], om[key])
    migrated = sum(st["migrated"] for st in all_stats)
    print(f"  {name}: ok  owned={got1['owned_per_rank']} ghosts={[st['n_ghost'] for st in all_stats]} "
          f"migrated={migrated} rebuilds={all_stats[0]['rebuilds']}", flush=True)
    return migrated


def perturbed_gas(n_side, seed):
    o = gas(n_side)
    o.pos += np.random.default_rng(seed).uniform(-1.45, 1.45, o.pos.shape)
    orc.apply_boundary_conditions(o)
    return o


def run_c3_rows(rank):
    """BASELINE config C3 (argon 100^3 = 10^6 atoms) decomposed over the ranks, MD_FORCE_FAST: sampled force rows against the
    reference's row scan (1e-10 bar relative to max(|F_i|, RMS) — the perturbed gas has no symmetric cancellation), the
    macro parameters against the oracle, then 30 NVT steps: every atom still owned exactly once, momentum conserved."""
    o = perturbed_gas(100, 1)
    lj = orc.LennardJones()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    with md.Solver(device=local) as s:
        mdd.init_solver_comm(s)
        s.upload_arrays(o.pos, o.vel, o.mass, o.box)
        s.update_force()
        got = mdd.gather_by_id(s.download_local(), o.n)
        m0 = s.macro()
        s.step(30, DT, thermostat=(md.Thermostat.Berendsen(10.0), 300.0))
        got1 = mdd.gather_by_id(s.download_local(), o.n)
        m1 = s.macro()
        stats = s.stats()
    if rank != 0:
        return
    starts = (0, o.n // 2 - 128, o.n - 256)
    rows = np.concatenate([np.arange(i0, i0 + 256) for i0 in starts])
    for i0 in starts:
        orc.update_force(lj, o, rows=(i0, i0 + 256))
    rms = np.sqrt((got["force"] ** 2).sum(axis=1).mean())
    scale = np.maximum(np.abs(o.force[rows]).max(axis=1), rms)
    assert np.all(np.abs(got["force"][rows] - o.force[rows]) <= 1e-10 * 100 * scale[:, None]), "c3 rows"
    assert np.all(np.abs(got["potential"][rows] - o.pot[rows]) <= 1e-10 * (np.abs(o.pot[rows]) + 4 * lj.eps))
    orc.update_force(lj, o, mode="cells")
    om = orc.macro(o)
    for key in ("kinetic", "thermal", "potential", "temperature", "pressure"):
        assert abs(m0[key] - om[key]) <= 1e-9 * max(1.0, abs(om[key])), ("c3", key, m0[key], om[key])
    assert np.abs(m1["momentum"]).max() <= 1e-9 * o.n
    assert stats["persistent_loop"] == 1 or not stats["peer_memory"]
    print(f"  c3 10^6 atoms fast rows+macro: ok  owned={got1['owned_per_rank']} T {m0['temperature']:.6f} -> {m1['temperature']:.6f}",
          flush=True)


def main():
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    hot = gas(10, temperature=3000.0)   # fast atoms: plenty of slab crossings in 200 steps
    cases = [
        ("liquid4096 exact NVE", liquid(16), True, 100, None, None),
        ("liquid4096 fast NVT", liquid(16), False, 100, (10.0, 120.0), None),
        ("liquid4096 exact NPT", liquid(16), True, 100, (10.0, 120.0), (1.0, 5.0, 1.01325)),
        ("dense_gas3000 exact NVE", dense_gas(3000), True, 100, None, None),
        ("hot gas1000 exact NVE", hot, True, 200, None, None),
        ("gas1000 fast NPT", gas(10), False, 100, (10.0, 300.0), (1.0, 5.0, 1.01325)),
        ("gas32768 (C2 size, perturbed) fast NVT", perturbed_gas(32, 2), False, 100, (10.0, 300.0), None),
        ("gas32768 (C2 size, perturbed) exact NPT", perturbed_gas(32, 4), True, 60, (10.0, 300.0), (1.0, 5.0, 1.01325)),
    ]
    if dist.get_world_size() > 2:
        # the liquid boxes are too small for more than two slabs (slab width >= 2.1 x halo): keep the gases, add a bigger
        # hot one so atoms cross several slab faces
        cases = [c for c in cases if "gas1000" in c[0] or "gas32768" in c[0]]
        cases.append(("hot gas4096 fast NVT", gas(16, temperature=2000.0), False, 150, (10.0, 300.0), None))
    total_migrated = 0
    for name, o, exact, n_steps, th, ba in cases:
        mig = run_case(name, o, exact, n_steps, th, ba, rank)
        if rank == 0:
            total_migrated += mig
        dist.barrier()
    run_c3_rows(rank)
    dist.barrier()
    if rank == 0:
        assert total_migrated > 0, "no atom ever changed rank: migration path untested"
        print("DIST_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
