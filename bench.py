#!/usr/bin/env python
"""bench.py — atom-steps/s of the moldyn `solve` step loop on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl ours|reference]

A "step" is one Integrator::calculate (velocity-Verlet + LJ forces + Berendsen thermostat) over the whole
system.  `value` = atoms × K / device time of K consecutive steps with the state resident in HBM (CUDA events on
the library's stream, list rebuilds included).  `e2e` = the same metric through the reference-facing per-call
C-ABI (md_calculate_host: State in host memory in, State out, every step).  `--impl reference` times the CPU
restatement of the reference's own Θ(N²) algorithm on the host cores (the reference is Rust; no toolchain here).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ARGON_MASS = 66.335
K_B = 1.380648528
GAS_CELL = 3.338339
LIQUID_CELL = 0.36165
DT = 0.002
ALGO_BYTES_STEP = 160       # r+w of x, v, F (144 B) + write U, W (16 B) per atom-step
# What the kernels' steady-state contracts move per atom, used for the per-kernel figures:
CONTRACT_BYTES = {
    "k_kick_drift": 72,     # read x, u (48); write x (24)
    "k_force": 80,          # read x, u (48) + list count and first row (8); write u' (24)
    "k_md_loop": 104,       # drift phase: read x, u (48) + list count (8), write x (24) and u' of the atoms without partners
                            # (24); the force phase adds ~(48 + 24 + 28 per partner) B for every atom WITH partners
}
ALGO_FLOP_PAIR = 42         # flop per directed in-range pair
ALGO_FLOP_ATOM = 30
FP64_FALLBACK_TFLOPS = 40.0  # B200 datasheet FP64 (used only if the in-run DFMA measurement fails)

WORKLOADS = {
    # name: side, lattice cell, T_init, (tau, T0), (beta, tau, P0) or None, (r_cut, u_cut) or None
    "c1": dict(desc="argon 10x10x10 (1000 atoms) NPT Berendsen", side=10, cell=GAS_CELL, t_init=273.15,
               thermostat=(10.0, 300.0), barostat=(1.0, 5.0, 1.01325), cut=None),
    "c2": dict(desc="argon 32x32x32 (32768 atoms) NVT Berendsen", side=32, cell=GAS_CELL, t_init=273.15,
               thermostat=(10.0, 300.0), barostat=None, cut=None),
    "c3": dict(desc="argon 100x100x100 (1M atoms) NVT Berendsen", side=100, cell=GAS_CELL, t_init=273.15,
               thermostat=(10.0, 300.0), barostat=None, cut=None),
    "c4": dict(desc="argon 216^3 (10.08M atoms) NPT Berendsen", side=216, cell=GAS_CELL, t_init=273.15,
               thermostat=(10.0, 300.0), barostat=(1.0, 5.0, 1.01325), cut=None),
    "c5": dict(desc="liquid argon 64^3 (262144 atoms) NVT, r_cut 3.5 sigma", side=64, cell=LIQUID_CELL, t_init=120.0,
               thermostat=(10.0, 120.0), barostat=None, cut=(1.1963, -0.003723224030513348)),
    "big": dict(desc="argon 200^3 (8M atoms) NVT Berendsen (state > L2)", side=200, cell=GAS_CELL, t_init=273.15,
                thermostat=(10.0, 300.0), barostat=None, cut=None),
}


def make_state(w, seed=42):
    """`moldyn-cli initialize -t u -s side side side -l cell -T t_init` with a seeded RNG (synthetic input):
    index = x*s*s + y*s + z, velocities N(0, K_B*T/100/m) with the second half the negated first half."""
    s = w["side"]
    g = np.arange(s, dtype=np.float64) * w["cell"]
    pos = np.empty((s, s, s, 3))
    pos[..., 0] = g[:, None, None]
    pos[..., 1] = g[None, :, None]
    pos[..., 2] = g[None, None, :]
    pos = pos.reshape(-1, 3)
    n = pos.shape[0]
    sigma_v = np.sqrt(K_B * (w["t_init"] * 0.01) / ARGON_MASS)
    half = np.random.default_rng(seed).standard_normal((n // 2, 3)) * sigma_v
    vel = np.concatenate([half, -half])
    if w["cell"] < 1.0:  # liquid: melt the perfect lattice faster
        pos = pos + np.random.default_rng(seed + 1).uniform(-0.03, 0.03, pos.shape)
        pos %= (w["cell"] * s)
    box = np.array([w["cell"] * s] * 3)
    return pos, vel, box


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index=0, period=0.02):
        super().__init__(daemon=True)
        self.period, self.index = period, index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` on `workload`, from profiles/ncu_traffic.json —
    written by scripts/ncu_traffic.py from an `ncu --set full` capture together with the commit it was taken at.  None when no
    capture of this kernel/workload is recorded: the line then prints traffic: null rather than a number from another build."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            rec = json.load(f).get(f"{workload}:{kernel}")
    except (OSError, ValueError):
        rec = None
    return rec or {}


def in_range_pairs(pos, box, r_cut, sample=4096, seed=0):
    """Mean number of partners within r_cut per atom (directed pairs / atom), from a random sample of atoms (periodic KD-tree)."""
    from scipy.spatial import cKDTree
    pos = np.mod(pos, box)
    pos = np.minimum(pos, np.nextafter(box, 0.0))
    tree = cKDTree(pos, boxsize=box)
    idx = np.random.default_rng(seed).choice(len(pos), size=min(sample, len(pos)), replace=False)
    cnt = tree.query_ball_point(pos[idx], r_cut, return_length=True)
    return float(np.mean(cnt) - 1.0)


def cpu_reference_rate(w, sample_rows, threads=None):
    """The reference's Θ(N²) update_force (potential.rs:158-216) on the host cores, rows [0, sample_rows) of the
    workload's own positions against ALL N partners → atoms/s of force evaluation ≈ atom-steps/s of the CPU
    solver (the Θ(N) parts of the step are negligible at these N)."""
    from oracle import oracle as orc
    # all the host threads the box offers — torchrun exports OMP_NUM_THREADS=1 to every rank, which is not what a CPU
    # baseline should be measured with
    orc.set_num_threads(threads or len(os.sched_getaffinity(0)))
    pos, vel, box = make_state(w)
    st = orc.State(pos, vel, ARGON_MASS, box)
    lj = orc.LennardJones() if w["cut"] is None else orc.LennardJones(r_cut=w["cut"][0], u_cut=w["cut"][1])
    rows = min(sample_rows, st.n)
    orc.update_force(lj, st, rows=(0, min(64, rows)))  # warm the threads
    t0 = time.perf_counter()
    orc.update_force(lj, st, rows=(0, rows))
    dt = time.perf_counter() - t0
    return rows / dt, dt, rows, st.n, orc.num_threads()


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_rows or max(256, int(2.0e9 / (w["side"] ** 3)))  # ≈ 2e9 pair tests per step
    # warm-up + K bounded steps
    for _ in range(max(args.warmup, 0) and 1):
        cpu_reference_rate(w, 64)
    # K bounded steps (each a sample of `sample` rows against all N partners); a wall-clock cap keeps a large K from
    # running for hours — `steps` in the line is what was actually timed
    rates, times = [], []
    t_begin = time.perf_counter()
    for _ in range(max(1, args.steps)):
        r, dt, rows, n, thr = cpu_reference_rate(w, sample)
        rates.append(r)
        times.append(dt)
        if time.perf_counter() - t_begin > 90.0:
            break
    value = float(np.mean(rates))
    out = {
        "impl": "reference", "metric": "atom-steps/s", "value": value, "unit": "atom-steps/s", "n_gpus": args.gpus,
        "steps": len(rates), "warmup": 1, "ms_per_step": float(np.mean(times)) * 1e3 * (n / rows),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "atoms": n, "dt": DT},
        "cpu_baseline": {"value": value, "unit": "atom-steps/s", "cores": thr, "kind": "port",
                         "sample": f"update_force rows [0,{rows}) of {n} per step, each against all {n} partners "
                                   f"(reference's Θ(N²) scan, potential.rs:158-216); ms_per_step extrapolated to N rows"},
        "e2e": {"value": value, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ns_per_day": value / n * 0.1728,
    }
    print(json.dumps(out))


def timed_steps(s, stream, steps, th, ba, dist=None):
    """K consecutive steps between two CUDA events on the library's stream, barrier + synchronize on both sides; ms (max over
    ranks).  The call between the events is the C ABI's md_step itself, its argument structs built beforehand: what is timed is
    the device's K steps (launch gaps of the library included), not the Python wrapper's bookkeeping around the call."""
    import ctypes as C

    import torch

    from moldyn_b200 import _ffi
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tc = th[0]._c(th[1]) if th else None
    bc = ba[0]._c(ba[1]) if ba else None
    md_step = _ffi.lib().md_step
    argv = (s._ctx, int(steps), float(DT), C.byref(tc) if tc else None, C.byref(bc) if bc else None)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
        torch.cuda.synchronize()
    e0.record(stream)
    rc = md_step(*argv)
    e1.record(stream)
    _ffi.check(s._ctx, rc)
    s.synchronize()
    torch.cuda.synchronize()
    if th:
        th[0].lambda_, th[0].psi = tc.lambda_, tc.psi
    if ba:
        ba[0].myu = bc.myu
    ms = e0.elapsed_time(e1)
    if dist:
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def run_ours(args, w):
    import torch

    import moldyn_b200 as md
    from moldyn_b200 import _ffi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pos, vel, box = make_state(w)
    n = pos.shape[0]
    th = lambda: (md.Thermostat.Berendsen(w["thermostat"][0]), w["thermostat"][1])  # noqa: E731
    ba = (lambda: (md.Barostat.Berendsen(w["barostat"][0], w["barostat"][1]), w["barostat"][2])) if w["barostat"] \
        else (lambda: None)
    r_cut = w["cut"][0] if w["cut"] else 0.8545

    s = md.Solver(device=local, skin=args.skin, cell_atoms=args.cell_atoms, cell_subdiv=args.cell_subdiv,
                  chunk_loop=args.loop == "chunk", host_loop=args.loop == "host")
    if world > 1:
        from moldyn_b200 import distributed as mdd
        mdd.init_solver_comm(s)
    if w["cut"]:
        s.set_potential(md.Potential(0.3418, 1.712, *w["cut"]))
    s.upload_arrays(pos, vel, ARGON_MASS, box)
    s.update_force()
    stream = torch.cuda.ExternalStream(s.stream(), device=local)
    t_th, t_ba = th(), ba()
    s.step(args.warmup, DT, thermostat=t_th, barostat=t_ba)
    st0 = s.stats()

    # ---- the timed region: K consecutive steps, state resident in HBM -----------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed_steps(s, stream, args.steps, t_th, t_ba, dist)
    clocks = sampler.stop()
    st1 = s.stats()
    value = n * args.steps / (ms * 1e-3)
    macro = s.macro()
    dense = st1["nbr_mean"] >= 8.0
    rebuilds = st1["rebuilds"] - st0["rebuilds"]

    # ---- where the time goes: device timing of the parts of a step (md_time_kernels), taken right after the timed region —
    # the same list state and the same step driver as the steps that were timed ----------------------------------------
    kt = s.time_kernels(min(max(args.steps, 50), 400), DT, thermostat=t_th, barostat=t_ba) if world == 1 else None
    st2 = s.stats()
    per = {k: (kt[k][0] / kt[k][1] if kt[k][1] else None) for k in kt} if kt else {}

    # ---- steady state: the driver's short runs start on the perfect gas lattice (no partner in range, no rebuild for the
    # first ~10^3 steps), so the same metric is taken again after the system has been run into its collisional steady
    # state, over a region long enough to contain list rebuilds ------------------------------------------------------
    steady = None
    if args.steady_steps > 0:
        s.step(args.steady_warmup, DT, thermostat=t_th, barostat=t_ba)
        sa = s.stats()
        ms_s = timed_steps(s, stream, args.steady_steps, t_th, t_ba, dist)
        sb = s.stats()
        steady = {"value": n * args.steady_steps / (ms_s * 1e-3), "unit": "atom-steps/s", "steps": args.steady_steps,
                  "steps_before": args.warmup + args.steps + args.steady_warmup, "us_per_step": ms_s / args.steady_steps * 1e3,
                  "rebuilds": sb["rebuilds"] - sa["rebuilds"], "nbr_mean": sb["nbr_mean"], "nbr_max": sb["nbr_max"],
                  "step_driver": "persistent loop (k_md_loop)" if sb["persistent_loop"] else
                                 ("tile kernels in graph chunks" if sb["tile_lists"] else "two-kernel step in graph chunks"),
                  "frac_of_step_roofline": ALGO_BYTES_STEP * n * args.steady_steps / (ms_s * 1e-3) / 1e9 / (peaks()[0] * world)}

    hbm, peak_src = peaks()
    fp64_peak, fp64_src = None, None
    rebuild_ms = per.get("rebuild")
    if world == 1 and rebuild_ms is None and args.time_rebuild:
        # no rebuild fell into the timed steps: force one and time it, so the line always carries its cost
        s.invalidate_lists()
        kr = s.time_kernels(1, DT, thermostat=t_th, barostat=t_ba)
        rebuild_ms = kr["rebuild"][0] / kr["rebuild"][1] if kr["rebuild"][1] else None
    step_us = ms / args.steps * 1e3
    loop = bool(st1["persistent_loop"]) and (world > 1 or (per.get("loop_barrier") is not None))
    if world == 1 and loop:
        phases = {"drift_phase_us": per["kick_drift"] * 1e3, "mid_step_barrier_us": per["loop_barrier"] * 1e3,
                  "force_phase_and_tail_us": per["force"] * 1e3}
        tr = ncu_traffic(args.workload, "k_md_loop")
        roofline = {
            "bound": "hbm", "kernel": "k_md_loop (one launch runs many steps; figures are per step)",
            "achieved": ALGO_BYTES_STEP * n / (step_us * 1e-6) / 1e9, "peak": hbm, "unit": "GB/s", "peak_source": peak_src,
            "algorithmic_bytes_per_atom": ALGO_BYTES_STEP, "algorithmic_bytes_per_launch": ALGO_BYTES_STEP * n,
            "avg_launch_ms": step_us * 1e-3, "contract_bytes_per_atom": CONTRACT_BYTES["k_md_loop"],
            "phases_us": phases, "traffic": tr.get("dram_bytes_per_launch"), "traffic_source": tr.get("source"),
            "traffic_commit": tr.get("commit"),
        }
        roofline["frac"] = roofline["achieved"] / hbm
    elif world == 1:
        f_ms, k_ms = per["force"], per["kick_drift"]
        roofline = {"kernel": "k_force", "avg_launch_ms": f_ms, "kernels_ms": {"k_force": f_ms, "k_kick_drift": k_ms},
                    "launches_timed": {k: kt[k][1] for k in kt}}
        tr = ncu_traffic(args.workload, "k_force")
        roofline.update({"traffic": tr.get("dram_bytes_per_launch"), "traffic_source": tr.get("source"),
                         "traffic_commit": tr.get("commit")})
        if dense:
            # FP64-bound: 42 flop per directed in-range pair + 30 per atom-step
            try:
                fp64_peak, fp64_src = s.measure_fp64_peak(), "measured in this run (register-only DFMA loop, md_measure_fp64_peak)"
            except Exception as e:  # noqa: BLE001
                fp64_peak, fp64_src = FP64_FALLBACK_TFLOPS, f"fallback datasheet figure ({e!r})"
            cur = np.empty(3 * n)
            s.download_arrays(pos=cur)
            pairs = in_range_pairs(cur.reshape(-1, 3), np.array(macro["box"]), r_cut)
            flop = (ALGO_FLOP_PAIR * pairs + ALGO_FLOP_ATOM) * n
            roofline.update({"bound": "fp64", "achieved": flop / (f_ms * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                             "peak_source": fp64_src, "pairs_in_range": pairs, "pairs_listed": st1["nbr_mean"],
                             "algorithmic_flop_per_launch": flop,
                             "hbm_view": {"contract_bytes_per_atom": CONTRACT_BYTES["k_force"] + 4 * st1["nbr_mean"],
                                          "achieved_gbs": (CONTRACT_BYTES["k_force"] + 4 * st1["nbr_mean"]) * n / (f_ms * 1e-3) / 1e9,
                                          "peak_gbs": hbm}})
            roofline["frac"] = roofline["achieved"] / fp64_peak
        else:
            dom, dom_ms = ("k_force", f_ms) if f_ms >= k_ms else ("k_kick_drift", k_ms)
            roofline.update({"bound": "hbm", "kernel": dom, "avg_launch_ms": dom_ms,
                             "achieved": CONTRACT_BYTES[dom] * n / (dom_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                             "peak_source": peak_src, "algorithmic_bytes_per_atom": CONTRACT_BYTES[dom],
                             "algorithmic_bytes_per_launch": CONTRACT_BYTES[dom] * n})
            roofline["frac"] = roofline["achieved"] / hbm
    else:
        roofline = {"bound": "hbm", "kernel": "k_md_loop per rank" if loop else "k_kick_drift + k_force per rank",
                    "achieved": ALGO_BYTES_STEP * value / 1e9, "peak": hbm * world, "unit": "GB/s", "peak_source": peak_src,
                    "frac": ALGO_BYTES_STEP * value / 1e9 / (hbm * world), "traffic": None,
                    "algorithmic_bytes_per_atom": ALGO_BYTES_STEP,
                    "note": "whole job against N x the single-GPU HBM peak; per-rank phase clocks are in per_rank"}
    roofline["step"] = {"algorithmic_bytes_per_atom_step": ALGO_BYTES_STEP, "achieved": ALGO_BYTES_STEP * value / 1e9,
                        "frac": ALGO_BYTES_STEP * value / 1e9 / (hbm * world), "us_per_step": step_us,
                        "state_mb": 48 * n / 1e6,
                        "note": "x, v planes the step streams vs the 126 MB L2: below ~2.6e6 atoms part of the traffic is "
                                "served from L2, so the step can exceed the HBM-only bound; `big` and c4 give the HBM view"}
    roofline["rebuild"] = {"ms_each": rebuild_ms, "in_timed_region": rebuilds,
                           "amortised_us_per_step": (rebuild_ms * 1e3 * steady["rebuilds"] / steady["steps"])
                           if (steady and rebuild_ms is not None) else None}

    # --- e2e: reference-facing per-call API, State in pinned host memory in and out every step ------------
    e2e = None
    if args.e2e_steps > 0 and world == 1:
        # One session: H2D 72 MB -> step -> D2H 88 MB, strictly in this order (the step needs the whole State, the results
        # exist only after it): PCIe carries one direction at a time.  The link is full duplex, so a host that has more than
        # one State to advance (an ensemble; or the next frame's State while the previous one is written out) drives two
        # sessions from two threads: one session's download overlaps the other's upload.  `value` is that; the single-session
        # number is beside it.
        import threading
        L = _ffi.lib()
        n_sessions = max(1, args.e2e_sessions)
        sessions = [s]
        for _ in range(n_sessions - 1):
            s2 = md.Solver(device=local, skin=args.skin, cell_atoms=args.cell_atoms, cell_subdiv=args.cell_subdiv,
                           chunk_loop=args.loop == "chunk", host_loop=args.loop == "host")
            if w["cut"]:
                s2.set_potential(md.Potential(0.3418, 1.712, *w["cut"]))
            sessions.append(s2)
        hboxes = [np.array(box) for _ in sessions]
        host = []
        for _ in sessions:
            hp = [torch.empty(sz, dtype=torch.float64, pin_memory=True) for sz in (3 * n, 3 * n, 3 * n, n, n)]
            s.download_arrays(*(t.data_ptr() for t in hp))
            host.append(hp)
        tcs = [t_th[0]._c(t_th[1]) for _ in sessions]
        bcs = [(t_ba[0]._c(t_ba[1]) if t_ba else None) for _ in sessions]

        def call(k):
            ss, hp, tc, bc, hbox = sessions[k], host[k], tcs[k], bcs[k], hboxes[k]
            rc = L.md_calculate_host(ss._ctx, n, hp[0].data_ptr(), hp[1].data_ptr(), hp[2].data_ptr(),
                                     hp[3].data_ptr(), hp[4].data_ptr(), ARGON_MASS,
                                     hbox.ctypes.data_as(C.c_void_p), DT, C.byref(tc), C.byref(bc) if bc else None)
            _ffi.check(ss._ctx, rc)

        def timed(n_thr, steps_each):
            errs = []
            gate = threading.Barrier(n_thr + 1)

            def work(k):
                try:
                    gate.wait()
                    for _ in range(steps_each):
                        call(k)
                except Exception as e:  # noqa: BLE001
                    errs.append(e)
            thr = [threading.Thread(target=work, args=(k,)) for k in range(n_thr)]
            for t in thr:
                t.start()
            torch.cuda.synchronize()
            gate.wait()
            t0 = time.perf_counter()
            for t in thr:
                t.join()
            torch.cuda.synchronize()
            dt_ = time.perf_counter() - t0
            if errs:
                raise errs[0]
            return dt_

        for k in range(n_sessions):
            for _ in range(3):
                call(k)
        t_one = timed(1, args.e2e_steps)
        t_e2e = timed(n_sessions, args.e2e_steps) if n_sessions > 1 else t_one
        total = n_sessions * args.e2e_steps
        e2e = {"value": n * total / t_e2e, "unit": "atom-steps/s",
               "h2d_bytes_per_step": (80 if t_ba else 72) * n + 24, "d2h_bytes_per_step": 88 * n + 24,
               "ms_per_step": t_e2e / total * 1e3, "steps": total, "sessions": n_sessions,
               "single_session": {"value": n * args.e2e_steps / t_one, "ms_per_step": t_one / args.e2e_steps * 1e3,
                                  "steps": args.e2e_steps},
               "api": "md_calculate_host (≡ Integrator::calculate on a host State: upload x,v,F → 1 step → "
                      "download x,v,F,U,W; the incoming U — and W without a barostat — are dead and not uploaded), pinned host "
                      f"buffers; {n_sessions} independent sessions (States) driven from {n_sessions} host threads so one's "
                      "download overlaps the other's upload on the duplex link; single_session = one State, strictly serial"}
    elif args.e2e_steps > 0:
        # decomposed form of the same call: every rank uploads the (pinned) host State, one collective step, every
        # rank downloads its own slab
        L = _ffi.lib()
        hp = [torch.empty(sz, dtype=torch.float64, pin_memory=True) for sz in (3 * n, 3 * n)]
        hp[0].copy_(torch.from_numpy(pos.reshape(-1)))
        hp[1].copy_(torch.from_numpy(vel.reshape(-1)))
        cap = int(1.6 * n / world) + 8192
        out = [torch.empty(sz, dtype=torch.float64, pin_memory=True) for sz in (3 * cap, 3 * cap, 3 * cap, cap, cap)]
        ids = torch.empty(cap, dtype=torch.int64, pin_memory=True)
        hbox = np.array(box)

        def call():
            s.upload_arrays(hp[0].data_ptr(), hp[1].data_ptr(), ARGON_MASS, box, n=n)
            s.update_force()
            s.step(1, DT, thermostat=t_th, barostat=t_ba)
            _ffi.check(s._ctx, L.md_download_local(s._ctx, ids.data_ptr(), out[0].data_ptr(), out[1].data_ptr(),
                                                   out[2].data_ptr(), out[3].data_ptr(), out[4].data_ptr(),
                                                   hbox.ctypes.data_as(C.c_void_p)))
        call()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            call()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
        n_loc = s.local_count()[0]
        e2e = {"value": n * args.e2e_steps / t_e2e, "unit": "atom-steps/s",
               "h2d_bytes_per_step": 48 * n * world, "d2h_bytes_per_step": 96 * n_loc * world,
               "ms_per_step": t_e2e / args.e2e_steps * 1e3, "steps": args.e2e_steps,
               "api": "per step on every rank: md_upload_state(full host State) → md_update_force → md_step(1) → "
                      "md_download_local(own slab), pinned host buffers"}

    cpu = None
    if args.cpu_rows >= 0 and rank == 0:
        rows = args.cpu_rows or max(256, int(6.0e9 / n))
        r, dt_cpu, rows, _, thr = cpu_reference_rate(w, rows)
        cpu = {"value": r, "unit": "atom-steps/s", "cores": thr, "kind": "port",
               "sample": f"oracle update_force rows [0,{rows}) of {n}, each against all {n} partners "
                         f"({dt_cpu:.1f} s; reference's Θ(N²) scan, potential.rs:158-216); the oracle is a restatement "
                         f"pinned to the reference's golden values, many-body behaviour self-pinned (no Rust toolchain)"}

    if world == 1:
        par = "single GPU"
    elif st1["peer_memory"]:
        par = (f"{world} x-slabs (spatial decomposition); per step: ghost positions stored into the neighbours' HBM by the "
               f"{'persistent step loop' if loop else 'drift kernel'}, rank sums exchanged through peer-memory mailboxes inside the "
               f"{'loop' if loop else 'force kernel'}")
    else:
        par = f"{world} x-slabs (spatial decomposition), NCCL halo send/recv + all-gather of 12 sums per step"
    out = {
        "metric": "atom-steps/s", "value": value, "unit": "atom-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "parallelism": par, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "atoms": n, "dt": DT, "thermostat": w["thermostat"],
                   "barostat": w["barostat"], "r_cut": r_cut, "skin": st1["skin"], "cells": st1["cells"],
                   "loop": "persistent cooperative kernel (k_md_loop)" if loop else "graph chunks of {k_kick_drift; k_force}",
                   "list_state": {"nbr_mean": st1["nbr_mean"], "nbr_max": st1["nbr_max"], "rebuilds_in_timed_region": rebuilds,
                                  "note": "a run started on the perfect gas lattice has no partner within the list radius "
                                          "for its first ~10^3 steps; see steady_state for the collisional regime"},
                   "l2": "one timed region of K consecutive, dependent MD steps of one trajectory (no input is "
                         "re-run, so there is no L2 flush between steps); per-step state "
                         f"{48 * n / 1e6:.0f} MB (x, v) vs 126 MB L2 — see roofline.step"},
        "ns_per_day": args.steps / (ms * 1e-3) * 0.1728,
        "roofline": roofline, "steady_state": steady, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
        "gpu_launches": st1["kernel_launches"] - st0["kernel_launches"],
        "rebuilds_in_timed_region": rebuilds,
        "graph_launches_in_timed_region": st1["graph_launches"] - st0["graph_launches"],
        "loop_launches_in_timed_region": st1["loop_launches"] - st0["loop_launches"],
        "state_check": {"temperature": macro["temperature"], "pressure": macro["pressure"],
                        "momentum_abs_max": float(np.abs(macro["momentum"]).max())},
    }
    if world > 1:
        alls = [None] * world if rank == 0 else None
        mine = {k: st1[k] for k in ("n_owned", "n_ghost", "migrated", "rebuilds", "peer_memory", "persistent_loop")}
        mine["wait_halo_us_per_step"] = (st1["wait_halo_ms"] - st0["wait_halo_ms"]) * 1e3 / args.steps
        mine["wait_sums_us_per_step"] = (st1["wait_sums_ms"] - st0["wait_sums_ms"]) * 1e3 / args.steps
        for k in ("force_atoms", "force_tail"):
            mine[k + "_us_per_step"] = (st1[k + "_ms"] - st0[k + "_ms"]) * 1e3 / args.steps
        mine["rebuild_ms_each"] = (st1["rebuild_ms"] - st0["rebuild_ms"]) / max(st1["rebuilds"] - st0["rebuilds"], 1)
        names = ("drift_phase", "mid_step_barrier", "force_phase", "tail_reduce_exchange_finalize")
        mine["loop_us_per_step"] = {k: (a - b) * 1e3 / args.steps
                                    for k, a, b in zip(names, st1["loop_phase_ms"], st0["loop_phase_ms"])}
        dist.gather_object(mine, alls, dst=0)
        out["per_rank"] = alls
    if rank == 0:
        print(json.dumps(out))
    s.close()
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--skin", type=float, default=0.0)
    ap.add_argument("--cell-atoms", type=float, default=0.0)
    ap.add_argument("--cell-subdiv", type=int, default=0)
    ap.add_argument("--loop", default="auto", choices=["auto", "chunk", "host"],
                    help="auto: persistent step loop for dilute systems, graph chunks for dense ones; chunk: the two-kernel "
                         "graph-chunk loop everywhere (A/B); host: one launch per step (ncu)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--e2e-sessions", type=int, default=3,
                    help="independent States advanced concurrently through md_calculate_host in the e2e leg (1 GPU)")
    ap.add_argument("--steady-steps", type=int, default=None,
                    help="steps of the extra steady-state region (0 = skip; default 4000, 1000 above 2e6 atoms, 0 for N > 1 "
                         "unless given)")
    ap.add_argument("--steady-warmup", type=int, default=None, help="steps run before the steady-state region")
    ap.add_argument("--no-time-rebuild", dest="time_rebuild", action="store_false",
                    help="do not force + time one list rebuild when none fell into the timed steps")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the CPU sample (0 = auto, -1 = skip)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 3
        args.warmup = args.warmup if args.warmup is not None else 1
        run_reference(args, w)
        return
    n = w["side"] ** 3
    if args.steps is None:
        args.steps = 20000 if n <= 2_000_000 else 2000
    if args.warmup is None:
        args.warmup = 500 if n <= 2_000_000 else 100
    args.warmup = max(args.warmup, 3)
    if args.steady_steps is None:
        args.steady_steps = (4000 if n <= 2_000_000 else 1000) if args.gpus == 1 else 0
    if args.steady_warmup is None:
        # a gas atom needs ~5 ps (2500 steps) to meet its first partner; the liquid is in its steady state at once
        args.steady_warmup = 0 if w["cell"] < 1.0 else (6000 if n <= 2_000_000 else 3000)
    run_ours(args, w)


if __name__ == "__main__":
    main()
