"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (the CPU restatement of the reference's
Θ(N²) update_force, the one leg of bench.py that may execute oracle/) prints the driver's JSON line, and the synthetic
inputs are the reference's own initializer layout (position.rs:24-104, velocity.rs:12-28)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_line():
    out = run_bench("--impl", "reference", "--workload", "c1", "--steps", "2", "--warmup", "1")
    line = json.loads(out.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "atom-steps/s" and line["dtype"] == "f64"
    assert line["steps"] == 2 and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_stay_silent():
    # under torchrun only rank 0 runs and prints the reference arm; the other ranks exit 0 without work
    out = run_bench("--impl", "reference", "--workload", "c1", "--gpus", "2", "--steps", "1",
                    env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert out.strip() == ""


def test_synthetic_state_is_the_initializer_layout():
    sys.path.insert(0, ROOT)
    import bench
    from oracle import oracle as orc
    w = bench.WORKLOADS["c1"]
    pos, vel, box = bench.make_state(w)
    ref = orc.argon_lattice(w["side"], w["cell"], w["t_init"], 42)
    assert np.array_equal(pos, ref.pos) and np.array_equal(box, ref.box)   # index = x*s*s + y*s + z, positions (x,y,z)*l
    n = pos.shape[0]
    assert np.array_equal(vel[n // 2:], -vel[:n // 2])                     # antisymmetric halves: sum v == 0 exactly
    sigma_v = np.sqrt(bench.K_B * (w["t_init"] * 0.01) / bench.ARGON_MASS)
    assert abs(vel[:n // 2].std() / sigma_v - 1.0) < 0.05
    # every BASELINE.json config is a named workload, and the default is C3 (the size the metric is quoted on)
    assert {"c1", "c2", "c3", "c4", "c5"} <= set(bench.WORKLOADS)
    assert bench.WORKLOADS["c3"]["side"] ** 3 == 1_000_000
