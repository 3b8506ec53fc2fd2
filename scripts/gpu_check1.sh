#!/bin/bash
# First GPU contact: staged so that a hang in one stage does not hide the others.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== stage 1: host-loop tiny" 
timeout 300 python - <<'PY' 2>&1 | tail -20
import numpy as np, moldyn_b200 as md
st = md.State([[0.75,0.75,0.5],[1.25,0.75,0.5]], [[1,1,0],[-1,1,0]], 66.335, [2,2,2])
for host_loop in (True, False):
    with md.Solver(host_loop=host_loop, exact=True) as s:
        s.upload(st, with_forces=False); s.update_force(); s.download(st)
        print("host_loop", host_loop, "f0", st.force[0], s.stats())
        s.step(3, 0.002); s.download(st)
        print(" pos", st.position[0], "vel", st.velocity[0], "f", st.force[0], s.stats())
        st = md.State([[0.75,0.75,0.5],[1.25,0.75,0.5]], [[1,1,0],[-1,1,0]], 66.335, [2,2,2])
PY
echo "== stage 2: smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15
echo "== stage 3: pytest gpu"
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -40
echo "== stage 4: memcheck on small tests"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -k "golden_verlet or liquid_small_box or test_errors" 2>&1 | tail -15
echo "== stage 5: bench c3"
timeout 900 python bench.py --workload c3 --steps 5000 --warmup 300 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 3000 gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
