#!/bin/bash
# Round 2, call Z (2 GPUs): the driver's exact N > 1 commands, both arms.
O=gpurun_out; mkdir -p $O
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/z_ref_2.json 2> $O/z_ref_2.err; echo "ref rc=$?"; grep '^{' $O/z_ref_2.json | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > $O/z_ours_2.json 2> $O/z_ours_2.err; echo "ours rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/z_ours_2.json') if l.startswith('{')][-1])
print("N=2: %.3e us/step %.2f" % (d["value"], d["ms_per_step"]*1e3), "scaling", d["scaling"], "roofline frac", d["roofline"].get("frac"))
print("steady:", d["steady_state"] and (d["steady_state"]["value"], d["steady_state"]["us_per_step"], d["steady_state"].get("step_driver")))
print("e2e:", d["e2e"] and (d["e2e"]["value"], d["e2e"]["ms_per_step"]), "launches", d.get("gpu_launches"), "clocks", d.get("clocks"))
PY
tail -c 400 $O/z_ours_2.err
