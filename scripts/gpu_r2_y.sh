#!/bin/bash
# Round 2, call Y: the driver's command once more (bench.py reordered: the parts of a step are timed right after the timed region).
O=gpurun_out; mkdir -p $O
timeout 200 python bench.py --steps 20 --warmup 5 > $O/y_default_driver.json 2> $O/y_default_driver.err; tail -c 300 $O/y_default_driver.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/y_default_driver.json').read().strip().splitlines()[-1])
r=d["roofline"]; e=d["e2e"]; s=d["steady_state"]
print("default: %.3e us/step %.2f" % (d["value"], d["ms_per_step"]*1e3), {k: r.get(k) for k in ("kernel","frac","traffic","traffic_commit","phases_us")})
print("steady:", s)
print("e2e: %.3f ms/step (%d sessions), single %.3f" % (e["ms_per_step"], e["sessions"], e["single_session"]["ms_per_step"]), "launches", d.get("gpu_launches"))
PY
