#!/bin/bash
# Round 2, call X: the last tree (builder pre-filter, loop/chunk rule for mid-size populated lists) — full GPU suite, the
# driver's command, smoke().
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu > $O/x_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/x_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/x_default_driver.json 2> $O/x_default_driver.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/x_default_driver.json').read().strip().splitlines()[-1])
r=d["roofline"]; e=d["e2e"]; s=d["steady_state"]
print("default: %.3e us/step %.2f frac %.3f traffic %s" % (d["value"], d["ms_per_step"]*1e3, r["frac"], r.get("traffic")))
print("steady:", s and ("%.3e" % s["value"], round(s["us_per_step"],2), s["rebuilds"], s.get("nbr_mean")))
print("e2e: %.3f ms/step (%d sessions), single %.3f" % (e["ms_per_step"], e["sessions"], e["single_session"]["ms_per_step"]))
print("launches", d.get("gpu_launches"), "clocks", d.get("clocks"))
PY
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 | cut -c1-200
