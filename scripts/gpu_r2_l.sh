#!/bin/bash
# Round 2, call L: full GPU suite, e2e with 1-4 sessions in flight, the driver's two arms.
O=gpurun_out; mkdir -p $O
timeout 2400 python -m pytest tests -x -q -m gpu > $O/l_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/l_pytest.log
for k in 1 2 3 4; do
  timeout 300 python bench.py --steps 20 --warmup 5 --e2e-sessions $k --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/l_e2e_$k.json 2> $O/l_e2e_$k.err
  python -c "
import json,sys
d=json.loads(open('$O/l_e2e_$k.json').read().strip().splitlines()[-1]); e=d['e2e']
print('e2e sessions $k: %.3f ms/step  %.3e atom-steps/s   single %.3f ms' % (e['ms_per_step'], e['value'], e['single_session']['ms_per_step']), ' value %.3e' % d['value'])"
done
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/l_reference.json 2> $O/l_reference.err; cut -c1-400 $O/l_reference.json
timeout 600 python bench.py --steps 20 --warmup 5 > $O/l_default_driver.json 2> $O/l_default_driver.err; cut -c1-3000 $O/l_default_driver.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
