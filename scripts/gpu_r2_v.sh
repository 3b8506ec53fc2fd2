#!/bin/bash
# Round 2, call V: single-precision pre-filter in the tile list builder — pair sets, parity, C5 bench, ncu of the builder.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile or dense or config_size or neighbour or full_size" 2>&1 | tail -3
timeout 300 python bench.py --workload c5 --steps 1000 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 > $O/v_c5.json 2> $O/v_c5.err
python -c "
import json
d=json.loads(open('$O/v_c5.json').read().strip().splitlines()[-1]); r=d['roofline']
print('c5 %.3e us/step %.2f' % (d['value'], d['ms_per_step']*1e3), r['kernels_ms'], 'rebuild', r['rebuild'])"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_build_tile' -s 3 -c 1 -o $O/r02_prof_c5_build_tile_v7 -f \
  python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/v_ncu_build.log 2>&1
python scripts/ncu_summary.py $O/r02_prof_c5_build_tile_v7.ncu-rep > $O/r02_prof_c5_build_tile_v7.txt 2>&1; head -31 $O/r02_prof_c5_build_tile_v7.txt
