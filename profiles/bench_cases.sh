#!/bin/bash
# usage: profiles/bench_cases.sh [case ...]   -- walks/s of the default mode per scene (bench.py, POINTS=100000 points x 500 walks)
cases=${@:-"karman taylorgreen_active smoke3d karman3d"}
for c in $cases; do
  timeout 600 python bench.py --case $c --points ${POINTS:-100000} --steps 10 --warmup 3 --no-cpu-baseline --no-sim-steps --no-python-e2e --no-also 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-20s %.3e walks/s  %.2f ms/step  lane occupancy %.3f' % (d['config']['case'], d['value'], d['ms_per_step'], d['config']['lane_occupancy_of_slice_loop']))"
done
