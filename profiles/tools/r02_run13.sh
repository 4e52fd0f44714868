set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest3.log 2>&1; tail -4 gpurun_out/r02_pytest3.log
for wl in box_sphere_100000pts_x500walks channel_circle_100000pts_x500walks; do
  timeout 600 python bench.py --workload $wl --steps 3 --warmup 1 --no-sim-steps --no-python-e2e --no-cpu-baseline --no-also 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'], '%.4g walks/s' % d['value'], '%.1f ms' % d['ms_per_step'])"
done
