set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02_parity_pytest.log 2>&1; tail -5 gpurun_out/r02_parity_pytest.log
for wl in smoke3d_1000000pts_x500walks karman_100000pts_x500walks karman3d_1000000pts_x500walks; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-sim-steps --no-python-e2e --no-cpu-baseline --no-also 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'], '%.4g walks/s' % d['value'], d['roofline_issue'].get('frac'), d['roofline_issue'].get('sm_clock_hz'))"
done
