set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_scale.py -x -q -k "channel_circle or box_sphere or random_mesh or big or bvh or large" 2>&1 | tail -5
for m in flat tree; do
 for wl in channel_circle_100000pts_x500walks box_sphere_100000pts_x500walks; do
  NMC_BIG_MESH=$m timeout 600 python bench.py --workload $wl --steps 3 --warmup 1 --no-sim-steps --no-python-e2e --no-cpu-baseline --no-also 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m', d['config']['workload'], '%.4g walks/s' % d['value'], '%.1f ms' % d['ms_per_step'])"
 done
done
