set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2>&1
B="python bench.py --steps 2 --warmup 3 --no-sim-steps --no-python-e2e --no-cpu-baseline --no-also"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_capture1.csv $B > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fastKernel -s 3 -c 1 -o gpurun_out/r02_fast3d_c1 -f $B > gpurun_out/ncu_3d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fastKernel -s 3 -c 1 -o gpurun_out/r02_fast2d_c1 -f $B --workload karman_100000pts_x500walks > gpurun_out/ncu_2d.log 2>&1
tail -5 gpurun_out/r02_pytest.log
cat gpurun_out/r02_bench_n1.json
