set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
B="python bench.py --steps 2 --warmup 3 --no-sim-steps --no-python-e2e --no-cpu-baseline --no-also"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_capture2.csv $B > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fastKernel -s 3 -c 1 -o gpurun_out/r02_fast3d_c2 -f $B > gpurun_out/ncu_3d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fastKernel -s 3 -c 1 -o gpurun_out/r02_fast2d_c2 -f $B --workload karman_100000pts_x500walks > gpurun_out/ncu_2d.log 2>&1
timeout 600 python bench_siren.py > gpurun_out/r02_siren_bench.jsonl 2> gpurun_out/r02_siren_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sirenForwardTc|sirenBackwardTc|sirenWeightGradTc" -c 8 -o gpurun_out/r02_siren_tc_c2 -f python profiles/siren_probe.py > gpurun_out/ncu_siren.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_fit_iteration_launches.csv python profiles/fit_iteration_probe.py > gpurun_out/ncu_fit.log 2>&1
for c in taylorgreen karman smoke3d karman3d smoke_obs; do
  timeout 400 python bench_step.py --case $c --iters 1000 --steps 2 2>/dev/null | tail -1 > gpurun_out/r02_step_${c}_K1000.json
  cut -c 1-80 gpurun_out/r02_step_${c}_K1000.json
  timeout 300 python profiles/step_phase_probe.py $c 100 > gpurun_out/r02_phase_${c}.txt 2>&1
done
timeout 400 python bench_step.py --case taylorgreen --iters 10000 --steps 1 2>/dev/null | tail -1 > gpurun_out/r02_step_taylorgreen_K10000.json
timeout 400 python bench_step.py --case karman --iters 10000 --steps 1 2>/dev/null | tail -1 > gpurun_out/r02_step_karman_K10000.json
bash profiles/bench_cases.sh karman taylorgreen_active smoke3d karman3d channel_circle box_sphere > gpurun_out/r02_bench_cases.txt 2>&1; cat gpurun_out/r02_bench_cases.txt
