# final measurement pass of round 2 (one gpurun call, one B200): GPU suite, bench line + reference arm, launch list, ncu captures of
# both walk kernels, step benches, fit-iteration timelines, per-scene rates, the 50-step Taylor-Green run.  Everything lands in
# gpurun_out/r02z_*; what is quoted is copied to profiles/ afterwards.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02z_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02z_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02z_bench_reference_arm.json 2>/dev/null
B="python bench.py --steps 2 --warmup 3 --no-sim-steps --no-python-e2e --no-cpu-baseline --no-also"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches_capture3.csv $B > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fastKernel -s 3 -c 1 -o gpurun_out/r02z_fast3d_c3 -f $B > gpurun_out/ncu_3d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fastKernel -s 3 -c 1 -o gpurun_out/r02z_fast2d_c3 -f $B --workload karman_100000pts_x500walks > gpurun_out/ncu_2d.log 2>&1
for c in taylorgreen karman smoke3d karman3d smoke_obs; do
  timeout 400 python bench_step.py --case $c --iters 1000 --steps 3 2>/dev/null | tail -1 > gpurun_out/r02z_step_${c}_K1000.json
  cut -c 1-80 gpurun_out/r02z_step_${c}_K1000.json
done
for c in taylorgreen karman smoke3d; do
  timeout 400 python bench_step.py --case $c --iters 10000 --steps 1 --cpu-sample 0 2>/dev/null | tail -1 > gpurun_out/r02z_step_${c}_K10000.json
  cut -c 1-80 gpurun_out/r02z_step_${c}_K10000.json
done
for c in taylorgreen karman smoke3d; do
  timeout 200 python profiles/tools/iteration_timeline.py $c advect 60 > gpurun_out/r02z_timeline_${c}_advect.txt 2>&1
done
bash profiles/bench_cases.sh karman taylorgreen_active smoke3d karman3d channel_circle box_sphere > gpurun_out/r02z_bench_cases.txt 2>&1; cat gpurun_out/r02z_bench_cases.txt
timeout 600 python profiles/tools/taylor_green_run.py 50 shipped > gpurun_out/r02z_taylor_green_50steps_shipped.log 2>&1; tail -2 gpurun_out/r02z_taylor_green_50steps_shipped.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; tail -3 gpurun_out/r02z_smoke.log
