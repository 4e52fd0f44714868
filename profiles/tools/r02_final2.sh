set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02f_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_reference_arm.json 2>/dev/null
timeout 600 python bench_siren.py > gpurun_out/r02f_siren_bench.jsonl 2> gpurun_out/r02f_siren_bench.err
timeout 300 python profiles/tools/time_fused_bwd.py > gpurun_out/r02f_fused_bwd_times.jsonl 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sirenBackwardFusedTc" -s 1 -c 1 -o gpurun_out/r02f_fused_bwd -f python profiles/siren_probe.py > gpurun_out/ncu_fused.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_fit_iteration_launches.csv python profiles/fit_iteration_probe_graph.py > gpurun_out/ncu_fit.log 2>&1
for c in taylorgreen karman smoke3d karman3d smoke_obs; do
  timeout 400 python bench_step.py --case $c --iters 1000 --steps 3 2>/dev/null | tail -1 > gpurun_out/r02f_step_${c}_K1000.json
  cut -c 1-80 gpurun_out/r02f_step_${c}_K1000.json
done
for c in taylorgreen karman smoke3d; do
  timeout 400 python bench_step.py --case $c --iters 10000 --steps 1 --cpu-sample 0 2>/dev/null | tail -1 > gpurun_out/r02f_step_${c}_K10000.json
  cut -c 1-80 gpurun_out/r02f_step_${c}_K10000.json
done
for c in taylorgreen karman smoke3d; do
  timeout 200 python profiles/tools/iteration_timeline.py $c advect 60 > gpurun_out/r02f_timeline_${c}_advect.txt 2>&1
  timeout 200 python profiles/tools/iteration_timeline.py $c project 60 > gpurun_out/r02f_timeline_${c}_project.txt 2>&1
done
timeout 600 python profiles/tools/taylor_green_run.py 50 shipped > gpurun_out/r02f_taylor_green_50steps_shipped.log 2>&1; tail -2 gpurun_out/r02f_taylor_green_50steps_shipped.log
