set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest2.log 2>&1; tail -3 gpurun_out/r02_pytest2.log
timeout 600 python bench_siren.py > gpurun_out/r02_siren_bench.jsonl 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sirenForwardTc|sirenBackwardTc|sirenWeightGradTc" -c 8 -o gpurun_out/r02_siren_tc_c1 -f python profiles/siren_probe.py > gpurun_out/ncu_siren.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_fit_iteration_launches.csv python profiles/fit_iteration_probe.py > gpurun_out/ncu_fit.log 2>&1
timeout 900 python bench.py > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err
cat gpurun_out/r02_bench_n1_b.json | cut -c 1-300
