set -x
timeout 900 python -m pytest tests/test_gpu_siren.py tests/test_gpu_stepper.py -x -q > gpurun_out/r02_siren_pytest.log 2>&1; tail -3 gpurun_out/r02_siren_pytest.log
for c in taylorgreen smoke3d karman karman3d; do
  for v in 1 0; do
    NMC_OVERLAP_TARGETS=$v timeout 300 python bench_step.py --case $c --iters 1000 --steps 2 > gpurun_out/r02_step2_${c}_ov$v.json 2>&1
    echo "$c overlap=$v $(tail -1 gpurun_out/r02_step2_${c}_ov$v.json | cut -c 1-70)"
  done
done
NMC_SIREN_TC_FWD_MIN=4096 timeout 300 python bench_step.py --case taylorgreen --iters 1000 --steps 2 2>&1 | tail -1 | cut -c 1-70
