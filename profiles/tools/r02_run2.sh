set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_siren.py -x -q -k "tensor_core_backward" > gpurun_out/r02_tcbwd_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tcbwd_pytest.log
tail -25 gpurun_out/r02_tcbwd_pytest.log
for c in taylorgreen smoke3d karman; do
  NMC_SIREN_TC_BWD=1 timeout 300 python profiles/step_phase_probe.py $c 100 > gpurun_out/r02_phase_${c}_tc.txt 2>&1
  NMC_SIREN_TC_BWD=0 timeout 300 python profiles/step_phase_probe.py $c 100 > gpurun_out/r02_phase_${c}_fp32.txt 2>&1
done
for c in taylorgreen smoke3d karman; do
  for v in 1 0; do
    NMC_SIREN_TC_BWD=$v timeout 300 python bench_step.py --case $c --iters 1000 --steps 2 > gpurun_out/r02_step_${c}_tc$v.json 2>&1
    tail -1 gpurun_out/r02_step_${c}_tc$v.json | cut -c 1-400
  done
done
