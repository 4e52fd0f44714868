set -x
P=neural-monte-carlo-fluid-simulation_b200
timeout 600 python -m pytest tests/test_gpu_siren.py -x -q > gpurun_out/r02_siren_pytest.log 2>&1; tail -3 gpurun_out/r02_siren_pytest.log
NMC_LIBNMCFS=$PWD/$P/build/variants/libnmcfs_trace.so timeout 300 python profiles/tools/tc_trace.py 64 5 16384 > gpurun_out/r02_tc_trace4_64.txt 2>&1
for c in taylorgreen smoke3d karman; do
  timeout 300 python profiles/step_phase_probe.py $c 100 > gpurun_out/r02_phase4_${c}.txt 2>&1
  grep -A7 "advect" gpurun_out/r02_phase4_${c}.txt | cut -c 1-120
done
for ch in 64 128 256; do
 for c in smoke3d karman; do
  echo "chunk $ch $c"
  NMC_WGRAD_CHUNK=$ch timeout 300 python profiles/step_phase_probe.py $c 100 2>&1 | grep -m1 "sirenWeightGradTc" | cut -c 1-100
 done
done
