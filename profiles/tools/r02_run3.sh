set -x
P=neural-monte-carlo-fluid-simulation_b200
NMC_LIBNMCFS=$PWD/$P/build/variants/libnmcfs_trace.so timeout 300 python profiles/tools/tc_trace.py 64 5 16384 > gpurun_out/r02_tc_trace_64.txt 2>&1
NMC_LIBNMCFS=$PWD/$P/build/variants/libnmcfs_trace.so timeout 300 python profiles/tools/tc_trace.py 128 2 16384 > gpurun_out/r02_tc_trace_128.txt 2>&1
cat gpurun_out/r02_tc_trace_64.txt
