set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "edge_cases or full_size" 2>&1 | tail -3
timeout 1200 python profiles/tools/taylor_green_run.py 50 shipped > gpurun_out/r02_taylor_green_50steps_shipped.log 2>&1; tail -4 gpurun_out/r02_taylor_green_50steps_shipped.log
timeout 1200 python profiles/tools/taylor_green_run.py 50 active > gpurun_out/r02_taylor_green_50steps_active.log 2>&1; tail -4 gpurun_out/r02_taylor_green_50steps_active.log
