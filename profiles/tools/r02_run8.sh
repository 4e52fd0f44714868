set -x
timeout 900 python -m pytest tests/test_gpu_bvc.py -x -q > gpurun_out/r02_bvc_pytest.log 2>&1; tail -30 gpurun_out/r02_bvc_pytest.log
timeout 300 python tests/bindings_check.py 2 gpu 2>&1 | tail -5
timeout 300 python tests/bindings_check.py 3 gpu 2>&1 | tail -3
