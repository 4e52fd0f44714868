set -x
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/r02_bench_n2.json
timeout 600 python -m pytest tests/test_gpu_stepper.py -x -q -k "distributed" 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench_step.py --case smoke3d --iters 1000 --steps 2 --gpus 2 2>&1 | tail -2 | cut -c 1-300
