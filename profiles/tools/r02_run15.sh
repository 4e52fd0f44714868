set -x
timeout 600 python -m pytest tests/test_gpu_stepper.py -x -q -k "distributed" 2>&1 | tail -6
for fp in replicated data; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench_step.py --case smoke3d --iters 1000 --steps 2 --gpus 2 --fit-parallel $fp 2>&1 | grep sim_steps | cut -c 1-90
done
timeout 600 python -m pytest tests/test_gpu_siren.py -x -q 2>&1 | tail -2
