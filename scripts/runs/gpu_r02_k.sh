#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 120 > gpurun_out/pytest_r02k.log 2>&1; echo "pytest ops rc=$?"; tail -3 gpurun_out/pytest_r02k.log | cut -c1-200
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_train_n2_r02.json 2> gpurun_out/bench_train_n2_r02.err; echo "train n2 rc=$?"; tail -c 300 gpurun_out/bench_train_n2_r02.err
timeout 100 python bench.py --workload frame --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_frame_r02k.json 2> gpurun_out/bench_frame_r02k.err; echo "frame rc=$?"
