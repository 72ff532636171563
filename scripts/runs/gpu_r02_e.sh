#!/bin/bash
mkdir -p gpurun_out
scripts/dbg/ts_mma > gpurun_out/ts_mma.log 2>&1; echo "ts_mma rc=$?"; cat gpurun_out/ts_mma.log
scripts/dbg/mma_rate > gpurun_out/mma_rate.log 2>&1; echo "mma_rate rc=$?"; head -20 gpurun_out/mma_rate.log
python -m pytest tests/test_gpu_chain.py tests/test_gpu_grad.py -m gpu -q --timeout 900 > gpurun_out/pytest_r02e.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r02e.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_r02e.json 2> gpurun_out/bench_train_r02e.err; echo "train rc=$?"
