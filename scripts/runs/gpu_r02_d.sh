#!/bin/bash
mkdir -p gpurun_out
scripts/dbg/ts_mma > gpurun_out/ts_mma.log 2>&1; echo "ts_mma rc=$?"; cat gpurun_out/ts_mma.log
python -m pytest tests/test_gpu_chain.py tests/test_gpu_grad.py tests/test_gpu_scale.py -m gpu -q --timeout 900 > gpurun_out/pytest_r02d.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r02d.log
