#!/bin/bash
mkdir -p gpurun_out
scripts/dbg/ts_mma > gpurun_out/ts_mma.log 2>&1; echo "ts_mma rc=$?"; grep rate gpurun_out/ts_mma.log
python -m pytest tests/test_gpu_ops.py tests/test_gpu_scale.py -m gpu -q --timeout 900 > gpurun_out/pytest_r02f.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_r02f.log
python bench.py --workload frame --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_frame_r02f.json 2> gpurun_out/bench_frame_r02f.err; echo "frame rc=$?"
python bench.py --workload stress --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stress_r02f.json 2> gpurun_out/bench_stress_r02f.err; echo "stress rc=$?"
