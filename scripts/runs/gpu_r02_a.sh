#!/bin/bash
# round-2 GPU pass A: whole gpu test-suite (incl. BASELINE-scale parity) + train / render bench lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA --timeout 900 > gpurun_out/pytest_r02a.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r02a.log
tail -5 gpurun_out/pytest_r02a.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_train_r02a.json 2> gpurun_out/bench_train_r02a.err
echo "bench train rc=$?"
python bench.py --workload render --steps 20 --warmup 5 > gpurun_out/bench_render_r02a.json 2> gpurun_out/bench_render_r02a.err
echo "bench render rc=$?"
