#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_grad.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_r02h.log 2>&1
rc=$?; echo "pytest chain/grad rc=$rc"; tail -12 gpurun_out/pytest_r02h.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_r02h.json 2> gpurun_out/bench_train_r02h.err; echo "train rc=$?"
timeout 300 python bench.py --workload render --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_render_r02h.json 2> gpurun_out/bench_render_r02h.err; echo "render rc=$?"
