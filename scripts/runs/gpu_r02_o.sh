#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_chain.py tests/test_gpu_grad.py -m gpu -q -x --timeout 200 > gpurun_out/pytest_r02o.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -6 gpurun_out/pytest_r02o.log | cut -c1-200
if [ $rc -ne 0 ]; then exit 0; fi
b() { name=$1; shift; timeout 120 python bench.py "$@" --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name rc=$?"; }
for m in 0 1; do
  MCF_PAIR_MERGE=$m b render_merge$m --workload render --steps 20 --warmup 5
  MCF_PAIR_MERGE=$m b train_merge$m --steps 20 --warmup 5
done
MCF_PAIR_MERGE=1 b frame_merge1 --workload frame --steps 4 --warmup 3
timeout 300 python -m pytest tests/test_gpu_scale.py -m gpu -q --timeout 300 > gpurun_out/pytest_r02o2.log 2>&1; echo "pytest scale rc=$?"; tail -3 gpurun_out/pytest_r02o2.log | cut -c1-200
