#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chain.py tests/test_gpu_grad.py -m gpu -q --timeout 200 > gpurun_out/pytest_r02n.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r02n.log | cut -c1-200
b() { name=$1; shift; timeout 120 python bench.py "$@" --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name rc=$?"; }
b train_h20000 --steps 20 --warmup 5
b render_h20000 --workload render --steps 20 --warmup 5
for h in 0 2000 200000; do
  MCF_LIB_PATH=$PWD/moco_flow_b200/csrc/libmoco_flow_b200_h$h.so b train_h$h --steps 20 --warmup 5
  MCF_LIB_PATH=$PWD/moco_flow_b200/csrc/libmoco_flow_b200_h$h.so b render_h$h --workload render --steps 20 --warmup 5
done
