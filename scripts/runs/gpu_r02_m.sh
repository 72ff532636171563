#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chain.py tests/test_gpu_grad.py -m gpu -q --timeout 200 > gpurun_out/pytest_r02m.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r02m.log | cut -c1-200
H=$PWD/moco_flow_b200/csrc/libmoco_flow_b200_hint.so
b() { name=$1; shift; timeout 120 python bench.py "$@" --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name rc=$?"; }
MCF_EARLY_ARRIVE=0 b train_early0 --steps 20 --warmup 5
MCF_EARLY_ARRIVE=1 b train_early1 --steps 20 --warmup 5
MCF_LIB_PATH=$H b train_hint --steps 20 --warmup 5
b render_nohint --workload render --steps 20 --warmup 5
MCF_LIB_PATH=$H b render_hint --workload render --steps 20 --warmup 5
MCF_LIB_PATH=$H b frame_hint --workload frame --steps 4 --warmup 3
