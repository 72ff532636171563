#!/bin/bash
# round-2 A/B: training saves stored per epilogue warp (default build) against the group-store build, plus the
# mixed variants; the group-store build is also the check of the semi-transparent bench weights
mkdir -p gpurun_out
V=moco_flow_b200/csrc/variants
run() { name=$1; shift; timeout 200 python bench.py "$@" --no-cpu-baseline > gpurun_out/bench_${name}.json 2> gpurun_out/bench_${name}.err; echo "bench $name rc=$?"; tail -c 300 gpurun_out/bench_${name}.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_${name}.json | head -1; }
MCF_LIB_PATH=$V/lib_groupstore.so timeout 300 python -m pytest tests/test_gpu_grad.py tests/test_gpu_scale.py -m gpu -q --timeout 200 > gpurun_out/pytest_q_groupstore.log 2>&1
echo "pytest groupstore rc=$?"; tail -3 gpurun_out/pytest_q_groupstore.log
MCF_LIB_PATH=$V/lib_groupstore.so run q_groupstore_train --steps 20 --warmup 5
MCF_LIB_PATH=$V/lib_groupstore.so run q_groupstore_render --workload render --steps 20 --warmup 5
timeout 400 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_q_warp.log 2>&1
echo "pytest warp rc=$?"; tail -3 gpurun_out/pytest_q_warp.log
run q_warp_train --steps 20 --warmup 5
MCF_LIB_PATH=$V/lib_nofwarp.so run q_nofwarp_train --steps 20 --warmup 5
MCF_LIB_PATH=$V/lib_chainwarp.so run q_chainwarp_train --steps 20 --warmup 5
