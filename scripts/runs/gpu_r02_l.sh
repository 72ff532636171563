#!/bin/bash
# 8-GPU pass: cfg3 at N=8 and cfg5 (1080x1080, 128+128, flow-consistency pass, gathered)
mkdir -p gpurun_out
port=29560
trun() { n=$1; name=$2; shift 2; port=$((port+1)); timeout 130 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > gpurun_out/bench_${name}.json 2> gpurun_out/bench_${name}.err; echo "bench $name rc=$?"; tail -c 300 gpurun_out/bench_${name}.json; echo; }
trun 8 stress_n8_r02 --workload stress --steps 4 --warmup 3
trun 8 train_n8_r02 --steps 20 --warmup 5
