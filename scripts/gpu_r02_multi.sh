#!/bin/bash
# round-2 multi-GPU pass (gpurun --gpus 8): 2-GPU DP gradient equality, train scaling, cfg4 (16 frames over 8 GPUs,
# gathered) and cfg5 (1080x1080, 128+128, flow-consistency pass, 8 GPUs)
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q --timeout 600 -s > gpurun_out/pytest_dp_r02.log 2>&1
echo "pytest dp rc=$?"; grep -E "\[dp\]|passed|failed|skipped" gpurun_out/pytest_dp_r02.log | cut -c1-400
port=29510
trun() { n=$1; name=$2; shift 2; port=$((port+1)); timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > gpurun_out/bench_${name}.json 2> gpurun_out/bench_${name}.err; echo "bench $name rc=$?"; tail -c 400 gpurun_out/bench_${name}.json | head -c 400; echo; }
trun 8 train_n8_r02 --steps 20 --warmup 5
trun 8 frame_n8_r02 --workload frame --steps 16 --warmup 3
trun 8 stress_n8_r02 --workload stress --steps 4 --warmup 3
trun 2 train_n2_r02 --steps 20 --warmup 5
trun 4 train_n4_r02 --steps 20 --warmup 5
