#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_chain.py tests/test_gpu_grad.py tests/test_gpu_ops.py tests/test_correspondence.py -m gpu -q --timeout 300 > gpurun_out/pytest_r02i.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/pytest_r02i.log | cut -c1-220
for sms in 0 32 48 64; do
  MCF_DW_OVERLAP_SMS=$sms timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_ov$sms.json 2> gpurun_out/bench_train_ov$sms.err; echo "train overlap=$sms rc=$?"
done
timeout 120 python bench.py --workload frame --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_frame_r02i.json 2> gpurun_out/bench_frame_r02i.err; echo "frame rc=$?"
