#!/bin/bash
mkdir -p gpurun_out
for sms in 16 24 32 40; do
  MCF_DW_OVERLAP_SMS=$sms timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_ov$sms.json 2> gpurun_out/bench_train_ov$sms.err; echo "train overlap=$sms rc=$?"
done
bash scripts/gpu_profile_r02.sh r02
