#!/bin/bash
mkdir -p gpurun_out
for sms in 0 24 32 40 48; do
  MCF_DW_OVERLAP_SMS=$sms timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_ovB$sms.json 2> gpurun_out/bench_train_ovB$sms.err; echo "train overlap=$sms rc=$?"
done
