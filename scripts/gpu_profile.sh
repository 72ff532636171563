#!/bin/bash
# Run on the GPU box (via gpurun): ncu launch list of the bench step + full captures of the top kernels.
# Outputs go to gpurun_out/; summaries are copied into profiles/ by scripts/summarise_profiles.py.
set -u
mkdir -p gpurun_out
R=${1:-r01}
# 1. every launch of a short train bench with its device time (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/launches_train_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_bench_train_$R.log 2>&1
# 2. same for the render workload
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_render_$R.csv python bench.py --workload render --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_bench_render_$R.log 2>&1
# 3. full capture of the chain kernel (one render step: NoF coarse, NeRF sigma coarse, NoF fine, NeRF fine)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 12 -c 4 \
    -o gpurun_out/prof_chain_render_$R -f python bench.py --workload render --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_chain_$R.log 2>&1
# 4. full capture of the training kernels: chain fwd(train)/bwd and the weight-gradient GEMM
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 72 -c 24 \
    -o gpurun_out/prof_chain_train_$R -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_chain_train_$R.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dw -s 264 -c 6 \
    -o gpurun_out/prof_dw_$R -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_dw_$R.log 2>&1
ls -la gpurun_out
