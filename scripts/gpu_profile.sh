#!/bin/bash
# Run on the GPU box (via gpurun): ncu launch lists of the bench step + full captures of the top kernels.
# Everything is exported to CSV on the box (gpurun_out/ is capped at 64 MiB); large .ncu-rep files are dropped.
set -u
mkdir -p gpurun_out
R=${1:-r01}
WHAT=${2:-all}
export_rep () {  # $1 = report basename
  local f=gpurun_out/$1.ncu-rep
  [ -f "$f" ] || return
  ncu -i "$f" --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i "$f" --page details --csv > gpurun_out/$1_details.csv 2>/dev/null
  if [ "${2:-}" = "source" ]; then ncu -i "$f" --page source --csv > gpurun_out/$1_source.csv 2>/dev/null; fi
  local sz=$(stat -c %s "$f")
  if [ "$sz" -gt 12000000 ]; then rm -f "$f"; fi
}
if [ "$WHAT" = "all" ] || [ "$WHAT" = "lists" ]; then
# 1. every launch of a short bench with its device time (cold-cache, serialised: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/launches_train_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph \
    > gpurun_out/ncu_bench_train_$R.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_render_$R.csv python bench.py --workload render --steps 2 --warmup 3 --no-cpu-baseline --no-graph \
    > gpurun_out/ncu_bench_render_$R.log 2>&1
fi
if [ "$WHAT" = "all" ] || [ "$WHAT" = "full" ]; then
# 2. full capture, render step (4 chain launches per step: NoF coarse, NeRF sigma coarse, NoF fine, NeRF fine)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 14 -c 2 \
    -o gpurun_out/prof_chain_render_$R -f python bench.py --workload render --steps 1 --warmup 3 --no-cpu-baseline --no-graph \
    > gpurun_out/ncu_chain_$R.log 2>&1
export_rep prof_chain_render_$R source
# 3. full capture, training step: 24 chain launches per step (12 fwd then 12 bwd); #82..85 of the 4th step =
#    last NoF fwd (fine), NeRF fwd fine, NeRF bwd fine, first NoF bwd (fine)
timeout 900 ncu --set full --clock-control none -k regex:k_chain -s 82 -c 4 \
    -o gpurun_out/prof_chain_train_$R -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph \
    > gpurun_out/ncu_chain_train_$R.log 2>&1
export_rep prof_chain_train_$R
# 4. weight-gradient GEMM: first launches of the 4th step's backward (fine NeRF layers)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dw -s 36 -c 3 \
    -o gpurun_out/prof_dw_$R -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph \
    > gpurun_out/ncu_dw_$R.log 2>&1
export_rep prof_dw_$R source
# 5. HBM-bound kernels at frame scale would need a bigger batch; capture them from the train step
timeout 600 ncu --set full --clock-control none -k regex:"k_composite|k_sample_pdf" -s 15 -c 5 \
    -o gpurun_out/prof_render_ops_$R -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph \
    > gpurun_out/ncu_ops_$R.log 2>&1
export_rep prof_render_ops_$R
fi
du -sh gpurun_out; ls -la gpurun_out | head -40
