#!/bin/bash
# Round-2 ncu evidence (run on the GPU box via gpurun, ONE GPU): launch lists of the bench step + full captures of the
# top kernels + stall breakdown of the chain kernels.  Everything is exported to CSV on the box (gpurun_out/ is capped
# at 64 MiB); the .ncu-rep files are dropped.
set -u
mkdir -p gpurun_out
R=${1:-r02}
export_rep () {  # $1 = report basename, $2 = "source" to also export the per-instruction page
  local f=gpurun_out/$1.ncu-rep
  [ -f "$f" ] || { echo "missing $f"; return; }
  ncu -i "$f" --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i "$f" --page details --csv > gpurun_out/$1_details.csv 2>/dev/null
  if [ "${2:-}" = "source" ]; then ncu -i "$f" --page source --csv > gpurun_out/$1_source.csv 2>/dev/null; fi
  rm -f "$f"
}
B="--no-cpu-baseline --no-graph --no-self-check"
# 1. every launch of a short bench with its device time (cold-cache, serialised: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
    --log-file gpurun_out/launches_train_$R.csv python bench.py --steps 2 --warmup 3 $B > gpurun_out/ncu_bench_train_$R.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_render_$R.csv python bench.py --workload render --steps 2 --warmup 3 $B > gpurun_out/ncu_bench_render_$R.log 2>&1
# 2. full capture, render step: per step k_nof (coarse), k_chain sigma (coarse), k_nof (fine), k_chain (fine)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_chain|k_nof" -s 12 -c 4 \
    -o gpurun_out/prof_chain_render_$R -f python bench.py --workload render --steps 1 --warmup 3 $B > gpurun_out/ncu_chain_$R.log 2>&1
export_rep prof_chain_render_$R source
# 3. full capture, training step: 24 chain launches per step (12 fwd then 12 bwd); #82..85 of the 4th step =
#    last NoF fwd (fine), NeRF fwd (fine), NeRF bwd (fine), first NoF bwd (fine)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_chain|k_nof" -s 82 -c 4 \
    -o gpurun_out/prof_chain_train_$R -f python bench.py --steps 1 --warmup 3 $B > gpurun_out/ncu_chain_train_$R.log 2>&1
export_rep prof_chain_train_$R source
# 4. weight-gradient GEMM: first launches of the 4th step's backward (fine NeRF, then fine NoF)
timeout 900 ncu --set full --clock-control none -k regex:k_dw -s 36 -c 3 \
    -o gpurun_out/prof_dw_$R -f python bench.py --steps 1 --warmup 3 $B > gpurun_out/ncu_dw_$R.log 2>&1
export_rep prof_dw_$R
# 5. HBM-bound kernels at frame scale (291 600 rays): composite (coarse sigma-only, fine) + sample_pdf of the 4th frame
timeout 900 ncu --set full --clock-control none -k regex:"k_composite|k_sample_pdf" -s 9 -c 3 \
    -o gpurun_out/prof_render_ops_$R -f python bench.py --workload frame --steps 1 --warmup 3 $B > gpurun_out/ncu_ops_$R.log 2>&1
export_rep prof_render_ops_$R
du -sh gpurun_out
