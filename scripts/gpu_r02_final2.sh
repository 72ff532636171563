#!/bin/bash
# round-2 closing pass on one GPU: whole gpu test-suite, smoke, one bench line per BASELINE config, then a fresh ncu
# capture of the training chain kernels (they changed last: per-warp save stores, register-selected mask words)
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -rA --timeout 150 > gpurun_out/pytest_r02_final.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_r02_final.log; grep -E "passed|failed" gpurun_out/pytest_r02_final.log | tail -3
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_r02.log
run() { name=$1; shift; timeout 120 python bench.py "$@" > gpurun_out/bench_${name}.json 2> gpurun_out/bench_${name}.err; echo "bench $name rc=$?"; tail -c 200 gpurun_out/bench_${name}.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_${name}.json | head -1; }
run train_final --steps 20 --warmup 5
run render_final --workload render --steps 20 --warmup 5
run cfg1_final --workload cfg1 --steps 20 --warmup 5
run frame_final --workload frame --steps 16 --warmup 3
run stress_final --workload stress --steps 3 --warmup 3 --no-cpu-baseline
MCF_DW_OVERLAP_SMS=0 run train_serial_final --steps 20 --warmup 5 --no-cpu-baseline
R=r02g
B="--no-cpu-baseline --no-graph --no-self-check"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k_chain|k_nof" -s 82 -c 4 \
    -o gpurun_out/prof_chain_train_$R -f python bench.py --steps 1 --warmup 3 $B > gpurun_out/ncu_chain_train_$R.log 2>&1
f=gpurun_out/prof_chain_train_$R.ncu-rep
if [ -f "$f" ]; then
  ncu -i "$f" --page raw --csv > gpurun_out/prof_chain_train_${R}_raw.csv 2>/dev/null
  ncu -i "$f" --page details --csv > gpurun_out/prof_chain_train_${R}_details.csv 2>/dev/null
  ncu -i "$f" --page source --csv > gpurun_out/prof_chain_train_${R}_source.csv 2>/dev/null
  rm -f "$f"
fi
echo "ncu chain_train done"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
    --log-file gpurun_out/launches_train_$R.csv python bench.py --steps 2 --warmup 3 $B > gpurun_out/ncu_bench_train_$R.log 2>&1
echo "ncu launches done"
MCF_LIB_PATH=moco_flow_b200/csrc/variants/lib_directsave.so run directsave_train --steps 20 --warmup 5 --no-cpu-baseline
