#!/bin/bash
# round-2 final single-GPU pass: the whole gpu test-suite + one bench line per BASELINE config
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA --timeout 600 > gpurun_out/pytest_r02_final.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_r02_final.log; grep -E "passed|failed" gpurun_out/pytest_r02_final.log | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_r02.log
run() { name=$1; shift; timeout 300 python bench.py "$@" > gpurun_out/bench_${name}.json 2> gpurun_out/bench_${name}.err; echo "bench $name rc=$?"; tail -c 200 gpurun_out/bench_${name}.err; }
run train_final --steps 20 --warmup 5
run render_final --workload render --steps 20 --warmup 5
run cfg1_final --workload cfg1 --steps 20 --warmup 5
run frame_final --workload frame --steps 16 --warmup 3
run stress_final --workload stress --steps 3 --warmup 3 --no-cpu-baseline
MCF_DW_OVERLAP_SMS=0 run train_serial_final --steps 20 --warmup 5 --no-cpu-baseline
