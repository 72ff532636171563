#!/bin/bash
# round-2 GPU pass B: gpu test-suite + one bench line per BASELINE config on one B200 + eager-PyTorch-on-GPU baseline
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA --timeout 900 > gpurun_out/pytest_r02b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r02b.log
tail -4 gpurun_out/pytest_r02b.log
run() { name=$1; shift; python bench.py "$@" > gpurun_out/bench_${name}.json 2> gpurun_out/bench_${name}.err; echo "bench $name rc=$?"; tail -c 300 gpurun_out/bench_${name}.err; }
run train_r02b --steps 20 --warmup 5
run render_r02b --workload render --steps 20 --warmup 5
run cfg1_r02b --workload cfg1 --steps 20 --warmup 5
run cfg1_ref_r02b --workload cfg1 --impl reference --steps 5 --warmup 2
run frame_r02b --workload frame --steps 16 --warmup 3
run stress_r02b --workload stress --steps 3 --warmup 3 --no-cpu-baseline
run train_ref_r02b --impl reference --steps 5 --warmup 2
run train_torchgpu_tf32off --impl reference --device cuda --tf32 0 --steps 5 --warmup 2
run train_torchgpu_tf32on --impl reference --device cuda --tf32 1 --steps 5 --warmup 2
run render_torchgpu_tf32off --impl reference --workload render --device cuda --tf32 0 --steps 5 --warmup 2
run render_torchgpu_tf32on --impl reference --workload render --device cuda --tf32 1 --steps 5 --warmup 2
