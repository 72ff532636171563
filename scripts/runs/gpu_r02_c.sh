#!/bin/bash
# round-2 GPU pass C: resident-weight NoF kernel A/B (tests, cycle breakdown, bench)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_chain.py tests/test_gpu_grad.py tests/test_gpu_scale.py -m gpu -q -x --timeout 900 > gpurun_out/pytest_r02c.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r02c.log
for res in 0 1; do
  MCF_NOF_RESIDENT=$res python scripts/chain_timing.py > gpurun_out/chain_timing_res$res.log 2>&1; echo "timing res=$res rc=$?"
  MCF_NOF_RESIDENT=$res python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_res$res.json 2> gpurun_out/bench_train_res$res.err; echo "train res=$res rc=$?"
  MCF_NOF_RESIDENT=$res python bench.py --workload render --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_render_res$res.json 2> gpurun_out/bench_render_res$res.err; echo "render res=$res rc=$?"
done
