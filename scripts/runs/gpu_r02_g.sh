#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_grad.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_r02g.log 2>&1
rc=$?; echo "pytest chain/grad rc=$rc"; tail -15 gpurun_out/pytest_r02g.log | cut -c1-200
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests/test_gpu_scale.py tests/test_gpu_ops.py tests/test_correspondence.py -m gpu -q --timeout 600 > gpurun_out/pytest_r02g2.log 2>&1
echo "pytest scale/ops/knn rc=$?"; tail -8 gpurun_out/pytest_r02g2.log | cut -c1-200
for k in smem ts; do
  MCF_NOF_KERNEL=$k timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_nof_$k.json 2> gpurun_out/bench_train_nof_$k.err; echo "train $k rc=$?"
  MCF_NOF_KERNEL=$k timeout 300 python bench.py --workload render --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_render_nof_$k.json 2> gpurun_out/bench_render_nof_$k.err; echo "render $k rc=$?"
done
timeout 300 python bench.py --workload frame --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_frame_r02g.json 2> gpurun_out/bench_frame_r02g.err; echo "frame rc=$?"
