#!/bin/bash
# Build a variant of the library for an A/B run on the GPU box:
#     scripts/build_variant.sh <name> <unit.cu[,unit.cu...]> <-D flags...>
# Recompiles the named translation units with the extra flags and links them with the objects of the regular build
# into moco_flow_b200/csrc/variants/lib_<name>.so (git-ignored, travels with gpurun).  Use: MCF_LIB_PATH=<that file>
set -e
cd "$(dirname "$0")/.."
name=$1; units=${2//,/ }; shift 2
python -m moco_flow_b200.build > /dev/null
C=moco_flow_b200/csrc
mkdir -p $C/variants
for unit in $units; do
  extra=""
  case $unit in render_ops.cu|camera.cu|correspondence.cu) extra="-fmad=false";; esac
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $extra "$@" \
    -c $C/$unit -o $C/variants/${name}_${unit%.cu}.o
done
objs=""
for o in $C/*.o; do
  v=$C/variants/${name}_$(basename $o)
  if [ -f "$v" ]; then objs="$objs $v"; else objs="$objs $o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $C/variants/lib_${name}.so $objs
echo $C/variants/lib_${name}.so
