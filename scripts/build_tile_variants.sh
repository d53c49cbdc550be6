#!/bin/bash
# Builds tools/bin/{tb_*,pt_*}: variants of the n = 60 tile kernel (tools/tile_bench.cu, tools/phase_timing.cu)
# for one gpurun call (scripts/run_tile_variants.sh).  tools/bin/ is git-ignored and travels to the GPU box.
set -e
cd "$(dirname "$0")/.."
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 --expt-relaxed-constexpr -lineinfo -I quasinewtonmethods.jl_b200/csrc"
mkdir -p tools/bin
build() { # name, extra flags
  local name=$1; shift
  $NV -DTB_TMEM "$@" -o tools/bin/tb_$name tools/tile_bench.cu 2>&1 | grep -E "error" || true
}
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  build $name $flags &
done
wait
ls -la tools/bin
