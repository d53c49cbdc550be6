#!/bin/bash
# Runs every tools/bin/tb_* (timing + output hash of the n = 60 tile kernel variants) and pt_* (per-phase
# cycles) on the GPU box; output to gpurun_out/$1.txt.  Built by scripts/build_tile_variants.sh.
out=gpurun_out/${1:-tile_variants}.txt
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out
for b in tools/bin/tb_*; do
  echo "== $b" >> $out
  timeout 120 $b 65536 3 >> $out 2>&1
done
for b in tools/bin/pt_*; do
  for np in 1 2 1184; do
    echo "== $b $np" >> $out
    timeout 60 $b $np >> $out 2>&1
  done
done
cat $out
