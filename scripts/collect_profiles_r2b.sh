#!/bin/bash
# Round 2, second collection: the leader/helper tensor-memory tile kernels of n = 128 and n = 160.
set -x
for n in 128 160; do
  python scripts/prof_case.py --n $n --batch 1184 --obj rosenbrock --max-iter 200 > gpurun_out/prof2_tilew${n}_plain.txt 2>&1 || exit 1
  ncu --set full --import-source on --clock-control none -k regex:qn_resident_tilew -s 1 -c 1 -f -o /tmp/r2_tilew${n}_full \
      python scripts/prof_case.py --n $n --batch 1184 --obj rosenbrock --max-iter 200 > gpurun_out/prof2_tilew${n}_ncu.log 2>&1
  python tools/ncu_summary.py /tmp/r2_tilew${n}_full.ncu-rep gpurun_out/r2_tilew${n}_ncu.txt "round 2: the kernel that runs Rosenbrock n=$n now (qn_resident_tilew_kernel: one CTA per problem, a warp per column block, the whole tile in tensor memory, leader warp), 1184 problems x 200 iterations" > /dev/null
done
