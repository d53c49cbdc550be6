#!/bin/bash
# Round 2: the kernels that run BASELINE.json configs[2] now — dense Gaussian (cooperative objective of the
# leader/helper tensor-memory tiles) and logistic regression (Team kernel) at n = 128.
set -x
for obj in gaussian logistic; do
  batch=4736; [ $obj = logistic ] && batch=592
  python scripts/prof_case.py --n 128 --batch $batch --obj $obj > gpurun_out/prof4_${obj}_plain.txt 2>&1 || exit 1
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:qn_resident -s 1 -c 1 -f -o /tmp/r2_${obj}128_full \
      python scripts/prof_case.py --n 128 --batch $batch --obj $obj > gpurun_out/prof4_${obj}_ncu.log 2>&1
  python tools/ncu_summary.py /tmp/r2_${obj}128_full.ncu-rep gpurun_out/r2_${obj}128_ncu.txt "round 2: the kernel that runs the n=128 $obj objective now, $batch problems to convergence (scripts/prof_case.py)" > /dev/null
  cat gpurun_out/prof4_${obj}_plain.txt
done
