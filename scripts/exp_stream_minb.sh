#!/bin/bash
# Build with different residency targets for the streaming sweeps and time both modes (GPU box).
for cfg in "4 4 2" "4 3 2" "4 2 2" "4 4 4" "4 4 3"; do
  set -- $cfg
  QN_EXTRA_NVCC="-DQN_PASS1_MINB=$1 -DQN_PASS2_MINB=$2 -DQN_PASSONE_MINB=$3" python quasinewtonmethods.jl_b200/build.py --force > /dev/null 2>&1
  echo "== pass1 MINB=$1 pass2 MINB=$2 pass_one MINB=$3"
  python scripts/stream_bench.py 2>&1 | grep ms_per_iter | sed 's/algorithmic.*//'
done
python quasinewtonmethods.jl_b200/build.py --force > /dev/null 2>&1
