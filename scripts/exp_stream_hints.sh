#!/bin/bash
# Time the streaming path for every cache-hint combination of the B^-1 accesses (runtime-selected
# kernel instantiations; see qn_stream.cu ld_h/st_h).  ld: 0 default 1 .cs 2 .lu 3 .cv 4 .cg ; st: 0 default 1 .cs 2 .cg 3 .wt
export QN_STREAM_ITERS=42
for ld in 0 1 2 3 4; do for st in 0 1 2 3; do
  echo -n "one-pass ld=$ld st=$st  "
  QN_STREAM_MODE=onepass QN_HINT_ONE_LD=$ld QN_HINT_ONE_ST=$st python scripts/stream_bench.py 2>&1 | grep ms_per_iter | sed 's/.*ms_per_iter.: \([0-9.]*\).*/\1/'
done; done
for p1 in 0 1 2 3 4; do
  echo -n "two-pass p1ld=$p1 p2ld=0 p2st=1  "
  QN_STREAM_MODE=twopass QN_HINT_P1_LD=$p1 QN_HINT_P2_LD=0 QN_HINT_P2_ST=1 python scripts/stream_bench.py 2>&1 | grep ms_per_iter | sed 's/.*ms_per_iter.: \([0-9.]*\).*/\1/'
done
for ld in 0 1 2 3 4; do for st in 0 1 2 3; do
  echo -n "two-pass p1ld=0 p2ld=$ld p2st=$st  "
  QN_STREAM_MODE=twopass QN_HINT_P1_LD=0 QN_HINT_P2_LD=$ld QN_HINT_P2_ST=$st python scripts/stream_bench.py 2>&1 | grep ms_per_iter | sed 's/.*ms_per_iter.: \([0-9.]*\).*/\1/'
done; done
