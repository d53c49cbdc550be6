#!/bin/bash
# Per-scheduler (SM sub-partition) counters of one resident kernel: how evenly the warps' work lands
# on the four schedulers of an SM.
# usage (on the GPU box): bash scripts/ncu_smsp.sh OUT "N BATCH [option=value ...]" ["N BATCH ..." ...]
M=smsp__cycles_elapsed.max,smsp__issue_active.avg,smsp__issue_active.max,smsp__issue_active.min,smsp__inst_executed_pipe_fp64.avg,smsp__inst_executed_pipe_fp64.max,smsp__inst_executed_pipe_fp64.min,smsp__warps_eligible.avg,smsp__warps_eligible.max,smsp__warps_eligible.min
out=gpurun_out/$1.txt; : > $out; shift
for spec in "$@"; do
  timeout 150 python scripts/one_launch.py $spec >> $out 2>&1 || { echo "plain run failed: $spec" >> $out; continue; }
  timeout 200 ncu --clock-control none --metrics $M -k regex:qn_resident -s 1 -c 1 --csv --page raw \
    --log-file /tmp/smsp.csv python scripts/one_launch.py $spec > /dev/null 2>&1
  python - >> $out <<'PY'
import csv
rows = [r for r in csv.reader(open("/tmp/smsp.csv")) if r]
hdr = [r for r in rows if r[0] == "ID"]
if not hdr:
    print("no ncu rows:", rows[-3:])
else:
    i = rows.index(hdr[0]); val = rows[i + 2]
    for h, v in zip(hdr[0], val):
        if "smsp__" in h and not h.endswith(".sum"): print("   %-45s %s" % (h, v))
PY
done
cat $out
