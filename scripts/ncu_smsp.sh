#!/bin/bash
# Per-scheduler (SM sub-partition) counters of the n = 60 tile kernel for two leader placements.
# usage (on the GPU box): bash scripts/ncu_smsp.sh   -> gpurun_out/r2_lead_smsp.txt
M=smsp__cycles_elapsed.max,smsp__issue_active.avg,smsp__issue_active.max,smsp__issue_active.min,smsp__inst_executed.avg,smsp__inst_executed.max,smsp__inst_executed.min,smsp__inst_executed_pipe_fp64.max,smsp__inst_executed_pipe_fp64.min,smsp__warps_eligible.avg,smsp__warps_eligible.max,smsp__warps_eligible.min,smsp__issue_inst0.max,smsp__issue_inst0.min
out=gpurun_out/r2_lead_smsp.txt; : > $out
for d in 0 -1; do
  timeout 150 python scripts/one_launch.py 60 16384 lead_parity=$d >> $out 2>&1 || { echo "plain run failed" >> $out; continue; }
  timeout 200 ncu --clock-control none --metrics $M -k regex:qn_resident -s 1 -c 1 --csv --page raw \
    --log-file /tmp/smsp_$d.csv python scripts/one_launch.py 60 16384 lead_parity=$d > /dev/null 2>&1
  python - $d >> $out <<'PY'
import csv, sys
d = sys.argv[1]
rows = [r for r in csv.reader(open("/tmp/smsp_%s.csv" % d)) if r]
hdr = [r for r in rows if r[0] == "ID"]
if not hdr:
    print("no ncu rows:", rows[-3:]); sys.exit(0)
i = rows.index(hdr[0]); val = rows[i + 2]
print("lead_parity", d)
for h, v in zip(hdr[0], val):
    if "smsp__" in h: print("   %-45s %s" % (h, v))
PY
done
cat $out
