#!/bin/bash
# Per-SOURCE-LINE instruction counts and stall samples of one kernel launch (-lineinfo build):
#   bash scripts/ncu_hot_lines.sh OUTNAME KERNEL_REGEX SKIP -- <command ...>
# ncu --set full --import-source on, then the source page (CUDA view) as CSV, reduced on the box to the
# hottest lines (the .ncu-rep is too big to travel).  Output: gpurun_out/OUTNAME.txt
out=$1; regex=$2; skip=$3; shift 4
"$@" > /dev/null 2>&1 || { echo "plain run failed"; exit 1; }
timeout 500 ncu --set full --import-source on --clock-control none -k regex:$regex -s $skip -c 1 -f -o /tmp/hotl "$@" > /tmp/hotl.log 2>&1
ncu -i /tmp/hotl.ncu-rep --page source --csv --print-source sass > /tmp/hotl.csv 2>/tmp/hotl_err.txt
python - "$out" <<'PY'
import csv, sys
rows = list(csv.reader(open('/tmp/hotl.csv', errors='replace')))
hdr = None
for i, r in enumerate(rows):
    if 'Source' in r and any('Sampling' in c for c in r):
        hdr = i; break
if hdr is None:
    print('no header found; first rows:', rows[:5]); sys.exit(0)
H = rows[hdr]
isrc = H.index('Source')
isamp = [k for k, c in enumerate(H) if 'Warp Stall Sampling (All' in c][0]
iexec = [k for k, c in enumerate(H) if c.strip() == 'Instructions Executed'][0]
data = []
for n, r in enumerate(rows[hdr + 1:]):
    if len(r) <= max(isamp, iexec): continue
    try: v = float(r[isamp].replace(',', '')); e = float(r[iexec].replace(',', ''))
    except ValueError: continue
    data.append((e, v, n, r[isrc].strip()[:110]))
te = sum(d[0] for d in data); tv = sum(d[1] for d in data)
def opcode(src):
    w = src.split()
    if not w: return '?'
    op = w[1] if w[0].startswith('@') and len(w) > 1 else w[0]
    return '.'.join(op.split('.')[:2]) if op.startswith(('LDG', 'LDS', 'STS', 'SHFL', 'BAR', 'MUFU')) else op.split('.')[0]
byop = {}
for e, v, n, src in data:
    k = opcode(src); a = byop.setdefault(k, [0.0, 0.0]); a[0] += e; a[1] += v
o = open('gpurun_out/%s.txt' % sys.argv[1], 'w')
o.write('# SASS of one launch (ncu --set full --import-source on): %.3e warp-instructions executed, %.0f stall samples\n' % (te, tv))
o.write('# by opcode: instructions executed (share) | stall samples (share)\n')
for k, (e, v) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:30]:
    o.write('  %-12s %12.0f %5.1f%% | %9.0f %5.1f%%\n' % (k, e, 100 * e / max(te, 1), v, 100 * v / max(tv, 1)))
o.write('# hottest instructions by stall samples: samples, share, executed, row, SASS\n')
for e, v, n, src in sorted(data, key=lambda d: -d[1])[:60]:
    o.write('%9.0f %5.2f%% %12.0f %6d  %s\n' % (v, 100 * v / max(tv, 1), e, n, src))
o.close()
print(open('gpurun_out/%s.txt' % sys.argv[1]).read()[:7000])
PY
