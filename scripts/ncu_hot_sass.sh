#!/bin/bash
# Per-instruction stall samples of the headline kernel: ncu --set full --import-source on, then the
# SASS page as CSV, reduced on the box to the hottest instructions (the .ncu-rep is too big to travel).
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:qn_resident_tile -s 3 -c 1 -f -o /tmp/hot \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /tmp/hot.log 2>&1
ncu -i /tmp/hot.ncu-rep --page source --csv --print-source sass > /tmp/hot_sass.csv 2>/tmp/hot_err.txt
python - <<'PY'
import csv, sys
rows = list(csv.reader(open('/tmp/hot_sass.csv', errors='replace')))
hdr = None
for i, r in enumerate(rows):
    if 'Source' in r and any('Sampling' in c for c in r):
        hdr = i; break
if hdr is None:
    print('no header found; first rows:', rows[:5]); sys.exit(0)
H = rows[hdr]
isrc = H.index('Source')
samp = [k for k, c in enumerate(H) if c.strip() == 'Warp Stall Sampling (All Samples)' or c.strip() == '# Samples' or 'Warp Stall Sampling (All' in c]
isamp = samp[0]
iexec = [k for k, c in enumerate(H) if 'Instructions Executed' in c]
data = []
for n, r in enumerate(rows[hdr + 1:]):
    if len(r) <= isamp: continue
    try: v = float(r[isamp].replace(',', ''))
    except ValueError: continue
    data.append((v, n, r[isrc].strip(), r[iexec[0]] if iexec else ''))
tot = sum(d[0] for d in data)
out = open('gpurun_out/r2_tile60_hot_sass.txt', 'w')
out.write('# hottest SASS instructions of the headline kernel by warp-stall samples (ncu --import-source on, bench launch); total samples %.0f\n' % tot)
out.write('# columns: samples, share, cumulative line index, instructions executed, SASS\n')
# also samples by opcode
byop = {}
for v, n, src, ex in data:
    op = src.split()[0] if src else '?'
    if op.startswith('@'): op = src.split()[1] if len(src.split()) > 1 else op
    op = op.split('.')[0]
    byop[op] = byop.get(op, 0) + v
out.write('# samples by opcode: ' + ', '.join('%s %.1f%%' % (k, 100 * v / tot) for k, v in sorted(byop.items(), key=lambda kv: -kv[1])[:25]) + '\n')
for v, n, src, ex in sorted(data, reverse=True)[:120]:
    out.write('%8.0f %5.2f%% %6d %14s  %s\n' % (v, 100 * v / tot, n, ex, src))
out.close()
print(open('gpurun_out/r2_tile60_hot_sass.txt').read()[:3000])
PY
