#!/bin/bash
# Round 2, last collection: the headline kernel as bench.py launches it after the leader-placement step.
# Every ncu command follows a clean run of the same program without ncu; only text summaries travel back.
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof3_bench_plain.json 2> gpurun_out/prof3_bench_plain.err || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof3_bench_ncu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qn_resident_tile -s 3 -c 1 -f -o /tmp/r2_tile60_final_bench_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof3_tile_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r2_tile60_final_bench_full.ncu-rep gpurun_out/r2_tile60_final_bench_ncu.txt "round 2 final: the headline kernel (leader/helper warps with the same warp of every pair leading, tiles in tensor memory, one fp64 element per tcgen05.st) at the bench's own launch (n=60, 2^16 problems, bench.py --steps 2 --warmup 3 --no-cpu-baseline, 4th launch)" > /dev/null
ls -la /tmp/*.ncu-rep; head -c 1500 gpurun_out/r2_tile60_final_bench_ncu.txt
