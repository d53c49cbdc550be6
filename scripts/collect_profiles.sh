#!/bin/bash
# Round-1 profile collection (run on the GPU box through gpurun; outputs under gpurun_out/).
# Every ncu command follows a clean run of the same program without ncu.
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_bench_plain.json 2> gpurun_out/prof_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_bench_ncu.log 2>&1
# the headline kernel at the bench's own launch size (2^16 problems): DRAM traffic per launch, stalls, pipes
ncu --set full --import-source on --clock-control none -k regex:qn_resident_tile -s 3 -c 1 -f -o gpurun_out/r1_tile60_bench_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_tile_ncu.log 2>&1
for mode in onepass twopass; do
  QN_STREAM_MODE=$mode python scripts/stream_bench.py > gpurun_out/prof_stream_plain_$mode.txt 2>&1 || exit 1
done
QN_STREAM_MODE=onepass ncu --set full --import-source on --clock-control none -k regex:stream_pass_one -s 4 -c 1 -f -o gpurun_out/r1_stream_pass_one_full \
    python scripts/stream_bench.py > gpurun_out/prof_stream_ncu1.log 2>&1
QN_STREAM_MODE=twopass ncu --set full --import-source on --clock-control none -k regex:stream_pass2 -s 4 -c 1 -f -o gpurun_out/r1_stream_pass2_full \
    python scripts/stream_bench.py > gpurun_out/prof_stream_ncu2.log 2>&1
QN_STREAM_MODE=twopass ncu --set full --import-source on --clock-control none -k regex:stream_pass1 -s 4 -c 1 -f -o gpurun_out/r1_stream_pass1_full \
    python scripts/stream_bench.py > gpurun_out/prof_stream_ncu3.log 2>&1
QN_STREAM_MODE=onepass ncu --set full --import-source on --clock-control none -k regex:stream_vec_cluster -s 4 -c 1 -f -o gpurun_out/r1_stream_cluster_full \
    python scripts/stream_bench.py > gpurun_out/prof_stream_ncu4.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/r1_stream_launches.csv \
    python scripts/stream_bench.py > gpurun_out/prof_stream_ncu5.log 2>&1
python scripts/prof_case.py --n 128 --batch 148 --obj rosenbrock --max-iter 100 > gpurun_out/prof_pair128_plain.txt 2>&1
ncu --set full --import-source on --clock-control none -k regex:qn_resident_kernel -s 1 -c 1 -f -o gpurun_out/r1_pair128_full \
    python scripts/prof_case.py --n 128 --batch 148 --obj rosenbrock --max-iter 100 > gpurun_out/prof_pair128_ncu.log 2>&1
python scripts/prof_case.py --n 8 --batch 262144 > gpurun_out/prof_packed8_plain.txt 2>&1
ncu --set full --import-source on --clock-control none -k regex:qn_resident_packed -s 1 -c 1 -f -o gpurun_out/r1_packed8_full \
    python scripts/prof_case.py --n 8 --batch 262144 > gpurun_out/prof_packed8_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
