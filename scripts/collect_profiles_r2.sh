#!/bin/bash
# Round-2 profile collection (run on the GPU box through gpurun; outputs under gpurun_out/).
# Every ncu command follows a clean run of the same program without ncu.
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof2_bench_plain.json 2> gpurun_out/prof2_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof2_bench_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:qn_resident_tile -s 3 -c 1 -f -o /tmp/r2_tile60_lh_bench_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof2_tile_ncu.log 2>&1
# config 1: ONE problem (README example) through the host entry, and the kernels that run n = 96 and n = 160 now
python scripts/single_problem.py > gpurun_out/r2_single_problem.txt 2>&1
for n in 96 160; do
  python scripts/prof_case.py --n $n --batch 296 --obj rosenbrock --max-iter 100 > gpurun_out/prof2_n${n}_plain.txt 2>&1
  ncu --set full --import-source on --clock-control none -k regex:qn_resident_kernel -s 1 -c 1 -f -o /tmp/r2_hybrid${n}_full \
      python scripts/prof_case.py --n $n --batch 296 --obj rosenbrock --max-iter 100 > gpurun_out/prof2_n${n}_ncu.log 2>&1
done
# streaming path launch lists at the floor (n = 1024) and at config 5's size
for n in 1024 8192; do
  QN_STREAM_MODE=twopass ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_stream_launches_n${n}.csv \
      python scripts/stream_bench.py $n > /dev/null 2>&1
done
# summaries only travel back (the .ncu-rep files are ~30 MB each; gpurun_out is capped at 64 MiB)
python tools/ncu_summary.py /tmp/r2_tile60_lh_bench_full.ncu-rep gpurun_out/r2_tile60_lh_bench_ncu.txt "round 2 final: the headline kernel (leader/helper warps, tiles in tensor memory, one fp64 element per tcgen05.st) at the bench's own launch (n=60, 2^16 problems, bench.py --steps 2 --warmup 3 --no-cpu-baseline, 4th launch)" > /dev/null
for n in 96 160; do
  python tools/ncu_summary.py /tmp/r2_hybrid${n}_full.ncu-rep gpurun_out/r2_hybrid${n}_ncu.txt "round 2: the kernel that runs Rosenbrock n=$n now (hybrid 4-row tiles, two rows in registers and two in shared memory, n at run time), 296 problems x 100 iterations" > /dev/null
done
ls -la /tmp/*.ncu-rep
