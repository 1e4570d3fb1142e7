#!/bin/bash
# Round-1 profile capture (run on the GPU box through gpurun).  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
# (1) launch list: per-launch durations of two decodes (after the bench's warm-up legs)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01b_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# (2) one full capture of each kernel of one decode
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 10 --launch-count 5 \
    -o gpurun_out/r01b_all5 -f python bench.py --steps 2 --warmup 2 --no-cpu-baseline >> gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
