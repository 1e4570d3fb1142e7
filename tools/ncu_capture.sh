#!/bin/bash
# Profile capture (run on the GPU box through gpurun).  Numbers printed by runs under ncu are never bench values.
#   tools/ncu_capture.sh <tag> [kernel-regex]
TAG=${1:-r02}
KRE=${2:-'k_polyphase_ut|k_lowpass_records|k_resolve_roots|k_pick_cluster|k_pick_links|k_gather_rows_lp'}
mkdir -p gpurun_out
# (1) launch list: per-launch durations of three decodes
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/one_decode.py > gpurun_out/ncu_${TAG}.log 2>&1
# (2) one full capture of each kernel of the last decode
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${KRE}" --launch-skip 10 --launch-count 5 \
    -o gpurun_out/${TAG}_full -f python tools/one_decode.py >> gpurun_out/ncu_${TAG}.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
# (3) DRAM traffic in application order: caches are NOT flushed between kernels and the two counters fit one pass (no replay),
#     so e -- written by the resampler, read by the record and gather kernels -- is served from L2 as in a real decode
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none \
    -k "regex:${KRE}" --launch-skip 10 --launch-count 5 --csv --log-file gpurun_out/${TAG}_dram_warm.csv \
    python tools/one_decode.py >> gpurun_out/ncu_${TAG}.log 2>&1
tail -12 gpurun_out/${TAG}_dram_warm.csv
