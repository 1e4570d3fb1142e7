#!/bin/bash
# compute-sanitizer memcheck + racecheck over small decodes of every resampler path; logs under gpurun_out/.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --print-limit 20 python tools/sanitize_decode.py > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
timeout 900 $CS --tool racecheck --print-limit 20 python tools/sanitize_decode.py 2 > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
tail -4 gpurun_out/sanitizer_memcheck.log gpurun_out/sanitizer_racecheck.log
