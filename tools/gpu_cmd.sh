bash tools/ncu_capture.sh
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 900 gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.json
timeout 400 python bench.py --workload c3 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 1200 gpurun_out/bench_c3.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
