# Round-end sanity on a GPU box: `gpurun --timeout 600 -- 'bash tools/gpu_cmd.sh'`
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1
