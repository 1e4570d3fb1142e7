timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b.json 2>gpurun_out/b.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["all_kernels_ms"])
PY
