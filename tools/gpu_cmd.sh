timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"value": [0-9.]*, "unit"\|"kernel_ms": [0-9.]*\|"frac": [0-9.]*' | head -4
echo "96k default"; RATE=96000 ITERS=8 timeout 100 python tools/time_front.py 2>&1 | tail -1
echo "96k q=2"; APTB200_UT_Q=2 RATE=96000 ITERS=8 timeout 100 python tools/time_front.py 2>&1 | tail -1
echo "96k ws"; APTB200_NO_UNIFORM_TAPS=1 RATE=96000 ITERS=8 timeout 100 python tools/time_front.py 2>&1 | tail -1
