timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
echo "48k"; ITERS=8 timeout 100 python tools/time_front.py 2>&1 | tail -1
echo "48k compute-only"; APTB200_TILE_DEBUG=2 ITERS=8 timeout 100 python tools/time_front.py 2>&1 | tail -1
echo "96k"; RATE=96000 ITERS=8 timeout 100 python tools/time_front.py 2>&1 | tail -1
APTB200_TILE_PROFILE=1 timeout 100 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep "ut profile" | tail -1
