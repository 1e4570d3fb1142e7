timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in lp4 lp5 lp6; do
  cp noaa-apt_b200/libaptb200_$v.so noaa-apt_b200/libaptb200.so
  echo "== $v"; timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*, "higher\|"all_kernels_ms": {[^}]*}'
done
cp noaa-apt_b200/libaptb200_lp4.so noaa-apt_b200/libaptb200.so
echo "== batch 16"; timeout 200 python bench.py --steps 10 --warmup 3 --batch 16 --no-cpu-baseline 2>&1 | grep -o '"value": [0-9.]*, "unit"' | head -1
