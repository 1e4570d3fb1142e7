for cfg in "4 0 0" "2 20 6" "2 24 3" "2 22 5"; do set -- $cfg; echo "q $1 warps $2 spare $3"; APTB200_UT_Q=$1 APTB200_UT_WARPS=$2 APTB200_UT_SPARE=$3 ITERS=8 timeout 100 python tools/time_front.py 2>&1 | tail -1; done
echo "96k q=4(n/a) q=2 warps 12 spare 2"; RATE=96000 APTB200_UT_WARPS=12 APTB200_UT_SPARE=2 ITERS=8 timeout 100 python tools/time_front.py 2>&1 | tail -1
