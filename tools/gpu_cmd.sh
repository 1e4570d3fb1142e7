bash tools/ncu_capture.sh
