#!/bin/bash
mkdir -p gpurun_out
QGD_FACE_TMA=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "100_steps and (hex_zg or adjust or prism_fixed or poly)" 2>&1 | tail -3
python scripts/gpu_tune.py 256 50 0,0,-1,0 > gpurun_out/tune.log 2>&1
QGD_FACE_TMA=3 python scripts/gpu_tune.py 256 50 0,0,-1,0 >> gpurun_out/tune.log 2>&1
cat gpurun_out/tune.log
