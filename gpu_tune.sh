#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python gpu_tune.py 256 50 0,0,-1,0 > gpurun_out/tune.log 2>&1
QGD_FACE_GEOM=0 python gpu_tune.py 256 50 0,0,-1,0 >> gpurun_out/tune.log 2>&1
cat gpurun_out/tune.log
