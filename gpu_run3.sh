#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python gpu_tune.py 256 1,2
QGD_FACE_TILE=0 python gpu_tune.py 256 1
QGD_FACE_TILE=256 python gpu_tune.py 256 1
