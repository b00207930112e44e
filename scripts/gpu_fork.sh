#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python scripts/gpu_envsweep.py 256 50 QGD_BND_FORK 0,1,0,1 > gpurun_out/fork.log 2>&1; cat gpurun_out/fork.log
timeout 300 python scripts/gpu_envsweep.py 128 200 QGD_BND_FORK 0,1,0,1 > gpurun_out/fork128.log 2>&1; cat gpurun_out/fork128.log
