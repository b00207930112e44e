#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python scripts/gpu_hint.py 256 50 > gpurun_out/hint.log 2>&1; cat gpurun_out/hint.log
