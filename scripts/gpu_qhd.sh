#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_qhd.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_qhd.log; cat gpurun_out/pytest_qhd.log
