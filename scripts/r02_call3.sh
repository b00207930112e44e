#!/bin/bash
# Round-2 call 3 on TWO B200s: whole GPU suite after the thermo / multi-GPU implicit / leastSquares / QHD additions
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rfE --timeout 1500 2>&1 | tail -60 > gpurun_out/r02c_pytest_gpu.log; cat gpurun_out/r02c_pytest_gpu.log
