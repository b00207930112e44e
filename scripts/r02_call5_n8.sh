#!/bin/bash
# Round-2 call 5 on EIGHT B200s: decomposed-run parity at N=8 and N=4 (QGDFoam incl. implicit / leastSquares / polyhedra / slip / 64^3,
# QHDFoam, PCG incl. block-local DIC), full per-field logs under gpurun_out/ -> profiles/.  No bench here (the driver's SCALE run does that).
mkdir -p gpurun_out
run() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) "$@" 2>&1 | grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" | tail -8; }
run 8 tests/multi_gpu_worker.py
run 8 tests/multi_gpu_qhd_worker.py
run 8 tests/multi_gpu_pcg_worker.py
QGD_MULTI_CASES=hex64_mixed_procrule,truncoct_mixed_serialrule,perturbed_mixed_implicit_serialrule,uniform_adjust_procrule,2d_leastSquares_serialrule run 4 tests/multi_gpu_worker.py
run 4 tests/multi_gpu_qhd_worker.py
run 4 tests/multi_gpu_pcg_worker.py
ls -la gpurun_out/*.log | tail
