#!/bin/bash
# Round-2 call 6 on ONE B200: block-local DIC with a warp per block; QHD bench lines
mkdir -p gpurun_out
python -m pytest tests/test_gpu_qhd.py tests/test_gpu_extra.py -m gpu -q -rfE --timeout 1500 2>&1 | tail -30 > gpurun_out/r02e_pytest_gpu.log; cat gpurun_out/r02e_pytest_gpu.log
for blk in 64 128 512; do
python bench.py --case qhd2d --precond DIC --pcg-blocks $blk --steps 3 --warmup 3 2>/dev/null | tee gpurun_out/r02e_bench_qhd2d_dic$blk.json | cut -c1-1200
done
python bench.py --case qhd2d --precond DIC --pcg-blocks 128 --p-tol 1e-6 --p-rel-tol 0.01 --steps 5 --warmup 3 2>/dev/null | tee gpurun_out/r02e_bench_qhd2d_dic128_reltol.json | cut -c1-1200
python bench.py --case qhd2d --precond diagonal --p-tol 1e-6 --p-rel-tol 0.01 --steps 5 --warmup 3 2>/dev/null | tee gpurun_out/r02e_bench_qhd2d_diag_reltol.json | cut -c1-1200
