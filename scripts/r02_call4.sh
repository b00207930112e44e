#!/bin/bash
# Round-2 call 4 on ONE B200: suite after block-local DIC + compact gradient record; bench lines; ncu of the new face kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rfE --timeout 1500 2>&1 | tail -40 > gpurun_out/r02d_pytest_gpu.log; cat gpurun_out/r02d_pytest_gpu.log
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r02d_bench256.json 2> gpurun_out/r02d_bench256.err; tail -3 gpurun_out/r02d_bench256.err; cut -c1-1800 gpurun_out/r02d_bench256.json
for blk in 256 512 1024; do
python bench.py --case qhd2d --precond DIC --pcg-blocks $blk --steps 3 --warmup 3 2>/dev/null | tee gpurun_out/r02d_bench_qhd2d_dic$blk.json | cut -c1-1200
done
python bench.py --case qhd2d --precond diagonal --steps 3 --warmup 3 2>/dev/null | tee gpurun_out/r02d_bench_qhd2d_diag.json | cut -c1-1200
python bench.py --case qhd2d --precond DIC --pcg-blocks 512 --p-tol 1e-6 --p-rel-tol 0.01 --steps 5 --warmup 3 2>/dev/null | tee gpurun_out/r02d_bench_qhd2d_dic512_reltol.json | cut -c1-1200
ncu --set full --clock-control none --import-source on -k regex:'k_face_flux|k_cell_update|k_points' -s 9 -c 3 -o gpurun_out/prof_r2d python scripts/gpu_tune.py 256 2 0,0,-1,0 > gpurun_out/ncu_full_d.log 2>&1
tail -2 gpurun_out/ncu_full_d.log
python bench.py --case qgd2d --steps 100 --warmup 5 2>/dev/null | tee gpurun_out/r02d_bench_qgd2d_step.json | cut -c1-600
python bench.py --case poly --poly-n 100 --steps 50 --warmup 5 2>/dev/null | tee gpurun_out/r02d_bench_poly.json | cut -c1-600
