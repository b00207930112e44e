#!/bin/bash
# Round-2 call 1 on ONE B200: full GPU suite (incl. slip / forward step / 64^3 / fields hand-off), bench line + reference arm,
# ncu launch list, ncu --set full of the three shipped step kernels, smoke, extra bench lines.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
python -m pytest tests -m gpu -q -rfE -x --timeout 1200 2>&1 | tail -25 > gpurun_out/r02a_pytest_gpu.log; cat gpurun_out/r02a_pytest_gpu.log
python bench.py --steps 100 --warmup 5 --e2e-full-state > gpurun_out/r02a_bench256.json 2> gpurun_out/r02a_bench256.err; tail -3 gpurun_out/r02a_bench256.err; cat gpurun_out/r02a_bench256.json
python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/r02a_bench_ref.json 2>> gpurun_out/r02a_bench256.err; cat gpurun_out/r02a_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_face_flux|k_cell_update|k_points' -s 9 -c 3 -o gpurun_out/prof_r2a python scripts/gpu_tune.py 256 2 0,0,-1,0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py --case qgd2d --steps 100 --warmup 5 2>/dev/null | tee gpurun_out/r02a_bench_qgd2d_step.json | cut -c1-400
python bench.py --case poly --poly-n 100 --steps 50 --warmup 5 2>/dev/null | tee gpurun_out/r02a_bench_poly.json | cut -c1-400
python bench.py --case qhd2d --steps 3 --warmup 3 2>/dev/null | tee gpurun_out/r02a_bench_qhd2d.json | cut -c1-400
