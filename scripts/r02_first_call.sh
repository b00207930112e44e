#!/bin/bash
# Round-2 opening evidence run on ONE B200 (about 5 GPU-minutes): full GPU suite incl. the cases written without GPU access,
# bench line + reference arm, ncu launch list, and the missing `ncu --set full` capture of the L2-policy face kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rxX 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --size 256 --steps 100 --warmup 5 > gpurun_out/bench256.json 2> gpurun_out/bench256.err; tail -3 gpurun_out/bench256.err; cat gpurun_out/bench256.json
python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench256.err; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/launches.csv python bench.py --size 256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_face_flux|k_cell_update|k_points' -s 9 -c 3 -o gpurun_out/prof_r2a python scripts/gpu_tune.py 256 2 0,0,-1,0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python __graft_entry__.py --smoke 2>&1 | tail -3
# extra lines: polyhedral mesh (configs[4] shape, ~2M cells) and the 2D / QHD cases
python bench.py --case poly --poly-n 100 --steps 50 --warmup 5 2>/dev/null | tee gpurun_out/bench_poly.json | cut -c1-300
