#!/bin/bash
# round-1 evidence run: tests, bench line, reference arm, ncu launch list + full capture of the step kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --size 256 --steps 100 --warmup 5 > gpurun_out/bench256.json 2> gpurun_out/bench256.err; tail -3 gpurun_out/bench256.err; cat gpurun_out/bench256.json
python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench256.err; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/launches.csv python bench.py --size 256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_face_flux|k_cell_update|k_points' -s 9 -c 3 -o gpurun_out/prof_r1g python scripts/gpu_tune.py 256 2 0,0,-1,0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python __graft_entry__.py --smoke 2>&1 | tail -3
