#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --size 256 --steps 20 --warmup 3 > gpurun_out/bench256.json 2> gpurun_out/bench256.err; tail -3 gpurun_out/bench256.err; cat gpurun_out/bench256.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --size 256 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
ncu --set full --clock-control none --import-source on -k regex:'k_face_flux|k_cell_update|k_points' -s 9 -c 3 -o gpurun_out/prof_r1a python bench.py --size 256 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
