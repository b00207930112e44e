#!/bin/bash
# round-1 evidence run: tests, bench line, reference arm, QHD line, ncu launch list + full capture of the step kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --size 256 --steps 100 --warmup 5 > gpurun_out/bench256.json 2> gpurun_out/bench256.err; tail -3 gpurun_out/bench256.err; cat gpurun_out/bench256.json
python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench256.err; cat gpurun_out/bench_ref.json
python bench.py --case qhd2d --steps 20 --warmup 3 > gpurun_out/bench_qhd2d.json 2> gpurun_out/bench_qhd2d.err; tail -3 gpurun_out/bench_qhd2d.err; cat gpurun_out/bench_qhd2d.json
python bench.py --case qhd2d --size 256 --precond DIC --steps 5 --warmup 1 > gpurun_out/bench_qhd2d_dic256.json 2>> gpurun_out/bench_qhd2d.err; cat gpurun_out/bench_qhd2d_dic256.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --size 256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_face_flux|k_cell_update|k_points' -s 9 -c 3 -o gpurun_out/prof_r1f python scripts/gpu_tune.py 256 2 0,0,-1,0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:'k_pcg' -s 2 -c 1 -o gpurun_out/prof_r1f_pcg python bench.py --case qhd2d --steps 2 --warmup 3 > gpurun_out/ncu_pcg.log 2>&1
tail -2 gpurun_out/ncu_pcg.log
