#!/bin/bash
mkdir -p gpurun_out
python bench.py --size 256 --steps 100 --warmup 5 > gpurun_out/bench256.json 2> gpurun_out/bench256.err; tail -3 gpurun_out/bench256.err; cat gpurun_out/bench256.json
python bench.py --case qhd2d --steps 20 --warmup 3 > gpurun_out/bench_qhd2d.json 2> gpurun_out/bench_qhd2d.err; tail -3 gpurun_out/bench_qhd2d.err; cat gpurun_out/bench_qhd2d.json
python bench.py --case qhd2d --steps 20 --warmup 3 --p-tol 1e-6 --p-rel-tol 0.01 > gpurun_out/bench_qhd2d_rel.json 2>> gpurun_out/bench_qhd2d.err; cat gpurun_out/bench_qhd2d_rel.json
python bench.py --case qhd2d --qhd-size 128 --precond DIC --steps 5 --warmup 2 > gpurun_out/bench_qhd2d_dic128.json 2>> gpurun_out/bench_qhd2d.err; cat gpurun_out/bench_qhd2d_dic128.json
