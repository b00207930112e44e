#!/bin/bash
# Round-2 call 2 on TWO B200s: the whole GPU suite (single-GPU tests not reached in call 1 + decomposed QGDFoam / QHDFoam / PCG
# at N=2), then the 2-GPU bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rfE --timeout 1500 2>&1 | tail -40 > gpurun_out/r02b_pytest_gpu.log; cat gpurun_out/r02b_pytest_gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus 2 --steps 100 --warmup 5 2> gpurun_out/r02b_bench_n2.err | tail -1 > gpurun_out/r02b_bench_n2.json
cut -c1-1500 gpurun_out/r02b_bench_n2.json; tail -3 gpurun_out/r02b_bench_n2.err
