#!/bin/bash
# Round-2 multi-GPU check (run with gpurun --gpus 2): parity of the decomposed run with the boundary kernels forked onto the side
# stream in multi-GPU mode (QGD_BND_FORK=2, not yet parity-tested), then an A/B of the 256^3 strong-scaling bench line.
mkdir -p gpurun_out
QGD_BND_FORK=2 timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
QGD_RUN_UNVERIFIED_MULTI=1 timeout 300 python -m pytest tests/test_zz_gpu_unverified.py -m gpu -q -rxX -k decomposed_pcg 2>&1 | tail -8
for v in 1 2; do
  QGD_BND_FORK=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
      bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_n2_fork$v.err | tail -1 > gpurun_out/bench_n2_fork$v.json
  cut -c1-260 gpurun_out/bench_n2_fork$v.json
done
