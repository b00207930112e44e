#!/bin/bash
mkdir -p gpurun_out
QGD_FACE_VARIANT=1 ncu --set full --clock-control none --import-source on -k regex:'k_face_flux|k_cell_update|k_points' -s 9 -c 3 -o gpurun_out/prof_r1b python gpu_tune.py 256 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
