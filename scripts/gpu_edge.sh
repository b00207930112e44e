#!/bin/bash
for w in 8 4; do
QGD_ELL_MAXW=$w python scripts/gpu_dbg.py poly 0 1 2>&1 | tail -1
QGD_ELL_MAXW=$w python scripts/gpu_dbg.py poly 1 30 2>&1 | tail -1
QGD_ELL_MAXW=$w python scripts/gpu_dbg.py polyzg 0 30 2>&1 | tail -1
QGD_ELL_MAXW=$w python scripts/gpu_dbg.py impl 0 1 2>&1 | tail -1
QGD_ELL_MAXW=$w python scripts/gpu_dbg.py impl 0 30 2>&1 | tail -1
done
