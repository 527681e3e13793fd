#!/bin/bash
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_parity_2d.py tests/test_gpu_parity_3d.py -x -q -k "gp_eval or unfused" > $O/r2o_adj.log 2>&1; echo "adj rc=$?"; tail -4 $O/r2o_adj.log
python tools/gp_probe.py 20 2>&1 | grep -v Warn | grep kernel | tee $O/r2o_gp_probe.txt
