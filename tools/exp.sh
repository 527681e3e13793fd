#!/bin/bash
O=gpurun_out; mkdir -p $O
python tools/fuzz_parity.py 60 11 > $O/r2t_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -3 $O/r2t_fuzz.log
python tools/fuzz_parity.py 40 5 > $O/r2t_fuzz2.log 2>&1; echo "fuzz2 rc=$?"; tail -2 $O/r2t_fuzz2.log
python -m pytest tests/test_gpu_parity_3d.py -x -q -k "gp_eval or unfused" 2>&1 | tail -2
python tools/gp_probe.py 20 2>&1 | grep "kernel"
