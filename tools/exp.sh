#!/bin/bash
python -m pytest tests/test_gpu_parity_3d.py tests/test_gpu_parity_2d.py -x -q -k "gp_eval or unfused" 2>&1 | tail -2
python tools/gp_probe.py 20 2>&1 | grep "kernel"
