#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests/test_gpu_parity_3d.py tests/test_gpu_parity_2d.py -m gpu -x -q -k "randomised or gauss_points or more_gauss" 2>&1 | tail -2
