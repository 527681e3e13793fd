#!/bin/bash
# tools/exp.sh -- scratch driver for ONE short GPU-box call (`gpurun -- 'bash tools/exp.sh'`): edit, run, read
# gpurun_out/.  The last use: the full gpu suite at the head of the tree.
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/exp_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/exp_pytest.log
