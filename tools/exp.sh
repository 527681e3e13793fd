#!/bin/bash
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_parity_3d.py -x -q -k "gp_eval or unfused" > $O/r2n_adj.log 2>&1; echo "adj rc=$?"; tail -8 $O/r2n_adj.log
python -m pytest tests/test_highorder.py tests/test_gpu_training.py -m gpu -x -q > $O/r2n_other.log 2>&1; echo "other rc=$?"; tail -3 $O/r2n_other.log
python tools/gp_probe.py 20 2>&1 | grep -v Warn | tee $O/r2n_gp_probe.txt
DN_GP_ADJ3=0 python tools/gp_probe.py 20 2>&1 | grep -v Warn | grep "3-D" | tee $O/r2n_gp_probe_old.txt
for zc in 8 16 32 64; do echo "ZC=$zc"; DN_GP_ADJ3_ZC=$zc python tools/gp_probe.py 20 2>&1 | grep "3-D" | grep -i adj; done
