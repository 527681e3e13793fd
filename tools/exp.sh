#!/bin/bash
O=gpurun_out; mkdir -p $O
C="python tools/gp_probe.py 2"
timeout 100 ncu --set full --clock-control none --import-source on -k regex:k_gp_eval_adj3 -c 2 -f -o $O/r2z_gp_eval_adj3 $C > $O/r2z_ncu_gp_adj3.log 2>&1
echo "ncu adj3 rc=$?"
