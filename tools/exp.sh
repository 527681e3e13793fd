C="python tools/gp_probe.py 2"
$C > gpurun_out/r2k_plain_gp.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_gp_eval_adj -c 12 -f -o gpurun_out/r2k_gp_adj $C > gpurun_out/r2k_ncu_gp.log 2>&1; echo "ncu rc=$?"
