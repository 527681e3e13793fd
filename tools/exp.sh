#!/bin/bash
O=gpurun_out; mkdir -p $O
python tools/sweep.py poisson3d_128_b1 --graph --n 40 \
  --cfg "DN_T3_STAGES=3" --cfg "DN_T3_STAGES=5" --cfg "DN_T3_STAGES=6" \
  --cfg "DN_T3_TY=6" --cfg "DN_T3_TY=5" --cfg "DN_T3_ZC=22" --cfg "DN_T3_ZC=26" --cfg "DN_T3_ZC=16" --cfg "DN_T3_ZC=32" \
  --cfg "DN_T3_LX=32" --cfg "DN_T3_LX=32 DN_T3_ZC=32" --cfg "DN_T3_LX=32 DN_T3_ZC=64" --cfg "DN_T3_LX=16 DN_T3_ZC=64" --cfg "DN_T3_THREADS=256" --cfg "DN_T3_THREADS=384" \
  --cfg "DN_T3_ISO=0" 2>&1 | grep -v Warn | tee $O/r2r_sweep128.txt
python tools/sweep.py poisson3d_param_64_b16 --graph --n 40 \
  --cfg "DN_T3_STAGES=3" --cfg "DN_T3_STAGES=6" --cfg "DN_T3_ZC=22" --cfg "DN_T3_ZC=16" --cfg "DN_T3_TY=8" --cfg "DN_T3_TY=11" --cfg "DN_T3_TY=13" --cfg "DN_T3_THREADS=320" --cfg "DN_T3_THREADS=416" 2>&1 | grep -v Warn | tee $O/r2r_sweep64.txt
