#!/bin/bash
# round-2 probe: does splitting the 3-D CTA into several smaller CTAs per SM (independent barriers) de-phase the warps?
cd "$(dirname "$0")/.."
for wl in poisson3d_256_b1 poisson3d_128_b1 poisson3d_param_64_b16; do
  for cfg in "DN_X=0" "DN_T3_THREADS=256" "DN_T3_THREADS=128" "DN_T3_THREADS=256 DN_T3_STAGES=4" \
             "DN_T3_THREADS=256 DN_T3_LX=16" "DN_T3_THREADS=256 DN_T3_LX=32" "DN_T3_THREADS=128 DN_T3_LX=16" \
             "DN_T3_THREADS=384" "DN_T3_THREADS=320" "DN_T3_THREADS=192" "DN_T3_STAGES=4" "DN_T3_STAGES=2"; do
    bash tools/bl3.sh "$cfg" --workload $wl --steps 50 --warmup 5 --reps 3
  done
done
