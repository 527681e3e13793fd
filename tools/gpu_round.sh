#!/bin/bash
# tools/gpu_round.sh <tag> -- one GPU-box session: gpu tests, the default bench line, the ncu launch list of the
# same command and one `ncu --set full` capture of each dominant kernel (every ncu run follows a plain run of the
# same command that exited 0).  Outputs land in gpurun_out/<tag>_*.
tag=${1:-r2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${tag}_smi.txt 2>&1
python -m pytest tests -m gpu -x -q > $O/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${tag}_pytest.log
tail -3 $O/${tag}_pytest.log
python tools/fuzz_parity.py ${FUZZ_SECONDS:-60} 7 > $O/${tag}_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -3 $O/${tag}_fuzz.log
python bench.py > $O/${tag}_bench_n1.json 2> $O/${tag}_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > $O/${tag}_bench_ref.json 2>> $O/${tag}_bench_n1.err; echo "ref rc=$?"
python tools/show_bench.py $O/${tag}_bench_n1.json 2>&1 | head -40
python tools/gp_probe.py 20 2>&1 | grep -v Warn | tee $O/${tag}_gp_probe.txt
L="python bench.py --steps 10 --warmup 3 --reps 1 --no-cpu --train-steps 0 --no-points --no-graph"
$L > $O/${tag}_plain_launch.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${tag}_launches_2d.csv $L > $O/${tag}_ncu_launch.log 2>&1
echo "launch list rc=$?"
for w in ${NCU_WORKLOADS:-poisson2d_param_256_b64 poisson3d_128_b1 poisson3d_256_b1 poisson3d_param_64_b16}; do
  C="python tools/sweep.py $w --n 6"
  $C > $O/${tag}_plain_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_fem -s 8 -c 2 -f -o $O/${tag}_$w $C > $O/${tag}_ncu_$w.log 2>&1
  echo "ncu $w rc=$?"; tail -2 $O/${tag}_plain_$w.log
done
C="python tools/gp_probe.py 2"
$C > $O/${tag}_plain_gp.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_gp_eval -c 12 -f -o $O/${tag}_gp_eval $C > $O/${tag}_ncu_gp.log 2>&1
echo "ncu gp rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_gp_eval_adj -c 10 -f -o $O/${tag}_gp_eval_adj $C > $O/${tag}_ncu_gp_adj.log 2>&1
echo "ncu gp adj rc=$?"
