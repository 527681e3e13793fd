for cfg in "DN_T2_BALANCE=0" ""; do
  echo "=== $cfg"
  env $cfg python tools/sweep.py --graph --n 50 poisson2d_param_256_b64 poisson2d_512_b16 ibn2d_512_b16 poisson2d_param_256_b16 2>&1 | grep -v Warning
done
python -m pytest tests/test_gpu_parity_2d.py -m gpu -x -q 2>&1 | tail -2
