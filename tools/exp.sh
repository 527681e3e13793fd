python tools/sweep.py --graph --n 60 poisson2d_param_256_b64 poisson2d_512_b16 ibn2d_512_b16 poisson2d_param_256_b16 poisson2d_param_256_b256 --cfg "DN_T2_FILL_PCT=100" --cfg "DN_T2_FILL_PCT=60" --cfg "DN_T2_FILL_PCT=80" 2>&1 | grep -v Warning
python -m pytest tests -m gpu -x -q 2>&1 | tail -1
python tools/fuzz_parity.py 60 11 2>&1 | tail -1
