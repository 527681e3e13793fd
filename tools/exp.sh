python -m pytest tests/test_gpu_parity_2d.py -m gpu -x -q -k balanced 2>&1 | tail -3
python tools/fuzz_parity.py 240 31 2>&1 | tail -1
