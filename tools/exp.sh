python -m pytest tests/test_gpu_parity_3d.py -m gpu -x -q -k "grad_nu" 2>&1 | tail -8
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
