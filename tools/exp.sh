python -m pytest tests/test_gpu_producers.py -m gpu -x -q 2>&1 | tail -15
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
