python -m pytest tests -m gpu -x -q 2>&1 | tail -1
python tools/fuzz_parity.py 90 23 2>&1 | tail -1
python bench.py --no-cpu --train-steps 0 > gpurun_out/r2h_bench.json 2>/dev/null; python tools/show_bench.py gpurun_out/r2h_bench.json | head -13
