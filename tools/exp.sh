# scratch experiment driver (one gpurun call)
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --no-cpu --train-steps 0 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r2c_bench.json | head -16
