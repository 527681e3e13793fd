for lib in variants/lib_prev.so diffnet_b200/lib/libdiffnet_fem.so variants/lib_prev.so diffnet_b200/lib/libdiffnet_fem.so; do
  echo "=== lib $lib"
  DIFFNET_FEM_LIB=$PWD/$lib python tools/sweep.py --graph --n 80 poisson2d_param_256_b64 poisson2d_512_b16 poisson2d_64_b1 2>&1 | grep -v Warning
done
python -m pytest tests/test_gpu_parity_2d.py -m gpu -x -q 2>&1 | tail -1
