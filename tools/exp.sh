for lib in variants/lib_prev.so diffnet_b200/lib/libdiffnet_fem.so; do
  echo "=== lib $lib"
  DIFFNET_FEM_LIB=$PWD/$lib python tools/sweep.py --graph --n 30 poisson3d_param_64_b16 poisson3d_256_b1 2>&1 | grep -v Warning
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
