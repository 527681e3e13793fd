for lib in "" variants/lib_o3.so variants/lib_o2.so; do
  echo "=== lib ${lib:-default}"
  DIFFNET_FEM_LIB=${lib:+$PWD/$lib} python tools/sweep.py --graph --n 20 --cfg "DN_T3_THREADS=256" --cfg "DN_T3_THREADS=384" poisson3d_256_b1 poisson3d_128_b1 poisson3d_param_64_b16 2>&1 | grep -v Warning
done
DIFFNET_FEM_LIB=$PWD/variants/lib_o3.so python -m pytest tests/test_gpu_parity_3d.py -m gpu -x -q 2>&1 | tail -3
