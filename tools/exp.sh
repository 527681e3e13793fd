for lib in diffnet_b200/lib/libdiffnet_fem.so variants/lib_unroll.so; do echo "== $lib"; DIFFNET_FEM_LIB=$PWD/$lib python tools/gp_probe.py 10 2>&1 | grep "kernel"; done
