#!/bin/bash
# scratch driver for one GPU-box call
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_load_vector.py -x -q > $O/r2l_lv.log 2>&1; echo "lv rc=$?"; tail -15 $O/r2l_lv.log
python -m pytest tests/test_known_answers.py tests/test_gpu_parity_2d.py tests/test_gpu_parity_3d.py -m gpu -x -q -k "gauss or fgp or KA3 or mms or more_gauss or known or golden" > $O/r2l_fgp.log 2>&1; echo "fgp rc=$?"; tail -5 $O/r2l_fgp.log
python - <<'PY'
import torch, time
from diffnet_b200 import ops, DiffNet2DFEM, DiffNet3DFEM
import bench
for name in ("mms2d_fgp_256_b64", "mms3d_fgp_64_b16"):
    for lv in (True, False):
        ops.USE_LOAD_VECTOR = lv
        r = bench.time_workload(name, torch.device("cuda:0"), 4321, 20, 5, 3)
        print(name, "load_vector" if lv else "general f_gp", "%.2f us" % (r["ms_step"] * 1e3))
        del r; torch.cuda.empty_cache()
# assembly cost
fem = DiffNet2DFEM(None, domain_size=256, batch_size=64)
f = torch.randn(64, 4, 255, 255, device="cuda")
for _ in range(3): ops.load_vector(fem.geometry, f)
torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.load_vector(fem.geometry, f)
e1.record(); torch.cuda.synchronize(); print("assembly 2d 256x64: %.1f us" % (e0.elapsed_time(e1)*100))
fem = DiffNet3DFEM(None, domain_size=64, batch_size=16)
f = torch.randn(16, 8, 63, 63, 63, device="cuda")
for _ in range(3): ops.load_vector(fem.geometry, f)
torch.cuda.synchronize(); e0.record()
for _ in range(10): ops.load_vector(fem.geometry, f)
e1.record(); torch.cuda.synchronize(); print("assembly 3d 64^3x16: %.1f us" % (e0.elapsed_time(e1)*100))
PY
