#!/bin/bash
python -m pytest tests/test_gpu_load_vector.py tests/test_known_answers.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import torch
from diffnet_b200 import ops, DiffNet2DFEM, DiffNet3DFEM
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
for fem, shp in ((DiffNet2DFEM(None, domain_size=256, batch_size=64), (64, 4, 255, 255)), (DiffNet3DFEM(None, domain_size=64, batch_size=16), (16, 8, 63, 63, 63))):
    f = torch.randn(shp, device="cuda")
    out = ops.load_vector(fem.geometry, f)
    for _ in range(3): ops.load_vector(fem.geometry, f, out=out)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): ops.load_vector(fem.geometry, f, out=out)
    e1.record(); torch.cuda.synchronize(); print("assembly", shp, "%.1f us" % (e0.elapsed_time(e1) * 50))
PY
