python - <<'PY'
import torch, time
from diffnet_b200 import DiffNet2DFEM, ops
dev="cuda:0"
fem=DiffNet2DFEM(None, domain_size=256)
u=torch.randn(64,1,256,256,device=dev,requires_grad=True)
out=ops.gp_eval(fem.geometry,u,"N")
cot=torch.randn_like(out)
for i in range(3):
    out=ops.gp_eval(fem.geometry,u,"N"); out.backward(cot)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
for i in range(3):
    out=ops.gp_eval(fem.geometry,u,"N")
    torch.cuda.synchronize(); t=time.time(); e0.record()
    out.backward(cot)
    e1.record(); torch.cuda.synchronize()
    print("backward: device %.1f us  wall %.1f us" % (e0.elapsed_time(e1)*1e3, (time.time()-t)*1e6))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    out=ops.gp_eval(fem.geometry,u,"N"); out.backward(cot); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8))
PY
