python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --workload poisson3d_256_slab --steps 200 --warmup 20 > gpurun_out/r2d_slab_n8.json 2> gpurun_out/r2d_slab_n8.err; echo "slab rc=$?"
tail -2 gpurun_out/r2d_slab_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d_slab_n8.json').read().strip().splitlines()[-1])
s=d['slab']
print({k:(round(v,4) if isinstance(v,float) else v) for k,v in s.items() if k not in ('desc','l2','roofline_per_gpu')})
PY
PROBE_CFGS=";DN_SLAB_DBG=7" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 tools/slab_link_probe.py 2>&1 | grep -v -i "warn\|OMP\|\*\*\*"
