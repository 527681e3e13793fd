time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 > gpurun_out/r2g_bench_n8.json 2> gpurun_out/r2g_bench_n8.err; echo "bench n8 rc=$?"
python tools/show_bench.py gpurun_out/r2g_bench_n8.json | head -30
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench_n8.json').read().strip().splitlines()[-1])
for k in ('e2e','e2e_compact_inputs','e2e_device_producers'):
    v=d.get(k); print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a!='note'} if v else v)
PY
