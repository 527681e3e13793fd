python -m pytest tests/test_highorder.py -m gpu -x -q 2>&1 | tail -8
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --train-steps 0 --no-points > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1])
for k in ('e2e','e2e_compact_inputs','e2e_device_producers'):
    v=d.get(k); print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a!='note'} if v else v)
print('value', d['value'], d['roofline']['frac'])
PY
