# scratch experiment driver (one gpurun call)
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for i in 1 2; do python -m pytest tests -m gpu -x -q 2>&1 | tail -1; done
python bench.py --no-cpu --train-steps 0 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r2b_bench.json | head -16
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
print('copy ref default:', d['roofline'].get('same_bytes_copy'), d['roofline'].get('frac_of_same_bytes_copy'))
for k,v in d['points'].items():
    if 'roofline' in v and 'same_bytes_copy' in v['roofline']:
        c=v['roofline']['same_bytes_copy']; print(k, 'copy %.1f us frac %.3f -> ours/copy %.3f'%(c['ms_per_launch']*1e3, c['frac_of_peak'], v['roofline']['frac_of_same_bytes_copy']))
PY
