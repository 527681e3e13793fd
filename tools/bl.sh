#!/bin/bash
# tools/bl.sh "<ENV=val ...>" <bench args...>  -- run bench.py under an env override, print value / ms / roofline frac
cfg="$1"; shift
out=$(env $cfg python bench.py --no-cpu "$@" 2>&1 | tail -1)
echo "$out" | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('[%s] %s: %.1f GDOF/s  %.4f ms  frac %.3f  e2e %.2f' % ('$cfg', d['config']['workload'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']))
except Exception as e:
    print('[$cfg] FAILED', e)
"
