#!/bin/bash
# tools/bl3.sh "<ENV=val ...>" <bench args...> -- like bl.sh but also prints the 3-D launch plan
cfg="$1"; shift
out=$(env DN_DEBUG_PLAN=1 $cfg python bench.py --no-cpu --train-steps 0 "$@" 2>&1)
echo "$out" | grep plan3t | tail -1 | cut -c1-200
echo "$out" | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('[%s] %s: %.1f GDOF/s  %.4f ms  frac %.3f' % ('$cfg', d['config']['workload'], d['value'], d['ms_per_step'], d['roofline']['frac']))
except Exception as e:
    print('[$cfg] FAILED', e)
"
