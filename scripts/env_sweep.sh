#!/bin/bash
# usage: env_sweep.sh "ENV=.. ENV=.." "..." : time the console under each set of run-time environment variables
for v in "$@"; do
  echo "=== $v"
  env $v python bench.py --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('step ms %.4f' % d['ms_per_step'], {k: (round(v,4) if v else v) for k,v in d['roofline']['kernel_ms'].items()})"
done
