#!/bin/bash
# usage: var_sweep.sh "ENV=.. ENV=.." "..." : rebuild with each set of build-time macros and time the console
for v in "$@"; do
  echo "=== $v"
  env $v python -m diffmst_b200.build --force > /dev/null || { echo build failed; continue; }
  python bench.py --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('step ms %.4f' % d['ms_per_step'], {k: (round(v,4) if v else v) for k,v in d['roofline']['kernel_ms'].items()})"
done
python -m diffmst_b200.build --force > /dev/null
