#!/bin/bash
# Rebuild the library with each tile geometry and time the console (tuning aid; run on the GPU box).
for cfg in "$@"; do
  echo "=== DMST_TRACK_CFG=$cfg"
  DMST_TRACK_CFG=$cfg python -m diffmst_b200.build --force > /dev/null || { echo build failed; continue; }
  python scripts/quick_time.py 2>&1 | grep -v Warning
  python bench.py --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('step ms', d['ms_per_step'], d['roofline']['kernel_ms'])"
done
python -m diffmst_b200.build --force > /dev/null
