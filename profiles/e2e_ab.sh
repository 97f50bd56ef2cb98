#!/bin/bash
# A/B of the prefetched host pipeline's chunk schedules inside ONE box session (the host link is shared with the box's other tenants)
python profiles/h2d_probe.py
for rep in 1 2; do
  for cfg in "1 0" "3 1" "3 0" "4 0" "6 1"; do
    set -- $cfg
    if [ $2 = 1 ]; then export ORBX_UNIFORM_CHUNKS=1; else unset ORBX_UNIFORM_CHUNKS; fi
    ORBX_PREFETCH_CHUNKS=$1 python bench.py --no-cpu-baseline --steps 40 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rep $rep chunks $1 uniform $2: e2e %.0f frames/s' % d['e2e']['value'])"
  done
done
python profiles/h2d_probe.py
