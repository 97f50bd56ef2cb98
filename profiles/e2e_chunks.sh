#!/bin/bash
# e2e leg of the bench under different chunkings of the host pipeline (ORBX_HOST_CHUNKS, ORBX_UNIFORM_CHUNKS)
for c in 2 3 4 6 8; do
  for u in 0 1; do
    if [ $u = 1 ]; then export ORBX_UNIFORM_CHUNKS=1; else unset ORBX_UNIFORM_CHUNKS; fi
    ORBX_HOST_CHUNKS=$c python bench.py --steps 30 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunks $c uniform $u: e2e %.0f frames/s, value %.0f' % (d['e2e']['value'], d['value']))"
  done
done
