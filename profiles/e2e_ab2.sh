#!/bin/bash
# does the occasional half-speed e2e run come from stream -> hardware-queue aliasing?  same config, with and without more connections
for rep in 1 2 3 4 5 6; do
  for conn in 8 32; do
    CUDA_DEVICE_MAX_CONNECTIONS=$conn python bench.py --no-cpu-baseline --steps 40 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rep $rep connections $conn: e2e %.0f frames/s' % d['e2e']['value'])"
  done
done
