#!/bin/bash
# quick iteration loop on one B200: extraction parity tests, then a short device-resident bench with per-stage times
python -m pytest tests/test_extract_gpu.py tests/test_ref_gpu.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 50 --no-e2e --no-cpu-baseline 2>/dev/null | python profiles/bench_short.py
