"""Condensed view of one bench.py JSON line (stdin): value, ms/step, e2e, latency and the per-stage times."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("value %.0f %s | %.3f ms/step | e2e %.0f | p50 %.3f ms" % (d["value"], d["unit"], d["ms_per_step"], (d.get("e2e") or {}).get("value") or 0, (d.get("latency") or {}).get("p50_ms") or 0))
for k, v in d["roofline"]["stages"].items():
    print("  %-32s %.3f ms" % (k[:32], v["ms_per_step"]))
