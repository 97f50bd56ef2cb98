"""Warm per-stage times of ONE frame through the host-buffer call (batch = 1) in the serial launch order: the extractor's event
profile (pyramid | FAST | octree | finalize+orient+describe; profiling turns the launch graph and the level-parallel order off)
and, by difference, what is left of the call (H2D, matcher kernels, D2H, synchronisation)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from multi_orbslam3_b200 import orbx, synth

W, H = 752, 480
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
frames = synth.rects_stream(W, H, 16, seed=0)
hf = torch.from_numpy(np.ascontiguousarray(frames)).pin_memory().numpy()
ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=1)
m = orbx.ORBmatcher(0.9, True, max_keypoints=ex.cap, max_batch=1)
cap = ex.cap
pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
out = {"kps": pin((1, cap, 7), torch.float32).view(np.uint8).reshape(1, cap, 28).view(orbx.KP_DTYPE).reshape(1, cap),
       "desc": pin((1, cap, 32), torch.uint8), "n": pin((1,), torch.int32), "mono": pin((1,), torch.int32),
       "matches12": pin((1, cap), torch.int32), "nmatches": pin((1,), torch.int32),
       "knn_idx": pin((1, cap, 2), torch.int32), "knn_dist": pin((1, cap, 2), torch.int32)}
bounds = (0.0, float(W), 0.0, float(H))
for i in range(20):
    orbx.extract_match_batch(ex, m, hf[i % 16:i % 16 + 1], (0, 0), bounds, 100, out)
ex.profile(1)
ts = []
for i in range(n):
    t = time.perf_counter(); orbx.extract_match_batch(ex, m, hf[i % 16:i % 16 + 1], (0, 0), bounds, 100, out); ts.append(time.perf_counter() - t)
ms, nb = ex.profile(0)
ts = np.array(ts) * 1e6
names = ["pyramid+blur (8 launches)", "FAST", "octree", "finalize+orient+describe"]
print("serial order, events between the stages, %d frames: call p50 %.1f us" % (nb, np.percentile(ts, 50)))
for nme, v in zip(names, ms):
    print("  %-28s %6.1f us" % (nme, v / nb * 1e3))
print("  %-28s %6.1f us" % ("rest (H2D, matcher, D2H, sync)", np.percentile(ts, 50) - sum(ms) / nb * 1e3))
