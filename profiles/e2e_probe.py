"""Where the host-to-host step goes: pinned H2D alone, the call without / with the kNN tables, one chunk (no overlap) vs the default."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from multi_orbslam3_b200 import orbx, synth
W, H, B = 752, 480, 512
base = synth.rects_stream(W, H, 64, seed=0)
frames = np.concatenate([base] * (B // 64))
hf = torch.from_numpy(frames).pin_memory(); hn = hf.numpy()
d = torch.empty_like(hf, device="cuda")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
print("H2D 185 MB alone: %.2f ms" % timeit(lambda: d.copy_(hf, non_blocking=True)))
ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
m = orbx.ORBmatcher(0.9, True, max_keypoints=ex.cap, max_batch=B)
cap = ex.cap
pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
out = {"kps": pin((B, cap, 7), torch.float32).view(np.uint8).reshape(B, cap, 28).view(orbx.KP_DTYPE).reshape(B, cap),
       "desc": pin((B, cap, 32), torch.uint8), "n": pin((B,), torch.int32), "mono": pin((B,), torch.int32),
       "matches12": pin((B, cap), torch.int32), "nmatches": pin((B,), torch.int32)}
full = dict(out); full["knn_idx"] = pin((B, cap, 2), torch.int32); full["knn_dist"] = pin((B, cap, 2), torch.int32)
bounds = (0.0, float(W), 0.0, float(H))
for tag, o in (("without kNN tables", out), ("with kNN tables", full)):
    for ch in ("1", "6"):
        os.environ["ORBX_HOST_CHUNKS"] = ch
        print("call %s, %s chunk(s): %.2f ms" % (tag, ch, timeit(lambda: orbx.extract_match_batch(ex, m, hn, (0, 0), bounds, 100, o), 10)))
d_m12 = torch.empty((B, m.K), dtype=torch.int32, device="cuda"); d_nm = torch.empty(B, dtype=torch.int32, device="cuda")
d_ki = torch.empty((B, m.K, 2), dtype=torch.int32, device="cuda"); d_kd = torch.empty_like(d_ki)
s = torch.cuda.current_stream().cuda_stream
print("device-resident step: %.2f ms" % timeit(lambda: orbx.extract_match_batch_device(ex, m, d.data_ptr(), B, W, H, W, W * H, (0, 0), bounds, 100, d_m12.data_ptr(), d_nm.data_ptr(), d_ki.data_ptr(), d_kd.data_ptr(), s)))
