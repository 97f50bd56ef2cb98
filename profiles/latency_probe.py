"""p50 latency of one frame through the host-buffer C-ABI call (batch = 1): H2D + extract + SearchForInitialization + BF kNN-2
against the previous frame + D2H.  Under `ncu --metrics gpu__time_duration.sum` it yields the per-kernel times of one frame."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from multi_orbslam3_b200 import orbx, synth

W, H = 752, 480
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
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
ts = []
for i in range(n):
    t = time.perf_counter(); orbx.extract_match_batch(ex, m, hf[i % 16:i % 16 + 1], (0, 0), bounds, 100, out); ts.append(time.perf_counter() - t)
ts = np.array(ts) * 1e3
print("batch-1 latency over %d frames: p50 %.3f ms  p95 %.3f ms  min %.3f ms  (n=%d kp, %d matches)" %
      (n, np.percentile(ts, 50), np.percentile(ts, 95), ts.min(), int(out["n"][0]), int(out["nmatches"][0])))

# ---- the extractor alone: ORBextractor::operator() of the class API = orbx_extract on one host image (pageable like a cv::Mat,
# then pinned), keypoints + descriptors back in host arrays
import ctypes as C
L = orbx.lib()
ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=1)
for tag, src in (("pageable image and results", np.ascontiguousarray(frames)), ("pinned image and results", hf)):
    if tag.startswith("pinned"):
        kps = out["kps"][0]; desc = out["desc"][0]
    else:
        kps = np.zeros(ex.cap, orbx.KP_DTYPE); desc = np.zeros((ex.cap, 32), np.uint8)
    nn, mono = C.c_int(0), C.c_int(0)
    ts = []
    for i in range(n + 20):
        img = src[i % 16]
        t = time.perf_counter()
        rc = L.orbx_extract(ex._h, img.ctypes.data_as(C.c_void_p), W, H, W, 0, 0, kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), ex.cap,
                            C.byref(nn), C.byref(mono))
        ts.append(time.perf_counter() - t)
        assert rc == 0
    ts = np.array(ts[20:]) * 1e3
    print("orbx_extract (operator()), one frame, %s: p50 %.3f ms  p95 %.3f ms  (n=%d kp)" % (tag, np.percentile(ts, 50), np.percentile(ts, 95), nn.value))

# ---- mvImagePyramid for the class API (ORBextractor::operator() leaves it on the host by default): 8 per-level downloads vs
# levels 1 .. 7 in one round trip (level 0 is the input, copied on the host)
bufs = [np.zeros((ex.level_size(l)[1], ex.level_size(l)[0]), np.uint8) for l in range(8)]
ptrs = (C.c_void_p * 7)(*[b.ctypes.data for b in bufs[1:]]); strides = (C.c_int * 7)(*[b.strides[0] for b in bufs[1:]])
t_old, t_new, t_stg = [], [], []
sp = (C.c_void_p * 7)(); ss = (C.c_int * 7)()
for i in range(220):
    L.orbx_extract(ex._h, hf[i % 16].ctypes.data_as(C.c_void_p), W, H, W, 0, 0, kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), ex.cap, C.byref(nn), C.byref(mono))
    t = time.perf_counter()
    for l in range(8):
        L.orbx_pyramid_to_host(ex._h, 0, l, bufs[l].ctypes.data_as(C.c_void_p), bufs[l].strides[0])
    t_old.append(time.perf_counter() - t)
    t = time.perf_counter()
    np.copyto(bufs[0], hf[i % 16])
    L.orbx_pyramid_levels_to_host(ex._h, 0, 1, 7, ptrs, strides)
    t_new.append(time.perf_counter() - t)
    t = time.perf_counter()
    np.copyto(bufs[0], hf[i % 16])
    L.orbx_pyramid_levels_staged(ex._h, 0, 1, 7, sp, ss)
    t_stg.append(time.perf_counter() - t)
print("pyramid to host after a frame: 8 per-level downloads p50 %.3f ms; level 0 from the input + levels 1..7 in one round trip p50 %.3f ms; "
      "the same as headers over the pinned staging (what the class layer does) p50 %.3f ms" %
      (np.percentile(np.array(t_old[20:]) * 1e3, 50), np.percentile(np.array(t_new[20:]) * 1e3, 50), np.percentile(np.array(t_stg[20:]) * 1e3, 50)))
