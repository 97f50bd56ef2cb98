import sys, os
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from multi_orbslam3_b200 import orbx, synth
from oracle import oracle as O
B, W, H = 64, 640, 480
frames = synth.rects_stream(W, H, B, seed=92)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
b0 = pin(frames)
def run(prefetch):
    ex = orbx.ORBextractor(800, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    m = orbx.ORBmatcher(0.9, True, max_keypoints=ex.cap, max_batch=B)
    cap = ex.cap
    out = {"kps": np.zeros((B, cap), orbx.KP_DTYPE), "desc": np.zeros((B, cap, 32), np.uint8), "n": np.zeros(B, np.int32),
           "mono": np.zeros(B, np.int32), "matches12": np.zeros((B, cap), np.int32), "nmatches": np.zeros(B, np.int32),
           "knn_idx": np.zeros((B, cap, 2), np.int32), "knn_dist": np.zeros((B, cap, 2), np.int32)}
    if prefetch: orbx.extract_match_batch_prefetch(ex, m, b0)
    orbx.extract_match_batch(ex, m, b0, (0, 0), (0, W, 0, H), 100, out)
    ex.close(); m.close()
    return out
for pf in (False, True):
    o = run(pf)
    bad = []
    for f in range(1, B):
        npk, n = int(o["n"][f-1]), int(o["n"][f])
        ri, rd = O.bf_knn2(o["desc"][f-1, :npk], o["desc"][f, :n])
        if not (np.array_equal(o["knn_idx"][f, :npk], ri) and np.array_equal(o["knn_dist"][f, :npk], rd)):
            bad.append((f, int((o["knn_idx"][f, :npk] != ri).sum())))
    print("prefetch", pf, "frames with wrong kNN:", bad[:12], len(bad))
