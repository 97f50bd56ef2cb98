"""Latency of the host-array search calls the class layer makes per frame (ORBmatcher::SearchByProjection, SearchForInitialization,
BF kNN-2): ~1000 queries against the ~1000 keypoints of a frame, pageable numpy arrays in and out like std::vector / cv::Mat."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from multi_orbslam3_b200 import orbx, synth

W, H = 752, 480
fr = synth.rects_stream(W, H, 2, seed=3)
ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H)
(_, k1, d1), (_, k2, d2) = ex(fr[0]), ex(fr[1])
scale = 1.2 ** np.arange(8)
rng = np.random.default_rng(1)
q = np.zeros(len(k1), orbx.PROJQ_DTYPE)
q["u"] = k1["x"] + rng.normal(0, 2, len(k1)); q["v"] = k1["y"] + rng.normal(0, 2, len(k1)); q["r"] = 15.0 * scale[k1["octave"]]
q["minl"] = k1["octave"] - 1; q["maxl"] = k1["octave"] + 1; q["angle"] = k1["angle"]; q["valid"] = 1; q["ur"] = q["u"] - 10
m = orbx.ORBmatcher(0.9, True, max_keypoints=2048)
bounds = (0, W, 0, H)
pre = np.full(len(k2), -1, np.int32)
prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)


def timeit(fn, n=300):
    for _ in range(20):
        fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return np.percentile(np.array(ts) * 1e3, 50)


print("nq %d, n2 %d" % (len(k1), len(k2)))
print("SearchByProjection (mode 0, th 15)        p50 %.3f ms" % timeit(lambda: m.SearchByProjection(0, q, d1, k2, d2, bounds, pre.copy(), None)))
print("SearchByProjection (mode 1)               p50 %.3f ms" % timeit(lambda: m.SearchByProjection(1, q, d1, k2, d2, bounds, pre.copy(), None)))
print("SearchForInitialization (window 100)      p50 %.3f ms" % timeit(lambda: m.SearchForInitialization(k1, d1, k2, d2, bounds, prev.copy(), 100)))
print("bf_knn2 (1000 x 1000)                     p50 %.3f ms" % timeit(lambda: m.knnMatch2(d1, d2)))
# SearchByBoW(KeyFrame*, Frame&) (Tracking::TrackReferenceKeyFrame): feature vectors from an ORBvoc-shaped synthetic vocabulary
vocab = synth.random_vocabulary(k=10, L=5, seed=13)
V = orbx.ORBVocabulary(*vocab, L=5)
fv1 = V.transform(d1, 4)[1]; fv2 = V.transform(d2, 4)[1]
v1 = np.ones(len(k1), np.uint8)
mb = orbx.ORBmatcher(0.7, True, max_keypoints=2048)
print("SearchByBoW (KF, F)                       p50 %.3f ms" % timeit(lambda: mb.SearchByBoW(0, k1, d1, v1, fv1, k2, d2, None, fv2)))
print("BoW transform of a frame (1000 desc)      p50 %.3f ms" % timeit(lambda: V.transform(d1, 4)))
print("orbx_bow_transform alone (C call)         p50 %.3f ms" % timeit(lambda: V.transform_features(d1, 4)))
