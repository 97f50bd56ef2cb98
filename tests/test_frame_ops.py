"""Frame-level helpers.  Frame::UndistortKeyPoints (R/src/Frame.cc:721-754): the oracle is pinned against cv2.undistortPoints (OpenCV 4.13) on the
CPU; the CUDA path is compared with the oracle on the GPU (both bit for bit)."""
import numpy as np
import pytest

from oracle import oracle as O

EUROC_K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]], np.float32)
EUROC_D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], np.float32)
TUM_K = np.array([[517.306408, 0, 318.643040], [0, 516.469215, 255.313989], [0, 0, 1]], np.float32)
TUM_D = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32)


def keypoints(n, w, h, seed):
    rng = np.random.default_rng(seed)
    k = np.zeros(n, O.KP_DTYPE)
    k["x"] = rng.uniform(-20, w + 20, n).astype(np.float32); k["y"] = rng.uniform(-20, h + 20, n).astype(np.float32)
    k["size"] = 31; k["angle"] = rng.uniform(0, 360, n).astype(np.float32); k["response"] = rng.integers(7, 200, n)
    k["octave"] = rng.integers(0, 8, n); k["class_id"] = -1
    return k


@pytest.mark.parametrize("K,D,w,h", [(EUROC_K, EUROC_D, 752, 480), (TUM_K, TUM_D, 640, 480)], ids=["euroc_4coef", "tum_5coef"])
def test_oracle_pinned_to_cv2(K, D, w, h):
    cv2 = pytest.importorskip("cv2")
    kps = keypoints(20000, w, h, 1)
    P = K.copy(); P[0, 0] *= np.float32(0.9); P[0, 2] += np.float32(3.5)          # a new camera matrix different from K
    for Pm in (K, P):
        ref = cv2.undistortPoints(np.stack([kps["x"], kps["y"]], 1).reshape(-1, 1, 2), K, D, None, Pm).reshape(-1, 2)
        got = O.undistort_keypoints(kps, K, D, Pm)
        np.testing.assert_array_equal(got["x"], ref[:, 0]); np.testing.assert_array_equal(got["y"], ref[:, 1])
        for f in ("size", "angle", "response", "octave", "class_id"):
            np.testing.assert_array_equal(got[f], kps[f])
    same = O.undistort_keypoints(kps, K, np.zeros(4, np.float32), K)              # mDistCoef[0] == 0: mvKeysUn = mvKeys
    assert same.tobytes() == kps.tobytes()


@pytest.mark.gpu
def test_gpu_matches_oracle_host_and_slots():
    import torch
    from multi_orbslam3_b200 import orbx, synth
    m = orbx.ORBmatcher(0.9, True, max_keypoints=1024)
    for K, D, w, h in ((EUROC_K, EUROC_D, 752, 480), (TUM_K, TUM_D, 640, 480)):
        kps = keypoints(50000, w, h, 2)
        got = m.UndistortKeyPoints(kps, K, D, K)
        assert got.tobytes() == O.undistort_keypoints(kps, K, D, K).tobytes()
    assert m.UndistortKeyPoints(kps, K, np.zeros(5, np.float32), K).tobytes() == kps.tobytes()
    # on the result slots of an extractor (mvKeysUn of a batch stays on the device)
    W, H, B = 752, 480, 3
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    res = ex.extract_batch(synth.rects_stream(W, H, B, seed=4))
    d_un = torch.zeros((B, ex.cap, 7), dtype=torch.float32, device="cuda")
    Kf = np.ascontiguousarray(EUROC_K.reshape(9)); Df = np.ascontiguousarray(EUROC_D)
    orbx._check(orbx.lib().orbx_undistort_slots_device(ex._h, 0, B, orbx._p(Kf), orbx._p(Df), 4, orbx._p(Kf), d_un.data_ptr(), None))
    ex.sync()
    torch.cuda.synchronize()
    h_un = d_un.cpu().numpy().view(np.uint8).reshape(B, ex.cap, 28).view(orbx.KP_DTYPE).reshape(B, ex.cap)
    for i, (_, k, _) in enumerate(res):
        assert h_un[i, :len(k)].tobytes() == O.undistort_keypoints(k, EUROC_K, EUROC_D, EUROC_K).tobytes()
    ex.close(); m.close()


MSG_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "u1"), ("angle", "<f4"), ("response", "u1"), ("octave", "i1")])   # packed: 15 bytes


def test_kf_msg_keypoint_records_oracle():
    """KF.msg keypoint records (msg/CvKeyPoint.msg in ROS1 wire layout): the oracle against a packed numpy dtype."""
    assert MSG_DTYPE.itemsize == 15
    kps = keypoints(3000, 752, 480, 5)
    kps["size"] = np.float32(31) * np.float32(1.2) ** kps["octave"]
    msg = O.keypoints_to_msg(kps)
    rec = msg.reshape(-1).view(MSG_DTYPE)
    np.testing.assert_array_equal(rec["x"], kps["x"]); np.testing.assert_array_equal(rec["angle"], kps["angle"])
    np.testing.assert_array_equal(rec["size"], kps["size"].astype(np.int32).astype(np.uint8))
    np.testing.assert_array_equal(rec["response"], kps["response"].astype(np.int32).astype(np.uint8))
    np.testing.assert_array_equal(rec["octave"], kps["octave"].astype(np.int8))
    back = O.keypoints_from_msg(msg)
    np.testing.assert_array_equal(back["x"], kps["x"]); np.testing.assert_array_equal(back["y"], kps["y"])
    np.testing.assert_array_equal(back["size"], np.floor(kps["size"])); np.testing.assert_array_equal(back["response"], kps["response"])
    assert (back["class_id"] == -1).all() and (back["octave"] == kps["octave"]).all()


@pytest.mark.gpu
def test_kf_msg_keypoint_records_gpu():
    import torch
    from multi_orbslam3_b200 import orbx, synth
    m = orbx.ORBmatcher(0.9, True, max_keypoints=1024)
    kps = keypoints(5000, 752, 480, 6)
    kps["size"] = np.float32(31) * np.float32(1.2) ** kps["octave"]
    msg = m.KeyPointsToMsg(kps)
    np.testing.assert_array_equal(msg, O.keypoints_to_msg(kps))
    assert m.KeyPointsFromMsg(msg).tobytes() == O.keypoints_from_msg(msg).tobytes()
    W, H = 640, 480
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H)
    _, k, _ = ex(synth.rects_frame(W, H, 3), None, (0, 0))
    d = torch.zeros(ex.cap * 15, dtype=torch.uint8, device="cuda")
    orbx._check(orbx.lib().orbx_slot_keypoints_to_msg_device(ex._h, 0, d.data_ptr(), None))
    ex.sync(); torch.cuda.synchronize()
    np.testing.assert_array_equal(d.cpu().numpy()[:len(k) * 15].reshape(-1, 15), O.keypoints_to_msg(k))
    ex.close(); m.close()


@pytest.mark.gpu
def test_device_pipeline_with_undistorted_keypoints():
    """Distorted camera, everything device-resident: extract a batch -> undistort the result slots -> SearchForInitialization
    between consecutive frames on mvKeysUn (as Tracking does) == oracle on undistorted keypoints."""
    import torch
    from multi_orbslam3_b200 import orbx, synth
    W, H, B = 752, 480, 3
    frames = synth.rects_stream(W, H, B, seed=12)
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    m = orbx.ORBmatcher(0.9, True, max_keypoints=ex.cap, max_batch=B)
    d_frames = torch.from_numpy(frames).cuda()
    ex.extract_batch_device(d_frames.data_ptr(), B, W, H, W, W * H, (0, 0), 0, None)
    d_un = torch.zeros((B + 1, ex.cap, 7), dtype=torch.float32, device="cuda")
    Kf = np.ascontiguousarray(EUROC_K.reshape(9)); Df = np.ascontiguousarray(EUROC_D)
    orbx._check(orbx.lib().orbx_undistort_slots_device(ex._h, 0, B, orbx._p(Kf), orbx._p(Df), 4, orbx._p(Kf), d_un.data_ptr(), None))
    orbx._check(orbx.lib().orbx_matcher_set_slot_keypoints(m._h, d_un.data_ptr()))
    a = torch.arange(0, B - 1, dtype=torch.int32, device="cuda"); b = torch.arange(1, B, dtype=torch.int32, device="cuda")
    m12 = torch.empty((B - 1, m.K), dtype=torch.int32, device="cuda"); nm = torch.empty(B - 1, dtype=torch.int32, device="cuda")
    bounds = (-30.0, W + 30.0, -30.0, H + 30.0)                # ComputeImageBounds of the undistorted corners (host side)
    m.match_slots_device(ex, (a.data_ptr(), B - 1), (b.data_ptr(), B - 1), bounds, 100, m12.data_ptr(), nm.data_ptr())
    ex.sync(); m.sync(); torch.cuda.synchronize()
    res = ex.download(0, B)
    hm12, hnm = m12.cpu().numpy(), nm.cpu().numpy()
    for i in range(B - 1):
        k1 = O.undistort_keypoints(res[i][1], EUROC_K, EUROC_D, EUROC_K); k2 = O.undistort_keypoints(res[i + 1][1], EUROC_K, EUROC_D, EUROC_K)
        rn, rm12, _ = O.search_for_initialization(k1, res[i][2], k2, res[i + 1][2], bounds, np.stack([k1["x"], k1["y"]], 1), 100, 0.9, True)
        assert hnm[i] == rn and rn > 20
        np.testing.assert_array_equal(hm12[i, :len(k1)], rm12)
    orbx._check(orbx.lib().orbx_matcher_set_slot_keypoints(m._h, None))
    ex.close(); m.close()


@pytest.mark.gpu
def test_stream_pipeline_with_camera():
    """orbx_extract_match_batch with a distorted camera set on the matcher: every frame is matched against its predecessor on
    mvKeysUn (predecessor carried across calls), equal to the oracle on undistorted keypoints."""
    from multi_orbslam3_b200 import orbx, synth
    W, H = 752, 480
    frames = synth.rects_stream(W, H, 5, seed=14)
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=3)
    m = orbx.ORBmatcher(0.9, True, max_keypoints=ex.cap, max_batch=3)
    Kf = np.ascontiguousarray(EUROC_K.reshape(9)); Df = np.ascontiguousarray(EUROC_D)
    orbx._check(orbx.lib().orbx_matcher_set_camera(m._h, orbx._p(Kf), orbx._p(Df), 4, orbx._p(Kf)))
    bounds = (-30.0, W + 30.0, -30.0, H + 30.0)
    got = []
    for lo, hi in ((0, 3), (3, 5)):                      # two calls: the second one starts from the carried predecessor
        nb, cap = hi - lo, ex.cap
        out = {"kps": np.zeros((nb, cap), orbx.KP_DTYPE), "desc": np.zeros((nb, cap, 32), np.uint8), "n": np.zeros(nb, np.int32),
               "mono": np.zeros(nb, np.int32), "matches12": np.zeros((nb, cap), np.int32), "nmatches": np.zeros(nb, np.int32)}
        orbx.extract_match_batch(ex, m, np.ascontiguousarray(frames[lo:hi]), (0, 0), bounds, 100, out)
        for i in range(hi - lo):
            n = int(out["n"][i])
            got.append((out["kps"][i, :n].copy(), out["desc"][i, :n].copy(), out["matches12"][i].copy(), int(out["nmatches"][i])))
    ref = O.Extractor(1000, 1.2, 8, 20, 7)
    prev = None
    for i in range(5):
        _, k, d = ref(frames[i], (0, 0))
        assert got[i][0].tobytes() == k.tobytes()                              # the returned keypoints are mvKeys (distorted)
        ku = O.undistort_keypoints(k, EUROC_K, EUROC_D, EUROC_K)
        if prev is not None:
            rn, rm12, _ = O.search_for_initialization(prev[0], prev[1], ku, d, bounds, np.stack([prev[0]["x"], prev[0]["y"]], 1), 100, 0.9, True)
            assert got[i][3] == rn and rn > 20
            np.testing.assert_array_equal(got[i][2][:len(prev[0])], rm12)
        prev = (ku, d)
    ex.close(); m.close()
