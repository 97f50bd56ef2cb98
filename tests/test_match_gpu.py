"""GPU parity tests of the matcher kernels (through the C ABI) vs the CPU oracle: bit-exact indices/distances."""
import os

import numpy as np
import pytest

from multi_orbslam3_b200 import orbx, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frames():
    """Two consecutive frames of a stream, extracted by the oracle (matcher tests do not depend on the GPU extractor)."""
    st = synth.rects_stream(752, 480, 2, seed=3)
    e = O.Extractor(1000, 1.2, 8, 20, 7)
    out = []
    for f in st:
        _, k, d = e(f, (0, 0))
        out.append((k, d))
    return out


def test_descriptor_distance_pairs():
    m = orbx.ORBmatcher(0.9, True, max_keypoints=4096)
    a = synth.random_descriptors(5000, 1); b = synth.random_descriptors(5000, 2, 0.3)
    a[0] = 0; b[0] = 255; b[1] = a[1]
    got = m.DescriptorDistance(a, b)
    ref = np.unpackbits(a ^ b, axis=1).sum(1)
    np.testing.assert_array_equal(got, ref)
    assert got[0] == 256 and got[1] == 0
    m.close()


@pytest.mark.parametrize("nq,nt", [(1, 1), (1, 2), (5, 3), (100, 257), (1000, 1000), (130, 70000), (3, 0)])
def test_bf_knn2_parity(nq, nt):
    m = orbx.ORBmatcher(0.7, True, max_keypoints=2048)
    q = synth.random_descriptors(nq, 10 + nq, 0.2)
    t = synth.random_descriptors(max(nt, 1), 20 + nt, 0.5)[:nt]
    if nt > 300:
        t[100] = q[0]; t[299] = q[0]; t[nt - 1] = q[0]      # exact ties -> lowest index first
    idx, dist = m.knnMatch2(q, t)
    ridx, rdist = O.bf_knn2(q, t)
    np.testing.assert_array_equal(idx, ridx)
    np.testing.assert_array_equal(dist, rdist)
    m.close()


def test_search_for_initialization_parity(frames):
    (k1, d1), (k2, d2) = frames
    m = orbx.ORBmatcher(0.9, True, max_keypoints=2048)
    prev = np.stack([k1["x"], k1["y"]], 1)
    for window, ratio, ori in ((100, 0.9, True), (10, 0.9, True), (100, 0.6, False), (300, 0.99, True)):
        m.mfNNratio, m.mbCheckOrientation = ratio, ori
        n, m12, p2 = m.SearchForInitialization(k1, d1, k2, d2, (0, 752, 0, 480), prev, window)
        rn, rm12, rp2 = O.search_for_initialization(k1, d1, k2, d2, (0, 752, 0, 480), prev, window, ratio, ori)
        assert n == rn
        np.testing.assert_array_equal(m12, rm12)
        np.testing.assert_array_equal(p2, rp2)
        # second call with the updated vbPrevMatched, as Tracking does on the next frame
        n2, m12b, _ = m.SearchForInitialization(k1, d1, k2, d2, (0, 752, 0, 480), p2, window)
        rn2, rm12b, _ = O.search_for_initialization(k1, d1, k2, d2, (0, 752, 0, 480), rp2, window, ratio, ori)
        assert n2 == rn2
        np.testing.assert_array_equal(m12b, rm12b)
    m.close()


def test_search_for_initialization_steal_back_and_ties():
    """Planted duplicates: several F1 keypoints compete for one F2 keypoint (ORBmatcher.cc:741, :760-767)."""
    rng = np.random.default_rng(4)
    n = 300
    k1 = np.zeros(n, O.KP_DTYPE); k2 = np.zeros(n, O.KP_DTYPE)
    k1["x"] = rng.uniform(20, 620, n).astype(np.float32); k1["y"] = rng.uniform(20, 460, n).astype(np.float32)
    k2["x"] = k1["x"] + 2; k2["y"] = k1["y"] + 1
    k1["angle"] = rng.uniform(0, 360, n).astype(np.float32); k2["angle"] = k1["angle"]
    k1["octave"][::7] = 1
    d2 = synth.random_descriptors(n, 8)
    d1 = d2.copy()
    for i in range(n):                      # a few flipped bits so that distances are small but distinct
        for b in rng.integers(0, 256, int(rng.integers(0, 12))):
            d1[i, b >> 3] ^= np.uint8(1 << (b & 7))
    d1[10] = d1[11] = d1[12] = d2[40]       # three queries want F2 keypoint 40 at distance 0
    k1["x"][10:13] = k2["x"][40]; k1["y"][10:13] = k2["y"][40]
    d2[50] = d2[51]; k2["x"][51] = k2["x"][50] + 1; k2["y"][51] = k2["y"][50]   # exact tie between candidates
    m = orbx.ORBmatcher(0.9, True, max_keypoints=1024)
    prev = np.stack([k1["x"], k1["y"]], 1)
    n_, m12, p2 = m.SearchForInitialization(k1, d1, k2, d2, (0, 640, 0, 480), prev, 60)
    rn, rm12, rp2 = O.search_for_initialization(k1, d1, k2, d2, (0, 640, 0, 480), prev, 60, 0.9, True)
    assert n_ == rn
    np.testing.assert_array_equal(m12, rm12)
    np.testing.assert_array_equal(p2, rp2)
    m.close()


def make_queries(k1, rng, th=15.0, scale=None):
    q = np.zeros(len(k1), O.PROJQ_DTYPE)
    q["u"] = k1["x"] + rng.normal(0, 2, len(k1)).astype(np.float32)
    q["v"] = k1["y"] + rng.normal(0, 2, len(k1)).astype(np.float32)
    q["r"] = (th * scale[k1["octave"]]).astype(np.float32)
    q["minl"] = k1["octave"] - 1; q["maxl"] = k1["octave"] + 1
    q["angle"] = k1["angle"]; q["valid"] = (rng.random(len(k1)) > 0.1).astype(np.int32)
    q["ur"] = q["u"] - 10
    return q


@pytest.mark.parametrize("mode", [0, 1])
def test_search_by_projection_parity(frames, mode):
    (k1, d1), (k2, d2) = frames
    rng = np.random.default_rng(12)
    scale = O.Extractor(1000, 1.2, 8, 20, 7).scale
    q = make_queries(k1, rng, 15.0 if mode == 0 else 4.0 * 3, scale)
    if mode == 1:
        q["minl"] = k1["octave"] - 1; q["maxl"] = k1["octave"]
    pre = np.full(len(k2), -1, np.int32); pre[::9] = 7777       # keypoints that already hold a map point
    uright = np.where(rng.random(len(k2)) > 0.5, k2["x"] - 10 + rng.normal(0, 8, len(k2)), -1).astype(np.float32)
    m = orbx.ORBmatcher(0.8, True, max_keypoints=2048)
    for ur in (None, uright):
        n, a = m.SearchByProjection(mode, q, d1, k2, d2, (0, 752, 0, 480), pre, ur)
        rn, ra = O.search_by_projection(mode, q, d1, k2, d2, (0, 752, 0, 480), pre, ur, 0.8, True)
        assert n == rn
        np.testing.assert_array_equal(a, ra)
    m.close()


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("seed", [5, 6, 7])
def test_search_by_projection_contention(mode, seed):
    """Many queries compete for few keypoints (ORBmatcher.cc:89-91 / :2045-2047): long displacement chains for the parallel fixed-point
    resolve (k_proj_resolve).  60 keypoints with near-identical descriptors (ties), 700 queries whose windows hold all of them,
    occupying, non-occupying (0-observation owners, valid bit 1) and invalid queries mixed, some keypoints occupied before the call.
    Also with ORBX-style tiny capacity headroom: max_keypoints just above the sizes."""
    rng = np.random.default_rng(seed)
    n2, nq = 60, 700
    k2 = np.zeros(n2, O.KP_DTYPE)
    k2["x"] = rng.uniform(300, 340, n2).astype(np.float32); k2["y"] = rng.uniform(200, 240, n2).astype(np.float32)
    k2["angle"] = rng.uniform(0, 360, n2).astype(np.float32); k2["octave"] = rng.integers(0, 3, n2)
    base = synth.random_descriptors(1, seed)[0]
    d2 = np.tile(base, (n2, 1))
    for i in range(n2):
        for b in rng.integers(0, 256, int(rng.integers(0, 6))):
            d2[i, b >> 3] ^= np.uint8(1 << (b & 7))
    q = np.zeros(nq, O.PROJQ_DTYPE)
    q["u"] = rng.uniform(310, 330, nq).astype(np.float32); q["v"] = rng.uniform(210, 230, nq).astype(np.float32)
    q["r"] = 60.0; q["minl"] = 0; q["maxl"] = 7 if mode == 0 else rng.integers(0, 3, nq)
    if mode == 1:
        q["minl"] = np.maximum(q["maxl"] - 1, 0)
    q["angle"] = rng.uniform(0, 360, nq).astype(np.float32); q["ur"] = q["u"] - 10
    q["valid"] = rng.choice([0, 1, 1, 1, 3, 3], nq).astype(np.int32)
    qd = np.tile(base, (nq, 1))
    for i in range(nq):
        for b in rng.integers(0, 256, int(rng.integers(0, 8))):
            qd[i, b >> 3] ^= np.uint8(1 << (b & 7))
    pre = np.full(n2, -1, np.int32); pre[::11] = 4242
    for K in (1024, 768):
        m = orbx.ORBmatcher(0.8, True, max_keypoints=K)
        n, a = m.SearchByProjection(mode, q, qd, k2, d2, (0, 752, 0, 480), pre, None)
        rn, ra = O.search_by_projection(mode, q, qd, k2, d2, (0, 752, 0, 480), pre, None, 0.8, True)
        assert n == rn and n > 0
        np.testing.assert_array_equal(a, ra)
        m.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_search_by_projection_two_camera_parity(mode):
    """Two-camera frame (Frame::Nleft != -1; R/src/ORBmatcher.cc:144-213, :2093-2160): left + right halves with their own grids, the
    queries of a point interleaved over one occupancy table, stereo partners (mode 1), 0-observation owners, one rotation histogram
    (mode 0).  Oracle: orc_search_by_projection_rig, a literal restatement of the reference's loops."""
    rng = np.random.default_rng(31 + mode)
    e = O.Extractor(800, 1.2, 8, 20, 7)
    L, R = synth.stereo_pair(752, 480, seed=5, disparity=9)
    _, kl, dl = e(L, (0, 0)); _, kr, dr = e(R, (0, 0))
    nL, nR = len(kl), len(kr)
    k2 = np.concatenate([kl, kr]); d2 = np.concatenate([dl, dr])
    # map points = the left keypoints of a slightly shifted view; descriptors a few bits away
    nq = nL
    qd = dl.copy()
    for i in range(nq):
        for b in rng.integers(0, 256, int(rng.integers(0, 20))):
            qd[i, b >> 3] ^= np.uint8(1 << (b & 7))
    ql = make_queries(kl, rng, 15.0 if mode == 0 else 12.0, e.scale)
    qr = make_queries(kl, rng, 15.0 if mode == 0 else 12.0, e.scale)
    qr["u"] -= 9                                               # the right camera sees the point 9 px to the left
    if mode == 1:
        for q in (ql, qr):
            q["minl"] = kl["octave"] - 1; q["maxl"] = kl["octave"]
    noobs = rng.random(nq) < 0.3                               # temporal points without observations: they do not occupy what they take
    ql["valid"] |= np.where(noobs, 2, 0).astype(np.int32); qr["valid"] |= np.where(noobs, 2, 0).astype(np.int32)
    # stereo partners from a brute-force match of the two halves (as Frame::ComputeStereoFishEyeMatches fills them)
    bi, bd = O.bf_knn2(dl, dr)
    l2r = np.where(bd[:, 0] < 40, bi[:, 0], -1).astype(np.int32)
    r2l = np.full(nR, -1, np.int32)
    for i in np.nonzero(l2r >= 0)[0]:
        r2l[l2r[i]] = i
    pre = np.full(nL + nR, -1, np.int32); pre[::11] = 4242
    m = orbx.ORBmatcher(0.8, True, max_keypoints=4096, max_batch=2)
    for partners in ((l2r, r2l), (None, None)):
        n, a = m.SearchByProjectionRig(mode, ql, qr, qd, k2, d2, nL, (0, 752, 0, 480), pre, partners[0], partners[1], 100)
        rn, ra = O.search_by_projection_rig(mode, ql, qr, qd, k2, d2, nL, (0, 752, 0, 480), pre, partners[0], partners[1], 0.8, True, 100)
        assert n == rn and n > 200
        np.testing.assert_array_equal(a, ra)
        assert (a[nL:] >= 0).sum() > 50 and (a[:nL] >= 0).sum() > 50      # both cameras took part
    m.close()


def test_stereo_band_match_parity():
    L, R = synth.stereo_pair(752, 480, seed=2, disparity=14)
    e = O.Extractor(1200, 1.2, 8, 20, 7)
    _, kl, dl = e(L, (0, 0)); _, kr, dr = e(R, (0, 0))
    m = orbx.ORBmatcher(0.9, True, max_keypoints=2048)
    bf, b = 47.9, 0.11
    bi, bd = m.StereoBandMatch(kl, dl, kr, dr, e.scale, 480, 0.0, bf / b * 0.1)
    ri, rd = O.stereo_band_match(kl, dl, kr, dr, e.scale, 480, 0.0, bf / b * 0.1)
    np.testing.assert_array_equal(bi, ri)
    np.testing.assert_array_equal(bd, rd)
    assert (bi >= 0).sum() > 100
    m.close()


def test_slots_pipeline_matches_oracle():
    """Device-resident path used by bench.py: extract a batch, match consecutive slots, compare with the oracle."""
    import torch
    B = 5
    frames = synth.rects_stream(640, 480, B, seed=77)
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=B)
    m = orbx.ORBmatcher(0.9, True, max_keypoints=ex.cap, max_batch=B)
    d_frames = torch.from_numpy(frames).cuda()
    stream = torch.cuda.current_stream().cuda_stream   # 0 = legacy default stream; the wrapper maps it to cudaStreamLegacy
    ex.extract_batch_device(d_frames.data_ptr(), B, 640, 480, 640, 640 * 480, (0, 0), first_slot=1, stream=stream)
    ex.copy_slot(B, 0, stream)        # slot 0 = predecessor of the batch's first frame (here: its last frame)
    a = torch.arange(0, B, dtype=torch.int32, device="cuda"); b = torch.arange(1, B + 1, dtype=torch.int32, device="cuda")
    K = m.K
    m12 = torch.full((B, K), -2, dtype=torch.int32, device="cuda"); nm = torch.zeros(B, dtype=torch.int32, device="cuda")
    kidx = torch.full((B, K, 2), -2, dtype=torch.int32, device="cuda"); kdist = torch.full((B, K, 2), -2, dtype=torch.int32, device="cuda")
    m.match_slots_device(ex, (a.data_ptr(), B), (b.data_ptr(), B), (0, 640, 0, 480), 100, m12.data_ptr(), nm.data_ptr(),
                         kidx.data_ptr(), kdist.data_ptr(), stream)
    ex.sync(stream); m.sync(stream)
    res = ex.download(0, B + 1, stream)
    ref = O.Extractor(1000, 1.2, 8, 20, 7)
    for p in range(B):
        (_, k1, d1), (_, k2, d2) = res[p], res[p + 1]
        if p > 0:
            _, rk, rd = ref(frames[p - 1], (0, 0))
            assert k1.tobytes()[:0] == b"" and len(k1) == len(rk)
        prev = np.stack([k1["x"], k1["y"]], 1)
        rn, rm12, _ = O.search_for_initialization(k1, d1, k2, d2, (0, 640, 0, 480), prev, 100, 0.9, True)
        assert int(nm[p]) == rn
        np.testing.assert_array_equal(m12[p, :len(k1)].cpu().numpy(), rm12)
        ridx, rdist = O.bf_knn2(d1, d2)
        np.testing.assert_array_equal(kidx[p, :len(k1)].cpu().numpy(), ridx)
        np.testing.assert_array_equal(kdist[p, :len(k1)].cpu().numpy(), rdist)
    ex.close(); m.close()


def test_popc_probe_runs():
    p, l = orbx.popc_peak(0)
    assert p > 1e11 and l > 1e11


def test_extract_match_batch_host_pipeline():
    """orbx_extract_match_batch (the e2e call): chunked H2D | kernels | D2H pipeline, predecessor carried across calls."""
    import torch
    B, W, H = 70, 640, 480
    frames = synth.rects_stream(W, H, 2 * B, seed=91)
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    m = orbx.ORBmatcher(0.9, True, max_keypoints=ex.cap, max_batch=B)
    cap = ex.cap
    ref = O.Extractor(1000, 1.2, 8, 20, 7)
    prev = None
    for call, pinned in ((0, True), (1, False)):
        chunk = np.ascontiguousarray(frames[call * B:(call + 1) * B])
        if pinned:
            hold = torch.from_numpy(chunk).pin_memory(); chunk = hold.numpy()
            mk = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
            out = {"kps": mk((B, cap, 7), torch.float32).view(np.uint8).reshape(B, cap, 28).view(orbx.KP_DTYPE).reshape(B, cap),
                   "desc": mk((B, cap, 32), torch.uint8), "n": mk((B,), torch.int32), "mono": mk((B,), torch.int32),
                   "matches12": mk((B, cap), torch.int32), "nmatches": mk((B,), torch.int32),
                   "knn_idx": mk((B, cap, 2), torch.int32), "knn_dist": mk((B, cap, 2), torch.int32)}
        else:
            out = {"kps": np.zeros((B, cap), orbx.KP_DTYPE), "desc": np.zeros((B, cap, 32), np.uint8), "n": np.zeros(B, np.int32),
                   "mono": np.zeros(B, np.int32), "matches12": np.zeros((B, cap), np.int32), "nmatches": np.zeros(B, np.int32),
                   "knn_idx": np.zeros((B, cap, 2), np.int32), "knn_dist": np.zeros((B, cap, 2), np.int32)}
        orbx.extract_match_batch(ex, m, chunk, (0, 0), (0, W, 0, H), 100, out)
        for f in (0, 1, 17, 18, 35, 36, B - 1):
            n = int(out["n"][f])
            rmono, rk, rd = ref(chunk[f], (0, 0))
            assert n == len(rk) and int(out["mono"][f]) == rmono
            k = out["kps"][f, :n]; d = out["desc"][f, :n]
            for name in ("x", "y", "size", "response", "octave"):
                np.testing.assert_array_equal(k[name], rk[name])
            same = k["angle"] == rk["angle"]
            np.testing.assert_array_equal(d[same], rd[same])
            # predecessor: previous frame of the stream (previous call's last frame for f == 0)
            if f > 0:
                _, pk, pd = ref(chunk[f - 1], (0, 0))
            elif prev is not None:
                pk, pd = prev
            else:
                continue
            rn, rm12, _ = O.search_for_initialization(pk, pd, rk, rd, (0, W, 0, H), np.stack([pk["x"], pk["y"]], 1), 100, 0.9, True)
            assert int(out["nmatches"][f]) == rn
            np.testing.assert_array_equal(out["matches12"][f, :len(pk)], rm12)
            # BF kNN-2 of the pair (the other half of the C1 "match"): rows = the predecessor's keypoints
            ridx, rdist = O.bf_knn2(pd, rd)
            np.testing.assert_array_equal(out["knn_idx"][f, :len(pk)], ridx)
            np.testing.assert_array_equal(out["knn_dist"][f, :len(pk)], rdist)
        _, lk, ld = ref(chunk[B - 1], (0, 0))
        prev = (lk, ld)
    ex.close(); m.close()


def test_extract_match_batch_prefetch_equals_plain_call():
    """orbx_extract_match_batch_prefetch: the input of the next call travels on its own stream into a second staging buffer; the
    results (and the predecessor carried from call to call) must be those of the plain calls, whatever is prefetched when."""
    import torch
    B, W, H = 64, 640, 480
    frames = synth.rects_stream(W, H, 3 * B, seed=92)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    batches = [pin(frames[i * B:(i + 1) * B]) for i in range(3)]

    def run(prefetch):
        ex = orbx.ORBextractor(800, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
        m = orbx.ORBmatcher(0.9, True, max_keypoints=ex.cap, max_batch=B)
        cap = ex.cap
        outs = []
        if prefetch:
            orbx.extract_match_batch_prefetch(ex, m, batches[0])
        for i, b in enumerate(batches):
            if prefetch and i + 1 < len(batches):
                orbx.extract_match_batch_prefetch(ex, m, batches[i + 1])          # two batches wait at this moment
            out = {"kps": np.zeros((B, cap), orbx.KP_DTYPE), "desc": np.zeros((B, cap, 32), np.uint8), "n": np.zeros(B, np.int32),
                   "mono": np.zeros(B, np.int32), "matches12": np.zeros((B, cap), np.int32), "nmatches": np.zeros(B, np.int32),
                   "knn_idx": np.zeros((B, cap, 2), np.int32), "knn_dist": np.zeros((B, cap, 2), np.int32)}
            orbx.extract_match_batch(ex, m, b, (0, 0), (0, W, 0, H), 100, out)
            outs.append(out)
        if prefetch:                                   # a stale prefetch is dropped by a call with another buffer
            orbx.extract_match_batch_prefetch(ex, m, batches[0])
            out = {k: np.zeros_like(v) for k, v in outs[0].items()}
            orbx.extract_match_batch(ex, m, batches[1], (0, 0), (0, W, 0, H), 100, out)
            outs.append(out)
            with pytest.raises(orbx.OrbxError):        # pageable frames cannot be prefetched
                orbx.extract_match_batch_prefetch(ex, m, np.array(batches[0]))
        ex.close(); m.close()
        return outs

    plain, pre = run(False), run(True)
    for ci, (a, b) in enumerate(zip(plain, pre[:3])):
        for k in ("n", "mono", "nmatches"):
            np.testing.assert_array_equal(a[k], b[k], err_msg=k)
        for f in range(B):                              # rows are defined up to the frame's / its predecessor's keypoint count
            n = int(a["n"][f])
            assert a["kps"][f, :n].tobytes() == b["kps"][f, :n].tobytes()
            np.testing.assert_array_equal(a["desc"][f, :n], b["desc"][f, :n])
            if f == 0 and ci == 0:
                continue                                # the very first frame has no predecessor
            npk = int(a["n"][f - 1]) if f else int(plain[ci - 1]["n"][B - 1])
            for k in ("matches12", "knn_idx", "knn_dist"):
                np.testing.assert_array_equal(a[k][f, :npk], b[k][f, :npk], err_msg="%s call %d frame %d" % (k, ci, f))
    # the fourth call of the prefetch run = batch 1 after batch 2: only its own frames matter for extraction
    np.testing.assert_array_equal(pre[3]["n"], plain[1]["n"]); np.testing.assert_array_equal(pre[3]["desc"], plain[1]["desc"])
    for f in range(1, B):                               # rows are defined for the predecessor's keypoints
        npk = int(plain[1]["n"][f - 1])
        np.testing.assert_array_equal(pre[3]["matches12"][f, :npk], plain[1]["matches12"][f, :npk])


@pytest.mark.parametrize("shape,nf,disp", [((752, 480), 1200, 14), ((1241, 376), 2000, 23)], ids=["C2_euroc", "C3_kitti"])
def test_compute_stereo_matches_full(shape, nf, disp):
    """BASELINE configs C2 / C3: stereo pair, extraction on both cameras + Frame::ComputeStereoMatches (SAD refinement,
    sub-pixel fit, outlier cut) entirely on the device, against the oracle: mvuRight / mvDepth bit-exact."""
    W, H = shape
    L, R = synth.stereo_pair(W, H, seed=8, disparity=disp)
    exl = orbx.ORBextractor(nf, 1.2, 8, 20, 7, max_width=W, max_height=H)
    exr = orbx.ORBextractor(nf, 1.2, 8, 20, 7, max_width=W, max_height=H)
    gl = exl(L, None, (0, 0)); gr = exr(R, None, (0, 0))
    m = orbx.ORBmatcher(0.9, True, max_keypoints=max(exl.cap, exr.cap))
    mb, mbf = 0.11, 47.9
    ur, dp, sd = m.ComputeStereoMatches(exl, exr, mb, mbf)
    ol, orr = O.Extractor(nf, 1.2, 8, 20, 7), O.Extractor(nf, 1.2, 8, 20, 7)
    _, kl, dl = ol(L, (0, 0)); _, kr, dr = orr(R, (0, 0))
    rur, rdp, rsd = O.compute_stereo_matches(ol, orr, kl, dl, kr, dr, mb, mbf)
    assert len(ur) == len(rur)
    np.testing.assert_array_equal(sd, rsd)
    np.testing.assert_array_equal(ur, rur)          # bit-exact fp32
    np.testing.assert_array_equal(dp, rdp)
    ok = rur >= 0
    assert ok.sum() > 0.4 * len(rur)
    assert abs(np.median(kl["x"][ok] - rur[ok]) - disp) < 0.5
    exl.close(); exr.close(); m.close()


def test_stereo_matches_batched_stream():
    """C3 as a stream: a batch of KITTI-shape stereo pairs, both cameras extracted in batch, ComputeStereoMatches for all
    pairs in three launches; every pair equals the oracle's per-pair result."""
    W, H, nf, B = 1241, 376, 2000, 4
    pairs = [synth.stereo_pair(W, H, seed=20 + i, disparity=15 + 3 * i) for i in range(B)]
    Ls = np.stack([p[0] for p in pairs]); Rs = np.stack([p[1] for p in pairs])
    exl = orbx.ORBextractor(nf, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    exr = orbx.ORBextractor(nf, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    gl = exl.extract_batch(Ls); gr = exr.extract_batch(Rs)
    m = orbx.ORBmatcher(0.9, True, max_keypoints=exl.cap)
    mb, mbf = 0.54, 386.1
    ur, dp = m.ComputeStereoMatchesBatch(exl, exr, mb, mbf, 0, B)
    for i in range(B):
        ol, orr = O.Extractor(nf, 1.2, 8, 20, 7), O.Extractor(nf, 1.2, 8, 20, 7)
        _, kl, dl = ol(Ls[i], (0, 0)); _, kr, dr = orr(Rs[i], (0, 0))
        rur, rdp, _ = O.compute_stereo_matches(ol, orr, kl, dl, kr, dr, mb, mbf)
        n = len(rur)
        assert n == len(gl[i][1])
        np.testing.assert_array_equal(ur[i, :n], rur)
        np.testing.assert_array_equal(dp[i, :n], rdp)
        assert (rur >= 0).sum() > 0.3 * n
    # single-pair entry point on a slot of the batch agrees with the batched one
    u1, d1, _ = m.ComputeStereoMatches(exl, exr, mb, mbf, slot_l=2, slot_r=2, frame_l=2, frame_r=2)
    np.testing.assert_array_equal(u1, ur[2, :len(u1)])
    exl.close(); exr.close(); m.close()


def test_match_candidates_generic():
    """orbx_match_candidates: the primitive behind SearchByBoW / SearchForTriangulation / Fuse (explicit candidate lists)."""
    rng = np.random.default_rng(3)
    q = synth.random_descriptors(700, 31, 0.2); t = synth.random_descriptors(3000, 32, 0.4)
    counts = rng.integers(0, 90, 700); counts[5] = 0; counts[6] = 1; counts[7] = 2500
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    ind = rng.integers(0, 3000, off[-1]).astype(np.int32)
    t[ind[off[10]]] = q[10]; t[ind[off[10] + 3]] = q[10]          # exact ties: first in list order wins
    m = orbx.ORBmatcher(0.9, True, max_keypoints=1024)
    gi, gd = m.MatchCandidates(q, t, off, ind)
    ri, rd = O.match_candidates(q, t, off, ind)
    np.testing.assert_array_equal(gi, ri)
    np.testing.assert_array_equal(gd, rd)
    assert list(gi[5]) == [-1, -1] and gi[6, 1] == -1
    bad = ind.copy(); bad[3] = 3000
    with pytest.raises(orbx.OrbxError):
        m.MatchCandidates(q, t, off, bad)              # out-of-range candidate index is refused, not read
    m.close()


def test_extract_stereo_batch_host_pipeline():
    """orbx_extract_stereo_batch (host frames of both cameras in, keypoints + mvuRight + mvDepth out, chunked over five
    streams) equals the step-by-step path and the oracle, with pageable and with pinned caller buffers."""
    import torch
    W, H, nf, B = 752, 480, 1200, 40                       # 40 pairs -> 4 chunks of 10
    pairs = [synth.stereo_pair(W, H, seed=60 + i, disparity=12 + (i % 5) * 4) for i in range(6)]
    Ls = np.stack([pairs[i % 6][0] for i in range(B)]); Rs = np.stack([pairs[i % 6][1] for i in range(B)])
    exl = orbx.ORBextractor(nf, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    exr = orbx.ORBextractor(nf, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    m = orbx.ORBmatcher(0.9, True, max_keypoints=exl.cap)
    mb, mbf = 0.11, 47.9
    out = m.ExtractStereoBatch(exl, exr, Ls, Rs, mb, mbf)
    cap = exl.cap
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    pout = {"kps_l": pin((B, cap, 7), torch.float32).view(np.uint8).reshape(B, cap, 28).view(orbx.KP_DTYPE).reshape(B, cap),
            "desc_l": pin((B, cap, 32), torch.uint8), "n_l": pin((B,), torch.int32),
            "kps_r": pin((B, cap, 7), torch.float32).view(np.uint8).reshape(B, cap, 28).view(orbx.KP_DTYPE).reshape(B, cap),
            "desc_r": pin((B, cap, 32), torch.uint8), "n_r": pin((B,), torch.int32),
            "uright": pin((B, cap), torch.float32), "depth": pin((B, cap), torch.float32)}
    m.ExtractStereoBatch(exl, exr, torch.from_numpy(Ls).pin_memory().numpy(), torch.from_numpy(Rs).pin_memory().numpy(), mb, mbf, pout)
    for i in range(6):
        ol, orr = O.Extractor(nf, 1.2, 8, 20, 7), O.Extractor(nf, 1.2, 8, 20, 7)
        _, kl, dl = ol(pairs[i][0], (0, 0)); _, kr, dr = orr(pairs[i][1], (0, 0))
        rur, rdp, _ = O.compute_stereo_matches(ol, orr, kl, dl, kr, dr, mb, mbf)
        for j in (i, i + 6 * ((B - 1 - i) // 6)):           # the first and the last copy of this pair in the batch (different chunks)
            for o in (out, pout):
                n = int(o["n_l"][j])
                assert n == len(kl) and int(o["n_r"][j]) == len(kr)
                assert o["kps_l"][j, :n].tobytes() == kl.tobytes() and o["kps_r"][j, :len(kr)].tobytes() == kr.tobytes()
                np.testing.assert_array_equal(o["desc_l"][j, :n], dl)
                np.testing.assert_array_equal(o["uright"][j, :n], rur)
                np.testing.assert_array_equal(o["depth"][j, :n], rdp)
    exl.close(); exr.close(); m.close()


def _proj_setup(seed=71):
    W, H = 752, 480
    st = synth.rects_stream(W, H, 2, seed=seed)
    e = O.Extractor(1000, 1.2, 8, 20, 7)
    _, k1, d1 = e(st[0], (0, 0)); _, k2, d2 = e(st[1], (0, 0))
    q = np.zeros(len(k1), O.PROJQ_DTYPE)
    q["u"] = k1["x"] + np.float32(3); q["v"] = k1["y"] + np.float32(2)
    q["r"] = np.float32(4.0) * np.asarray(e.scale, np.float32)[k1["octave"]]
    q["minl"] = k1["octave"] - 1; q["maxl"] = k1["octave"]
    q["ur"] = q["u"]; q["angle"] = k1["angle"]; q["valid"] = (np.arange(len(k1)) % 9 != 0).astype(np.int32)
    return W, H, e, k1, d1, k2, d2, q


def test_projection_search_with_distance_bound():
    """The Sim3 / relocalisation SearchByProjection overloads = mode 0 with their own acceptance bound and a preset
    'already matched' table (ORBmatcher.cc:473-700, :2188-2310)."""
    W, H, e, k1, d1, k2, d2, q = _proj_setup()
    pre = np.full(len(k2), -1, np.int32); pre[::17] = len(k1)                 # keypoints that already have a map point
    for max_dist, ori in ((50, False), (25, False), (100, True), (64, True)):
        m = orbx.ORBmatcher(0.9, ori, max_keypoints=2048)
        n, a = m.SearchByProjectionEx(0, q, d1, k2, d2, (0, W, 0, H), assigned=pre, max_dist=max_dist)
        rn, ra = O.search_by_projection_ex(0, q, d1, k2, d2, (0, W, 0, H), assigned=pre, nnratio=0.9, check_ori=ori, max_dist=max_dist)
        assert n == rn and n > 50
        np.testing.assert_array_equal(a, ra)
        m.close()
    # independent check of the rule with the grid query of the oracle: nothing above the bound is accepted
    n25, a25 = O.search_by_projection_ex(0, q, d1, k2, d2, (0, W, 0, H), assigned=pre, check_ori=False, max_dist=25)
    for i2 in np.nonzero((a25 >= 0) & (a25 < len(k1)))[0]:
        assert O.hamming256(d1[a25[i2]], d2[i2]) <= 25


def test_fuse_style_independent_best_with_chi2_gate():
    """Fuse (ORBmatcher.cc:1395-1742): best candidate of every projected map point on its own, candidates gated by the
    reprojection error e2 * invSigma2[level] > 5.99; compared with the oracle and with a brute-force numpy restatement
    built on the oracle's GetFeaturesInArea."""
    W, H, e, k1, d1, k2, d2, q = _proj_setup(seed=72)
    inv_sigma2 = np.asarray(e.inv_sigma2, np.float32)
    m = orbx.ORBmatcher(0.9, True, max_keypoints=2048)
    for chi2, sg in ((5.99, inv_sigma2), (0.0, None)):
        n, bi, bd = m.SearchByProjectionEx(3, q, d1, k2, d2, (0, W, 0, H), inv_sigma2=sg, chi2=chi2)
        rn, rbi, rbd = O.search_by_projection_ex(3, q, d1, k2, d2, (0, W, 0, H), inv_sigma2=sg, chi2=chi2)
        assert n == rn and n > 100
        np.testing.assert_array_equal(bi, rbi); np.testing.assert_array_equal(bd, rbd)
    m.close()


def test_pipeline_rejects_batch_beyond_extractor():
    """the pipeline must validate the batch against BOTH handles (the extractor's result slots / scratch are sized
    by its own max_batch), reject a stride below the width and handles on different devices; nothing is launched."""
    W, H = 320, 240
    ex = orbx.ORBextractor(300, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=2)
    m = orbx.ORBmatcher(0.9, True, max_keypoints=ex.cap, max_batch=8)
    cap = ex.cap
    frames = synth.rects_stream(W, H, 4, seed=3)
    out = {"kps": np.zeros((4, cap), orbx.KP_DTYPE), "desc": np.zeros((4, cap, 32), np.uint8), "n": np.zeros(4, np.int32),
           "mono": np.zeros(4, np.int32), "matches12": np.zeros((4, cap), np.int32), "nmatches": np.zeros(4, np.int32)}
    with pytest.raises(orbx.OrbxError) as e:
        orbx.extract_match_batch(ex, m, frames, (0, 0), (0, W, 0, H), 100, out)
    assert e.value.code == orbx.ORBX_E_INVALID and "extractor" in str(e.value)
    # a legal batch on the same handles still works afterwards
    orbx.extract_match_batch(ex, m, np.ascontiguousarray(frames[:2]), (0, 0), (0, W, 0, H), 100, {k: v[:2] for k, v in out.items()})
    assert out["n"][0] > 0
    ex.close(); m.close()


def test_projection_options_occupancy_origin_and_stereo_gate():
    """orbx_search_by_projection_opts: (1) claims by 0-observation MapPoints (valid bit 1) leave the keypoint free, later queries
    take it over and both count (R/src/ORBmatcher.cc:89-91, :2045-2047), claims cleared by the rotation check come back as -2;
    (2) the query origin of a KeyFrame grid (int-truncated mnMinX / mnMinY, R/src/KeyFrame.cc:897-911); (3) Fuse's 3-dof gate for
    keypoints with a right coordinate (:1525-1540).  Against the C oracle (whose class-level behaviour tests/test_dropin_scenario.py
    pins to the reference's ORBmatcher.cc)."""
    W, H, e, k1, d1, k2, d2, q = _proj_setup(seed=73)
    rng = np.random.default_rng(5)
    # (1) duplicate every query so that two queries compete for the same keypoint; a third of them do not occupy
    q2 = np.concatenate([q, q]); dd = np.concatenate([d1, d1])
    q2["valid"] = np.where(q2["valid"] > 0, np.where(rng.random(len(q2)) < 0.35, 3, 1), 0).astype(np.int32)
    pre = np.full(len(k2), -1, np.int32); pre[::23] = len(q2)
    for mode, ori in ((0, True), (0, False), (1, False)):
        m = orbx.ORBmatcher(0.9, ori, max_keypoints=4096)
        n, a = m.SearchByProjectionOpts(mode, q2, dd, k2, d2, (0, W, 0, H), assigned=pre)
        rn, ra = O.search_by_projection_full(mode, q2, dd, k2, d2, (0, W, 0, H), assigned=pre, nnratio=0.9, check_ori=ori)
        assert n == rn and n > 100
        np.testing.assert_array_equal(a, ra)
        owners = a[(a >= 0) & (a < len(q2))]
        assert n > len(owners) or mode == 0 and ori            # takeovers: more accepting queries than owned keypoints
        if mode == 0 and ori:
            assert (a == -2).any()                              # something was claimed and then cleared by the rotation check
        m.close()
    # (2) fractional grid bounds (a distorted camera) with the truncated query origin
    bounds = (-3.7, W + 2.4, -2.6, H + 1.9)
    m = orbx.ORBmatcher(0.9, False, max_keypoints=4096)
    n, a = m.SearchByProjectionOpts(0, q, d1, k2, d2, bounds, query_origin=(-3.0, -2.0), max_dist=60)
    rn, ra = O.search_by_projection_full(0, q, d1, k2, d2, bounds, query_origin=(-3.0, -2.0), nnratio=0.9, check_ori=False, max_dist=60)
    assert n == rn and n > 100
    np.testing.assert_array_equal(a, ra)
    # (3) both chi-square gates: two thirds of the keypoints have a right coordinate
    ur = np.where(np.arange(len(k2)) % 3 != 0, k2["x"] - np.float32(9.0), np.float32(-1)).astype(np.float32)
    q["ur"] = q["u"] - np.float32(9.0) + rng.normal(0, 0.7, len(q)).astype(np.float32)
    sg = np.asarray(e.inv_sigma2, np.float32)
    n, bi, bd = m.SearchByProjectionOpts(3, q, d1, k2, d2, (0, W, 0, H), uright=ur, inv_sigma2=sg, chi2=5.99, chi2_stereo=7.8)
    rn, rbi, rbd = O.search_by_projection_full(3, q, d1, k2, d2, (0, W, 0, H), uright=ur, inv_sigma2=sg, chi2=5.99, chi2_stereo=7.8)
    n1, bi1, _ = O.search_by_projection_full(3, q, d1, k2, d2, (0, W, 0, H), uright=ur, inv_sigma2=sg, chi2=5.99, chi2_stereo=0.0)
    assert n == rn and n > 40 and not np.array_equal(rbi, bi1)       # the stereo form changes the outcome on this input
    np.testing.assert_array_equal(bi, rbi); np.testing.assert_array_equal(bd, rbd)
    m.close()


def test_stream_submit_wait_equals_plain_calls():
    """orbx_stream_submit / orbx_stream_wait: two batches in flight (input copy, kernels and result copy of three consecutive
    batches overlap); results and the predecessor carried from batch to batch must be those of the plain synchronous calls."""
    import torch
    B, W, H = 48, 640, 480
    nb = 5
    frames = synth.rects_stream(W, H, nb * B, seed=93)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    batches = [pin(frames[i * B:(i + 1) * B]) for i in range(nb)]

    def buffers(cap, pinned):
        mk = (lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory().numpy()) if pinned else (lambda shape, dt: torch.zeros(shape, dtype=dt).numpy())
        return {"kps": mk((B, cap, 7), torch.float32).view(np.uint8).reshape(B, cap, 28).view(orbx.KP_DTYPE).reshape(B, cap),
                "desc": mk((B, cap, 32), torch.uint8), "n": mk((B,), torch.int32), "mono": mk((B,), torch.int32),
                "matches12": mk((B, cap), torch.int32), "nmatches": mk((B,), torch.int32),
                "knn_idx": mk((B, cap, 2), torch.int32), "knn_dist": mk((B, cap, 2), torch.int32)}

    ex = orbx.ORBextractor(800, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    m = orbx.ORBmatcher(0.9, True, max_keypoints=ex.cap, max_batch=B)
    plain = []
    for b in batches:
        out = buffers(ex.cap, False)
        orbx.extract_match_batch(ex, m, b, (0, 0), (0, W, 0, H), 100, out)
        plain.append(out)
    ex.close(); m.close()

    ex = orbx.ORBextractor(800, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    m = orbx.ORBmatcher(0.9, True, max_keypoints=ex.cap, max_batch=B)
    outs = [buffers(ex.cap, True) for _ in range(nb)]
    tickets = [orbx.stream_submit(ex, m, batches[0], (0, 0), (0, W, 0, H), 100, outs[0])]
    for k in range(1, nb):
        tickets.append(orbx.stream_submit(ex, m, batches[k], (0, 0), (0, W, 0, H), 100, outs[k]))     # two in flight
        if k == 2:
            with pytest.raises(orbx.OrbxError) as e:      # a third one is refused
                orbx.stream_submit(ex, m, batches[k], (0, 0), (0, W, 0, H), 100, outs[k])
            assert e.value.code == orbx.ORBX_E_CAPACITY
        orbx.stream_wait(ex, m, tickets[k - 1])
    orbx.stream_wait(ex, m, tickets[-1])
    orbx.stream_wait(ex, m, tickets[0])                   # waiting again is harmless
    for ci, (a, b) in enumerate(zip(plain, outs)):
        for k in ("n", "mono", "nmatches"):
            np.testing.assert_array_equal(a[k], b[k], err_msg="%s batch %d" % (k, ci))
        for f in range(B):
            n = int(a["n"][f])
            assert a["kps"][f, :n].tobytes() == b["kps"][f, :n].tobytes()
            np.testing.assert_array_equal(a["desc"][f, :n], b["desc"][f, :n])
            if f == 0 and ci == 0:
                continue
            npk = int(a["n"][f - 1]) if f else int(plain[ci - 1]["n"][B - 1])
            for k in ("matches12", "knn_idx", "knn_dist"):
                np.testing.assert_array_equal(a[k][f, :npk], b[k][f, :npk], err_msg="%s batch %d frame %d" % (k, ci, f))
    # the plain call still works on the same handles afterwards (predecessor = last streamed frame)
    out = buffers(ex.cap, False)
    orbx.extract_match_batch(ex, m, batches[0], (0, 0), (0, W, 0, H), 100, out)
    np.testing.assert_array_equal(out["n"], plain[0]["n"])
    ex.close(); m.close()


def test_full_c1_batch_matching_properties():
    """BASELINE config C1 at the bench size through the host-buffer call (512 frames of 752x480, extract + SearchForInitialization +
    BF kNN-2 against the previous frame).  The stream repeats with period 32, so pair (i-1, i) must give byte-identical matches and
    kNN tables as pair (i-33, i-32) wherever it sits in the batch and in whichever chunk of the pipeline; pairs of one period are
    compared with the oracle."""
    W, H, B, PER = 752, 480, 512, 32
    base = synth.rects_stream(W, H, PER, seed=321)
    frames = np.ascontiguousarray(np.concatenate([base] * (B // PER)))
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    m = orbx.ORBmatcher(0.9, True, max_keypoints=ex.cap, max_batch=B)
    cap = ex.cap
    out = {"kps": np.zeros((B, cap), orbx.KP_DTYPE), "desc": np.zeros((B, cap, 32), np.uint8), "n": np.zeros(B, np.int32),
           "mono": np.zeros(B, np.int32), "matches12": np.zeros((B, cap), np.int32), "nmatches": np.zeros(B, np.int32),
           "knn_idx": np.zeros((B, cap, 2), np.int32), "knn_dist": np.zeros((B, cap, 2), np.int32)}
    orbx.extract_match_batch(ex, m, frames, (0, 0), (0, W, 0, H), 100, out)
    n = out["n"]
    assert n.min() >= 900
    for i in range(PER + 1, B):
        j = i - PER
        npk = int(n[i - 1])                                  # rows of the pair's tables = keypoints of the predecessor
        assert n[i] == n[j] and n[i - 1] == n[j - 1] and out["nmatches"][i] == out["nmatches"][j], i
        assert np.array_equal(out["matches12"][i, :npk], out["matches12"][j, :npk]), i
        assert np.array_equal(out["knn_idx"][i, :npk], out["knn_idx"][j, :npk]) and np.array_equal(out["knn_dist"][i, :npk], out["knn_dist"][j, :npk]), i
    ref = O.Extractor(1000, 1.2, 8, 20, 7)
    for i in (1, 17, PER):                                   # PER: frame 0 of the second period against frame 31 of the first
        _, pk, pd = ref(frames[i - 1], (0, 0))
        _, rk, rd = ref(frames[i], (0, 0))
        rn, rm12, _ = O.search_for_initialization(pk, pd, rk, rd, (0, W, 0, H), np.stack([pk["x"], pk["y"]], 1), 100, 0.9, True)
        assert int(out["nmatches"][i]) == rn and np.array_equal(out["matches12"][i, :len(pk)], rm12)
        ri, rdist = O.bf_knn2(pd, rd)
        assert np.array_equal(out["knn_idx"][i, :len(pk)], ri) and np.array_equal(out["knn_dist"][i, :len(pk)], rdist)
    assert int(out["nmatches"][1:].min()) > 0
    ex.close(); m.close()


def test_single_frame_call_graph_survives_reconfiguration():
    """The batch-1 host call replays its kernels from a CUDA graph after two identical calls (captured in the level-parallel order
    of run_batch_dag: per-level FAST / octree / blur branches beside the resize chain, the brute-force search beside the window
    search).  A reconfiguration of the extractor through ANOTHER entry point (a different image size) re-allocates its buffers: the cached graph must not be replayed."""
    W, H = 640, 480
    frames = synth.rects_stream(W, H, 6, seed=94)
    small = synth.rects_stream(320, 240, 1, seed=95)[0]
    ex = orbx.ORBextractor(800, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=1)
    m = orbx.ORBmatcher(0.9, True, max_keypoints=ex.cap, max_batch=1)
    ref = O.Extractor(800, 1.2, 8, 20, 7)
    cap = ex.cap

    def call(f):
        out = {"kps": np.zeros((1, cap), orbx.KP_DTYPE), "desc": np.zeros((1, cap, 32), np.uint8), "n": np.zeros(1, np.int32),
               "mono": np.zeros(1, np.int32), "matches12": np.zeros((1, cap), np.int32), "nmatches": np.zeros(1, np.int32),
               "knn_idx": np.zeros((1, cap, 2), np.int32), "knn_dist": np.zeros((1, cap, 2), np.int32)}
        orbx.extract_match_batch(ex, m, np.ascontiguousarray(frames[f:f + 1]), (0, 0), (0, W, 0, H), 100, out)
        return out

    def check(out, f, prev):
        rmono, rk, rd = ref(frames[f], (0, 0))
        n = int(out["n"][0])
        assert n == len(rk) and int(out["mono"][0]) == rmono
        for name in ("x", "y", "size", "response", "octave"):
            np.testing.assert_array_equal(out["kps"][0, :n][name], rk[name])
        assert np.array_equal(out["desc"][0, :n][out["kps"][0, :n]["angle"] == rk["angle"]], rd[out["kps"][0, :n]["angle"] == rk["angle"]])
        if prev is not None:
            pk, pd = prev
            rn, rm12, _ = O.search_for_initialization(pk, pd, rk, rd, (0, W, 0, H), np.stack([pk["x"], pk["y"]], 1), 100, 0.9, True)
            assert int(out["nmatches"][0]) == rn and np.array_equal(out["matches12"][0, :len(pk)], rm12)
            ri, rdist = O.bf_knn2(pd, rd)
            assert np.array_equal(out["knn_idx"][0, :len(pk)], ri) and np.array_equal(out["knn_dist"][0, :len(pk)], rdist)
        return rk, rd

    prev = None
    for f in range(4):                                   # direct, captured, replayed, replayed
        prev = check(call(f), f, prev)
    _, sk, _ = ex(small, None, (0, 0))                   # reconfigures the extractor (320x240), then back
    assert len(sk) > 0
    prev = None                                          # slot 0 was rewritten by the plain call
    out = call(4)
    prev = check(out, 4, None)
    check(call(5), 5, prev)
    ex.close(); m.close()
