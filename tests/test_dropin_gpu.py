"""The drop-in C++ classes (dropin/ORBextractor, dropin/ORBmatcher: the reference's public signatures over the C ABI)
driven the way Frame/Tracking drive them, compared with the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from multi_orbslam3_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dropin_classes_match_oracle(tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "dropin")])
    W, H = 752, 480
    st = synth.rects_stream(W, H, 2, seed=5)
    for k in range(2):
        st[k].tofile(tmp_path / ("f%d.raw" % k))
    out = tmp_path / "out.bin"
    # a synthetic vocabulary in the text format of ORBvoc.txt ("k L scoring weighting", then parent isLeaf 32 bytes weight)
    vocab = synth.random_vocabulary(k=10, L=5, seed=21)
    with open(tmp_path / "voc.txt", "w") as fh:
        fh.write("10 5 0 0\n")
        for i in range(1, len(vocab[0])):
            fh.write("%d %d %s %r\n" % (vocab[0][i], vocab[1][i], " ".join(str(int(b)) for b in vocab[2][i]), float(vocab[3][i])))
    subprocess.check_call([os.path.join(ROOT, "dropin", "_build", "test_dropin"), str(W), str(H),
                           str(tmp_path / "f0.raw"), str(tmp_path / "f1.raw"), str(out), str(tmp_path / "voc.txt")])
    buf = out.read_bytes()
    off = 0
    ref = O.Extractor(1000, 1.2, 8, 20, 7)
    frames = []
    for k in range(2):
        n, mono = struct.unpack_from("<ii", buf, off); off += 8
        kps = np.frombuffer(buf, O.KP_DTYPE, n, off); off += 28 * n
        desc = np.frombuffer(buf, np.uint8, 32 * n, off).reshape(n, 32); off += 32 * n
        rmono, rk, rd = ref(st[k], (0, 0))
        assert (mono, n) == (rmono, len(rk))
        for name in ("x", "y", "size", "response", "octave", "class_id"):
            np.testing.assert_array_equal(kps[name], rk[name])
        assert np.abs(kps["angle"] - rk["angle"]).max() <= 1e-3
        same = kps["angle"] == rk["angle"]
        np.testing.assert_array_equal(desc[same], rd[same])
        frames.append((kps, desc))
        if k == 1:
            lvl3 = ref.level_image(3)
    nm, = struct.unpack_from("<i", buf, off); off += 4
    n0 = len(frames[0][0])
    m12 = np.frombuffer(buf, np.int32, n0, off); off += 4 * n0
    (k1, d1), (k2, d2) = frames
    rn, rm12, _ = O.search_for_initialization(k1, d1, k2, d2, (0, W, 0, H), np.stack([k1["x"], k1["y"]], 1), 100, 0.9, True)
    assert nm == rn
    np.testing.assert_array_equal(m12, rm12)
    lw, lh = struct.unpack_from("<ii", buf, off); off += 8
    pyr3 = np.frombuffer(buf, np.uint8, lw * lh, off).reshape(lh, lw); off += lw * lh
    np.testing.assert_array_equal(pyr3, lvl3)           # mvImagePyramid after SyncPyramidToHost()
    d, = struct.unpack_from("<i", buf, off); off += 4
    assert d == O.hamming256(d1[0], d2[0])

    # SearchByProjection(Frame&, vector<MapPoint*>&, th = 3): the drop-in marshals MapPoints into flat queries
    scale = np.asarray(ref.scale, np.float32)
    n1 = len(k1); n2 = len(k2)
    q = np.zeros(n1, O.PROJQ_DTYPE)
    idx = np.arange(n1)
    view = np.where(idx % 3 != 0, np.float32(0.9995), np.float32(0.9))
    r = np.where(view > 0.998, np.float32(2.5), np.float32(4.0)).astype(np.float32) * np.float32(3.0)
    q["u"] = k1["x"] + np.float32(3); q["v"] = k1["y"] + np.float32(2)
    q["r"] = r * scale[k1["octave"]]
    q["minl"] = k1["octave"] - 1; q["maxl"] = k1["octave"]
    q["valid"] = (idx % 7 != 0).astype(np.int32)
    nA, = struct.unpack_from("<i", buf, off); off += 4
    gotA = np.frombuffer(buf, np.int32, n2, off); off += 4 * n2
    rnA, rA = O.search_by_projection(1, q, d1, k2, d2, (0, W, 0, H), nnratio=0.8, check_ori=True)
    assert nA == rnA and nA > 100
    np.testing.assert_array_equal(gotA, rA)

    # SearchByProjection(Frame& Current, const Frame& Last, th = 7, bMono): projection of the last frame's map points
    q = np.zeros(n1, O.PROJQ_DTYPE)
    q["u"] = k1["x"] + np.float32(3); q["v"] = k1["y"] + np.float32(2)
    q["r"] = np.float32(7.0) * scale[k1["octave"]]
    q["minl"] = k1["octave"] - 1; q["maxl"] = k1["octave"] + 1
    q["ur"] = q["u"]; q["angle"] = k1["angle"]
    q["valid"] = (idx % 11 != 0).astype(np.int32)
    nB, = struct.unpack_from("<i", buf, off); off += 4
    gotB = np.frombuffer(buf, np.int32, n2, off); off += 4 * n2
    rnB, rB = O.search_by_projection(0, q, d1, k2, d2, (0, W, 0, H), nnratio=0.9, check_ori=True)
    assert nB == rnB and nB > 100
    np.testing.assert_array_equal(gotB, rB)

    # ORBmatcher::ComputeStereoMatches(left extractor, right extractor, ...): frame 1 as the left view, frame 0 as the right
    ns, = struct.unpack_from("<i", buf, off); off += 4
    uR = np.frombuffer(buf, np.float32, ns, off); off += 4 * ns
    dep = np.frombuffer(buf, np.float32, ns, off); off += 4 * ns
    ol, orr = O.Extractor(1000, 1.2, 8, 20, 7), O.Extractor(1000, 1.2, 8, 20, 7)
    _, skl, sdl = ol(st[1], (0, 0)); _, skr, sdr = orr(st[0], (0, 0))
    rur, rdp, _ = O.compute_stereo_matches(ol, orr, skl, sdl, skr, sdr, 0.11, 47.9)
    assert ns == len(rur)
    np.testing.assert_array_equal(uR, rur)
    np.testing.assert_array_equal(dep, rdp)
    assert (rur >= 0).sum() > 10

    # ORBVocabulary::loadFromTextFile + transform(features, BowVector, FeatureVector, 4) on frame 0's descriptors
    RV = O.Vocabulary(*vocab, L=5)
    (rbw, rbv), (rfn, rff) = RV.transform(d1, 4)
    nbow, = struct.unpack_from("<i", buf, off); off += 4
    assert nbow == len(rbw)
    rec = np.frombuffer(buf, np.dtype([("w", "<u4"), ("v", "<f8")]), nbow, off); off += 12 * nbow
    np.testing.assert_array_equal(rec["w"], rbw)
    np.testing.assert_array_equal(rec["v"], rbv)                  # doubles, bit for bit (values survive the text round trip via repr)
    nfv, = struct.unpack_from("<i", buf, off); off += 4
    assert nfv == len(rfn)
    for k in range(nfv):
        nid, c = struct.unpack_from("<Ii", buf, off); off += 8
        feats = np.frombuffer(buf, np.uint32, c, off); off += 4 * c
        assert nid == rfn[k]
        np.testing.assert_array_equal(feats, rff[k])
    nw, = struct.unpack_from("<I", buf, off); off += 4
    assert nw == int(vocab[1].sum())

    # ORBmatcher::SearchByBoW(KeyFrame*, Frame&): result per frame feature = index of the keyframe's map point
    fv1 = RV.transform(d1, 4)[1]; fv2 = RV.transform(d2, 4)[1]
    i1 = np.arange(n1); i2 = np.arange(n2)
    v1 = ((i1 % 6 != 0) & (i1 % 10 != 3)).astype(np.uint8)
    v2 = ((i2 % 5 != 0) & (i2 % 13 != 4)).astype(np.uint8)
    nR, = struct.unpack_from("<i", buf, off); off += 4
    got = np.frombuffer(buf, np.int32, n2, off); off += 4 * n2
    rn, rm12 = O.search_by_bow(0, k1, d1, v1, fv1, k2, d2, None, fv2, 0.7, True)
    want = np.full(n2, -1, np.int32); want[rm12[rm12 >= 0]] = np.nonzero(rm12 >= 0)[0]
    assert nR == rn and nR > 5
    np.testing.assert_array_equal(got, want)
    # ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*): result per feature of keyframe 1 = index of keyframe 2's map point
    nL, = struct.unpack_from("<i", buf, off); off += 4
    got = np.frombuffer(buf, np.int32, n1, off); off += 4 * n1
    rn, rm12 = O.search_by_bow(1, k1, d1, v1, fv1, k2, d2, v2, fv2, 0.8, True)
    assert nL == rn and nL > 5
    np.testing.assert_array_equal(got, rm12)
