"""GPU parity tests of the bag-of-words transform (orbx_vocab_* / orbx_bow_transform through the C ABI) against the oracle
and the committed golden vectors."""
import os

import numpy as np
import pytest
import torch

from multi_orbslam3_b200 import orbx, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "bow.npz")


@pytest.mark.parametrize("name", ["k10_L3", "irregular_k6_L4", "wide_k20_L2"])
def test_bow_golden(name):
    g = np.load(GOLD)
    vocab = tuple(g["%s_%s" % (name, k)] for k in ("parent", "leaf", "desc", "weight"))
    q, L = g["%s_queries" % name], int(g["%s_L" % name])
    V = orbx.ORBVocabulary(*vocab, L=L)
    for ls in (1, L - 1, L + 2):
        w, wt, nd = V.transform_features(q, ls)
        np.testing.assert_array_equal(w, g["%s_ls%d_word" % (name, ls)])
        np.testing.assert_array_equal(nd, g["%s_ls%d_node" % (name, ls)])
        (bw, bv), _ = V.transform(q, ls)
        np.testing.assert_array_equal(bw, g["%s_ls%d_bow_words" % (name, ls)])
        np.testing.assert_array_equal(bv, g["%s_ls%d_bow_values" % (name, ls)])
    V.close()


@pytest.mark.parametrize("kw,n,ls", [(dict(k=10, L=6), 20000, 4), (dict(k=9, L=5, irregular=True, shuffle=True), 5000, 3),
                                     (dict(k=20, L=3), 3000, 2), (dict(k=2, L=9), 2000, 4)],
                         ids=["orbvoc_shape_k10_L6", "irregular_shuffled", "k20", "binary_deep"])
def test_bow_matches_oracle(kw, n, ls):
    """ORBvoc-shaped vocabulary (1 111 111 nodes, one million words) and odd shapes: every word / node id and both maps
    equal the oracle's (doubles bit for bit)."""
    vocab = synth.random_vocabulary(seed=3, **kw)
    q = synth.vocabulary_queries(vocab, n, seed=4)
    q[5] = vocab[2][np.nonzero(vocab[1])[0][17]]                 # an exact word descriptor: distance 0 at the leaf
    V = orbx.ORBVocabulary(*vocab, L=kw["L"]); R = O.Vocabulary(*vocab, L=kw["L"])
    w, wt, nd = V.transform_features(q, ls)
    rw, rwt, rnd = R.transform_features(q, ls)
    np.testing.assert_array_equal(w, rw); np.testing.assert_array_equal(wt, rwt); np.testing.assert_array_equal(nd, rnd)
    (bw, bv), (fn, ff) = V.transform(q[:1500], ls)
    (rbw, rbv), (rfn, rff) = R.transform(q[:1500], ls)
    np.testing.assert_array_equal(bw, rbw); np.testing.assert_array_equal(bv, rbv); np.testing.assert_array_equal(fn, rfn)
    assert all(np.array_equal(a, b) for a, b in zip(ff, rff))
    assert V.n_words == int(vocab[1].sum())
    V.close()


def test_bow_on_extractor_slots():
    """Frame::ComputeBoW straight after extraction: the descriptors stay in the extractor's result slots on the device."""
    W, H, B = 640, 480, 3
    frames = synth.rects_stream(W, H, B, seed=9)
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    res = ex.extract_batch(frames)
    vocab = synth.random_vocabulary(k=10, L=4, seed=5)
    V = orbx.ORBVocabulary(*vocab, L=4); R = O.Vocabulary(*vocab, L=4)
    dw = torch.full((B, ex.cap), -7, dtype=torch.int32, device="cuda"); dn = torch.full((B, ex.cap), -7, dtype=torch.int32, device="cuda")
    V.transform_slots_device(ex, 0, B, dw.data_ptr(), dn.data_ptr(), levelsup=2)
    ex.sync()
    torch.cuda.synchronize()
    hw, hn = dw.cpu().numpy(), dn.cpu().numpy()
    for i, (_, kps, desc) in enumerate(res):
        rw, _, rnd = R.transform_features(desc, 2)
        np.testing.assert_array_equal(hw[i, :len(kps)], rw)
        np.testing.assert_array_equal(hn[i, :len(kps)], rnd)
        assert (hw[i, len(kps):] == -7).all()                  # rows past the keypoint count are untouched
    V.close(); ex.close()


def test_bow_rejects_bad_tables():
    parent = np.array([-1, 0, 0], np.int32); desc = np.zeros((3, 32), np.uint8); w = np.ones(3)
    with pytest.raises(orbx.OrbxError):
        orbx.ORBVocabulary(np.array([-1, 2, 0], np.int32), np.array([0, 1, 1], np.uint8), desc, w, L=1)      # parent after child
    with pytest.raises(orbx.OrbxError):
        orbx.ORBVocabulary(parent, np.array([0, 0, 1], np.uint8), desc, w, L=1)                               # childless inner node
    V = orbx.ORBVocabulary(parent, np.array([0, 1, 1], np.uint8), desc, w, L=1)
    w0, _, n0 = V.transform_features(np.zeros((0, 32), np.uint8), 0)
    assert len(w0) == 0 and len(n0) == 0
    V.close()


@pytest.mark.parametrize("levelsup", [3, 2])
@pytest.mark.parametrize("mode,ratio", [(0, 0.7), (1, 0.8), (0, 0.95)])
def test_search_by_bow_golden(levelsup, mode, ratio):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "search_by_bow.npz"))
    vocab = tuple(g["voc_" + k] for k in ("parent", "leaf", "desc", "weight"))
    V = orbx.ORBVocabulary(*vocab, L=4)
    k1, d1, k2, d2, v1, v2 = g["k1"], g["d1"], g["k2"], g["d2"], g["valid1"], g["valid2"]
    fv1 = V.transform(d1, levelsup)[1]; fv2 = V.transform(d2, levelsup)[1]
    tag = "ls%d_m%d_r%d" % (levelsup, mode, int(ratio * 100))
    for ori, suffix in ((True, ""), (False, "_noori")):
        m = orbx.ORBmatcher(ratio, ori, max_keypoints=2048)
        n, m12 = m.SearchByBoW(mode, k1, d1, v1, fv1, k2, d2, v2 if mode == 1 else None, fv2)
        assert n == int(g[tag + suffix + "_n"])
        np.testing.assert_array_equal(m12, g[tag + suffix + "_m12"])
        m.close()
    V.close()


def test_search_by_bow_matches_oracle_c1_and_duplicates():
    """C1-size frames with planted duplicate descriptors (ties inside a node: first in list order wins, a claimed feature
    is skipped by later keyframe features), BoW on the GPU for both sides."""
    W, H = 752, 480
    st = synth.rects_stream(W, H, 2, seed=41)
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H)
    (_, k1, d1), (_, k2, d2) = ex(st[0], None, (0, 0)), ex(st[1], None, (0, 0))
    d2 = d2.copy(); d1 = d1.copy()
    d2[10:40] = d1[10:40]; d2[40:70] = d1[10:40]              # exact copies, twice: best == second -> ratio test fails
    d1[100:130] = d1[130:160]                                  # two keyframe features compete for the same frame feature
    d2[100:130] = d1[100:130]
    vocab = synth.random_vocabulary(k=10, L=5, seed=13)
    V = orbx.ORBVocabulary(*vocab, L=5); R = O.Vocabulary(*vocab, L=5)
    fv1 = V.transform(d1, 4)[1]; fv2 = V.transform(d2, 4)[1]
    rf1 = R.transform(d1, 4)[1]; rf2 = R.transform(d2, 4)[1]
    assert all(np.array_equal(a, b) for a, b in zip(fv1[1], rf1[1]))
    v1 = np.ones(len(k1), np.uint8); v1[::9] = 0
    v2 = np.ones(len(k2), np.uint8); v2[::7] = 0
    for mode, ratio in ((0, 0.7), (1, 0.9)):
        m = orbx.ORBmatcher(ratio, True, max_keypoints=2048)
        n, m12 = m.SearchByBoW(mode, k1, d1, v1, fv1, k2, d2, v2 if mode == 1 else None, fv2)
        rn, rm12 = O.search_by_bow(mode, k1, d1, v1, rf1, k2, d2, v2 if mode == 1 else None, rf2, ratio, True)
        assert n == rn and n > 20
        np.testing.assert_array_equal(m12, rm12)
        m.close()
    # malformed tables are refused
    m = orbx.ORBmatcher(0.7, True, max_keypoints=2048)
    bad = (fv2[0][::-1].copy(), fv2[1][::-1])
    with pytest.raises(orbx.OrbxError):
        m.SearchByBoW(0, k1, d1, v1, fv1, k2, d2, None, bad)
    m.close(); V.close(); ex.close()


@pytest.mark.parametrize("only_stereo,coarse,ep,use_stereo", [(False, False, (1e6, 1e6), False), (False, False, (240.0, 180.0), True),
                                                             (True, False, (1e6, 1e6), True), (False, True, (100.0, 100.0), False)])
def test_search_for_triangulation_matches_oracle(only_stereo, coarse, ep, use_stereo):
    """ORBmatcher::SearchForTriangulation: epipole gate, epipolar-line gate, last-of-equals rule, rotation filter."""
    from test_bow import triangulation_case
    e, k1, d1, k2, d2, fv1, fv2, F12, free1, free2, st1, st2, _ = triangulation_case()
    d2 = d2.copy(); d2[40:60] = d1[40:60]                      # exact duplicates: equal distances inside a node
    s1, s2 = (st1, st2) if use_stereo else (None, None)
    for ori in (True, False):
        m = orbx.ORBmatcher(0.6, ori, max_keypoints=2048)
        n, m12 = m.SearchForTriangulation(k1, d1, free1, s1, fv1, k2, d2, free2, s2, fv2, F12, ep, e.scale, e.sigma2, only_stereo, coarse)
        rn, rm12 = O.search_for_triangulation(k1, d1, free1, s1, fv1, k2, d2, free2, s2, fv2, F12, ep, e.scale, e.sigma2, only_stereo, coarse, ori)
        assert n == rn
        np.testing.assert_array_equal(m12, rm12)
        m.close()
