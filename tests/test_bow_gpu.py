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
