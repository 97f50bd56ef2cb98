"""CPU tests of the bag-of-words oracle (DBoW2 transform as Frame::ComputeBoW uses it): golden vectors produced by the
independent Python restatement in oracle/gen_golden.py, invariants, and the host-side BowVector / FeatureVector assembly
the product uses (multi_orbslam3_b200.orbx.assemble_bow) on the oracle's per-feature output."""
import os

import numpy as np
import pytest

from multi_orbslam3_b200 import orbx, synth
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "bow.npz")
CASES = ["k10_L3", "irregular_k6_L4", "wide_k20_L2"]


def load(name):
    g = np.load(GOLD)
    vocab = tuple(g["%s_%s" % (name, k)] for k in ("parent", "leaf", "desc", "weight"))
    return g, vocab, g["%s_queries" % name], int(g["%s_L" % name])


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_golden(name):
    g, vocab, q, L = load(name)
    V = O.Vocabulary(*vocab, L=L)
    for ls in (1, L - 1, L + 2):
        w, wt, nd = V.transform_features(q, ls)
        np.testing.assert_array_equal(w, g["%s_ls%d_word" % (name, ls)])
        np.testing.assert_array_equal(nd, g["%s_ls%d_node" % (name, ls)])
        (bw, bv), (fn, ff) = V.transform(q, ls)
        np.testing.assert_array_equal(bw, g["%s_ls%d_bow_words" % (name, ls)])
        np.testing.assert_array_equal(bv, g["%s_ls%d_bow_values" % (name, ls)])       # bit-exact doubles
        if ls >= L:
            assert list(fn) == [0]                     # level L - levelsup <= 0: every feature is filed under the root


@pytest.mark.parametrize("name", CASES)
def test_invariants_and_host_assembly(name):
    g, vocab, q, L = load(name)
    V = O.Vocabulary(*vocab, L=L)
    w, wt, nd = V.transform_features(q, 1)
    parent, leaf = vocab[0], vocab[1]
    words_of_nodes = np.cumsum(leaf) - 1
    leaves = np.nonzero(leaf)[0]
    assert ((w >= 0) & (w < len(leaves))).all()
    np.testing.assert_array_equal(wt, vocab[3][leaves[w]])
    # the node at level L - 1 is an ancestor of (or equal to) the word's leaf
    for i in range(0, len(q), 7):
        n = leaves[w[i]]
        chain = [n]
        while parent[chain[-1]] >= 0:
            chain.append(parent[chain[-1]])
        assert nd[i] in chain
    (bw, bv), (fn, ff) = V.transform(q, 1)
    assert (np.diff(bw) > 0).all() and (np.diff(fn) > 0).all()
    assert abs(bv.sum() - 1.0) < 1e-12
    kept = np.sort(np.concatenate(ff))
    np.testing.assert_array_equal(kept, np.nonzero(wt > 0)[0])         # every non-stopped feature exactly once
    # the product's host assembly reproduces the oracle's maps from the per-feature arrays
    (pw, pv), (pn, pf) = orbx.assemble_bow(w, wt, nd)
    np.testing.assert_array_equal(pw, bw); np.testing.assert_array_equal(pv, bv)
    np.testing.assert_array_equal(pn, fn)
    assert all(np.array_equal(a, b) for a, b in zip(pf, ff))
    assert words_of_nodes[leaves[w[0]]] == w[0]


def test_first_child_wins_ties_and_invalid_tables():
    # two identical children: the first one in file order must win (strict '<' in the reference's scan)
    parent = np.array([-1, 0, 0, 0], np.int32); leaf = np.array([0, 1, 1, 1], np.uint8)
    desc = np.zeros((4, 32), np.uint8); desc[1] = 0xF0; desc[2] = 0x0F; desc[3] = 0x0F
    V = O.Vocabulary(parent, leaf, desc, np.array([0, 1.0, 2.0, 3.0]), L=1)
    w, wt, nd = V.transform_features(np.full((1, 32), 0x0F, np.uint8), 0)
    assert (w[0], wt[0], nd[0]) == (1, 2.0, 2)
    with pytest.raises(ValueError):
        O.Vocabulary(np.array([-1, 2, 0], np.int32), np.array([0, 1, 0], np.uint8), np.zeros((3, 32), np.uint8), np.zeros(3), L=2)


def _sbb_fixture():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "search_by_bow.npz"))
    vocab = tuple(g["voc_" + k] for k in ("parent", "leaf", "desc", "weight"))
    return g, O.Vocabulary(*vocab, L=4), vocab


@pytest.mark.parametrize("levelsup", [3, 2])
@pytest.mark.parametrize("mode,ratio", [(0, 0.7), (1, 0.8), (0, 0.95)])
def test_search_by_bow_oracle_matches_golden(levelsup, mode, ratio):
    """ORBmatcher::SearchByBoW (KeyFrame-Frame and KeyFrame-KeyFrame): the C oracle against the vectors written by the
    literal Python restatement in oracle/gen_golden.py (both with and without the rotation-consistency filter)."""
    g, V, _ = _sbb_fixture()
    k1, d1, k2, d2, v1, v2 = g["k1"], g["d1"], g["k2"], g["d2"], g["valid1"], g["valid2"]
    fv1 = V.transform(d1, levelsup)[1]; fv2 = V.transform(d2, levelsup)[1]
    tag = "ls%d_m%d_r%d" % (levelsup, mode, int(ratio * 100))
    for ori, suffix in ((True, ""), (False, "_noori")):
        n, m12 = O.search_by_bow(mode, k1, d1, v1, fv1, k2, d2, v2 if mode == 1 else None, fv2, ratio, ori)
        assert n == int(g[tag + suffix + "_n"])
        np.testing.assert_array_equal(m12, g[tag + suffix + "_m12"])
        assert n == (m12 >= 0).sum()
        got = m12[m12 >= 0]
        assert len(np.unique(got)) == len(got)                     # a set-2 feature is claimed at most once
        assert v1[m12 >= 0].all() and (mode == 0 or v2[got].all())


def py_search_for_triangulation(k1, d1, free1, st1, fv1, k2, d2, free2, st2, fv2, F, ep, scale2, sigma2, only_stereo, coarse, check_ori):
    """Literal restatement of ORBmatcher::SearchForTriangulation (ORBmatcher.cc:961-1202, pinhole, mpCamera2 == NULL) with
    np.float32 arithmetic; FeatureVectors as dicts."""
    f32 = np.float32
    F = np.asarray(F, f32).reshape(3, 3)
    bits1 = np.unpackbits(d1, axis=1); bits2 = np.unpackbits(d2, axis=1)
    m12 = [-1] * len(k1); rot = [[] for _ in range(30)]; n = 0
    f1 = dict(zip(fv1[0].tolist(), fv1[1])); f2 = dict(zip(fv2[0].tolist(), fv2[1]))
    for node in sorted(set(f1) & set(f2)):
        for i1 in f1[node]:
            if not free1[i1]:
                continue
            s1 = bool(st1[i1]) if st1 is not None else False
            if only_stereo and not s1:
                continue
            x1, y1 = f32(k1["x"][i1]), f32(k1["y"][i1])
            a = f32(f32(f32(x1 * F[0, 0]) + f32(y1 * F[1, 0])) + F[2, 0])
            b = f32(f32(f32(x1 * F[0, 1]) + f32(y1 * F[1, 1])) + F[2, 1])
            c = f32(f32(f32(x1 * F[0, 2]) + f32(y1 * F[1, 2])) + F[2, 2])
            best, bidx = 50, -1
            for i2 in f2[node]:
                if not free2[i2]:
                    continue
                s2 = bool(st2[i2]) if st2 is not None else False
                if only_stereo and not s2:
                    continue
                dist = int((bits1[i1] != bits2[i2]).sum())
                if dist > 50 or dist > best:
                    continue
                x2, y2, o2 = f32(k2["x"][i2]), f32(k2["y"][i2]), int(k2["octave"][i2])
                if not s1 and not s2:
                    ex, ey = f32(f32(ep[0]) - x2), f32(f32(ep[1]) - y2)
                    if f32(f32(ex * ex) + f32(ey * ey)) < f32(f32(100) * f32(scale2[o2])):
                        continue
                ok = coarse
                if not ok:
                    num = f32(f32(f32(a * x2) + f32(b * y2)) + c)
                    den = f32(f32(a * a) + f32(b * b))
                    if den != 0:
                        dsqr = f32(f32(num * num) / den)
                        ok = float(dsqr) < 3.84 * float(f32(sigma2[o2]))
                if ok:
                    best, bidx = dist, int(i2)
            if bidx >= 0:
                m12[i1] = bidx; n += 1
                if check_ori:
                    r = f32(f32(k1["angle"][i1]) - f32(k2["angle"][bidx]))
                    if r < 0:
                        r = f32(r + f32(360.0))
                    x = f32(r * f32(f32(1.0) / f32(30)))
                    bn = int(np.floor(x + f32(0.5)))
                    rot[0 if bn == 30 else bn].append(i1)
    if check_ori:
        sizes = [len(r) for r in rot]
        mx = [0, 0, 0]; ind = [-1, -1, -1]
        for i, s_ in enumerate(sizes):
            if s_ > mx[0]:
                mx = [s_, mx[0], mx[1]]; ind = [i, ind[0], ind[1]]
            elif s_ > mx[1]:
                mx = [mx[0], s_, mx[1]]; ind = [ind[0], i, ind[1]]
            elif s_ > mx[2]:
                mx[2] = s_; ind[2] = i
        if mx[1] < f32(0.1) * f32(mx[0]):
            ind[1] = ind[2] = -1
        elif mx[2] < f32(0.1) * f32(mx[0]):
            ind[2] = -1
        for i in range(30):
            if i not in ind:
                for i1 in rot[i]:
                    m12[i1] = -1; n -= 1
    return n, np.array(m12, np.int32)


def triangulation_case(seed=91):
    """two frames related by an image shift of (3, 2) px and the fundamental matrix of that 'translation' (K = I)"""
    W, H = 480, 360
    st = synth.rects_stream(W, H, 2, seed=seed)
    e = O.Extractor(600, 1.2, 8, 20, 7)
    _, k1, d1 = e(st[0], (0, 0)); _, k2, d2 = e(st[1], (0, 0))
    vocab = synth.random_vocabulary(k=10, L=4, seed=17)
    V = O.Vocabulary(*vocab, L=4)
    fv1 = V.transform(d1, 3)[1]; fv2 = V.transform(d2, 3)[1]
    t = np.array([3.0, 2.0, 0.0])
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]], np.float32)
    F12 = (-tx).astype(np.float32)                              # l2 = x1' F12 = [t]x x1
    free1 = (np.arange(len(k1)) % 4 != 0).astype(np.uint8); free2 = (np.arange(len(k2)) % 5 != 0).astype(np.uint8)
    st1 = (np.arange(len(k1)) % 3 == 0).astype(np.uint8); st2 = (np.arange(len(k2)) % 2 == 0).astype(np.uint8)
    return e, k1, d1, k2, d2, fv1, fv2, F12, free1, free2, st1, st2, vocab


@pytest.mark.parametrize("only_stereo,coarse,ep,use_stereo", [(False, False, (1e6, 1e6), False), (False, False, (240.0, 180.0), True),
                                                             (True, False, (1e6, 1e6), True), (False, True, (100.0, 100.0), False)])
def test_search_for_triangulation_oracle_vs_python(only_stereo, coarse, ep, use_stereo):
    e, k1, d1, k2, d2, fv1, fv2, F12, free1, free2, st1, st2, _ = triangulation_case()
    s1, s2 = (st1, st2) if use_stereo else (None, None)
    for ori in (True, False):
        n, m12 = O.search_for_triangulation(k1, d1, free1, s1, fv1, k2, d2, free2, s2, fv2, F12, ep, e.scale, e.sigma2, only_stereo, coarse, ori)
        pn, pm12 = py_search_for_triangulation(k1, d1, free1, s1, fv1, k2, d2, free2, s2, fv2, F12, ep, e.scale, e.sigma2, only_stereo, coarse, ori)
        assert n == pn
        np.testing.assert_array_equal(m12, pm12)
    if not only_stereo:
        assert n > 10
