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
