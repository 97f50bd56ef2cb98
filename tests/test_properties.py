"""Property tests of the oracle with hypothesis: popcount against numpy bit unpacking, brute-force kNN
order, octree selection invariants, fixed-point image primitives on constant images, distinctive-descriptor optimality."""
import numpy as np
from hypothesis import given, settings, strategies as st
from hypothesis.extra import numpy as hnp

from oracle import oracle as O

desc_rows = lambda lo, hi: hnp.arrays(np.uint8, st.tuples(st.integers(lo, hi), st.just(32)))


@settings(max_examples=40, deadline=None, derandomize=True)
@given(desc_rows(1, 40), desc_rows(1, 40))
def test_hamming_and_knn2(q, t):
    D = (np.unpackbits(q, axis=1)[:, None, :] != np.unpackbits(t, axis=1)[None, :, :]).sum(2)
    assert O.hamming256(q[0], t[0]) == D[0, 0]
    idx, dist = O.bf_knn2(q, t)
    for i in range(len(q)):
        order = np.lexsort((np.arange(len(t)), D[i]))           # by distance, ties by the lowest train index
        assert idx[i, 0] == order[0] and dist[i, 0] == D[i, order[0]]
        if len(t) > 1:
            assert idx[i, 1] == order[1] and dist[i, 1] == D[i, order[1]]
        else:
            assert idx[i, 1] == -1


@settings(max_examples=30, deadline=None, derandomize=True)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 600), st.integers(1, 120))
def test_octree_selection_invariants(seed, n, N):
    rng = np.random.default_rng(seed)
    W, H = 720, 448
    xy = np.unique(np.stack([rng.integers(0, W, n), rng.integers(0, H, n)], 1), axis=0)
    resp = rng.integers(7, 255, len(xy))
    cand = np.concatenate([xy, resp[:, None]], 1).astype(np.float32)
    out = O.distribute_octree(cand, 0, W, 0, H, N)
    out2 = O.distribute_octree(cand, 0, W, 0, H, N)
    assert out.tobytes() == out2.tobytes()                                        # deterministic
    have = {tuple(r) for r in cand.tolist()}
    assert all(tuple(r) in have for r in out.tolist())                             # a subset of the candidates
    assert len({(r[0], r[1]) for r in out.tolist()}) == len(out)                   # at most one keypoint per node
    assert len(out) <= max(N, 2) + 3 * 2 + 2                                       # N plus the last expansion's surplus (:736)
    assert len(out) >= min(len(cand), 1)
    # NB: even far below the quota a candidate can be dropped: the reference stops as soon as a sweep does not increase the
    # node count (ORBextractor.cc:667), e.g. when two points share the only non-empty child of every multi-point node


@settings(max_examples=25, deadline=None, derandomize=True)
@given(st.integers(0, 255), st.integers(8, 70), st.integers(8, 70), st.integers(8, 60), st.integers(8, 60))
def test_constant_images_stay_constant(v, sw, sh, dw, dh):
    img = np.full((sh, sw), v, np.uint8)
    assert (O.resize_linear(img, dw, dh) == v).all()                               # coefficients sum to 2048 on both axes
    assert (O.gaussian_blur7(img) == v).all()                                      # kernel sums to 256


@settings(max_examples=30, deadline=None, derandomize=True)
@given(desc_rows(1, 24))
def test_distinctive_descriptor_is_a_median_minimiser(d):
    off = np.array([0, len(d)], np.int32)
    best = int(O.distinctive_descriptors(d, off)[0])
    bits = np.unpackbits(d, axis=1).astype(np.int32)
    D = (bits[:, None, :] != bits[None, :, :]).sum(2)
    med = np.sort(D, axis=1)[:, int(0.5 * (len(d) - 1))]
    assert med[best] == med.min() and best == int(np.argmin(med))
