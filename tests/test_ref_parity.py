"""Parity pinned to the reference ITSELF: oracle/_ref is the reference's own source (R/src/ORBextractor.cc, R/src/ORBmatcher.cc,
R/src/CameraModels/Pinhole.cpp, DBoW2, and the Frame / KeyFrame / MapPoint bodies the front-end calls), compiled unmodified from
/root/reference by oracle/ref/Makefile.  These tests demand  _ref == the C restatement (oracle/orb_oracle.c)  byte for byte on the
BASELINE configs C1-C4, the committed goldens and the tie / steal-back cases; the `-m gpu` suites then compare the CUDA path with
both.  (oracle/_ref/ is built in the container that holds /root/reference and travels to the GPU box as a prebuilt library.)"""
import glob
import os

import numpy as np
import pytest

from multi_orbslam3_b200 import synth
from oracle import oracle as O
from oracle import ref as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built and /root/reference absent")

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def same_extraction(a, b):
    return a[0] == b[0] and a[1].tobytes() == b[1].tobytes() and a[2].tobytes() == b[2].tobytes()


def test_constructor_tables_equal():
    """ORBextractor::ORBextractor (R/src/ORBextractor.cc:408-468): scale tables, per-level quotas, umax."""
    for nf, sf, nl in ((1000, 1.2, 8), (1200, 1.2, 8), (2000, 1.2, 8), (5000, 1.2, 8), (500, 1.5, 5), (300, 1.1, 12), (7, 2.0, 3)):
        a, b = O.Extractor(nf, sf, nl, 20, 7), R.Extractor(nf, sf, nl, 20, 7)
        for name in ("scale", "inv_scale", "sigma2", "inv_sigma2", "features_per_level", "umax"):
            np.testing.assert_array_equal(getattr(a, name), getattr(b, name), err_msg=name)


@pytest.mark.parametrize("w,h,nf,lap", [(752, 480, 1000, (0, 1000)), (752, 480, 5000, (0, 1000)), (752, 480, 1200, (0, 0)),
                                         (1241, 376, 2000, (0, 0)), (640, 480, 1000, (0, 1000)), (376, 240, 500, (100, 250))],
                         ids=["C1_mono", "C1_init5000", "C2_stereo", "C3_kitti", "C4_tum", "lapping_area"])
def test_extractor_equals_reference(w, h, nf, lap):
    """operator() (R/src/ORBextractor.cc:1068-1150) on S-rects and S-noise frames: monoIndex, every keypoint field (angles
    included) and every descriptor byte of the reference == the C restatement; pyramid levels too."""
    a, b = O.Extractor(nf, 1.2, 8, 20, 7), R.Extractor(nf, 1.2, 8, 20, 7)
    frames = list(synth.rects_stream(w, h, 3, seed=7)) + [synth.noise_frame(w, h, seed=3)]
    for f in frames:
        ra, rb = a(f, lap), b(f, lap)
        assert len(ra[1]) > nf // 2
        assert same_extraction(ra, rb)
        for l in range(8):
            np.testing.assert_array_equal(a.level_image(l), b.level_image(l))
        for l in (0, 3, 7):
            assert a.level_keypoints(l).tobytes() == b.octree_keypoints(f, l).tobytes()


def test_extractor_edge_cases_equal_reference():
    """empty image -> -1; constant image -> no keypoints; tiny pyramids; minThFAST fallback-only frames"""
    a, b = O.Extractor(300, 1.2, 8, 20, 7), R.Extractor(300, 1.2, 8, 20, 7)
    assert a(np.zeros((0, 0), np.uint8))[0] == -1 and b(np.zeros((0, 0), np.uint8))[0] == -1
    flat = np.full((240, 320), 97, np.uint8)
    assert same_extraction(a(flat), b(flat)) and len(b(flat)[1]) == 0
    rng = np.random.default_rng(5)
    low = (128 + rng.integers(-12, 13, (200, 260))).astype(np.uint8)          # only minThFAST (7) fires
    ra, rb = a(low), b(low)
    assert same_extraction(ra, rb) and len(ra[1]) > 0
    small = synth.rects_stream(160, 120, 1, seed=2)[0]
    assert same_extraction(a(small), b(small))


def test_goldens_equal_reference():
    """every committed extraction fixture (tests/golden/extract_*.npz) is reproduced by the reference's own code"""
    files = sorted(glob.glob(os.path.join(GOLD, "extract_*.npz")))
    assert files
    for path in files:
        g = np.load(path)
        p = g["params"]
        r = R.Extractor(int(p[0]), float(p[1]), int(p[2]), int(p[3]), int(p[4]))
        mono, k, d = r(g["image"], tuple(int(v) for v in g["lapping"]))
        assert mono == int(g["mono_index"]), path
        assert k.tobytes() == g["keypoints"].tobytes(), path
        np.testing.assert_array_equal(d, g["descriptors"], err_msg=path)


def test_octree_tie_rule_is_the_monotonic_allocator():
    """R/src/ORBextractor.cc:682 sorts equal-sized nodes by heap address.  Under a monotonic arena (address order =
    creation order) the reference's DistributeOctTree equals the canonical rule of the oracle and the CUDA kernel on every input;
    on glibc malloc it keeps a different (allocator-dependent) set.  Both facts are asserted / reported."""
    arena, heap = R.Extractor(1000, 1.2, 8, 20, 7), R.Extractor(1000, 1.2, 8, 20, 7, variant="malloc")
    assert R.lib("arena").ref_uses_arena() == 1 and R.lib("malloc").ref_uses_arena() == 0
    rng = np.random.default_rng(11)
    differs = 0
    cases = 0
    for trial in range(40):
        n = int(rng.integers(50, 6000))
        H = int(rng.integers(60, 500)); W = int(rng.integers(H, 4 * H))      # landscape levels, as every camera of the reference has (nIni = round(W / H) >= 1)
        pts = np.stack([rng.integers(0, W, n), rng.integers(0, H, n), rng.integers(1, 200, n)], 1).astype(np.float32)
        pts = pts[np.lexsort((pts[:, 0], pts[:, 1]))]
        N = int(rng.integers(5, 1200))
        want = O.distribute_octree(pts, 0, W, 0, H, N)
        got = arena.distribute_octree(pts, 0, W, 0, H, N)
        np.testing.assert_array_equal(got, want)
        other = heap.distribute_octree(pts, 0, W, 0, H, N)
        cases += 1
        differs += int(other.shape != want.shape or not np.array_equal(other, want))
    print("DistributeOctTree on glibc malloc differs from the canonical rule in %d of %d random cases" % (differs, cases))
    # the largest-first phase runs on most inputs, so heap order matters on most inputs: the nondeterminism is real
    assert differs > 0


def test_hamming_and_grid_equal_reference():
    rng = np.random.default_rng(0)
    d = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    for i in range(0, 200, 2):
        assert R.hamming256(d[i], d[i + 1]) == O.hamming256(d[i], d[i + 1]) == int(np.unpackbits(d[i] ^ d[i + 1]).sum())
    ex = O.Extractor(1000, 1.2, 8, 20, 7)
    _, k, _ = ex(synth.rects_stream(752, 480, 1, seed=3)[0], (0, 1000))
    bounds = (0.0, 752.0, 0.0, 480.0)
    for _ in range(200):
        x, y = float(rng.uniform(-50, 800)), float(rng.uniform(-50, 530))
        r = float(rng.choice([3.0, 10.0, 15.5, 40.0, 100.0]))
        lv = [(-1, -1), (0, 0), (0, 2), (3, 7), (2, -1)][int(rng.integers(0, 5))]
        np.testing.assert_array_equal(R.features_in_area(k, bounds, x, y, r, *lv), O.features_in_area(k, bounds, x, y, r, *lv))


@pytest.mark.parametrize("window,ratio,ori", [(100, 0.9, True), (30, 0.9, True), (100, 0.6, False), (200, 1.0, True)])
def test_search_for_initialization_equals_reference(window, ratio, ori):
    """ORBmatcher::SearchForInitialization (R/src/ORBmatcher.cc:702-817) on consecutive stream frames, two calls in a row with the
    vbPrevMatched the first one updated; plus a descriptor set with planted duplicates (ties, steal-back)."""
    ex = O.Extractor(1000, 1.2, 8, 20, 7)
    fr = synth.rects_stream(752, 480, 3, seed=21)
    e = [ex(f, (0, 1000)) for f in fr]
    bounds = (0.0, 752.0, 0.0, 480.0)
    prev = np.stack([e[0][1]["x"], e[0][1]["y"]], 1)
    for t in (1, 2):
        a = O.search_for_initialization(e[0][1], e[0][2], e[t][1], e[t][2], bounds, prev, window, ratio, ori)
        b = R.search_for_initialization(e[0][1], e[0][2], e[t][1], e[t][2], bounds, prev, window, ratio, ori)
        assert a[0] == b[0] and a[0] > 0
        np.testing.assert_array_equal(a[1], b[1]); np.testing.assert_array_equal(a[2], b[2])
        prev = b[2]
    # ties: frame 2 = frame 1 with every third descriptor duplicated onto its neighbour
    k1, d1 = e[0][1], e[0][2]
    d2 = d1.copy(); d2[1::3] = d2[0:-1:3][:len(d2[1::3])]
    a = O.search_for_initialization(k1, d1, k1, d2, bounds, np.stack([k1["x"], k1["y"]], 1), window, ratio, ori)
    b = R.search_for_initialization(k1, d1, k1, d2, bounds, np.stack([k1["x"], k1["y"]], 1), window, ratio, ori)
    assert a[0] == b[0]
    np.testing.assert_array_equal(a[1], b[1])


def test_search_init_golden_equals_reference():
    g = np.load(os.path.join(GOLD, "search_init_320x240.npz"))
    prev0 = np.stack([g["k1"]["x"], g["k1"]["y"]], 1)
    n, m12, prev = R.search_for_initialization(g["k1"], g["d1"], g["k2"], g["d2"], (0, 320, 0, 240), prev0, 100, 0.9, True)
    assert n == int(g["nmatches"])
    np.testing.assert_array_equal(m12, g["matches12"]); np.testing.assert_array_equal(prev, g["prev"])


@pytest.mark.parametrize("shape,nf,disp,mb,mbf", [((752, 480), 1200, 14, 0.11, 47.9), ((1241, 376), 2000, 23, 0.54, 386.1)], ids=["C2_euroc", "C3_kitti"])
def test_compute_stereo_matches_equals_reference(shape, nf, disp, mb, mbf):
    """Frame::ComputeStereoMatches (R/src/Frame.cc:785-962, descriptor search + SAD refinement + outlier cut) through the reference's
    own extractors: mvuRight / mvDepth bit-equal to the C restatement"""
    W, H = shape
    L, Rimg = synth.stereo_pair(W, H, seed=8, disparity=disp)
    ol, orr = O.Extractor(nf, 1.2, 8, 20, 7), O.Extractor(nf, 1.2, 8, 20, 7)
    rl, rr = R.Extractor(nf, 1.2, 8, 20, 7), R.Extractor(nf, 1.2, 8, 20, 7)
    _, kl, dl = ol(L, (0, 0)); _, kr, dr = orr(Rimg, (0, 0))
    rl(L, (0, 0)); rr(Rimg, (0, 0))                     # the reference extractors hold THIS pair's pyramids (mvImagePyramid)
    u0, z0, _ = O.compute_stereo_matches(ol, orr, kl, dl, kr, dr, mb, mbf)
    u1, z1 = R.compute_stereo_matches(rl, rr, kl, dl, kr, dr, mb, mbf)
    assert (u0 >= 0).sum() > 100
    assert u0.tobytes() == u1.tobytes() and z0.tobytes() == z1.tobytes()


def test_bow_equals_reference():
    """DBoW2 transform (TemplatedVocabulary.h:1127-1259), BowVector / FeatureVector, and both SearchByBoW overloads
    (R/src/ORBmatcher.cc:269-471, 819-959) against the reference's own DBoW2 + ORBmatcher."""
    vocab = synth.random_vocabulary(k=10, L=3, seed=4)
    ov, rv = O.Vocabulary(*vocab, L=3), R.Vocabulary(*vocab, L=3, k=10)
    ex = O.Extractor(1000, 1.2, 8, 20, 7)
    fr = synth.rects_stream(640, 480, 2, seed=33)
    (_, k1, d1), (_, k2, d2) = ex(fr[0], (0, 0)), ex(fr[1], (0, 0))
    for levelsup in (0, 1, 2, 4):
        (bw0, bv0), (fn0, ff0) = ov.transform(d1, levelsup)
        (bw1, bv1), (fn1, ff1) = rv.transform(d1, levelsup)
        np.testing.assert_array_equal(bw0, bw1); assert bv0.tobytes() == bv1.tobytes()
        np.testing.assert_array_equal(fn0, fn1)
        assert all(np.array_equal(a, b) for a, b in zip(ff0, ff1)) and len(ff0) == len(ff1)
    w0, wt0, n0 = ov.transform_features(d1[:200], 2)
    w1, wt1, n1 = rv.transform_features(d1[:200], 2)
    np.testing.assert_array_equal(w0, w1); np.testing.assert_array_equal(n0, n1); assert wt0.tobytes() == wt1.tobytes()
    fv1, fv2 = ov.transform(d1, 2)[1], ov.transform(d2, 2)[1]
    rng = np.random.default_rng(9)
    valid1 = rng.random(len(k1)) < 0.7; valid2 = rng.random(len(k2)) < 0.8
    for mode, v2, ratio, ori in ((0, None, 0.75, True), (0, None, 0.9, False), (1, valid2, 0.8, True), (1, valid2, 0.9, False)):
        a = O.search_by_bow(mode, k1, d1, valid1, fv1, k2, d2, v2, fv2, ratio, ori)
        b = R.search_by_bow(mode, k1, d1, valid1, fv1, k2, d2, v2, fv2, ratio, ori)
        assert a[0] == b[0] and a[0] > 0
        np.testing.assert_array_equal(a[1], b[1])


def test_distinctive_descriptors_equal_reference():
    """MapPoint::ComputeDistinctiveDescriptors (R/src/MapPoint.cc:448-524)"""
    rng = np.random.default_rng(13)
    counts = rng.integers(0, 12, 60); counts[:4] = (0, 1, 2, 3)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    base = rng.integers(0, 256, (len(counts), 32), dtype=np.uint8)
    desc = np.repeat(base, counts, axis=0)
    flip = rng.random((len(desc), 256)) < 0.08
    desc ^= np.packbits(flip, axis=1, bitorder="little")
    np.testing.assert_array_equal(R.distinctive_descriptors(desc, offsets), O.distinctive_descriptors(desc, offsets))
