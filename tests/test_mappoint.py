"""MapPoint::ComputeDistinctiveDescriptors (R/src/MapPoint.cc:448-524), batched: the C oracle against a numpy restatement
(CPU), and the CUDA path against the oracle (GPU)."""
import numpy as np
import pytest

from multi_orbslam3_b200 import synth
from oracle import oracle as O


def numpy_distinctive(desc, offsets):
    out = []
    for p in range(len(offsets) - 1):
        d = desc[offsets[p]:offsets[p + 1]]
        N = len(d)
        if N == 0:
            out.append(-1); continue
        bits = np.unpackbits(d, axis=1).astype(np.int32)
        D = (bits[:, None, :] != bits[None, :, :]).sum(2)
        med = np.sort(D, axis=1)[:, int(0.5 * (N - 1))]
        out.append(int(np.argmin(med)))                       # first minimum, as `median < BestMedian`
    return np.array(out, np.int32)


def make_points(seed, npoints, nmax):
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, nmax + 1, npoints); counts[0] = 0; counts[1] = 1; counts[2] = 2
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    desc = np.empty((offsets[-1], 32), np.uint8)
    for p in range(npoints):
        n = counts[p]
        if n:
            base = rng.integers(0, 256, 32, dtype=np.uint8)
            flips = np.zeros((n, 256), np.uint8)
            k = rng.integers(0, 40, n)
            for i in range(n):
                flips[i, rng.choice(256, k[i], replace=False)] = 1
            desc[offsets[p]:offsets[p + 1]] = base ^ np.packbits(flips, axis=1)
    return desc, offsets


def test_oracle_matches_numpy():
    desc, off = make_points(1, 120, 40)
    desc[off[5]:off[6]] = desc[off[5]]                        # all observations identical: every median is 0, index 0 wins
    got = O.distinctive_descriptors(desc, off)
    np.testing.assert_array_equal(got, numpy_distinctive(desc, off))
    assert got[0] == -1 and got[1] == 0 and got[5] == 0


@pytest.mark.gpu
def test_gpu_matches_oracle():
    from multi_orbslam3_b200 import orbx
    m = orbx.ORBmatcher(0.7, True, max_keypoints=1024)
    for seed, npoints, nmax in ((2, 500, 30), (3, 40, 300), (4, 3000, 12)):     # incl. points with more than 256 observations
        desc, off = make_points(seed, npoints, nmax)
        np.testing.assert_array_equal(m.DistinctiveDescriptors(desc, off), O.distinctive_descriptors(desc, off))
    e = np.zeros((0, 32), np.uint8)
    assert list(m.DistinctiveDescriptors(e, np.array([0, 0, 0], np.int32))) == [-1, -1]
    with pytest.raises(orbx.OrbxError):
        m.DistinctiveDescriptors(desc, np.array([0, 5, 3], np.int32))
    m.close()
