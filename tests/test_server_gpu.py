"""GPU test of the sharded DB on one device (world size 1 and simulated shards) against the oracle."""
import numpy as np
import pytest
import torch

from multi_orbslam3_b200 import orbx, synth
from multi_orbslam3_b200.server import ShardedDescriptorDB, gpu_fns
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_partial_top2_merge_equals_full_search():
    m = orbx.ORBmatcher(0.7, True, max_keypoints=2048)
    match_fn, merge_fn = gpu_fns(m)
    db = synth.random_descriptors(40000, 5, 0.2); q = synth.random_descriptors(500, 6)
    db[5] = q[3]; db[39999] = q[3]; db[20000] = q[3]
    want_i, want_d = O.bf_knn2(q, db)
    dq = torch.from_numpy(q).cuda(); ddb = torch.from_numpy(db).cuda()
    # 8 simulated shards -> partial tables -> merge kernel
    parts = [match_fn(dq, ddb[r * 5000:(r + 1) * 5000].contiguous(), r * 5000) for r in range(8)]
    pi = torch.stack([p[0] for p in parts]); pd = torch.stack([p[1] for p in parts])
    gi, gd = merge_fn(pi, pd)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(gi.cpu().numpy(), want_i)
    np.testing.assert_array_equal(gd.cpu().numpy(), want_d)
    sdb = ShardedDescriptorDB(ddb, match_fn, merge_fn)
    for fn in (sdb.knn2_allgather_top2, sdb.knn2_allgather_db):
        i2, d2 = fn(dq)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(i2.cpu().numpy(), want_i)
        np.testing.assert_array_equal(d2.cpu().numpy(), want_d)
    m.close()
