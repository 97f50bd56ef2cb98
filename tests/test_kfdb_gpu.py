"""keyframes arrive as KF.msg byte runs (R/msg/KF.msg:24-29, CvKeyPoint.msg = 15 packed bytes, Descriptor.msg =
uint8[32]) and land in the GPU-resident shard that the server's brute-force search reads (include/orbx.h, orbx_kfdb_*).
Oracle: the wire records by oracle.keypoints_to_msg / keypoints_from_msg (pinned to a packed numpy dtype in test_frame_ops.py),
the search by oracle.bf_knn2 over the concatenated descriptors."""
import threading

import numpy as np
import pytest
import torch

from multi_orbslam3_b200 import orbx, synth
from multi_orbslam3_b200.server import ShardedDescriptorDB, gpu_fns
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _keyframes(ex, nkf, W, H, seed):
    frames = synth.rects_stream(W, H, nkf, seed=seed)
    out = []
    for f in range(nkf):
        _, k, d = ex(frames[f], None, (0, 0))
        out.append((k.copy(), d.copy()))
    return out


def test_kf_msgs_land_in_the_device_shard_and_are_searchable():
    W, H = 640, 480
    ex = orbx.ORBextractor(800, 1.2, 8, 20, 7, max_width=W, max_height=H)
    m = orbx.ORBmatcher(0.7, True, max_keypoints=2048)
    kfs = _keyframes(ex, 6, W, H, 11)
    db = orbx.KeyframeDB(capacity_rows=8000, max_keyframes=16)
    firsts, ids = [], [70, 3, 41, 9000000000, 5, 6]
    for kf_id, (k, d) in zip(ids[:5], kfs[:5]):
        firsts.append(db.ingest_msg(kf_id, O.keypoints_to_msg(k), d))                 # the bytes a KF.msg carries
    assert db.ingest_msg(99, np.zeros((0, 15), np.uint8), np.zeros((0, 32), np.uint8)) == sum(len(k) for k, _ in kfs[:5])   # an empty keyframe
    # the sixth keyframe is extracted on this GPU: slot -> DB without a host hop
    _, k5, d5 = ex(synth.rects_stream(W, H, 6, seed=11)[5], None, (0, 0))
    firsts.append(db.ingest_slot(ids[5], ex, 0))
    rows, nkf, cap = db.size()
    assert nkf == 7 and cap == 8000 and rows == sum(len(k) for k, _ in kfs)
    assert firsts == list(np.cumsum([0] + [len(k) for k, _ in kfs[:5]]))
    db.sync()
    dd, dk = db.device_views()
    t_desc = torch.as_tensor(dd, device="cuda")[:rows].cpu().numpy()
    t_kps = torch.as_tensor(dk, device="cuda")[:rows].cpu().numpy().view(orbx.KP_DTYPE).reshape(-1)
    all_desc = np.concatenate([d for _, d in kfs])
    np.testing.assert_array_equal(t_desc, all_desc)
    # keyframes that came over the wire hold the msg-quantised keypoints (u8 size / response, class_id -1), the local one is untouched
    want_k = np.concatenate([O.keypoints_from_msg(O.keypoints_to_msg(k)) for k, _ in kfs[:5]] + [kfs[5][0]])
    assert t_kps.tobytes() == want_k.tobytes()
    # search: every descriptor of keyframe 2 as a query against the whole DB
    q = kfs[2][1]
    gi, gd = db.knn2(m, q, idx_base=100000)
    wi, wd = O.bf_knn2(q, all_desc)
    np.testing.assert_array_equal(gi, wi + 100000); np.testing.assert_array_equal(gd, wd)
    kf, ft = db.locate(gi[:, 0] - 100000)
    assert (kf == ids[2]).all() and np.array_equal(ft, np.arange(len(q)))             # each query finds itself at distance 0
    kf, ft = db.locate([0, firsts[3], firsts[3] - 1, rows - 1, rows, -1])
    assert kf.tolist() == [70, 9000000000, 41, 6, -1, -1] and ft.tolist() == [0, 0, len(kfs[2][0]) - 1, len(kfs[5][0]) - 1, -1, -1]
    # the sharded multi-GPU search runs on the same storage (world size 1 here; tests/test_server_nccl_gpu.py covers ranks)
    match_fn, merge_fn = gpu_fns(m)
    sdb = ShardedDescriptorDB.from_kfdb(db, match_fn, merge_fn)
    assert sdb.n_valid == rows and sdb.shard.data_ptr() == torch.as_tensor(dd, device="cuda").data_ptr()
    for fn in (sdb.knn2_allgather_top2, sdb.knn2_allgather_db):
        i2, d2 = fn(torch.from_numpy(q).cuda())
        torch.cuda.synchronize()
        np.testing.assert_array_equal(i2.cpu().numpy(), wi); np.testing.assert_array_equal(d2.cpu().numpy(), wd)
    db.close(); ex.close(); m.close()


def test_capacity_and_concurrent_ingest():
    db = orbx.KeyframeDB(capacity_rows=4096, max_keyframes=64)
    rng = np.random.default_rng(3)
    msgs = [(i, rng.integers(0, 256, (100, 15), dtype=np.uint8), rng.integers(0, 256, (100, 32), dtype=np.uint8)) for i in range(40)]
    firsts = {}

    def client(part):                                  # one communication thread per client, as Communicator.cc:110-148
        for i, k, d in part:
            firsts[i] = db.ingest_msg(i, k, d)
    ths = [threading.Thread(target=client, args=(msgs[c::4],)) for c in range(4)]
    [t.start() for t in ths]; [t.join() for t in ths]
    rows, nkf, _ = db.size()
    assert rows == 4000 and nkf == 40 and sorted(firsts.values()) == list(range(0, 4000, 100))
    db.sync()
    dd, dk = db.device_views()
    desc = torch.as_tensor(dd, device="cuda").cpu().numpy()
    kps = torch.as_tensor(dk, device="cuda").cpu().numpy().view(orbx.KP_DTYPE).reshape(-1)
    for i, k, d in msgs:
        np.testing.assert_array_equal(desc[firsts[i]:firsts[i] + 100], d)
        assert kps[firsts[i]:firsts[i] + 100].tobytes() == O.keypoints_from_msg(k).tobytes()
    with pytest.raises(orbx.OrbxError) as e:           # 97 rows are left: a 100-row keyframe does not fit and nothing is appended
        db.ingest_msg(1000, msgs[0][1], msgs[0][2])
    assert e.value.code == orbx.ORBX_E_CAPACITY and db.size()[0] == 4000
    db.close()
