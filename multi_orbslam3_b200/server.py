"""Server-side cross-agent descriptor matching over a keyframe-descriptor DB sharded across the GPUs of one box
(BASELINE config C5).

The reference's server matches keyframes of different agents through BoW buckets (LoopClosing.cc:605-657); the
north star asks for brute-force Hamming kNN-2 against the whole DB with cv::BFMatcher semantics (Frame.cc:1127-1137:
top-2 by (distance, trainIdx)).  The DB is sharded by owning agent = GPU (rank r holds global rows
[r*shard, (r+1)*shard)).  Two exchange strategies, identical results:

  * allgather_db  (the north star's wording): all-gather the descriptor shards over NVLink (32 B per descriptor), then
    every rank matches its own queries against the full DB.
  * allgather_top2 (less traffic): every rank matches ALL queries against its local shard, then all-gathers only the
    per-query partial top-2 tables (16 B per query and rank) and merges them by (distance, global index).

The device work (orbx_bf_knn2_device / orbx_knn2_merge_device) is injected as `match_fn` / `merge_fn` so that the
collective and index logic can be exercised on CPU with gloo (tests/test_server_dist.py injects the oracle).
"""
import torch
import torch.distributed as dist


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _all_gather(t, group):
    w, _ = _world(group)
    if w == 1:
        return t.unsqueeze(0).clone()
    out = torch.empty((w * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)   # concatenated along dim 0
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out.view((w,) + tuple(t.shape))


class ShardedDescriptorDB:
    """shard: uint8 tensor [n_local, 32]; every rank allocates the same n_local rows (the all-gather needs equal shapes) and says
    how many of them are real with n_valid (default: all).  Rows beyond n_valid are never matched: padding with copies of a real
    row would plant duplicates that become the second-best match and defeat the ratio test.  Global row index = rank * n_local +
    local row."""

    def __init__(self, shard, match_fn, merge_fn, group=None, n_valid=None):
        assert shard.dtype == torch.uint8 and shard.dim() == 2 and shard.shape[1] == 32
        self.shard, self.match_fn, self.merge_fn, self.group = shard.contiguous(), match_fn, merge_fn, group
        self.world, self.rank = _world(group)
        self.n_local = shard.shape[0]
        self.n_valid = self.n_local if n_valid is None else int(n_valid)
        assert 0 <= self.n_valid <= self.n_local
        counts = torch.tensor([self.n_valid], dtype=torch.int64, device=shard.device)
        self.valid_counts = [int(v) for v in _all_gather(counts, group).reshape(-1).tolist()]     # every rank's n_valid

    @classmethod
    def from_kfdb(cls, kfdb, match_fn, merge_fn, group=None):
        """The shard IS the device storage of an orbx.KeyframeDB (keyframes ingested from KF.msg byte runs):
        no copy, n_valid = the rows ingested so far (every rank allocates the same capacity).  Call again after further ingests."""
        kfdb.sync()
        rows, _, _ = kfdb.size()
        dd, _ = kfdb.device_views()
        shard = torch.as_tensor(dd, device=torch.device("cuda", kfdb.device))
        db = cls(shard, match_fn, merge_fn, group=group, n_valid=rows)
        db.kfdb = kfdb
        return db

    def knn2_allgather_top2(self, queries):
        """queries: the same uint8 [nq, 32] tensor on every rank (the query keyframe is broadcast by its owner).
        Returns (idx [nq,2] global row indices, dist [nq,2])."""
        idx, d = self.match_fn(queries, self.shard[:self.n_valid], self.rank * self.n_local)      # partial top-2 on the local shard
        parts_i = _all_gather(idx, self.group)                                      # [world, nq, 2]
        parts_d = _all_gather(d, self.group)
        return self.merge_fn(parts_i, parts_d)

    def knn2_allgather_db(self, queries):
        """all-gather the shards (world * n_local * 32 bytes into every rank), then every rank matches ITS slice of the
        queries against the full DB and the per-slice results are all-gathered (nq is padded to a multiple of world)."""
        full = _all_gather(self.shard, self.group)                   # [world, n_local, 32]
        nq = queries.shape[0]
        per = (nq + self.world - 1) // self.world
        lo = min(self.rank * per, nq); hi = min(lo + per, nq)
        mine = queries[lo:hi]
        if hi - lo < per:                                            # pad the last slice with copies of query 0 (QUERIES: harmless, cut off below)
            mine = torch.cat([mine, queries[:1].expand(per - (hi - lo), 32)], 0)
        mine = mine.contiguous()
        if all(v == self.n_local for v in self.valid_counts):
            idx, d = self.match_fn(mine, full.reshape(-1, 32), 0)
        else:                                                        # ragged shards: one pass per owner over its real rows, then the merge
            parts = [self.match_fn(mine, full[r, :self.valid_counts[r]], r * self.n_local) for r in range(self.world)]
            idx, d = self.merge_fn(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
        idx = _all_gather(idx, self.group).reshape(-1, 2)[:nq]
        d = _all_gather(d, self.group).reshape(-1, 2)[:nq]
        return idx.contiguous(), d.contiguous()

    def broadcast_queries(self, queries, src):
        if self.world > 1:
            dist.broadcast(queries, src=src, group=self.group)
        return queries


def gpu_fns(matcher):
    """match_fn / merge_fn backed by the CUDA kernels of an orbx.ORBmatcher (tensors on its device)."""
    import ctypes as C
    from . import orbx

    def match_fn(q, t, idx_base):
        nq = q.shape[0]
        idx = torch.empty((nq, 2), dtype=torch.int32, device=q.device)
        d = torch.empty((nq, 2), dtype=torch.int32, device=q.device)
        s = torch.cuda.current_stream(q.device).cuda_stream
        orbx._check(orbx.lib().orbx_bf_knn2_device(matcher._h, C.c_void_p(q.data_ptr()), nq, C.c_void_p(t.data_ptr()), t.shape[0],
                                                   C.c_void_p(idx.data_ptr()), C.c_void_p(d.data_ptr()), int(idx_base), orbx._s(s)))
        return idx, d

    def merge_fn(parts_i, parts_d):
        nparts, nq = parts_i.shape[0], parts_i.shape[1]
        idx = torch.empty((nq, 2), dtype=torch.int32, device=parts_i.device)
        d = torch.empty((nq, 2), dtype=torch.int32, device=parts_i.device)
        s = torch.cuda.current_stream(parts_i.device).cuda_stream
        orbx._check(orbx.lib().orbx_knn2_merge_device(matcher._h, C.c_void_p(parts_i.data_ptr()), C.c_void_p(parts_d.data_ptr()), nparts, nq,
                                                      C.c_void_p(idx.data_ptr()), C.c_void_p(d.data_ptr()), orbx._s(s)))
        return idx, d

    return match_fn, merge_fn
