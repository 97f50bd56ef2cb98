"""world_size-2 gloo test (CPU) of the sharded server DB protocol: offsets, all-gather layout and merge give the
same top-2 as a single-process brute force.  The per-shard matcher is the oracle here; on GPUs it is the CUDA kernel
(tests/test_server_gpu.py)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_orbslam3_b200 import synth
from multi_orbslam3_b200.server import ShardedDescriptorDB
from oracle import oracle as O


def oracle_match(q, t, base):
    idx, d = O.bf_knn2(q.numpy(), t.numpy())
    idx = np.where(idx >= 0, idx + base, -1).astype(np.int32)
    return torch.from_numpy(idx), torch.from_numpy(d)


def numpy_merge(parts_i, parts_d):
    pi = parts_i.numpy().transpose(1, 0, 2).reshape(parts_i.shape[1], -1)
    pd = parts_d.numpy().transpose(1, 0, 2).reshape(parts_d.shape[1], -1)
    pd = np.where(pi < 0, np.iinfo(np.int32).max, pd)
    order = np.lexsort((pi, pd), axis=1)[:, :2]
    idx = np.take_along_axis(pi, order, 1); d = np.take_along_axis(pd, order, 1)
    d = np.where(idx < 0, -1, d)
    return torch.from_numpy(idx.astype(np.int32)), torch.from_numpy(d.astype(np.int32))


def worker(rank, world, port, db, q, ret, ragged=0):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_local = len(db) // world
    shard = torch.from_numpy(db[rank * n_local:(rank + 1) * n_local].copy())
    n_valid = n_local - ragged * (rank + 1)                 # ragged: the tail rows of every shard are padding (garbage = query 0!)
    if ragged:
        shard[n_valid:] = torch.from_numpy(q[0])
    sdb = ShardedDescriptorDB(shard, oracle_match, numpy_merge, n_valid=n_valid)
    queries = torch.from_numpy(q.copy()) if rank == 0 else torch.zeros_like(torch.from_numpy(q))
    sdb.broadcast_queries(queries, src=0)
    a = sdb.knn2_allgather_top2(queries)
    b = sdb.knn2_allgather_db(queries)
    if rank == 1:
        ret["a"] = (a[0].numpy(), a[1].numpy()); ret["b"] = (b[0].numpy(), b[1].numpy())
    dist.destroy_process_group()


def test_sharded_db_two_ranks_gloo():
    db = synth.random_descriptors(600, 3, 0.3)
    q = synth.random_descriptors(40, 4)
    db[10] = q[0]; db[450] = q[0]; db[299] = q[1]; db[300] = q[1]       # ties inside and across shards
    want_i, want_d = O.bf_knn2(q, db)
    mgr = mp.Manager(); ret = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(worker, args=(2, port, db, q, ret), nprocs=2, join=True)
    for k in ("a", "b"):
        np.testing.assert_array_equal(ret[k][0], want_i)
        np.testing.assert_array_equal(ret[k][1], want_d)


def test_ragged_shards_ignore_padding_rows():
    """shards with fewer real rows than allocated: the padding (deliberately filled with copies of a query) must never be matched"""
    db = synth.random_descriptors(600, 13, 0.2)
    q = synth.random_descriptors(24, 14)
    ragged = 7
    n_local = 300
    keep = np.ones(600, bool)
    for r in range(2):
        keep[r * n_local + n_local - ragged * (r + 1):(r + 1) * n_local] = False
    real = np.nonzero(keep)[0]
    wi, wd = O.bf_knn2(q, db[real])
    want_i = np.where(wi >= 0, real[np.maximum(wi, 0)], -1).astype(np.int32)          # back to global row numbers
    mgr = mp.Manager(); ret = mgr.dict()
    port = 29500 + ((os.getpid() + 77) % 500)
    mp.spawn(worker, args=(2, port, db, q, ret, ragged), nprocs=2, join=True)
    for k in ("a", "b"):
        np.testing.assert_array_equal(ret[k][0], want_i)
        np.testing.assert_array_equal(ret[k][1], wd)
