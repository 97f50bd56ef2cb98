"""Worker of tests/test_server_nccl_gpu.py: launched by torchrun, one rank per GPU, NCCL.  Both C5 exchanges (all-gather of the
partial top-2 tables, all-gather of the descriptor shards) on a DB small enough for the CPU oracle, with ties planted inside and
across shards and ragged shards; every rank compares its merged top-2 with the oracle's single-process answer."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from multi_orbslam3_b200 import orbx, synth  # noqa: E402
from multi_orbslam3_b200.server import ShardedDescriptorDB, gpu_fns  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_local = 6000
    db = synth.random_descriptors(n_local * world, 31, 0.2)
    q = synth.random_descriptors(300, 32)
    # ties: the same descriptor in two shards, at a shard boundary, and three times in one shard
    db[7] = q[0]; db[n_local * (world - 1) + 9] = q[0]
    db[n_local - 1] = q[1]; db[n_local] = q[1]
    db[100] = q[2]; db[200] = q[2]; db[300] = q[2]
    m = orbx.ORBmatcher(0.7, True, max_keypoints=2048, device=local)
    match_fn, merge_fn = gpu_fns(m)
    for ragged in (0, 11):
        n_valid = n_local - ragged * (rank + 1)
        keep = np.ones(len(db), bool)
        for r in range(world):
            keep[r * n_local + n_local - ragged * (r + 1):(r + 1) * n_local] = False
        real = np.nonzero(keep)[0]
        wi, wd = O.bf_knn2(q, db[real])
        want_i = np.where(wi >= 0, real[np.maximum(wi, 0)], -1).astype(np.int32)
        shard = torch.from_numpy(db[rank * n_local:(rank + 1) * n_local].copy()).cuda()
        if ragged:
            shard[n_valid:] = torch.from_numpy(q[0]).cuda()          # padding rows hold a query: they must never be matched
        sdb = ShardedDescriptorDB(shard, match_fn, merge_fn, n_valid=n_valid)
        queries = torch.from_numpy(q.copy()).cuda() if rank == 0 else torch.zeros((len(q), 32), dtype=torch.uint8, device="cuda")
        sdb.broadcast_queries(queries, src=0)
        for name, fn in (("top2", sdb.knn2_allgather_top2), ("db", sdb.knn2_allgather_db)):
            idx, d = fn(queries)
            torch.cuda.synchronize()
            assert np.array_equal(idx.cpu().numpy(), want_i), (rank, ragged, name, "indices differ from the oracle")
            assert np.array_equal(d.cpu().numpy(), wd), (rank, ragged, name, "distances differ from the oracle")
    dist.barrier()
    if rank == 0:
        print("nccl c5 ok: world=%d, both exchanges == oracle (even and ragged shards)" % world, flush=True)
    m.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
