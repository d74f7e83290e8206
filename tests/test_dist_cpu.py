"""CPU, world_size 2 over gloo: the multi-GPU plumbing of bench.py (necat_b200/dist.py) -- query-range
sharding, the all-gather of candidate records and the template ownership rule."""
import os

import numpy as np
import torch.multiprocessing as mp


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    from necat_b200 import dist as nbdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(99)
    full = rng.integers(0, 2**31, size=(1001, 7), dtype=np.uint32)            # the records "found" for all reads
    qid = full[:, 4] % 977
    lo, hi = nbdist.shard_range(977, rank, world)
    mine = full[(qid >= lo) & (qid < hi)]                                      # what this rank's query shard yields
    allrec = nbdist.all_gather_records(mine, torch.device("cpu"), world)
    assert allrec.shape == full.shape
    assert np.array_equal(allrec[np.lexsort(allrec.T[::-1])], full[np.lexsort(full.T[::-1])])
    ids = np.arange(977)
    own = nbdist.owns_template(ids, rank, world)
    np.save(os.path.join(tmp, f"own{rank}.npy"), own)
    # an empty shard must not break the exchange
    empty = nbdist.all_gather_records(np.zeros((0, 7), np.uint32) if rank == 1 else mine, torch.device("cpu"), world)
    assert empty.shape[0] == ((qid >= nbdist.shard_range(977, 0, world)[0]) & (qid < nbdist.shard_range(977, 0, world)[1])).sum()
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_and_ownership_world2(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    own = [np.load(os.path.join(str(tmp_path), f"own{r}.npy")) for r in range(world)]
    assert np.all(own[0] ^ own[1])                                              # every template owned exactly once


def test_shard_ranges_cover_everything():
    from necat_b200 import dist as nbdist
    for n in (0, 1, 7, 23762):
        for world in (1, 2, 4, 8):
            r = [nbdist.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
