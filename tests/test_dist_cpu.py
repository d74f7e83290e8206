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


# ---- the volume-sharded layout and the all-to-all-v exchange (necat_b200/pipeline.py)
def test_mapping_plan_covers_every_volume_pair_once():
    from necat_b200 import pipeline
    for n_vol in (1, 2, 3, 4, 7):
        pairs = {(i, j) for i in range(n_vol) for j in range(i, n_vol)}
        for world in (1, 2, 3, 4, 8):
            plan = pipeline.mapping_plan(n_vol, world)
            seen = {}
            for r, items in enumerate(plan):
                for (i, j, s, shares) in items:
                    seen.setdefault((i, j), []).append((s, shares))
                    if n_vol >= world:
                        assert i % world == r               # index volumes round-robin (pm_worker.c:372 per volume)
            assert set(seen) == pairs
            for v in seen.values():                          # the shares of a pair partition its read range
                shares = v[0][1]
                assert sorted(s for s, _ in v) == list(range(shares)) and all(n == shares for _, n in v)


def test_correction_plan_owns_every_unit_once():
    from necat_b200 import pipeline
    for n_part in (1, 2, 3, 10):
        for world in (1, 2, 4, 8):
            S, units = pipeline.correction_plan(n_part, world)
            flat = [u for r in units for u in r]
            assert sorted(flat) == [(p, s) for p in range(n_part) for s in range(S)]
            assert S == max(1, world // n_part)


def _exchange_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    from necat_b200 import formats, pipeline
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    n_reads, part = 2500, 700                                 # 4 partitions over 2 ranks, and 1 partition shared by 2
    full = np.zeros((3001, 7), np.uint32)
    full[:, 0] = rng.integers(0, 2**29, 3001)
    full[:, 1] = rng.integers(0, n_reads, 3001)               # subject id
    full[:, 4] = rng.integers(0, n_reads, 3001)               # query id
    full[:, 2] = 10; full[:, 3] = 500; full[:, 5] = 20; full[:, 6] = 600
    mine = full[rank::world]                                   # what this rank's mapping "found"
    for ps in (part, 5000):
        got = pipeline.exchange_records(mine, n_reads, world, rank, torch.device("cpu"), ps)
        n_part = (n_reads + ps - 1) // ps
        S, units = pipeline.correction_plan(n_part, world)
        both = np.concatenate([full, formats.swap_roles(full)])
        exp = [both[(both[:, 1].astype(np.int64) // ps) == p] for (p, s) in units[rank]]
        exp = np.concatenate(exp) if exp else np.zeros((0, 7), np.uint32)
        if S > 1:                                              # a shared partition: every sharing rank holds all its records
            assert len(units[rank]) == 1
        assert got.shape == exp.shape, (ps, got.shape, exp.shape)
        assert np.array_equal(got[np.lexsort(got.T[::-1])], exp[np.lexsort(exp.T[::-1])])
    dist.barrier()
    dist.destroy_process_group()


def test_all_to_all_exchange_world2(tmp_path):
    world = 2
    port = 29700 + (os.getpid() % 1000)
    mp.spawn(_exchange_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
